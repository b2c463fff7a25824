// gk_encode.cu -- K4 `encode_states`: packed boards -> the 6 x 15 x 15 uint8 feature planes of
// Board.encoded_states() (reference core/py_ext/src/game_ext.hpp:87-104), optionally in the 8
// rotations / reflections of augment_game_data (network/data_helper.py:36-55), for self-play data.
//   planes: [stones of the side to move, stones of the opponent, empty cells, last move (one-hot),
//            second-to-last move (one-hot), 1 if black is to move]
//   variants (augment = 1), in the reference's order: for i in 0..3: rot90(i), fliplr(rot90(i))
// Pure data movement: 64 B in, 1 350 B (or 10 800 B) out per position -- HBM-write bound.
//
// One warp per position, everything as BIT planes until the last moment:
//   1. lanes 0..14 hold the 15-bit rows of "my stones" / "opponent's stones" (and the two one-hot planes);
//      their transposes come from 15 ballots each;
//   2. every variant of every plane is one of {plane, transpose} with rows and / or bits reversed, so a
//      variant row is a shuffle plus an optional bit reversal; the 15-bit rows are OR-ed into the position's
//      output BIT stream in shared memory (10 800 bits for 8 variants);
//   3. the stream is expanded to bytes 16 at a time (4 bits -> 4 bytes by one multiply and one mask) and
//      written with 128-bit stores.  No byte ever goes through shared memory.
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kPlaneBytes = 6 * kCells;        // 1350
constexpr int kWarps = 8;
constexpr int kStreamWords = 344;              // 8 x 1350 bits = 337.5 words, padded to a 16-byte multiple + slack for funnel reads

__device__ __forceinline__ uint32_t squeeze_even15(uint32_t x) {                // even bits of a 30-bit field -> 15 contiguous bits
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

// transpose of a 15 x 15 bit matrix held as rows in lanes 0..14: row j of the result = column j of the input
__device__ __forceinline__ uint32_t transpose15(uint32_t row, int lane) {
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < kWidth; ++j) {
        const uint32_t col = __ballot_sync(0xffffffffu, (row >> j) & 1u) & 0x7fffu;
        if (lane == j) t = col;
    }
    return t;
}

__device__ __forceinline__ uint32_t spread4(uint32_t nibble) {                  // 4 bits -> 4 bytes of 0 / 1
    return (nibble * 0x00204081u) & 0x01010101u;
}

__global__ void __launch_bounds__(kWarps * 32)
encode_states_kernel(EncodeArgs a) {
    __shared__ uint16_t s_perm[8 * kCells];                // (variant, cell) -> source cell, for the probabilities
    __shared__ __align__(16) uint32_t s_stream[kWarps][kStreamWords];
    const int variants = a.augment ? 8 : 1;
    if (a.probs) {
        for (int f = threadIdx.x; f < variants * kCells; f += blockDim.x) {
            const int v = f / kCells, i = f - v * kCells;
            const int r = i / kWidth, c = i - r * kWidth, cf = (v & 1) ? kWidth - 1 - c : c;   // fliplr after the rotation
            int r0, c0;
            switch (v >> 1) {                              // np.rot90(m, k)[r][c]
                case 0: r0 = r; c0 = cf; break;
                case 1: r0 = cf; c0 = kWidth - 1 - r; break;
                case 2: r0 = kWidth - 1 - r; c0 = kWidth - 1 - cf; break;
                default: r0 = kWidth - 1 - cf; c0 = r; break;
            }
            s_perm[f] = uint16_t(r0 * kWidth + c0);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* stream = s_stream[warp];
    const int out_bytes = variants * kPlaneBytes;
    for (long long b = (long long)blockIdx.x * kWarps + warp; b < a.n; b += (long long)gridDim.x * kWarps) {
        uint32_t w = 0;
        if (lane < kBoardWords) w = __ldg(a.boards + b * kBoardWords + lane);
        w &= ~((w >> 1) & 0x55555555u);                                       // a cell holding the invalid value 3 reads as white (as in ac_eval)
        if (lane == 14) w &= 3u;                                              // only cell 224 lives in word 14
        if (lane == 15) w = 0u;
        // stone counts decide the side to move: black iff #black == #white (Game.h:128, Game.cpp:52)
        const int nb = __reduce_add_sync(0xffffffffu, __popc(w & 0x55555555u & ~(w >> 1)));
        const int nw = __reduce_add_sync(0xffffffffu, __popc((w >> 1) & 0x55555555u & ~w));
        const bool black_to_move = nb == nw;
        // ---- 1. bit rows (lanes 0..14) and their transposes --------------------------------------------------
        const int y = lane < kHeight ? lane : 0, off = 30 * y;
        const uint32_t lo = __shfl_sync(0xffffffffu, w, off >> 5), hi = __shfl_sync(0xffffffffu, w, (off >> 5) + 1);
        const uint32_t v30 = __funnelshift_r(lo, hi, off & 31) & 0x3fffffffu;
        uint32_t blk = squeeze_even15(v30 & ~(v30 >> 1)), wht = squeeze_even15((v30 >> 1) & ~v30);
        if (lane >= kHeight) blk = wht = 0;
        const int last1 = a.last_moves ? a.last_moves[b * 2] : -1, last2 = a.last_moves ? a.last_moves[b * 2 + 1] : -1;
        uint32_t base[4], tr[4];                                              // mine, opponent's, last, second-to-last
        base[0] = black_to_move ? blk : wht;
        base[1] = black_to_move ? wht : blk;
        base[2] = (last1 >= 0 && last1 / kWidth == lane) ? 1u << (last1 % kWidth) : 0u;
        base[3] = (last2 >= 0 && last2 / kWidth == lane) ? 1u << (last2 % kWidth) : 0u;
        tr[0] = transpose15(base[0], lane);
        tr[1] = transpose15(base[1], lane);
        tr[2] = (last1 >= 0 && last1 % kWidth == lane) ? 1u << (last1 / kWidth) : 0u;
        tr[3] = (last2 >= 0 && last2 % kWidth == lane) ? 1u << (last2 / kWidth) : 0u;
        // ---- 2. the output bit stream ---------------------------------------------------------------------------
        for (int i = lane; i < kStreamWords / 4; i += 32) reinterpret_cast<uint4*>(stream)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint32_t side = black_to_move ? 0x7fffu : 0u;
        for (int v = 0; v < variants; ++v) {
            const int k = v >> 1;
            const bool use_t = k & 1, reversed = (k >= 2) != bool(v & 1);
            const int src_row = (k == 1 || k == 2) ? kHeight - 1 - y : y;     // lanes >= 15 read row 0 and are masked below
            uint32_t rows[6];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t r = __shfl_sync(0xffffffffu, use_t ? tr[p] : base[p], src_row);
                if (reversed) r = __brev(r) >> 17;
                rows[p < 2 ? p : p + 1] = r;
            }
            rows[2] = ~(rows[0] | rows[1]) & 0x7fffu;                         // empty cells
            rows[5] = side;
            if (lane < kHeight) {
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const int bit = (v * 6 + p) * kCells + lane * kWidth;
                    const uint32_t r = rows[p];
                    if (r) {
                        atomicOr(&stream[bit >> 5], r << (bit & 31));
                        if ((bit & 31) > 32 - kWidth) atomicOr(&stream[(bit >> 5) + 1], r >> (32 - (bit & 31)));
                    }
                }
            }
        }
        __syncwarp();
        // ---- 3. bits -> bytes, 128-bit stores.  The position's output range need not be 16-byte aligned
        //         (1350 = 84 x 16 + 6): byte-wise head up to the next boundary, 128-bit body, byte-wise tail
        uint8_t* out = a.planes + b * out_bytes;
        const int head = int((16u - unsigned(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
        const int body = (out_bytes - head) / 16;
        if (lane < head) out[lane] = (stream[lane >> 5] >> (lane & 31)) & 1u;
        for (int q = lane; q < body; q += 32) {
            const int bit = head + q * 16;
            const uint32_t m = __funnelshift_r(stream[bit >> 5], stream[(bit >> 5) + 1], bit & 31);   // 16 stream bits in the low half
            *reinterpret_cast<uint4*>(out + bit) = make_uint4(spread4(m & 15u), spread4((m >> 4) & 15u), spread4((m >> 8) & 15u),
                                                              spread4((m >> 12) & 15u));
        }
        for (int f = head + body * 16 + lane; f < out_bytes; f += 32) out[f] = (stream[f >> 5] >> (f & 31)) & 1u;
        if (a.probs) {                                                        // rot_probs / flip_probs: the cell permutation of the variant
            const float* pin = a.probs + b * kCells;
            float* pout = a.probs_out + b * variants * kCells;
            for (int f = lane; f < variants * kCells; f += 32) pout[f] = __ldg(pin + s_perm[f]);
        }
        __syncwarp();
    }
}

// Self-play bookkeeping: the positions a batch of games went through.  Game g started from boards0[g] and played
// moves[g][0 .. length[g]); the position BEFORE ply k goes to out_boards[starts[g] + k], with the two moves that led
// to it in out_last (what Board.encoded_states() reads from the move record, game_ext.hpp:96-101) and the game's
// outcome from the view of the side to move in out_z (the z of an AlphaZero-style sample).  One warp per game.
__global__ void expand_games_kernel(ExpandArgs a) {
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= a.n) return;
    uint32_t w = lane < kBoardWords ? __ldg(a.boards0 + g * kBoardWords + lane) : 0u;
    const int nb = __reduce_add_sync(0xffffffffu, __popc(w & 0x55555555u & ~(w >> 1)));
    const int nw = __reduce_add_sync(0xffffffffu, __popc((w >> 1) & 0x55555555u & ~w));
    uint32_t colour = nb == nw ? 1u : 2u;                                   // black moves first (Game.h:128)
    const int len = a.lengths[g];
    const long long start = a.starts[g];
    const int winner = a.winners ? a.winners[g] : 0;
    int last1 = -1, last2 = -1;
    for (int k = 0; k < len; ++k) {
        const long long o = start + k;
        if (lane < kBoardWords) a.out_boards[o * kBoardWords + lane] = w;
        if (a.out_last && lane < 2) a.out_last[o * 2 + lane] = (int16_t)(lane == 0 ? last1 : last2);
        if (a.out_z && lane == 0) a.out_z[o] = (int8_t)(colour == 1u ? winner : -winner);
        const int cell = a.moves[g * a.max_moves + k];
        if (lane == (cell >> 4)) w |= colour << ((cell & 15) * 2);
        last2 = last1; last1 = cell;
        colour ^= 3u;
    }
}

}  // namespace

cudaError_t launch_expand_games(const ExpandArgs& a, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    expand_games_kernel<<<(unsigned)((a.n + 7) / 8), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_encode(const EncodeArgs& a, int sm_count, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    const long long want = (a.n + kWarps - 1) / kWarps;
    const long long resident = (long long)sm_count * 8;
    encode_states_kernel<<<int(want < resident ? want : resident), kWarps * 32, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gk

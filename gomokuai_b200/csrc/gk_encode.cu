// gk_encode.cu -- K4 `encode_states`: packed boards -> the 6 x 15 x 15 uint8 feature planes of
// Board.encoded_states() (reference core/py_ext/src/game_ext.hpp:87-104), optionally in the 8
// rotations / reflections of augment_game_data (network/data_helper.py:36-55), for self-play data.
//   planes: [stones of the side to move, stones of the opponent, empty cells, last move (one-hot),
//            second-to-last move (one-hot), 1 if black is to move]
//   variants (augment = 1), in the reference's order: for i in 0..3: rot90(i), fliplr(rot90(i))
// Pure data movement: 64 B in, 1 350 B (or 10 800 B) out per position -- HBM-write bound.  One warp
// stages the six base planes as bytes in shared memory, then streams the variants out as 16-byte
// stores through a precomputed (variant, byte) -> base-byte table.
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kPlaneBytes = 6 * kCells;        // 1350
constexpr int kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32)
encode_states_kernel(EncodeArgs a) {
    __shared__ uint16_t s_map[8 * kPlaneBytes];            // (variant, byte) -> index into the base planes
    __shared__ __align__(16) uint8_t s_base[kWarps][kPlaneBytes + 10];
    const int variants = a.augment ? 8 : 1;
    for (int f = threadIdx.x; f < variants * kPlaneBytes; f += blockDim.x) {
        const int v = f / kPlaneBytes, rem = f - v * kPlaneBytes, p = rem / kCells, i = rem - p * kCells;
        const int r = i / kWidth, c = i - r * kWidth, cf = (v & 1) ? kWidth - 1 - c : c;   // fliplr after the rotation
        int r0, c0;
        switch (v >> 1) {                                  // np.rot90(m, k)[r][c]
            case 0: r0 = r; c0 = cf; break;
            case 1: r0 = cf; c0 = kWidth - 1 - r; break;
            case 2: r0 = kWidth - 1 - r; c0 = kWidth - 1 - cf; break;
            default: r0 = kWidth - 1 - cf; c0 = r; break;
        }
        s_map[f] = uint16_t(p * kCells + r0 * kWidth + c0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* base = s_base[warp];
    const long long out_bytes = (long long)variants * kPlaneBytes;
    for (long long b = (long long)blockIdx.x * kWarps + warp; b < a.n; b += (long long)gridDim.x * kWarps) {
        uint32_t w = 0;
        if (lane < kBoardWords) w = __ldg(a.boards + b * kBoardWords + lane);
        // stone counts decide the side to move: black iff #black == #white (Game.h:128, Game.cpp:52)
        uint32_t v = lane == 14 ? (w & 3u) : lane == 15 ? 0u : w;
        const int blk = __popc(v & 0x55555555u & ~(v >> 1)), wht = __popc((v >> 1) & 0x55555555u & ~v);
        const int nb = __reduce_add_sync(0xffffffffu, blk), nw = __reduce_add_sync(0xffffffffu, wht);
        const uint32_t mine = nb == nw ? 1u : 2u;
        const int last1 = a.last_moves ? a.last_moves[b * 2] : -1, last2 = a.last_moves ? a.last_moves[b * 2 + 1] : -1;
        for (int c0 = 0; c0 < kCells; c0 += 32) {                          // warp-uniform trip count (shuffles inside)
            const int c = c0 + lane;
            const uint32_t word = __shfl_sync(0xffffffffu, w, (c >> 4) & 15);
            const uint32_t val = (word >> ((c & 15) * 2)) & 3u;
            if (c < kCells) {
                base[c] = val == mine;
                base[kCells + c] = val == (3u - mine);
                base[2 * kCells + c] = val == 0u;
                base[3 * kCells + c] = c == last1;
                base[4 * kCells + c] = c == last2;
                base[5 * kCells + c] = mine == 1u;
            }
        }
        __syncwarp();
        // the position's output range need not be 16-byte aligned (1350 = 84 x 16 + 6): byte-wise head up
        // to the next 16-byte boundary, 128-bit body, byte-wise tail
        uint8_t* out = a.planes + b * out_bytes;
        const int head = int((16u - unsigned(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
        const int body = (int(out_bytes) - head) / 16;
        if (lane < head) out[lane] = base[s_map[lane]];
        for (int q = lane; q < body; q += 32) {
            uint32_t pack[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int f = head + q * 16 + k * 4;
                pack[k] = uint32_t(base[s_map[f]]) | uint32_t(base[s_map[f + 1]]) << 8 | uint32_t(base[s_map[f + 2]]) << 16 |
                          uint32_t(base[s_map[f + 3]]) << 24;
            }
            *reinterpret_cast<uint4*>(out + head + q * 16) = make_uint4(pack[0], pack[1], pack[2], pack[3]);
        }
        for (int f = head + body * 16 + lane; f < out_bytes; f += 32) out[f] = base[s_map[f]];
        if (a.probs) {                                                    // rot_probs / flip_probs: plane 0's cell permutation
            const float* pin = a.probs + b * kCells;
            float* pout = a.probs_out + b * variants * kCells;
            for (int f = lane; f < variants * kCells; f += 32) {
                const int v = f / kCells;
                pout[f] = __ldg(pin + s_map[v * kPlaneBytes + (f - v * kCells)]);
            }
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_encode(const EncodeArgs& a, int sm_count, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    const long long want = (a.n + kWarps - 1) / kWarps;
    const long long resident = (long long)sm_count * 6;
    encode_states_kernel<<<int(want < resident ? want : resident), kWarps * 32, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gk

// gk_rollout_warp.cu -- K2w `rollout_warp`: the LATENCY form of K2 (gk_rollout.cu), one WARP per rollout.
//
// Same job and same results as rollout_small_kernel -- Default::RandomRollout (reference include/algorithms/
// MonteCarlo.hpp:37-47) = per move Board::getRandomMove (src/Game.cpp:64-73), Board::applyMove (:37-47) and
// Board::checkGameEnd (:88-136), c_rollouts times per position with the Philox streams of include/gomoku_b200.h -- but
// built for the case a tree search actually has: a few hundred leaves x 5 playouts, where the GPU is empty and the only
// cost is the SERIAL chain of one game's moves.  K2 keeps a rollout in one thread: every move is a chain of three
// dependent shared-memory reads (row -> cell table -> diagonal slot), a Philox block every fourth move, ~330 cycles.
// Here the 72 line slots of a rollout live in the REGISTERS of a warp (lane l owns slots l, l + 32, l + 64): the row the
// move lands in comes from its owner by one shuffle, every lane then updates the slots it owns with plain arithmetic (a
// slot is "a x + b y == c", the stone's bit is "min(p, q)"), and the win test is one vote.  No shared memory in the loop,
// one Philox block per LANE gives the random numbers of 128 moves at once.
//
// One CTA per position, one warp per rollout (rollouts_per_pos <= 32).  Used for batches of up to ten warps per SM
// (rollout_warp_fits): a move costs ~90 warp instructions PER ROLLOUT here, so beyond that the warps wait for issue slots and
// K2's thread-per-rollout forms win.
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kLineSlots = 72;          // 15 rows, 15 columns, 21 + 21 diagonals of length >= 5 (the layout of gk_rollout.cu)
constexpr uint32_t kFull = 0xffffffffu;

// Philox4x32-10 (the generator of gk_rollout.cu; restated here so that file stays as profiled)
__device__ __forceinline__ void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                             uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// bit i of the result is set iff bits i-4..i of v are: black stones sit in bits 0..14, white in 16..30, bit 15 is never
// set, so a run cannot leak from one colour's half into the other's
__device__ __forceinline__ uint32_t run_of_five(uint32_t v) {
    uint32_t t = v & (v << 1);
    t &= t << 2;
    return t & (v << 4);
}

// A line slot as arithmetic: cell (x, y) lies on it iff ax * x + by * y == c, and its stone is bit min(mx * x + cx, my * y + cy).
struct Line {
    int ax, by, c, mx, cx, my, cy;
};

__device__ __forceinline__ Line line_of(int slot) {
    Line g{};
    if (slot < 15) {                     // row y = slot: bit x
        g.ax = 0; g.by = 1; g.c = slot; g.mx = 1; g.cx = 0; g.my = 0; g.cy = 99;
    } else if (slot < 30) {              // column x = slot - 15: bit y
        g.ax = 1; g.by = 0; g.c = slot - 15; g.mx = 0; g.cx = 99; g.my = 1; g.cy = 0;
    } else if (slot < 51) {              // (+1,+1) diagonal x - y = k: bit min(x, y)
        const int k = slot - 40;
        g.ax = 1; g.by = -1; g.c = k; g.mx = 1; g.cx = 0; g.my = 1; g.cy = 0;
    } else if (slot < kLineSlots) {      // (-1,+1) diagonal x + y = t: bit min(14 - x, y)
        const int t = slot - 47;
        g.ax = 1; g.by = 1; g.c = t; g.mx = -1; g.cx = 14; g.my = 1; g.cy = 0;
    } else {                             // lanes 8..31 own two slots only: a third that no cell lies on
        g.ax = 0; g.by = 0; g.c = 1; g.mx = 0; g.cx = 0; g.my = 0; g.cy = 0;
    }
    return g;
}

// kMaxThreads: 256 for up to 8 rollouts per position (the search's 5), 1024 otherwise -- the tighter bound lets ptxas use
// more registers, i.e. interleave the three slot updates of a move instead of running them one after the other.
// A single leaf (the single-tree search: one launch per playout) rides in the kernel's parameters, which arrive with the
// launch: reading it through the pointer would be a round trip over PCIe to the caller's page-locked memory first.
struct InlineBoard { uint32_t w[kBoardWords]; int valid; };

template <bool kMoves, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads) rollout_warp_kernel(RolloutArgs a, int32_t* __restrict__ out, const InlineBoard b0) {
    __shared__ uint32_t s_board[kBoardWords], s_cnt[3], s_slot[kLineSlots];
    const int pos = blockIdx.x, tid = threadIdx.x, lane = tid & 31, roll = tid >> 5;
    if (tid < kBoardWords) s_board[tid] = b0.valid ? b0.w[tid] : a.boards[(size_t)pos * kBoardWords + tid];
    if (tid < 3) s_cnt[tid] = 0;
    for (int i = tid; i < kLineSlots; i += blockDim.x) s_slot[i] = 0;
    __syncthreads();
    // ---- the position's 72 line words, built once by the whole CTA: a stone goes onto its four lines ----------
    for (int c = tid; c < kCells; c += blockDim.x) {
        const uint32_t v = (s_board[c >> 4] >> ((c & 15) * 2)) & 3u;
        if (v == 1u || v == 2u) {
            const int y = c / 15, x = c - 15 * y;
            const uint32_t half = v == 1u ? 1u : 0x10000u;
            atomicOr(&s_slot[y], half << x);
            atomicOr(&s_slot[15 + x], half << y);
            const uint32_t k = uint32_t(x - y + 10), t = uint32_t(x + y - 4);
            if (k <= 20u) atomicOr(&s_slot[30 + k], half << min(x, y));
            if (t <= 20u) atomicOr(&s_slot[51 + t], half << min(14 - x, y));
        }
    }
    __syncthreads();
    // ---- this lane's three slots: geometry, and the stones already on them --------------------------------
    Line g[3];
    uint32_t sw[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g[k] = line_of(lane + 32 * k);
        sw[k] = lane + 32 * k < kLineSlots ? s_slot[lane + 32 * k] : 0u;
    }
    // ---- the position: stones, side to move, decided already? (the image meta words of gk_rollout.cu) -----
    uint32_t fives = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) fives |= run_of_five(sw[k]);
    const bool black_five = __any_sync(kFull, (fives & 0xffffu) != 0), white_five = __any_sync(kFull, (fives >> 16) != 0);
    const bool is_row = lane < 15;
    const uint32_t blk = __reduce_add_sync(kFull, is_row ? __popc(sw[0] & 0x7fffu) : 0);
    const uint32_t wht = __reduce_add_sync(kFull, is_row ? __popc(sw[0] >> 16) : 0);
    uint32_t left = uint32_t(kCells) - blk - wht;
    uint32_t colour = blk == wht ? 0u : 1u;                      // black moves first, Game.h:128; Game.cpp:52
    uint32_t played = 0;
    int result;                                                   // 0 white won, 1 draw, 2 black won (the wdb columns)
    if (black_five || white_five || left == 0) {
        result = black_five ? 2 : white_five ? 0 : 1;
    } else {
        const uint32_t c1 = uint32_t(roll), c2 = uint32_t(a.pos_base) + uint32_t(pos);
        uint8_t* trace = kMoves && a.moves ? a.moves + (size_t(pos) * a.rollouts_per_pos + roll) * kCells : nullptr;
        // The loop is software-pipelined so that the only chain carried from one move to the next is "row word -> first
        // empty cell -> stone into the NEXT move's row word" (~9 dependent instructions):
        //  * the start index of move m+1 does not depend on the board, so its row word is fetched from the owner lane at
        //    the top of iteration m, BEFORE move m's stone is on the board, and patched with that stone if it is the same
        //    row; the index itself (one byte of some lane's Philox block) is fetched an iteration before that;
        //  * the fallback row -- the next row that still has an empty cell, when the start index lies behind the last
        //    empty cell of its row -- is looked up only when needed (a uniform branch: one ballot of the row owners);
        //  * the win vote of move m is looked at one iteration later: move m+1 is played speculatively and thrown away
        //    when move m turns out to have ended the game (registers only, nothing to undo).
        auto draw_block = [&](uint32_t first_move) {              // lane j: the four start indices of Philox block first_move / 4 + j, one byte each
            uint32_t rnd[4];
            philox_block((first_move >> 2) + uint32_t(lane), c1, c2, a.ctr_hi, a.key_lo, a.key_hi, rnd);
            return __umulhi(rnd[0], uint32_t(kCells)) | __umulhi(rnd[1], uint32_t(kCells)) << 8 |
                   __umulhi(rnd[2], uint32_t(kCells)) << 16 | __umulhi(rnd[3], uint32_t(kCells)) << 24;
        };
        asm volatile("" : "+r"(played));                         // opaque: keeps the move counters out of the uniform datapath,
                                                                  // whose results reach the vector registers a dozen cycles late
        uint32_t won_prev = 0;
        for (;;) {                                                // one pass = one draw of 128 start indices
            if (won_prev || left == 0) { result = won_prev ? (colour ? 2 : 0) : 1; break; }   // Game.cpp:125-132; the last mover is colour ^ 1
            const uint32_t rpack = draw_block(played);
            auto start_index = [&](uint32_t move) {               // of a move of this draw (anything for a move beyond it)
                return (__shfl_sync(kFull, rpack, (move >> 2) & 31u) >> ((move & 3u) * 8u)) & 0xffu;
            };
            const uint32_t r0 = start_index(played);
            uint32_t y = (r0 * 137u) >> 11;
            uint32_t xmask = kFull << (r0 - 15u * y);             // cells of row y at or after the start index
            uint32_t w = __shfl_sync(kFull, sw[0], y);
            uint32_t rn = start_index(played + 1u);
            uint32_t cnt = min(left, 128u);
            uint32_t stop;
            do {
                // ---- move m + 1: its row as it is BEFORE move m; move m + 2: its start index ----------------------------
                const uint32_t yn = (rn * 137u) >> 11, xn = rn - 15u * yn;
                uint32_t wn = __shfl_sync(kFull, sw[0], yn);
                rn = start_index(played + 2u);
                // ---- move m: Board::getRandomMove = the first empty cell at or after the start index, cyclically --------
                uint32_t avail = ~(w | (w >> 16)) & 0x7fffu & xmask;
                if (avail == 0) {                                 // nothing left in row y from there on: all of the next row that has a cell
                    const uint32_t occ0 = (sw[0] | (sw[0] >> 16)) & 0x7fffu;   // (the rows that still have one: asked of their owners now)
                    const uint32_t rowmask = __ballot_sync(kFull, lane < 15 && occ0 != 0x7fffu);
                    uint32_t m = rowmask & ~((2u << y) - 1u);
                    if (m == 0) m = rowmask;
                    y = 31u - __clz(m & (0u - m));
                    w = __shfl_sync(kFull, sw[0], y);
                    avail = ~(w | (w >> 16)) & 0x7fffu;
                }
                const uint32_t xbit = avail & (0u - avail);
                const uint32_t stone = 1u + 0xffffu * colour;
                if (yn == y) wn |= xbit * stone;
                const uint32_t x = 31u - __clz(xbit);
                if (kMoves && trace && lane == 0) trace[played] = uint8_t(y * 15u + x);
                // ---- Board::applyMove + checkGameEnd: every lane puts the stone on the slots it owns that pass through (x, y)
                const int xi = int(x), yi = int(y);
                bool on[3];
                int bit[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) on[k] = g[k].ax * xi + g[k].by * yi == g[k].c;
#pragma unroll
                for (int k = 0; k < 3; ++k) bit[k] = min(g[k].mx * xi + g[k].cx, g[k].my * yi + g[k].cy);
#pragma unroll
                for (int k = 0; k < 3; ++k) sw[k] = on[k] ? sw[k] | (stone << bit[k]) : sw[k];
                // No line held five before this move (the game would be over), so any run now is the mover's and is new.
                // One PRMT packs the mover's halves of two slots into one word (bit 15 of a half is never set).
                const uint32_t pair = __byte_perm(sw[0], sw[1], 0x5410u + 0x2222u * colour);
                const uint32_t five = run_of_five(pair) | run_of_five(sw[2]);
                const uint32_t won_now = __ballot_sync(kFull, five != 0);
                stop = won_prev;                                  // move m - 1 had won: this move never happened (undone below)
                won_prev = won_now;
                left -= 1;
                played += 1;
                colour ^= 1u;
                y = yn; xmask = kFull << xn; w = wn;
            } while (!stop && --cnt != 0);
            if (stop) { played -= 1; result = colour ? 0 : 2; break; }       // colour was flipped once too often: winner = colour
        }
    }
    if (lane == 0) {
        atomicAdd(&s_cnt[result], 1u);
        if (kMoves) {
            const size_t gidx = size_t(pos) * a.rollouts_per_pos + roll;
            if (a.winners) a.winners[gidx] = int8_t(result == 2 ? 1 : result == 0 ? -1 : 0);
            if (a.lengths) a.lengths[gidx] = uint8_t(played);
        }
    }
    __syncthreads();
    if (out && tid < 3) out[(size_t)pos * 3 + tid] = int(s_cnt[tid]);
}

}  // namespace

bool rollout_warp_fits(int n, int rollouts_per_pos, int sm_count) {
    // A move costs ~90 warp instructions PER ROLLOUT here against ~85 per 32 rollouts in K2, so the form only pays while
    // the warps are few enough that none waits for an issue slot: about 2-3 warps per SM sub-partition.
    return n > 0 && rollouts_per_pos > 0 && rollouts_per_pos <= 32 &&
           (long long)n * rollouts_per_pos <= (long long)sm_count * 10;
}

cudaError_t launch_rollout_warp(const RolloutArgs& a, int32_t* out, cudaStream_t stream, const uint32_t* h_board0) {
    if (a.n <= 0 || a.rollouts_per_pos <= 0) return cudaSuccess;
    if (a.rollouts_per_pos > 32) return cudaErrorInvalidValue;
    const int threads = a.rollouts_per_pos * 32;
    InlineBoard b0{};
    if (h_board0 && a.n == 1) {
        for (int i = 0; i < kBoardWords; ++i) b0.w[i] = h_board0[i];
        b0.valid = 1;
    }
    const bool trace = a.moves || a.winners || a.lengths;
    if (threads <= 256) {
        if (trace) rollout_warp_kernel<true, 256><<<a.n, threads, 0, stream>>>(a, out, b0);
        else rollout_warp_kernel<false, 256><<<a.n, threads, 0, stream>>>(a, out, b0);
    } else {
        if (trace) rollout_warp_kernel<true, 1024><<<a.n, threads, 0, stream>>>(a, out, b0);
        else rollout_warp_kernel<false, 1024><<<a.n, threads, 0, stream>>>(a, out, b0);
    }
    return cudaGetLastError();
}

}  // namespace gk

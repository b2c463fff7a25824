// gk_rollout.cu -- K2 `rollout`: batched random playouts, one thread per rollout.
//
// Replaces Default::RandomRollout / Simulate (reference include/algorithms/MonteCarlo.hpp:
// 37-47,83-88), i.e. per move: Board::getRandomMove (src/Game.cpp:64-73: start index r, then
// the first empty cell at or after r, cyclically), Board::applyMove (:37-47) and
// Board::checkGameEnd (:88-136: five-or-more through the new stone, else draw when full).
//
// State of one rollout = 72 "line slots" of 32 bits in shared memory, laid out
// [slot][thread] so that every access of a warp is bank-conflict free whatever slots the
// lanes touch: 15 rows, 15 columns and the 2 x 21 diagonals of length >= 5, each holding the
// black stones of that line in bits 0..14 and the white stones in bits 16..30.  A stone is
// therefore stored four times, and the win test of a move is four independent
// "read slot, or the bit in, store, m & m>>1 & m>>2 & m>>3 & m>>4" sequences in registers.
// Legal-move selection works on the row slots plus a 15-bit "row still has an empty cell" mask
// kept in a register: at most two shared-memory probes per move, no loop.
//
// Threads are persistent: a finished lane waits for the next refill point (every kRefill = 16
// steps, a multiple of 4 so that all lanes draw a fresh Philox4x32-10 block on the same steps),
// takes the next rollout ticket and re-copies the position's 72-word slot image.
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kSlots = 72;
constexpr int kImageWords = 80;          // 72 slots + meta, 320 B per position
constexpr int kMetaInfo = 72;            // empties | to_move << 8 | decided << 9 | winner code << 10
constexpr int kMetaRows = 73;            // rows that still have an empty cell
constexpr int kThreads = 256;            // per CTA: 256 x 73 x 4 B = 73 KiB of slot state, 3 CTAs per SM
constexpr int kRefillDefault = 16;       // move steps between refill points (multiple of 4)
constexpr int kTicketBlock = 256;        // consecutive rollouts a CTA claims at a time

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ bool has_five(uint32_t m) {      // five or more consecutive bits
    uint32_t t = m & (m >> 1);
    t &= t >> 2;
    return (t & (m >> 4)) != 0;
}

// ---- position -> slot image ------------------------------------------------------------------
// one warp per position; lane l builds slots l, l+32, l+64 (and clears the position's wdb counters)
// `b`: the 16 board words (global or shared memory); `img`: kImageWords words (global or shared)
__device__ __forceinline__ void build_image(const uint32_t* b, uint32_t* img, int lane) {
    uint32_t any_five = 0;                                   // bit 0: black, bit 1: white
    for (int slot = lane; slot < kSlots; slot += 32) {
        int cell0, stride, len;
        if (slot < 15) { cell0 = slot * 15; stride = 1; len = 15; }
        else if (slot < 30) { cell0 = slot - 15; stride = 15; len = 15; }
        else if (slot < 51) { const int k = slot - 40; cell0 = k > 0 ? k : -k * 15; stride = 16; len = 15 - (k > 0 ? k : -k); }
        else { const int s = slot - 47; const int x0 = s < 14 ? s : 14; cell0 = (s - x0) * 15 + x0; stride = 14; len = (s < 15 ? s : 28 - s) + 1; }
        uint32_t w = 0;
        for (int i = 0; i < len; ++i) {
            const int c = cell0 + i * stride;
            const uint32_t v = (b[c >> 4] >> ((c & 15) * 2)) & 3u;
            if (v == 1u) w |= 1u << i;
            else if (v == 2u) w |= 1u << (16 + i);
        }
        img[slot] = w;
        if (has_five(w & 0x7fffu)) any_five |= 1u;
        if (has_five(w >> 16)) any_five |= 2u;
    }
    any_five = __reduce_or_sync(0xffffffffu, any_five);
    // rows: empties, stone counts
    uint32_t blk = 0, wht = 0, rowmask = 0;
    if (lane < 15) {
        uint32_t w = 0;
        for (int i = 0; i < 15; ++i) {
            const int c = lane * 15 + i;
            const uint32_t v = (b[c >> 4] >> ((c & 15) * 2)) & 3u;
            if (v == 1u) { w |= 1u << i; ++blk; }
            else if (v == 2u) { w |= 1u << i; ++wht; }
        }
        if (w != 0x7fffu) rowmask = 1u << lane;
    }
    blk = __reduce_add_sync(0xffffffffu, blk);
    wht = __reduce_add_sync(0xffffffffu, wht);
    rowmask = __reduce_or_sync(0xffffffffu, rowmask);
    if (lane == 0) {
        const uint32_t empties = kCells - blk - wht;
        const uint32_t to_move = blk == wht ? 0u : 1u;         // black moves first, Game.h:128; Game.cpp:52
        const uint32_t decided = (any_five || empties == 0) ? 1u : 0u;
        const uint32_t wcode = (any_five & 1u) ? 1u : (any_five & 2u) ? 2u : 0u;
        img[kMetaInfo] = empties | to_move << 8 | decided << 9 | wcode << 10;
        img[kMetaRows] = rowmask;
    }
}

__global__ void build_images_kernel(const uint32_t* __restrict__ boards, int n, uint32_t* __restrict__ images,
                                    int32_t* __restrict__ wdb) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    if (wdb && lane < 3) wdb[(size_t)warp * 3 + lane] = 0;
    build_image(boards + (size_t)warp * kBoardWords, images + (size_t)warp * kImageWords, lane);
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Lane {
    uint32_t rowmask;      // rows with at least one empty cell
    uint32_t left;         // empty cells left
    uint32_t colour;       // 0 black to move, 1 white
    uint32_t start;        // empty cells when the rollout started: moves played = start - left
    uint32_t pos, roll;    // position index (batch-local), rollout index within the position
};

// raw five-in-a-row detector on a whole slot word: bit i of the result is set iff bits i-4..i are.
// Bit 15 of a 16-bit half is never set, so runs never leak from one half into the other.  Left shifts on purpose:
// they compile to IMAD.SHL on the FMA pipe, and this kernel is bound by the ALU pipe (LOP3 / SHF / PRMT).
__device__ __forceinline__ uint32_t five_bits(uint32_t v) {
    uint32_t t = v & (v << 1);
    t &= t << 2;
    return t & (v << 4);
}

// Outcome counters of a lane, packed in one register: black wins in bits 0..9, white wins in 10..19, draws in
// 20..29 (flushed before a field can reach 1024).
constexpr uint32_t kIncBlack = 1u, kIncWhite = 1u << 10, kIncDraw = 1u << 20;

// per-cell diagonal addressing, one byte each: slot of the (+1,+1) diagonal, slot of the (-1,+1) diagonal; cells on
// diagonals shorter than five map to the scratch slot.  Their bit positions min(x, y) / min(14 - x, y) are <= 3
// there, so the scratch slot only ever holds bits 0..3 of each half and can never show a run of five.
__device__ __forceinline__ uint32_t cell_lut_entry(uint32_t c) {
    const uint32_t y = c / 15u, x = c - 15u * y;
    const uint32_t k = x - y + 10u, t = x + y - 4u;
    const uint32_t dslot = k <= 20u ? 30u + k : uint32_t(kSlots);
    const uint32_t aslot = t <= 20u ? 51u + t : uint32_t(kSlots);
    return dslot | aslot << 8;
}

// One move of an active rollout.  Returns the outcome increment: 0 = game goes on, else kIncBlack / kIncWhite / kIncDraw.
// Stone masks are built as (1 << position) * base with base = 1 (black half) or 65536 (white half): the multiply
// runs on the FMA pipe.
// shared-window loads / stores on 32-bit addresses held in registers (a generic pointer into shared memory makes the
// compiler re-derive the window base with four uniform-datapath instructions per access)
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

constexpr uint32_t kSlotStride = kThreads * 4u;                                   // bytes between consecutive slots of one thread (the big kernel)

__device__ __forceinline__ uint32_t play_move(uint32_t my /* shared address of slots[0][tid] */, uint32_t cell_addr /* of the cell LUT */,
                                              Lane& L, uint32_t r, const uint32_t kSlotStride = kThreads * 4u /* bytes between a thread's slots */,
                                              uint32_t* cell_out = nullptr /* the cell played (trace variants only; dead code otherwise) */) {
    uint32_t y = (r * 137u) >> 11, x = r - 15u * y;                              // r / 15, r % 15 for r < 225
    // Both candidate rows are read up front -- row y, and the next row after it that still has an empty cell (known from the
    // row mask, cyclically) -- so the fallback needs no second, dependent probe and no divergent branch: in most steps
    // SOME lane of the warp lands behind its row's last empty cell.
    uint32_t m = L.rowmask & ~((2u << y) - 1u);
    if (m == 0) m = L.rowmask;
    const uint32_t y2 = 31u - __clz(m & (0u - m));
    uint32_t w = lds32(my + y * kSlotStride);
    const uint32_t w2 = lds32(my + y2 * kSlotStride);
    uint32_t empty = ~(w | (w >> 16)) & 0x7fffu;                                 // empty cells of the row
    uint32_t avail = empty & (0xffffffffu << x);
    if (avail == 0) {                                                            // first empty cell after r: all of row y2's
        y = y2;
        w = w2;
        avail = empty = ~(w2 | (w2 >> 16)) & 0x7fffu;
    }
    const uint32_t xbit = avail & (0u - avail), ybit = 1u << y;                  // 1 << x, 1 << y
    x = 31u - __clz(xbit);
    uint32_t lut;                                                                // s_cell[15 y + x]; the table is read-only after the CTA's first barrier
    asm("ld.shared.u32 %0, [%1];" : "=r"(lut) : "r"(cell_addr + (y * 15u + x) * 4u));
    if (cell_out) *cell_out = y * 15u + x;
    const uint32_t base = 1u + 0xffffu * L.colour;
    // row
    w |= xbit * base;
    sts32(my + y * kSlotStride, w);
    if ((empty & (empty - 1u)) == 0) L.rowmask ^= ybit;                          // that was the row's last empty cell
    // column
    const uint32_t pc = my + (15u + x) * kSlotStride;
    const uint32_t vc = lds32(pc) | ybit * base;
    sts32(pc, vc);
    // the two diagonals: bit min(x, y) of (+1,+1), bit min(14 - x, y) of (-1,+1)
    const uint32_t pd = my + __byte_perm(lut, 0u, 0x4440u) * kSlotStride;
    const uint32_t vd = lds32(pd) | min(xbit, ybit) * base;
    sts32(pd, vd);
    const uint32_t pa = my + __byte_perm(lut, 0u, 0x4441u) * kSlotStride;
    const uint32_t va = lds32(pa) | min(0x4000u >> x, ybit) * base;
    sts32(pa, va);
    // win test on the mover's halves only, two lines per register: one PRMT packs the 16-bit halves of two slots
    const uint32_t sel = 0x5410u + 0x2222u * L.colour;
    const uint32_t fives = five_bits(__byte_perm(w, vc, sel)) | five_bits(__byte_perm(vd, va, sel));
    L.left -= 1;
    uint32_t inc = L.left == 0 ? kIncDraw : 0u;                                  // Game.cpp:129-132
    if (fives) inc = kIncBlack + (kIncWhite - kIncBlack) * L.colour;             // winner = player of the last stone, Game.cpp:125-128
    L.colour ^= 1u;
    return inc;
}

// kTrace: the caller wants per-rollout winners / lengths (tests, tooling); the counting-only kernel has none of that code
template <bool kInjected, int kRefill, bool kTrace>
__global__ void __launch_bounds__(kThreads, 3)
rollout_kernel(RolloutArgs a, const uint32_t* __restrict__ images) {
    extern __shared__ __align__(16) uint32_t s_slots[];                          // [kSlots + 1][kThreads]
    __shared__ uint32_t s_ticket;
    __shared__ uint32_t s_cell[kCells];
    if (threadIdx.x == 0) s_ticket = 0;
    if (threadIdx.x < kCells) s_cell[threadIdx.x] = cell_lut_entry(threadIdx.x);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, lt = lanemask_lt();
    uint32_t cell_addr = (uint32_t)__cvta_generic_to_shared(s_cell);             // kept in a register: re-deriving it costs 5 uniform
    asm volatile("" : "+r"(cell_addr));                                          // instructions per move
    uint32_t my = (uint32_t)__cvta_generic_to_shared(s_slots + threadIdx.x);     // shared address of slots[0][tid]
    asm volatile("" : "+r"(my));
    sts32(my + kSlots * kSlotStride, 0u);                                        // scratch slot: must never hold a run
    const uint32_t total = uint32_t(a.n) * uint32_t(a.rollouts_per_pos);

    Lane L{};
    bool active = false, dead = false;
    uint32_t acc = 0, acc_n = 0;                                                 // packed outcomes / rollouts started since the last flush, of position acc_pos
    uint32_t acc_pos = 0xffffffffu;
    uint32_t rnd[4] = { 0, 0, 0, 0 };
    const uint8_t* inj = nullptr;

    // Branch-free on the hot path: one add per step for every lane; the per-rollout trace is written only when the
    // caller asked for it.
    auto finish = [&](uint32_t inc) {
        acc += inc;
        if (inc) {
            if (kTrace) {
                const size_t g = size_t(L.pos) * a.rollouts_per_pos + L.roll;
                if (a.winners) a.winners[g] = (int8_t)(inc == kIncBlack ? 1 : inc == kIncWhite ? -1 : 0);
                if (a.lengths) a.lengths[g] = (uint8_t)(L.start - L.left);
            }
            active = false;
        }
    };
    auto flush = [&]() {
        if (a.wdb && acc_pos != 0xffffffffu) {
            const uint32_t white = (acc >> 10) & 1023u, draw = acc >> 20, black = acc & 1023u;
            if (white) atomicAdd(a.wdb + size_t(acc_pos) * 3 + 0, int(white));
            if (draw) atomicAdd(a.wdb + size_t(acc_pos) * 3 + 1, int(draw));
            if (black) atomicAdd(a.wdb + size_t(acc_pos) * 3 + 2, int(black));
        }
        acc = 0;
        acc_n = 0;
    };

    for (;;) {
        // ---- refill point --------------------------------------------------------------------------
        const bool need = !active && !dead;
        const uint32_t m = __ballot_sync(0xffffffffu, need);
        if (m) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(m) - 1)) base = atomicAdd(&s_ticket, uint32_t(__popc(m)));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (need) {
                const uint32_t t = base + __popc(m & lt);
                const unsigned long long g64 =
                    ((unsigned long long)(t / kTicketBlock) * gridDim.x + blockIdx.x) * kTicketBlock + t % kTicketBlock;
                if (g64 >= total) {
                    dead = true;
                } else {
                    const uint32_t g = uint32_t(g64);
                    L.pos = g / uint32_t(a.rollouts_per_pos);
                    L.roll = g - L.pos * uint32_t(a.rollouts_per_pos);
                    if (L.pos != acc_pos || acc_n >= 1000u) { flush(); acc_pos = L.pos; }
                    acc_n += 1;
                    const uint32_t* img = images + size_t(L.pos) * kImageWords;
                    const uint32_t info = __ldg(img + kMetaInfo);
                    L.left = L.start = info & 0xffu;
                    L.colour = (info >> 8) & 1u;
                    L.rowmask = __ldg(img + kMetaRows);
                    if ((info >> 9) & 1u) {                                      // already decided: 0 moves
                        const uint32_t wc = (info >> 10) & 3u;
                        active = true;
                        finish(wc == 1u ? kIncBlack : wc == 2u ? kIncWhite : kIncDraw);
                    } else {
                        const uint4* img4 = reinterpret_cast<const uint4*>(img);    // 320-byte images: 16-byte aligned
#pragma unroll 6
                        for (int s = 0; s < kSlots / 4; ++s) {
                            const uint4 q = __ldg(img4 + s);
                            sts32(my + (4 * s + 0) * kSlotStride, q.x); sts32(my + (4 * s + 1) * kSlotStride, q.y);
                            sts32(my + (4 * s + 2) * kSlotStride, q.z); sts32(my + (4 * s + 3) * kSlotStride, q.w);
                        }
                        active = true;
                        if (kInjected) inj = a.r_stream + size_t(g) * a.stream_stride;
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0) {
            if (__ballot_sync(0xffffffffu, !dead) == 0) break;
            continue;
        }
        // ---- kRefill move steps ------------------------------------------------------------------------
#pragma unroll 1
        for (int quad = 0; quad < kRefill / 4; ++quad) {
            if (!kInjected) {
                if (active)
                    philox4x32_10((L.start - L.left) >> 2, L.roll, uint32_t(a.pos_base) + L.pos, a.ctr_hi, a.key_lo, a.key_hi, rnd);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (active) {
                    uint32_t r;
                    if (kInjected) {
                        const uint32_t played = L.start - L.left;
                        if (int(played) >= a.stream_stride) {                    // stream exhausted: report length 255, winner 0
                            L.start = L.left + 255u;
                            finish(kIncDraw);
                            continue;
                        }
                        r = inj[played];
                    } else {
                        r = __umulhi(rnd[s], uint32_t(kCells));
                    }
                    finish(play_move(my, cell_addr, L, r));
                }
            }
        }
    }
    flush();
}

// ---- a handful of positions, one launch ----------------------------------------------------------
// The MCTS mirror simulates ONE leaf per call (RandomPolicy: 5 rollouts): pure latency.  One CTA per position builds the
// slot image in shared memory, plays one rollout per thread (same Philox streams and the same move code as the big
// kernel, so the counts are identical) and writes the three counts with plain stores -- `boards` and `out` may be
// page-locked host memory, so that the whole call is one launch and one synchronisation.
// kMoves: also write every rollout's winner, length and the cells it played (a.winners / a.lengths / a.moves, which
// may be page-locked host memory too) -- PoolRAVEPolicy::defaultSimulate leaves the board at the END of its playout
// (policies/PoolRAVE.h:27-48), so the host mirror replays the playout's moves on its Board.
template <bool kMoves>
__global__ void __launch_bounds__(kThreads) rollout_small_kernel(RolloutArgs a, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t s_slots[];                          // [kSlots + 1][blockDim.x]
    __shared__ uint32_t s_board[kBoardWords], s_img[kImageWords], s_cell[kCells], s_cnt[3];
    const int pos = blockIdx.x, tid = threadIdx.x;
    if (tid < kBoardWords) s_board[tid] = a.boards[(size_t)pos * kBoardWords + tid];
    if (tid < 3) s_cnt[tid] = 0;
    for (int c = tid; c < kCells; c += blockDim.x) s_cell[c] = cell_lut_entry(c);
    __syncthreads();
    if (tid < 32) build_image(s_board, s_img, tid);
    __syncthreads();
    const uint32_t info = s_img[kMetaInfo];
    const uint32_t stride = blockDim.x * 4u;
    uint32_t cell_addr = (uint32_t)__cvta_generic_to_shared(s_cell);
    uint32_t my = (uint32_t)__cvta_generic_to_shared(s_slots + tid);
    asm volatile("" : "+r"(cell_addr), "+r"(my));
    uint32_t inc = 0;
    if (tid < a.rollouts_per_pos) {
        if ((info >> 9) & 1u) {                                                  // already decided: 0 moves
            const uint32_t wc = (info >> 10) & 3u;
            inc = wc == 1u ? kIncBlack : wc == 2u ? kIncWhite : kIncDraw;
            if (kMoves && a.lengths) a.lengths[size_t(pos) * a.rollouts_per_pos + tid] = 0;
        } else {
            Lane L{};
            L.left = L.start = info & 0xffu;
            L.colour = (info >> 8) & 1u;
            L.rowmask = s_img[kMetaRows];
            L.pos = pos; L.roll = tid;
            for (int s = 0; s < kSlots; ++s) sts32(my + s * stride, s_img[s]);
            sts32(my + kSlots * stride, 0u);
            uint32_t rnd[4];
            uint8_t* trace = kMoves && a.moves ? a.moves + (size_t(pos) * a.rollouts_per_pos + tid) * kCells : nullptr;
            while (inc == 0) {
                philox4x32_10((L.start - L.left) >> 2, L.roll, uint32_t(a.pos_base) + L.pos, a.ctr_hi, a.key_lo, a.key_hi, rnd);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (inc == 0) {
                        uint32_t cell = 0;
                        const uint32_t k = L.start - L.left;
                        inc = play_move(my, cell_addr, L, __umulhi(rnd[s], uint32_t(kCells)), stride, kMoves ? &cell : nullptr);
                        if (kMoves && trace) trace[k] = uint8_t(cell);
                    }
            }
            if (kMoves && a.lengths) a.lengths[size_t(pos) * a.rollouts_per_pos + tid] = uint8_t(L.start - L.left);
        }
        if (kMoves && a.winners) a.winners[size_t(pos) * a.rollouts_per_pos + tid] = int8_t(inc == kIncBlack ? 1 : inc == kIncWhite ? -1 : 0);
        atomicAdd(&s_cnt[inc == kIncWhite ? 0 : inc == kIncDraw ? 1 : 2], 1u);
    }
    __syncthreads();
    if (out && tid < 3) out[(size_t)pos * 3 + tid] = int(s_cnt[tid]);
}

}  // namespace

cudaError_t launch_rollout_small(const RolloutArgs& a, int32_t* out, cudaStream_t stream) {
    if (a.n <= 0 || a.rollouts_per_pos <= 0) return cudaSuccess;
    if (a.rollouts_per_pos > kThreads) return cudaErrorInvalidValue;
    const int threads = (a.rollouts_per_pos + 31) / 32 * 32;
    const size_t smem = size_t(kSlots + 1) * threads * sizeof(uint32_t);
    static bool configured = false;
    if (!configured) {
        const int most = int(size_t(kSlots + 1) * kThreads * sizeof(uint32_t));
        cudaError_t e = cudaFuncSetAttribute(rollout_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(rollout_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (a.moves || a.winners || a.lengths) rollout_small_kernel<true><<<a.n, threads, smem, stream>>>(a, out);
    else rollout_small_kernel<false><<<a.n, threads, smem, stream>>>(a, out);
    return cudaGetLastError();
}

int rollout_launches(const RolloutArgs& a) { return a.n > 0 ? 2 : 0; }
size_t rollout_scratch_bytes(int n) { return size_t(n > 0 ? n : 0) * kImageWords * sizeof(uint32_t); }

cudaError_t launch_rollout(const RolloutArgs& a, int sm_count, uint32_t* images, cudaStream_t stream) {
    if (a.n <= 0 || a.rollouts_per_pos <= 0) return cudaSuccess;
    if (!images) return cudaErrorInvalidValue;
    cudaError_t err;
    build_images_kernel<<<(a.n * 32 + 255) / 256, 256, 0, stream>>>(a.boards, a.n, images, a.wdb);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const size_t smem = size_t(kSlots + 1) * kThreads * sizeof(uint32_t);   // + one scratch slot per thread
    const unsigned long long total = (unsigned long long)a.n * a.rollouts_per_pos;
    unsigned long long blocks = (total + kTicketBlock - 1) / kTicketBlock;
    const unsigned long long resident = (unsigned long long)sm_count * 3;         // 3 CTAs per SM (73 KB of slot state each)
    const int grid = int(blocks < resident ? blocks : resident);
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kThreads, smem, stream>>>(a, images);
        return cudaGetLastError();
    };
    if (a.r_stream) return launch(rollout_kernel<true, kRefillDefault, true>);
    if (a.winners != nullptr || a.lengths != nullptr) return launch(rollout_kernel<false, kRefillDefault, true>);
    return launch(rollout_kernel<false, kRefillDefault, false>);
}

}  // namespace gk

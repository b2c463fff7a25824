// gk_eval.cu -- K1 `ac_eval`: from-scratch evaluation of a batch of boards, one warp per board.
//
// Replaces, for a batch, Evaluator::applyMove replay + read-out of m_scores / totals
// (reference src/Pattern.cpp:128-302,418-550; semantics restated in SURVEY.md Appendix A).
//
//   phase 0  load the 64-byte board (one word per lane, kept in registers), clear the warp's
//            shared-memory accumulators
//   phase 1  stone-density block score (+160 where a player has a stone on a weighted offset of
//            the 7x7 neighbourhood, Pattern.cpp:236-272,598-609) from 15-bit row masks
//   phase 2  the Aho-Corasick scan: 32 lanes walk their chains of whole lines in lock step; per
//            symbol: tape entry, board word by warp shuffle, rotate, ONE dependent 16-bit table
//            load.  A step that emitted appends one 16-bit entry to the lane's private list
//            (predicated store, no votes); the lists cannot overflow (capacity proven by the
//            table compiler over all boards)
//   phase 3  emission scatter: the lists are concatenated by a prefix sum and entry i of the
//            concatenation goes to lane i % 32 (binary search by shuffle over the 32 prefix counts), so the
//            work is balanced whatever lines the stones sit on; pattern scores go to the '_' / '^'
//            cells with shared-memory atomics, totals are bumped, per-cell pattern counts kept
//   phase 4  compounds (double-three / four-three / double-four, Pattern.cpp:418-550) from the flags;
//            the 13-symbol window rescans of Compound::updateAntis are spread one per lane
//   phase 5  coalesced 128-bit store of the four 225-cell score maps + totals + winner
//
// Bound by integer issue and shared-memory latency, not HBM (64 B in, 3.6 KB out per board).
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

#ifndef GK_HEADS_INLINE
#define GK_HEADS_INLINE __forceinline__
#endif
#ifndef GK_HEADS_UNROLL
#define GK_HEADS_UNROLL 1
#endif
#define GK_PRAGMA(x) _Pragma(#x)
#define GK_UNROLL(n) GK_PRAGMA(unroll n)
constexpr int kWarpsPerCta = 32;
constexpr size_t kSmemLimit = 227 * 1024;
constexpr int kScoreWords = 4 * kCells;        // 900
constexpr int kFlagWords = 2 * kCells + 2;     // 452 (16-byte multiple)
constexpr int kBoardSmem = 20;                 // 17 words used (cells up to 271 read as pad)
constexpr int kTotalWords = 24;                // 16 pattern + 6 compound + spare

// Per-warp shared-memory block; the emission lists (32 lanes x list_cap uint16) follow it.
struct WarpSmem {
    int scores[kScoreWords];                   // [group][cell]
    uint32_t flags[kFlagWords];                // [cell][player grp]: 3 classes x 4 dirs x 2-bit count (0..2)
    uint32_t board[kBoardSmem];
    uint32_t totals[kTotalWords];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp block must keep 16-byte alignment");

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 32-bit shared-window addressing for the scan loop (keeps the generic->shared conversion out of it)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.shared.u16 [%0], t; }" :: "r"(addr), "r"(v) : "memory");
}

// ---- bulk asynchronous copy shared -> global (the TMA engine's 1-D form, sm_90+): one elected lane hands a whole 3 600-byte
// score block to the copy engine and the warp goes on with its next board; the block may not be overwritten before
// bulk_wait_read(), and is in global memory (visible to every thread) after bulk_wait_all().
__device__ __forceinline__ void bulk_store(void* dst_global, uint32_t src_shared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                 :: "l"(dst_global), "r"(src_shared), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// x / d for d > 0, IEEE-rounded like the reference's float division.  A zero numerator (most cells: no stone nearby, or
// occupied) skips the division: 0 / d is +0 anyway, and the correctly rounded division's special-case path -- which a
// zero operand takes -- costs ~40 instructions for the whole warp every time ANY lane needs it.
// (A plain `x != 0 ? x / d : 0` does not help: the compiler divides first and selects afterwards.  So the division itself
// never sees the zero: it divides 1 instead and the quotient is dropped.)
__device__ __forceinline__ float div_pos(float x, float d) {
    const bool nz = x != 0.f;
    float xs = nz ? x : 1.f;
    asm("" : "+f"(xs));                                   // opaque: or the compiler folds the two selects back into x / d
    const float q = xs / d;
    return nz ? q : 0.f;
}

__device__ __forceinline__ uint32_t cell_value(const uint32_t* board, uint32_t cell) {
    return (board[cell >> 4] >> ((cell & 15u) * 2u)) & 3u;
}

// even bits of a 30-bit field -> 15 contiguous bits
__device__ __forceinline__ uint32_t squeeze_even(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

// One emission of pattern `rec` whose last char sits on virtual cell `vend` of a line with the
// given direction (Updater::updatePatterns, Pattern.cpp:138-165).  Returns winner bits.
// dflags (policy-head variants only): per cell, which (type >= DeadThree, favour, perspective) pattern flags and
// which (compound type, favour, perspective) compound flags are set in ANY direction -- what
// Heuristic::DecisiveFilter reads through Record::get(favour, perspective) (Heuristic.hpp:147-151):
//   bit (type - 4) * 4 + Group(favour, perspective), bit 16 + compound type * 4 + Group(favour, perspective)
// delta = -1 (incremental guided playouts only) takes the emission back: Updater::updatePatterns with delta = -1.
__device__ __forceinline__ uint32_t apply_emission(WarpSmem& ws, uint32_t* dflags, const PatRec rec, int vend, uint32_t dir,
                                                   int stride, int delta = 1) {
    const uint32_t type = pr_type(rec.w0), black = pr_black(rec.w0);
    if (type == kTypeFive) return black ? 1u : 2u;                              // :140-145
    atomicAdd(&ws.totals[black * 8 + type], uint32_t(delta));                   // :147
    const int score = delta * (dir >= 2 ? int(rec.w1 >> 16) : int(rec.w1 & 0xffffu));   // :151-152
    int* self = ws.scores + black * 3 * kCells;                                 // Group(f, f)
    int* rival = ws.scores + (black + 1) * kCells;                              // Group(f, -f)
    const uint32_t ncells = pr_ncells(rec.w0);
    // compound classes count their '_' cells per direction (Record::set's saturating 00 -> 01 -> 11, :395-400, as a binary
    // count: an exhaustive enumeration of lines shows it never passes 2); lo = 0 for the other patterns
    const uint32_t cclass = pr_cclass(rec.w0);
    const uint32_t lo = cclass ? 1u << (cclass * 8 - 8 + dir * 2) : 0u;
    // decisive flags (policy-head variants): update_pose marks '_' cells for both perspectives, '^' cells for the rival's, :153-161
    const uint32_t d_rival = (dflags && type >= 4u) ? 1u << ((type - 4u) * 4u + black + 1u) : 0u;
    const uint32_t d_both = d_rival | ((dflags && type >= 4u) ? 1u << ((type - 4u) * 4u + black * 3u) : 0u);
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {                                          // at most four scored cells: straight-line, predicated
        const uint32_t nib = (rec.w0 >> (4 * s)) & 15u;
        const int cell = vend - int(nib & 7u) * stride;
        if (s < ncells) {
            atomicAdd(&rival[cell], score);                                     // '_' and '^', :158-161
            if (d_rival) atomicOr(&dflags[cell], (nib & 8u) ? d_both : d_rival);
            if (nib & 8u) {
                atomicAdd(&self[cell], score);
                if (lo) {
                    uint32_t* word = &ws.flags[cell * 2 + black];
                    atomicAdd(word, delta > 0 ? lo : 0u - lo);                  // 2-bit BINARY count per (class, direction); it never exceeds 2 (theorem T2)
                }
            }
        }
    }
    return 0;
}

// Compound::updateAntis (Pattern.cpp:520-543): rescan the 13-symbol window centred on `cell`
// from the root state and give +600 (rival's perspective) to the other '_' / '^' cells of the
// first emission of class `cclass` that has `cell` on a '_'.
// kRawBoard: `board` is the caller's packed board in global memory (deferred tasks, see the kernel): cells
// holding the invalid value 3 read as white there too; `rival` may then point into the global score block.
template <bool kRawBoard>
__device__ __forceinline__ void anti_cells(const uint32_t* board, uint32_t* dflags, uint32_t anti_bit, uint32_t next_addr,
                                           uint32_t root_off, uint32_t emit_thr, const uint32_t* s_erec, const PatRec* s_patrec,
                                           int cell, uint32_t dir, uint32_t cclass, int* rival, int amount = 600) {
    const int cx = cell % kWidth, cy = cell / kWidth;
    const int stride = dir_stride(dir);
    // steps i in [lo, hi] of the window are on the board, the rest reads as '?'
    int before, after;                                                          // cells available towards -/+ along the line
    if (dir == 0) { before = cx; after = kWidth - 1 - cx; }
    else if (dir == 1) { before = cy; after = kHeight - 1 - cy; }
    else if (dir == 2) { before = min(cx, cy); after = kWidth - 1 - max(cx, cy); }
    else { before = min(kWidth - 1 - cx, cy); after = min(cx, kHeight - 1 - cy); }
    const int lo = 6 - min(before, 6), hi = 6 + min(after, 6);
    // gather the 13 cell values first (independent loads), value * 2 at bit 2 * i + 1
    uint32_t window = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        uint32_t v = 3u;
        if (i >= lo && i <= hi) {
            v = cell_value(board, cell + (i - 6) * stride);
            if (kRawBoard && v == 3u) v = 2u;
        }
        window |= v << (2 * i + 1);
    }
    uint32_t nx = root_off;
#pragma unroll 1
    for (int i = 0; i < 13; ++i, window >>= 2) {
        nx = lds_u16(next_addr + nx + (window & 6u));
        if (nx >= emit_thr) continue;
        const uint32_t er = s_erec[nx >> 3];
        if (er_cclass0(er) != cclass && er_pid(er, 1) == kDevNoPid) continue;   // wrong class, no second emission: cheap reject
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t pid = er_pid(er, k);
            if (pid == kDevNoPid) break;
            const PatRec rec = s_patrec[pid];
            const int off = i - int(er_prev(er, k)) - 6;                        // `cell` is the off-th char from the pattern's end
            if (pr_cclass(rec.w0) != cclass || off < 0 || off >= int(pr_len(rec.w0))) continue;   // HasCovered, :22-25
            // `cell` must sit on a '_': one of the four cell nibbles equals 8 | off (unused nibbles are 0 and cannot match);
            // a nibble of x is zero iff it matches -- the has-a-zero-nibble test replaces a loop over the cells
            const uint32_t x = (rec.w0 & 0xffffu) ^ ((8u | uint32_t(off)) * 0x1111u);
            if (((x - 0x1111u) & ~x & 0x8888u) == 0) continue;
            uint32_t cells = rec.w0;
            for (uint32_t n = pr_ncells(rec.w0); n != 0; --n, cells >>= 4) {
                const int j = int(cells & 7u);
                if (j != off) {
                    atomicAdd(&rival[cell + (off - j) * stride], amount);       // updatePose(current, component, -favour), :536
                    if (dflags) atomicOr(&dflags[cell + (off - j) * stride], anti_bit);
                }
            }
            return;                                                             // only the first such pattern, :540
        }
    }
}

// Compound::locate + updateCritical for one (cell, player) whose flags passed Compound::Test
// (Pattern.cpp:440-518).  Returns the two updateAntis tasks in t0 / t1 (0 = none):
// cell | black << 8 | dir << 9 | class << 11 | 1 << 13.
// `scores` / `totals32` are the warp's shared accumulators, or (deferred form) the board's stored score block and
// the 32-bit word holding its first compound total in global memory (`totals_stride_bits` = 16 there: uint16 fields).
__device__ __forceinline__ void compound_at(int* scores, uint32_t* totals32, int totals_field_bits, uint32_t* dflags,
                                            uint32_t f, uint32_t idx, uint32_t& t0, uint32_t& t1, int amount = 600) {
    const int cell = idx >> 1;
    const uint32_t black = idx & 1u;
    // states: S0 0, L2 1, LD3 2, To33 3, To43 4, To44 5
    int state = 0, l3 = 0, ncomp = 0;
    bool triple = false;
    uint32_t first = 0, second = 0;
#pragma unroll
    for (uint32_t dir = 0; dir < 4; ++dir) {
        const uint32_t c1 = (f >> (dir * 2)) & 3u, c2 = (f >> (8 + dir * 2)) & 3u, c3 = (f >> (16 + dir * 2)) & 3u;
        const uint32_t cls = c1 ? 1u : c2 ? 2u : c3 ? 3u : 0u;                  // LiveThree > DeadThree > LiveTwo
        if (!cls) continue;
        const uint32_t bits = cls == 1u ? c1 : cls == 2u ? c2 : c3;
        const int count = bits == 2u ? 2 : 1, cond = cls == 3u ? 1 : 2;   // bits: binary count 1 or 2 (Record::set's unary 01 / 11)
        l3 += cls == 1u;
        const uint32_t task = uint32_t(cell) | black << 8 | dir << 9 | cls << 11 | 1u << 13;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i < count) {
                if (ncomp == 0) first = task;
                else if (ncomp == 1) second = task;
                ++ncomp;
                int offset;
                if (state == 0) offset = 0;
                else if (state <= 2) offset = 1;
                else { triple = true; offset = state == 5 ? -cond : -1; }
                state += cond + offset;
            }
        }
    }
    t0 = t1 = 0;
    const int type = state - 3;
    if (type < 0) return;   // needs an L3 plus two lower-class patterns on one cell of one line: excluded by exhaustive line enumeration (tests)
    atomicAdd(&scores[black * 3 * kCells + cell], amount * ncomp);              // updateCritical, :515-518
    atomicAdd(&scores[(black + 1) * kCells + cell], amount * ncomp);
    if (totals32) {                                                             // one compound, :505-508
        const uint32_t field = black * 3 + uint32_t(type);
        if (totals_field_bits == 32) atomicAdd(&totals32[field], 1u);
        else atomicAdd(&totals32[field >> 1], 1u << (16 * (field & 1u)));       // two uint16 totals per word
    }
    if (dflags) atomicOr(&dflags[cell], (1u << (16 + type * 4 + black * 3)) | (1u << (16 + type * 4 + black + 1)));   // updateCritical
    if (!triple && l3 == 0) { t0 = first | uint32_t(type) << 14; t1 = second | uint32_t(type) << 14; }   // exactly two components here, :500-502
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
constexpr int kDflagWords = 228;                                               // 225 used (16-byte multiple)
// guided playouts do not filter (no decisive flags): the same 912 bytes hold their density accumulators, uint16 [2][225] + idle-lane slots
__host__ __device__ inline size_t warp_bytes(int list_cap, bool heads) {
    return sizeof(WarpSmem) + size_t(list_cap) * 64 + (heads ? kDflagWords * 4 : 0);
}

// Heuristic::DensityWeight / EvaluationProbs / EvaluationValue (include/algorithms/Heuristic.hpp:16-45) for the
// side to move, from the finished score maps of one board.  `mine` is the lane's 15-bit row mask
// (lanes 0..14 white rows, 15..29 black rows).  Each lane owns the cells lane, lane + 32, ...
//   density(P)[c]  = { count, weight } of P's stones on the weighted offsets of the clipped 7x7 window
//                    (Pattern.cpp:236-272; occupied cells are filtered to 0 by DensityWeight's max(x, 0))
//   DW(P)          = normalized( 3 W / (1 + 2 N) )                                       (:40-45)
//   probs          = normalized( 0.6 S(p,p) . DW(p) + 0.4 S(-p,p) . DW(-p) ), one-hot centre on an empty board (:16-27)
//   value          = tanh( (1.2 <S(p,p), DW(p)> - <S(-p,-p), DW(-p)>) / 500 )            (:32-36)
// Floating point: sums run lane-strided then by warp shuffle, so they differ from Eigen's packet order in
// the last bits (tolerance in tests/test_heads.py).
// The loops over cells are deliberately NOT unrolled and keep their per-cell values in shared memory (the flag
// words and emission lists are dead by now): a fully unrolled version is 115 KB of code, and with 28 warps at
// different places the kernel then waits on instruction fetch (ncu: stall_no_instruction 9 per issue).
//   dwv  [2][225] floats over ws.flags: DensityWeight numerators 3W / (1 + 2N), then normalised
//   prob [225]    floats over the emission lists: the move probabilities
__device__ GK_HEADS_INLINE void policy_heads(WarpSmem& ws, float* prob, const uint16_t* s_lut, uint32_t mine, int lane,
                                          float* value_out, int& n_stones, int& to_move, uint16_t* dacc, bool dacc_valid) {
    float* dwv = reinterpret_cast<float*>(ws.flags);
    const uint32_t cnt = lane < 30 ? __popc(mine) : 0u;
    const int n_white = int(__reduce_add_sync(0xffffffffu, lane < 15 ? cnt : 0u));
    const int n_black = int(__reduce_add_sync(0xffffffffu, lane >= 15 ? cnt : 0u));
    const int p = n_black == n_white ? 1 : 0;                      // Group(player to move): black moves first (Game.h:128)
    // Density by columns: lane (colour, x) walks the rows once.  A row's 7-bit slice around x is looked up for
    // |dy| = 0..3 (4 table loads) and added to a sliding window of accumulators (rows y - 3 .. y + 2 in w0 .. w5); the
    // oldest row is complete after each step and is turned into 3 W / (1 + 2 N) at once.  60 table loads per lane
    // instead of 8 cells x 7 rows x 2 colours = 112, no per-cell row shuffles, and a loop body of ~45 instructions
    // (this variant of the kernel is sensitive to code size, see above).
    const int dc = lane >= 15, dx = lane - 15 * dc;                // lanes 0..14 white, 15..29 black (lanes 30, 31 idle along)
    const uint32_t occ = mine | __shfl_sync(0xffffffffu, mine, lane < 15 ? lane + 15 : lane - 15);   // occupancy of row y on lanes y and 15 + y
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;
    float n2 = 0.f;
    // branch-free body: every lane computes and stores (lanes 30, 31 into the two spare words behind the maps)
    const bool live = lane < 30;
    const uint32_t xbit = live ? 1u << dx : 0u;                    // 0: lanes 30, 31 always read "open" but their w7 slices are empty
    float* dst = live ? dwv + dc * kCells + dx : dwv + 2 * kCells + (lane - 30);
    const int dst_step = live ? kWidth : 0;
    const int src0 = 15 * dc;
    // Guided playouts keep the packed (count, weight) accumulators of every cell in shared memory (`dacc`): a game adds
    // one stone per evaluation, so after the first full pass only the 7 x 7 neighbourhood of the new stone is updated
    // (by the kernel, when the move is made) and this pass just turns the accumulators into weights.
    uint16_t* acc_p = dacc ? dacc + (live ? dc * kCells + dx : 2 * kCells + (lane - 30)) : nullptr;
    if (dacc != nullptr && dacc_valid) {
#pragma unroll 1
        for (int y = 0; y < kHeight; ++y) {
            const uint32_t full = live ? uint32_t(*acc_p) : 0u;
            acc_p += dst_step;
            const uint32_t orow = __shfl_sync(0xffffffffu, occ, y);
            float v = div_pos(3.f * float(full >> 8), 1.f + 2.f * float(full & 0xffu));
            v = (orow & xbit) ? 0.f : v;
            *dst = v;
            dst += dst_step;
            n2 += v * v;
        }
    } else {
#pragma unroll 1
        for (int y = 0; y < kHeight + 3; ++y) {
            uint32_t row = __shfl_sync(0xffffffffu, mine, src0 + y);    // y >= 15 reads some other lane: replaced by an empty row
            row = (y < kHeight && live) ? row : 0u;                     // flushing steps: slice 0 looks up { 0, 0 }
            const uint32_t w7 = ((row << 3) >> dx) & 0x7fu;             // cells dx - 3 .. dx + 3 of row y
            const uint32_t t0 = s_lut[w7], t1 = s_lut[128 + w7], t2 = s_lut[256 + w7], t3 = s_lut[384 + w7];
            const uint32_t full = w0 + t3;                              // row y - 3 has seen rows y - 6 .. y
            w0 = w1 + t2; w1 = w2 + t1; w2 = w3 + t0; w3 = w4 + t1; w4 = w5 + t2; w5 = t3;
            if (y >= 3) {                                               // warp-uniform
                const uint32_t orow = __shfl_sync(0xffffffffu, occ, y - 3);
                float v = div_pos(3.f * float(full >> 8), 1.f + 2.f * float(full & 0xffu));
                v = (orow & xbit) ? 0.f : v;                            // occupied cells are filtered to 0 (DensityWeight's max(x, 0))
                *dst = v;
                dst += dst_step;
                n2 += v * v;
                if (dacc != nullptr) { *acc_p = uint16_t(full); acc_p += dst_step; }
            }
        }
    }
    float n2w = lane < 15 ? n2 : 0.f, n2b = lane >= 15 ? n2 : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n2w += __shfl_xor_sync(0xffffffffu, n2w, d);
        n2b += __shfl_xor_sync(0xffffffffu, n2b, d);
    }
    const float nrm_w = n2w > 0.f ? sqrtf(n2w) : 1.f, nrm_b = n2b > 0.f ? sqrtf(n2b) : 1.f;
    const int* s_self = ws.scores + 3 * p * kCells;                // S(p, p)
    const int* s_anti = ws.scores + (2 * (1 - p) + p) * kCells;    // S(-p, p)
    const int* s_rival = ws.scores + 3 * (1 - p) * kCells;         // S(-p, -p)
    float a2 = 0.f, sdot = 0.f, rdot = 0.f;
    __syncwarp();
#pragma unroll 1
    for (int c = lane; c < kCells; c += 32) {
        const float w0 = div_pos(dwv[c], nrm_w), w1 = div_pos(dwv[kCells + c], nrm_b);    // normalized DW(white), DW(black)
        const float wp = p ? w1 : w0, wr = p ? w0 : w1;
        const float self_worthy = float(s_self[c]) * wp, rival_anti = float(s_anti[c]) * wr;
        const float av = 0.6f * self_worthy + 0.4f * rival_anti;
        prob[c] = av;
        a2 += av * av;
        sdot += float(s_self[c]) * wp;
        rdot += float(s_rival[c]) * wr;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a2 += __shfl_xor_sync(0xffffffffu, a2, d);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, d);
        rdot += __shfl_xor_sync(0xffffffffu, rdot, d);
    }
    const bool empty_board = n_black + n_white == 0;
    const float an = a2 > 0.f ? sqrtf(a2) : 1.f;
    n_stones = n_black + n_white;
    to_move = p;
#pragma unroll 1
    for (int c = lane; c < kCells; c += 32)
        prob[c] = empty_board ? (c == (kHeight / 2) * kWidth + kWidth / 2 ? 1.f : 0.f) : div_pos(prob[c], an);
    if (value_out && lane == 0) *value_out = float(tanh((1.2 * double(sdot) - double(rdot)) / 500.0));
    __syncwarp();
}

// Heuristic::DecisiveFilter (include/algorithms/Heuristic.hpp:93-161): walk the priority automaton
// +4 > -4 > +L3 == +To44 > -L3 == -To44 >= +To43 > -To43 > +To33 > -To33 over the pattern / compound totals; at the
// first class with a non-zero count keep only the cells flagged for one of the remaining candidates
// (Record::get(player, cur_player), any direction) and re-normalise.  p = Group(side to move).
__device__ GK_HEADS_INLINE void decisive_filter(const WarpSmem& ws, const uint32_t* dflags, float* prob, int p, int lane) {
    // the automaton table (:101-105) visits (state, anti) in this fixed order until a candidate has a count
    //   state: 0 = _4 {LiveFour, DeadFour}, 1 = L3 {LiveThree}, 2..4 = To44 / To43 / To33 {compound 2 / 1 / 0}
    uint32_t mask = 0;
#pragma unroll 1
    for (int i = 0; i < 10 && !mask; ++i) {
        const int state = (0x4433212100 >> (4 * i)) & 15, anti = (0x2b2 >> i) & 1;   // (0,0)(0,1)(1,0)(2,0)(1,1)(2,1)(3,0)(3,1)(4,0)(4,1)
        const int pl = anti ? 1 - p : p, g = 2 * pl + p;                        // Group(favour = pl, perspective = cur)
        if (state == 0) {
            const uint32_t l4 = ws.totals[pl * 8 + 7], d4 = ws.totals[pl * 8 + 6];
            if (l4) mask = (1u << (3 * 4 + g)) | (1u << (2 * 4 + g));           // [LiveFour, DeadFour] both stay queued
            else if (d4) mask = 1u << (2 * 4 + g);
        } else if (state == 1) {
            if (ws.totals[pl * 8 + 5]) mask = 1u << (1 * 4 + g);
        } else {
            const int ct = 4 - state;                                           // Pattern::Size + (To33 - state)
            if (ws.totals[16 + pl * 3 + ct]) mask = 1u << (16 + ct * 4 + g);
        }
        if (mask && anti && state != 0) mask |= 1u << (0 * 4 + 2 * (1 - pl) + p);   // own DeadThree also counters, :133-135
    }
    if (!mask) return;
    float n2 = 0.f;
#pragma unroll 1
    for (int c = lane; c < kCells; c += 32) {
        const float v = (dflags[c] & mask) ? prob[c] : 0.f;
        prob[c] = v;
        n2 += v * v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, d);
    if (n2 > 0.f) {
        const float nrm = sqrtf(n2);
#pragma unroll 1
        for (int c = lane; c < kCells; c += 32) prob[c] = div_pos(prob[c], nrm);
    }
    __syncwarp();
}

__device__ __forceinline__ uint32_t philox_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                               uint32_t which) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return which == 0 ? c0 : which == 1 ? c1 : which == 2 ? c2 : c3;
}

// The move of one step of Heuristic::EvaluatedRollout (include/algorithms/Heuristic.hpp:61-91) from the
// probabilities held by the warp (lane owns cells lane + 32 k).  Returns -1 when no cell has weight.
//   mode 1  MaxEvaluatedRollout: the most probable cell, lowest index among equals (Eigen maxCoeff)
//   mode 2  RandomEvaluatedRollout: a draw from the discrete distribution `probs` (Board::getRandomMove(probs),
//           Game.cpp:75-78).  The reference's std::discrete_distribution / mt19937 stream is implementation
//           defined; here the weights are quantised to w = round(p * 2^20), cells are ordered by index, and the
//           draw is r = mulhi32(philox word, sum w): the first cell whose running sum exceeds r.
__device__ GK_HEADS_INLINE int select_move(const float* prob, int lane, int mode, uint32_t rnd) {
    if (mode == 1) {
        float best = 0.f;
        int arg = 0x7fffffff;
#pragma unroll 1
        for (int c = lane; c < kCells; c += 32)
            if (prob[c] > best) { best = prob[c]; arg = c; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        return best > 0.f ? arg : -1;
    }
    // lane l owns the eight consecutive cells 8 l .. 8 l + 7, so ONE warp scan of the lanes' sums orders all 225 weights
    const int c0 = lane * 8;
    uint32_t w[8], mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        w[k] = c0 + k < kCells ? uint32_t(prob[c0 + k] * 1048576.f + 0.5f) : 0u;
        mine += w[k];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return -1;
    const uint32_t r = __umulhi(rnd, total);
    int chosen = 0x7fffffff;
    uint32_t run = incl - mine;                                    // weight of all cells below this lane's
    if (run <= r && r < incl) {                                    // exactly one lane
#pragma unroll
        for (int k = 0; k < 8; ++k) {                              // the first cell whose running sum exceeds r
            run += w[k];
            if (run > r && chosen == 0x7fffffff) chosen = c0 + k;
        }
    }
    return __reduce_min_sync(0xffffffffu, chosen);
}

// ======================= incremental guided playouts (BASELINE config 5) ===================================
// A playout adds ONE stone per evaluation, and a stone only changes the four lines through it (exhaustive theorem T5,
// an emission that does not cover the changed cell is unchanged).  guided_kernel therefore evaluates a
// game's START position from scratch (all 72 lines, as ac_eval_kernel) and afterwards, per move, rescans only the four
// 13-symbol windows around the stone twice -- as they were (emissions that cover the stone taken back,
// Updater::updatePatterns with delta = -1) and as they are (+1) -- which is the reference's own incremental scheme
// (Updater::updateMove, Pattern.cpp:274-302; theorem T4: the covering emissions of a window are the whole line's).
// Compounds follow the same scheme: they are a function of a cell's counts and of its own four 13-symbol windows, so a move
// can only change the compounds of the cells within 6 steps of it on its four lines; those are taken back before the
// lines change and added again afterwards (Updater::updateCompound, Pattern.cpp:169-197, does the same).
// The block score (+160 where a player's density weight is positive) is not stored: the heads add it as they read the
// maps, from the density accumulators the guided variant keeps anyway.
// Every float the heads compute is produced by the same operations in the same order as ac_eval_kernel<true, true>,
// so both variants play IDENTICAL games (tests/test_guided.py).

// Compound::Test (Pattern.cpp:424-433) on the two count words of a cell: per direction the classes' counts are OR-ed and
// a binary 10 is widened to 11 before the "at least two bits" test; bit 0 = white passes, bit 1 = black passes
__device__ __forceinline__ uint32_t compound_test(uint2 f) {
    const bool maybe = (((f.x & (f.x - 1)) | (f.y & (f.y - 1))) | ((f.x | f.y) & 0xaaaaaau)) != 0;   // two fields set, or a count of 2
    if (!maybe) return 0u;
    uint32_t bw0 = (f.x | (f.x >> 8) | (f.x >> 16)) & 0xffu, bw1 = (f.y | (f.y >> 8) | (f.y >> 16)) & 0xffu;
    bw0 |= (bw0 >> 1) & 0x55u;
    bw1 |= (bw1 >> 1) & 0x55u;
    return ((bw0 & (bw0 - 1)) != 0 ? 1u : 0u) | ((bw1 & (bw1 - 1)) != 0 ? 2u : 0u);
}

// every (cell, player) of the board that passes Compound::Test -> clist (cell * 2 + player), returns their number
__device__ __forceinline__ int compound_candidates_all(const WarpSmem& ws, unsigned short* clist, int lane, uint32_t lt) {
    int cn = 0;
    const uint2* f2 = reinterpret_cast<const uint2*>(ws.flags);
#pragma unroll 2
    for (int r = 0; r < (kCells + 31) / 32; ++r) {
        const int cell = r * 32 + lane;
        const uint32_t pass = compound_test(f2[cell < kCells ? cell : kCells]);      // flags[450..451] stay zero
        const uint32_t m0 = __ballot_sync(0xffffffffu, pass & 1u), m1 = __ballot_sync(0xffffffffu, pass & 2u);
        if (pass & 1u) clist[cn + __popc(m0 & lt)] = (unsigned short)(cell * 2);
        cn += __popc(m0);
        if (pass & 2u) clist[cn + __popc(m1 & lt)] = (unsigned short)(cell * 2 + 1);
        cn += __popc(m1);
    }
    __syncwarp();
    return cn;
}

// The cells a move at m can touch: its four 13-cell windows (Mapping.cpp:31-34), slot = direction * 13 + (k + 6) for the
// cell k steps from m; a lane holds slots lane and lane + 32 (52 slots).  wc[r] = the cell of slot r * 32 + lane, -1 off the board.
__device__ __forceinline__ void window_slots(int m, int lane, int wc[2]) {
    const int my = m / kWidth, mx = m - my * kWidth;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int slot = r * 32 + lane;
        const int dir = slot / 13, k = slot - dir * 13 - 6;
        const int x = mx + k * (dir == 1 ? 0 : dir == 3 ? -1 : 1), y = my + k * (dir == 0 ? 0 : 1);
        wc[r] = (slot < 52 && x >= 0 && x < kWidth && y >= 0 && y < kHeight) ? y * kWidth + x : -1;
    }
}
__device__ __forceinline__ bool slot_is_centre(int slot) { return slot == 6 || slot == 19 || slot == 32 || slot == 45; }

// compound_candidates_all over the cells whose compounds a move at m can change: the (up to 12) neighbours of m within 6
// steps on each of its four lines -- a compound's counts and its updateAntis windows only see its own four 13-symbol
// windows (Updater::updateCompound walks exactly these cells, Pattern.cpp:169-197) -- and, with_centre, m itself.
__device__ __forceinline__ int compound_candidates_window(const WarpSmem& ws, unsigned short* clist, const int wc[2], bool with_centre,
                                                          int lane, uint32_t lt) {
    const uint2* f2 = reinterpret_cast<const uint2*>(ws.flags);
    int cn = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int slot = r * 32 + lane, cell = wc[r];
        const bool on = cell >= 0 && (!slot_is_centre(slot) || (with_centre && slot == 6));
        const uint32_t pass = compound_test(f2[on ? cell : kCells]);
        const uint32_t m0 = __ballot_sync(0xffffffffu, pass & 1u), m1 = __ballot_sync(0xffffffffu, pass & 2u);
        if (pass & 1u) clist[cn + __popc(m0 & lt)] = (unsigned short)(cell * 2);
        cn += __popc(m0);
        if (pass & 2u) clist[cn + __popc(m1 & lt)] = (unsigned short)(cell * 2 + 1);
        cn += __popc(m1);
    }
    __syncwarp();
    return cn;
}

// Compound::locate / update / updateAntis for the candidates in clist, added (amount = +600) or taken back (-600):
// ac_eval_kernel's phase 4, in place.
__device__ __noinline__ void compounds_apply(WarpSmem& ws, const unsigned short* clist, int cn, int lane, uint32_t next_addr,
                                                uint32_t root_off, uint32_t emit_thr, const uint32_t* s_erec, const PatRec* s_patrec,
                                                int amount) {
    for (int base = 0; base < cn; base += 32) {
        uint32_t t0 = 0, t1 = 0;
        if (base + lane < cn) {
            const uint32_t idx = clist[base + lane];
            compound_at(ws.scores, nullptr, 32, nullptr, ws.flags[idx], idx, t0, t1, amount);
        }
        const uint32_t owners = __ballot_sync(0xffffffffu, t0 != 0);
        const int ntask = 2 * __popc(owners);
        for (int s0 = 0; s0 < ntask; s0 += 32) {
            const int s = s0 + lane;
            const int owner = s < ntask ? int(__fns(owners, 0, (s >> 1) + 1)) : 0;
            const uint32_t ta = __shfl_sync(0xffffffffu, t0, owner), tb = __shfl_sync(0xffffffffu, t1, owner);
            const uint32_t task = (s & 1) ? tb : ta;
            if (s < ntask) {
                const uint32_t black = (task >> 8) & 1u;
                anti_cells<false>(ws.board, nullptr, 0u, next_addr, root_off, emit_thr, s_erec, s_patrec, int(task & 0xffu),
                                  (task >> 9) & 3u, (task >> 11) & 3u, ws.scores + (black + 1) * kCells, amount);
            }
        }
    }
    __syncwarp();
}

// The heads of policy_heads() for the incremental variant: the density accumulators are always valid, the DensityWeight
// values are recomputed from them where policy_heads() parks them in ws.flags (the counts must survive here), and the
// block score the maps do not hold is added as the maps are read: +160 on S(P, P) where P's weight is positive
// (Pattern.cpp:244,268).  Same float operations in the same order as policy_heads().
// s_vlut[N * 101 + W] = (3 W) / (1 + 2 N) for every possible accumulator (N <= 24 weighted cells, W <= 100 = the sum of
// Evaluator::BlockWeights): the IEEE quotients, computed once per CTA, instead of four divisions per cell and move.
constexpr int kVlutW = 101, kVlutN = 25;
// occ8: bit k = cell lane + 32 k is occupied (kept by the kernel as stones are placed).
__device__ GK_HEADS_INLINE void policy_heads_inc(WarpSmem& ws, float* prob, uint32_t mine, uint32_t occ8, int lane, const uint16_t* dacc,
                                                 const float* s_vlut, int& n_stones, int& to_move) {
    const uint32_t cnt = lane < 30 ? __popc(mine) : 0u;
    const int n_white = int(__reduce_add_sync(0xffffffffu, lane < 15 ? cnt : 0u));
    const int n_black = int(__reduce_add_sync(0xffffffffu, lane >= 15 ? cnt : 0u));
    const int p = n_black == n_white ? 1 : 0;
    const int dc = lane >= 15, dx = lane - 15 * dc;
    const uint32_t occ = mine | __shfl_sync(0xffffffffu, mine, lane < 15 ? lane + 15 : lane - 15);
    const bool live = lane < 30;
    const uint32_t xbit = live ? 1u << dx : 0u;
    const uint16_t* acc_p = dacc + (live ? dc * kCells + dx : 2 * kCells + (lane - 30));
    const int step = live ? kWidth : 0;
    float n2 = 0.f;
#pragma unroll 1
    for (int y = 0; y < kHeight; ++y) {
        const uint32_t full = live ? uint32_t(*acc_p) : 0u;
        acc_p += step;
        const uint32_t orow = __shfl_sync(0xffffffffu, occ, y);
        float v = s_vlut[(full & 0xffu) * kVlutW + (full >> 8)];
        v = (orow & xbit) ? 0.f : v;
        n2 += v * v;
    }
    float n2w = lane < 15 ? n2 : 0.f, n2b = lane >= 15 ? n2 : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n2w += __shfl_xor_sync(0xffffffffu, n2w, d);
        n2b += __shfl_xor_sync(0xffffffffu, n2b, d);
    }
    const float nrm_w = n2w > 0.f ? sqrtf(n2w) : 1.f, nrm_b = n2b > 0.f ? sqrtf(n2b) : 1.f;
    const int* s_self = ws.scores + 3 * p * kCells;
    const int* s_anti = ws.scores + (2 * (1 - p) + p) * kCells;
    const int* s_rival = ws.scores + 3 * (1 - p) * kCells;
    float a2 = 0.f, sdot = 0.f, rdot = 0.f;
#pragma unroll 1
    for (int c = lane; c < kCells; c += 32, occ8 >>= 1) {
        const bool empty = (occ8 & 1u) == 0u;
        const uint32_t fw = dacc[c], fb = dacc[kCells + c];
        float vw = s_vlut[(fw & 0xffu) * kVlutW + (fw >> 8)], vb = s_vlut[(fb & 0xffu) * kVlutW + (fb >> 8)];
        vw = empty ? vw : 0.f;
        vb = empty ? vb : 0.f;
        const float w0 = div_pos(vw, nrm_w), w1 = div_pos(vb, nrm_b);
        const float wp = p ? w1 : w0, wr = p ? w0 : w1;
        const float vp = p ? vb : vw, vr = p ? vw : vb;                     // un-normalised weights: > 0 <=> the block score applies
        const int self_i = s_self[c] + (vp > 0.f ? 160 : 0), rival_i = s_rival[c] + (vr > 0.f ? 160 : 0);
        const float self_worthy = float(self_i) * wp, rival_anti = float(s_anti[c]) * wr;
        const float av = 0.6f * self_worthy + 0.4f * rival_anti;
        prob[c] = av;
        a2 += av * av;
        sdot += float(self_i) * wp;
        rdot += float(rival_i) * wr;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a2 += __shfl_xor_sync(0xffffffffu, a2, d);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, d);
        rdot += __shfl_xor_sync(0xffffffffu, rdot, d);
    }
    const bool empty_board = n_black + n_white == 0;
    const float an = a2 > 0.f ? sqrtf(a2) : 1.f;
    n_stones = n_black + n_white;
    to_move = p;
#pragma unroll 1
    for (int c = lane; c < kCells; c += 32)
        prob[c] = empty_board ? (c == (kHeight / 2) * kWidth + kWidth / 2 ? 1.f : 0.f) : div_pos(prob[c], an);
    __syncwarp();
}

// ---- pieces shared by guided_kernel and guided_pair_kernel ----------------------------------------------------------
// what the kernels read of the compiled table and their own shared tables
struct GuidedTables {
    const uint32_t* s_erec; const PatRec* s_patrec; const uint16_t* s_lut; const float* s_vlut;
    uint32_t next_addr, src_addr, emit_thr, root_off, start_off;
    const uint16_t* tape_info; int tape_steps, cap;
};

// the board of game b: one word per lane (lanes 14 / 15 carry the pads), cells holding the invalid value 3 read as white
__device__ __forceinline__ uint32_t guided_load_board(const uint32_t* boards, long long b, int lane) {
    uint32_t bw = 0xffffffffu;
    if (lane < kBoardWords) {
        bw = __ldg(boards + b * kBoardWords + lane);
        bw &= ~((bw >> 1) & 0x55555555u);
    }
    if (lane == 14) bw |= 0xfffffffcu;
    if (lane == 15) bw = 0xffffffffu;
    return bw;
}

// lanes 0..14: the white stones of row `lane`, lanes 15..29: the black stones of row `lane - 15` (from the shared board copy)
__device__ __forceinline__ uint32_t guided_rows(const uint32_t* board, int lane) {
    uint32_t mine = 0;
    if (lane < 30) {
        const int y = lane - 15 * (lane >= 15), off = 30 * y;
        const uint32_t lo = board[off >> 5], hi = board[(off >> 5) + 1];
        const uint32_t v = __funnelshift_r(lo, hi, off & 31) & 0x3fffffffu;
        mine = lane >= 15 ? squeeze_even(v) : squeeze_even(v >> 1);
    }
    return mine;
}

// bit k: cell lane + 32 k holds a stone
__device__ __forceinline__ uint32_t guided_occ8(const uint32_t* board, int lane) {
    uint32_t occ8 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (lane + 32 * k < kCells && cell_value(board, lane + 32 * k) != 0u) occ8 |= 1u << k;
    return occ8;
}

// density accumulators of the start position: policy_heads()' column walk, accumulators only
__device__ __forceinline__ void guided_density_init(uint32_t mine, int lane, const uint16_t* s_lut, uint16_t* dacc) {
    const int dc = lane >= 15, dx = lane - 15 * dc;
    const bool live = lane < 30;
    uint16_t* acc_p = dacc + (live ? dc * kCells + dx : 2 * kCells + (lane - 30));
    const int step = live ? kWidth : 0, src0 = 15 * dc;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;
#pragma unroll 1
    for (int y = 0; y < kHeight + 3; ++y) {
        uint32_t row = __shfl_sync(0xffffffffu, mine, src0 + y);
        row = (y < kHeight && live) ? row : 0u;
        const uint32_t w7 = ((row << 3) >> dx) & 0x7fu;
        const uint32_t t0 = s_lut[w7], t1 = s_lut[128 + w7], t2 = s_lut[256 + w7], t3 = s_lut[384 + w7];
        const uint32_t full = w0 + t3;
        w0 = w1 + t2; w1 = w2 + t1; w2 = w3 + t0; w3 = w4 + t1; w4 = w5 + t2; w5 = t3;
        if (y >= 3) { *acc_p = uint16_t(full); acc_p += step; }
    }
}

// a new stone of colour to_move (1 black) at (mx, my): its contribution to the density accumulators of its colour
__device__ __forceinline__ void guided_density_add(uint16_t* dacc, const uint16_t* s_lut, int lane, int mx, int my, int to_move) {
    const int dc = lane >= 15, dx = lane - 15 * dc, j = mx - dx + 3;
    if (lane < 30 && dc == to_move && j >= 0 && j < 7) {
#pragma unroll 1
        for (int d = -3; d <= 3; ++d) {
            const int y = my + d;
            if (y >= 0 && y < kHeight) dacc[dc * kCells + y * kWidth + dx] += s_lut[(d < 0 ? -d : d) * 128 + (1 << j)];
        }
    }
}

// the start position from scratch: phases 2 and 3 of ac_eval_kernel (all 72 lines; no block score, no compounds).  Returns winner bits.
__device__ __forceinline__ uint32_t guided_start_position(WarpSmem& ws, const uint16_t* lists, uint32_t list_addr, const GuidedTables& T,
                                                          uint32_t bw, int lane) {
    uint32_t win = 0;
    uint32_t nx = T.start_off, lp = list_addr, src = T.src_addr;
    for (int t = 0; t < T.tape_steps; t += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u, src += 64) {
            const uint32_t e = lds_u16(src);
            const uint32_t w = __shfl_sync(0xffffffffu, bw, e);
            const uint32_t v2 = __funnelshift_r(w, w, e >> 8) & 6u;
            nx = lds_u16(T.next_addr + nx + v2);
            if (nx < T.emit_thr) {
                sts_u16(lp, nx * 8u + uint32_t(t + u));
                lp += 2;
            }
        }
    }
    uint32_t incl = (lp - list_addr) >> 1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int total = int(__shfl_sync(0xffffffffu, incl, 31));
    const int excl = int(incl) - int((lp - list_addr) >> 1);
    __syncwarp();
    for (int i0 = 0; i0 < total; i0 += 32) {
        const int i = i0 + lane;
        int j = __shfl_sync(0xffffffffu, excl, 16) <= i ? 16 : 0;
        if (__shfl_sync(0xffffffffu, excl, j + 8) <= i) j += 8;
        if (__shfl_sync(0xffffffffu, excl, j + 4) <= i) j += 4;
        if (__shfl_sync(0xffffffffu, excl, j + 2) <= i) j += 2;
        if (__shfl_sync(0xffffffffu, excl, j + 1) <= i) j += 1;
        const int first = __shfl_sync(0xffffffffu, excl, j);
        if (i >= total) continue;
        const uint32_t ent = lists[j * T.cap + (i - first)];
        const uint32_t er = T.s_erec[ent >> 6];
        const uint32_t inf = __ldg(T.tape_info + (ent & 63u) * 32u + uint32_t(j));
        const uint32_t dir = (inf >> 9) & 3u;
        const int vcell = inf & 0x1ff, stride = int(inf >> 11);
        win |= apply_emission(ws, nullptr, T.s_patrec[er_pid(er, 0)], vcell - int(er_prev(er, 0)) * stride, dir, stride);
        const uint32_t p1 = er_pid(er, 1);
        if (p1 != kDevNoPid) win |= apply_emission(ws, nullptr, T.s_patrec[p1], vcell - int(er_prev(er, 1)) * stride, dir, stride);
    }
    __syncwarp();
    return win;
}

// A stone was placed at `cell` (ws.board already holds it): the four 13-symbol windows around it (Updater::matchPatterns,
// Pattern.cpp:128-136) are scanned from the root state, lane & 3 = direction, lanes 0..3 as they are now (emissions added),
// lanes 4..7 as they were (taken back).  Only emissions that cover the stone count (HasCovered, Pattern.cpp:22-25); by
// theorem T4 those are the whole line's, by T5 nothing else changes.  wc: window_slots(cell).  Returns winner bits.
__device__ __forceinline__ uint32_t guided_move_patterns(WarpSmem& ws, const uint16_t* lists, uint32_t list_addr, const GuidedTables& T,
                                                         int cell, const int wc[2], int lane) {
    uint32_t p0, p1;                                                           // the lane's window, one bit plane per symbol bit
    {
        uint32_t v[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) v[r] = wc[r] >= 0 ? cell_value(ws.board, uint32_t(wc[r])) : 3u;   // off the board: '?'
        const uint32_t a0 = __ballot_sync(0xffffffffu, v[0] & 1u), b0 = __ballot_sync(0xffffffffu, v[1] & 1u);
        const uint32_t a1 = __ballot_sync(0xffffffffu, v[0] & 2u), b1 = __ballot_sync(0xffffffffu, v[1] & 2u);
        const uint32_t sh = 13u * (uint32_t(lane) & 3u);                       // direction d owns slots 13 d .. 13 d + 12
        p0 = (sh < 32u ? __funnelshift_r(a0, b0, sh) : b0 >> (sh - 32u)) & 0x1fffu;
        p1 = (sh < 32u ? __funnelshift_r(a1, b1, sh) : b1 >> (sh - 32u)) & 0x1fffu;
        if (lane >= 4) { p0 &= ~0x40u; p1 &= ~0x40u; }                        // before the move the centre was empty
    }
    uint32_t nx = T.root_off, lp = list_addr;
    if (lane < 8) {
#pragma unroll 1
        for (int i = 0; i < 13; ++i, p0 >>= 1, p1 >>= 1) {
            nx = lds_u16(T.next_addr + nx + ((p0 & 1u) * 2u + (p1 & 1u) * 4u));
            if (nx < T.emit_thr) {
                sts_u16(lp, nx * 8u + uint32_t(i));
                lp += 2;
            }
        }
    }
    __syncwarp();
    uint32_t win = 0;
    // balanced scatter of the handful of emissions (ac_eval_kernel's phase 3; the owner lane names direction and sign)
    uint32_t incl = (lp - list_addr) >> 1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int total = int(__shfl_sync(0xffffffffu, incl, 31));
    const int excl = int(incl) - int((lp - list_addr) >> 1);
    __syncwarp();
    for (int i0 = 0; i0 < total; i0 += 32) {
        const int i = i0 + lane;
        int j = __shfl_sync(0xffffffffu, excl, 4) <= i ? 4 : 0;               // only lanes 0..7 hold entries
        if (__shfl_sync(0xffffffffu, excl, j + 2) <= i) j += 2;
        if (__shfl_sync(0xffffffffu, excl, j + 1) <= i) j += 1;
        const int first = __shfl_sync(0xffffffffu, excl, j);
        if (i >= total) continue;
        const uint32_t ent = lists[j * T.cap + (i - first)];
        const uint32_t er = T.s_erec[ent >> 6];
        const uint32_t dir = uint32_t(j) & 3u;
        const int stride = dir_stride(int(dir)), delta = j < 4 ? 1 : -1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t pid = er_pid(er, k);
            if (pid == kDevNoPid) break;
            const PatRec rec = T.s_patrec[pid];
            const int off = int(ent & 63u) - int(er_prev(er, k));              // window index of the pattern's last symbol; the stone sits at 6
            if (off < 6 || off - int(pr_len(rec.w0)) + 1 > 6) continue;
            const uint32_t w = apply_emission(ws, nullptr, rec, cell + (off - 6) * stride, dir, stride, delta);
            if (delta > 0) win |= w;
        }
    }
    __syncwarp();
    return win;
}

// the compiled table and the two look-up tables of the heads -> shared memory (all threads of the CTA)
__device__ __forceinline__ void guided_load_tables(const EvalArgs& a, uint16_t* s_next, uint32_t* s_erec, PatRec* s_patrec, uint16_t* s_src,
                                                   uint16_t* s_lut, float* s_vlut) {
    const int n_rows = a.n_clones + a.n_states;
    for (int i = threadIdx.x; i < kVlutN * kVlutW; i += blockDim.x)
        s_vlut[i] = (3.f * float(i % kVlutW)) / (1.f + 2.f * float(i / kVlutW));     // the expression of policy_heads(), Heuristic.hpp:39-45
    for (int i = threadIdx.x; i < n_rows * 4; i += blockDim.x) s_next[i] = a.next16[i];
    for (int i = threadIdx.x; i < a.n_clones; i += blockDim.x) s_erec[i] = a.erec[i];
    for (int i = threadIdx.x; i < a.n_patterns; i += blockDim.x) s_patrec[i] = a.patrec[i];
    for (int i = threadIdx.x; i < a.tape_steps * 32; i += blockDim.x) s_src[i] = a.tape_src[i];
    // density of one window row: count | weight << 8 of the stones in a 7-bit row slice, by |dy| (Evaluator::BlockWeights, Pattern.cpp:598-609)
    const int wts[4][7] = { { 1, 3, 4, 0, 4, 3, 1 }, { 0, 3, 5, 4, 5, 3, 0 }, { 0, 4, 3, 3, 3, 4, 0 }, { 2, 0, 0, 1, 0, 0, 2 } };
    for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) {
        int n = 0, w = 0;
        for (int bit = 0; bit < 7; ++bit)
            if ((i >> bit) & 1) { w += wts[i >> 7][bit]; n += wts[i >> 7][bit] > 0; }
        s_lut[i] = uint16_t(n | w << 8);
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
guided_kernel(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_rows = a.n_clones + a.n_states;
    uint16_t* s_next = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* s_erec = reinterpret_cast<uint32_t*>(smem_raw + align16(size_t(n_rows) * 8));
    PatRec* s_patrec = reinterpret_cast<PatRec*>(reinterpret_cast<unsigned char*>(s_erec) + align16(size_t(a.n_clones) * 4));
    uint16_t* s_src = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_patrec) + align16(size_t(a.n_patterns) * sizeof(PatRec)));
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_src) + align16(size_t(a.tape_steps) * 64));
    float* s_vlut = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_lut) + 4 * 128 * sizeof(uint16_t));
    unsigned char* s_warps = reinterpret_cast<unsigned char*>(s_vlut) + align16(size_t(kVlutN) * kVlutW * sizeof(float));
    guided_load_tables(a, s_next, s_erec, s_patrec, s_src, s_lut, s_vlut);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = a.list_cap;
    WarpSmem& ws = *reinterpret_cast<WarpSmem*>(s_warps + size_t(warp) * warp_bytes(cap, true));
    uint16_t* lists = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(&ws) + sizeof(WarpSmem));
    uint16_t* dacc = reinterpret_cast<uint16_t*>(lists + 32 * cap);        // density accumulators, uint16 [2][225] (+ 2 idle-lane slots)
    float* prob = reinterpret_cast<float*>(lists);                          // 225 floats over the emission lists (dead while the heads run)
    const uint32_t lt = lanemask_lt();
    GuidedTables T{ s_erec, s_patrec, s_lut, s_vlut, smem_addr(s_next), smem_addr(s_src + lane), uint32_t(a.n_clones) * 8u,
                    uint32_t(a.root_off), uint32_t(a.start_off), a.tape_info, a.tape_steps, cap };
    uint32_t list_addr = smem_addr(lists + lane * cap);
    asm volatile("" : "+r"(T.next_addr), "+r"(T.src_addr), "+r"(list_addr));

    const int warps = blockDim.x >> 5;
    for (long long b = (long long)blockIdx.x * warps + warp; b < a.n; b += (long long)gridDim.x * warps) {
        // ---- the start position, from scratch -------------------------------------------------------------------------
        uint32_t bw = guided_load_board(a.boards, b, lane);
        if (lane < kBoardSmem) ws.board[lane] = bw;
        {
            int4* z = reinterpret_cast<int4*>(ws.scores);
            for (int i = lane; i < (kScoreWords + kFlagWords) / 4; i += 32) z[i] = make_int4(0, 0, 0, 0);
            if (lane < kTotalWords) ws.totals[lane] = 0;
        }
        __syncwarp();
        uint32_t mine = guided_rows(ws.board, lane);                         // lanes 0..14 white rows, 15..29 black rows
        uint32_t occ8 = guided_occ8(ws.board, lane);
        guided_density_init(mine, lane, s_lut, dacc);
        uint32_t win = guided_start_position(ws, lists, list_addr, T, bw, lane);
        {   // the start position's compounds, from all its counts
            const int cn = compound_candidates_all(ws, lists, lane, lt);
            compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, 600);
        }

        // ---- the game: Heuristic::EvaluatedRollout (Heuristic.hpp:61-72) ------------------------------------------------
        int played = 0, result = 0;
        for (;;) {
            const uint32_t won = __reduce_or_sync(0xffffffffu, win);
            if (won) { result = (won & 1u) ? 1 : -1; break; }               // a Five emission ended the game, Pattern.cpp:140-145
            int n_stones, to_move;
            policy_heads_inc(ws, prob, mine, occ8, lane, dacc, s_vlut, n_stones, to_move);
            int cell = -1;
            if (n_stones < kCells && played < a.g_max_moves) {               // Evaluator::checkGameEnd, Pattern.cpp:343-353
                uint32_t rnd = 0;
                if (a.g_mode == 2)
                    rnd = philox_word(uint32_t(played) >> 2, 0u, uint32_t(a.g_game_base) + uint32_t(b), a.g_ctr_hi, a.g_key_lo,
                                      a.g_key_hi, uint32_t(played) & 3u);
                cell = select_move(prob, lane, a.g_mode, rnd);
            }
            if (cell < 0) break;
            __syncwarp();
            int wc[2];
            window_slots(cell, lane, wc);
            {   // the compounds this move can change are taken back while the lines still are as they were
                const int cn = compound_candidates_window(ws, lists, wc, true, lane, lt);
                compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, -600);
            }
            // ---- place the stone ----------------------------------------------------------------------------------------
            if (a.g_moves && lane == 0) a.g_moves[b * a.g_max_moves + played] = (int16_t)cell;
            const int my = cell / kWidth, mx = cell - my * kWidth;
            if (lane == (cell >> 4)) { bw |= (to_move ? 1u : 2u) << ((cell & 15) * 2); ws.board[lane] = bw; }
            if (lane == (to_move ? 15 : 0) + my) mine |= 1u << mx;
            if (lane == (cell & 31)) occ8 |= 1u << (cell >> 5);
            guided_density_add(dacc, s_lut, lane, mx, my, to_move);
            ++played;
            __syncwarp();
            win = guided_move_patterns(ws, lists, list_addr, T, cell, wc, lane);
            {   // ... and added again from the new counts and the new board
                const int cn = compound_candidates_window(ws, lists, wc, false, lane, lt);
                compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, 600);
            }
        }
        if (a.g_winner && lane == 0) a.g_winner[b] = (int8_t)result;
        if (a.g_length && lane == 0) a.g_length[b] = (int16_t)played;
        if (a.g_final && lane < kBoardWords) a.g_final[b * kBoardWords + lane] = lane == 14 ? (bw & 3u) : lane == 15 ? 0u : bw;
        __syncwarp();
    }
}

// ---- two warps per game: the latency-bound regime (at most a few games per SM) ------------------------------------------
// With 1 024 games on 148 SMs a guided game is the only warp of its scheduler and a move is one serial chain of ~2 900
// instructions.  Two of its parts do not depend on each other: (A) the evaluator update -- compounds taken back, the four
// windows scanned, emissions scattered, compounds added -- and (B) the density weights -- accumulators of the new stone, the
// two norms, one normalised weight per cell and colour.  guided_pair_kernel gives a game two warps: warp A does (A), warp B
// does (B) at the same time, then both halves of the per-cell work of the heads (each warp takes four of a lane's eight
// cells), warp A picks the move.  Every float is produced by the operations of policy_heads_inc() and SUMMED IN ITS ORDER
// (the a2 sum is taken from the stored per-cell values in cell order), so the games are those of guided_kernel, bit for bit.
// Four named barriers per move (bar.sync id, 64); the loop is the same for both warps, so they cannot miss each other.
constexpr int kPairMax = 15;                                       // pairs per CTA: barrier ids 1..15
constexpr int kWnWords = 2 * kCells + 2;                           // normalised density weights, float [2][225] (16-byte multiple)
struct PairMail { int cell; uint32_t won; int pad0, pad1; };
__host__ __device__ inline size_t pair_bytes(int list_cap) { return warp_bytes(list_cap, true) + kWnWords * sizeof(float) + sizeof(PairMail); }

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }

__global__ void __launch_bounds__(kPairMax * 64, 1)
guided_pair_kernel(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_rows = a.n_clones + a.n_states;
    uint16_t* s_next = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* s_erec = reinterpret_cast<uint32_t*>(smem_raw + align16(size_t(n_rows) * 8));
    PatRec* s_patrec = reinterpret_cast<PatRec*>(reinterpret_cast<unsigned char*>(s_erec) + align16(size_t(a.n_clones) * 4));
    uint16_t* s_src = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_patrec) + align16(size_t(a.n_patterns) * sizeof(PatRec)));
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_src) + align16(size_t(a.tape_steps) * 64));
    float* s_vlut = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_lut) + 4 * 128 * sizeof(uint16_t));
    unsigned char* s_pairs = reinterpret_cast<unsigned char*>(s_vlut) + align16(size_t(kVlutN) * kVlutW * sizeof(float));
    guided_load_tables(a, s_next, s_erec, s_patrec, s_src, s_lut, s_vlut);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1, role = warp & 1, bar = 1 + pair;
    const int cap = a.list_cap;
    unsigned char* base = s_pairs + size_t(pair) * pair_bytes(cap);
    WarpSmem& ws = *reinterpret_cast<WarpSmem*>(base);
    uint16_t* lists = reinterpret_cast<uint16_t*>(base + sizeof(WarpSmem));
    uint16_t* dacc = reinterpret_cast<uint16_t*>(lists + 32 * cap);
    float* wn = reinterpret_cast<float*>(base + warp_bytes(cap, true));    // wn[c] white, wn[225 + c] black: v / norm
    PairMail* mail = reinterpret_cast<PairMail*>(wn + kWnWords);
    float* prob = reinterpret_cast<float*>(lists);
    const uint32_t lt = lanemask_lt();
    GuidedTables T{ s_erec, s_patrec, s_lut, s_vlut, smem_addr(s_next), smem_addr(s_src + lane), uint32_t(a.n_clones) * 8u,
                    uint32_t(a.root_off), uint32_t(a.start_off), a.tape_info, a.tape_steps, cap };
    uint32_t list_addr = smem_addr(lists + lane * cap);
    asm volatile("" : "+r"(T.next_addr), "+r"(T.src_addr), "+r"(list_addr));

    const int pairs = blockDim.x >> 6;
    for (long long b = (long long)blockIdx.x * pairs + pair; b < a.n; b += (long long)gridDim.x * pairs) {
        pair_sync(bar);                                                      // the pair's block is free (previous game finished on both warps)
        uint32_t bw = guided_load_board(a.boards, b, lane);                  // both warps keep the board words and the row masks
        if (role == 0) {
            if (lane < kBoardSmem) ws.board[lane] = bw;
            int4* z = reinterpret_cast<int4*>(ws.scores);
            for (int i = lane; i < (kScoreWords + kFlagWords) / 4; i += 32) z[i] = make_int4(0, 0, 0, 0);
            if (lane < kTotalWords) ws.totals[lane] = 0;
        }
        pair_sync(bar);                                                      // ws.board is there
        uint32_t mine = guided_rows(ws.board, lane);
        uint32_t occ8 = guided_occ8(ws.board, lane);
        uint32_t win = 0;
        if (role == 0) {
            win = guided_start_position(ws, lists, list_addr, T, bw, lane);
            const int cn = compound_candidates_all(ws, lists, lane, lt);
            compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, 600);
        } else {
            guided_density_init(mine, lane, s_lut, dacc);
        }
        int played = 0, result = 0;
        for (;;) {
            // ---- (B) the density weights of this position, while (A) finishes the evaluator update ------------------------------
            const uint32_t cnt = lane < 30 ? __popc(mine) : 0u;
            const int n_white = int(__reduce_add_sync(0xffffffffu, lane < 15 ? cnt : 0u));
            const int n_black = int(__reduce_add_sync(0xffffffffu, lane >= 15 ? cnt : 0u));
            const int p = n_black == n_white ? 1 : 0, n_stones = n_black + n_white;
            if (role == 1) {
                __syncwarp();
                const int dc = lane >= 15, dx = lane - 15 * dc;
                const uint32_t occ = mine | __shfl_sync(0xffffffffu, mine, lane < 15 ? lane + 15 : lane - 15);
                const bool live = lane < 30;
                const uint32_t xbit = live ? 1u << dx : 0u;
                const uint16_t* acc_p = dacc + (live ? dc * kCells + dx : 2 * kCells + (lane - 30));
                const int step = live ? kWidth : 0;
                float n2 = 0.f;
#pragma unroll 1
                for (int y = 0; y < kHeight; ++y) {                          // policy_heads_inc(): the column sums, row by row
                    const uint32_t full = live ? uint32_t(*acc_p) : 0u;
                    acc_p += step;
                    const uint32_t orow = __shfl_sync(0xffffffffu, occ, y);
                    float v = s_vlut[(full & 0xffu) * kVlutW + (full >> 8)];
                    v = (orow & xbit) ? 0.f : v;
                    n2 += v * v;
                }
                float n2w = lane < 15 ? n2 : 0.f, n2b = lane >= 15 ? n2 : 0.f;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    n2w += __shfl_xor_sync(0xffffffffu, n2w, d);
                    n2b += __shfl_xor_sync(0xffffffffu, n2b, d);
                }
                const float nrm_w = n2w > 0.f ? sqrtf(n2w) : 1.f, nrm_b = n2b > 0.f ? sqrtf(n2b) : 1.f;
                uint32_t o = occ8;
#pragma unroll 1
                for (int c = lane; c < kCells; c += 32, o >>= 1) {
                    const bool empty = (o & 1u) == 0u;
                    const uint32_t fw = dacc[c], fb = dacc[kCells + c];
                    float vw = s_vlut[(fw & 0xffu) * kVlutW + (fw >> 8)], vb = s_vlut[(fb & 0xffu) * kVlutW + (fb >> 8)];
                    vw = empty ? vw : 0.f;
                    vb = empty ? vb : 0.f;
                    wn[c] = div_pos(vw, nrm_w);
                    wn[kCells + c] = div_pos(vb, nrm_b);
                }
            } else {
                const uint32_t won = __reduce_or_sync(0xffffffffu, win);
                if (lane == 0) mail->won = won;
            }
            pair_sync(bar);                                                  // (1) scores, counts and weights of this position are complete
            const uint32_t won = mail->won;
            if (won) { result = (won & 1u) ? 1 : -1; break; }               // a Five emission ended the game, Pattern.cpp:140-145
            // ---- the heads' per-cell values: warp A takes a lane's cells k = 0..3, warp B k = 4..7 ------------------------------------
            {
                const int* s_self = ws.scores + 3 * p * kCells;
                const int* s_anti = ws.scores + (2 * (1 - p) + p) * kCells;
#pragma unroll 1
                for (int k = 4 * role; k < 4 * role + 4; ++k) {
                    const int c = lane + 32 * k;
                    if (c < kCells) {
                        const float w0 = wn[c], w1 = wn[kCells + c];
                        const float wp = p ? w1 : w0, wr = p ? w0 : w1;
                        const int self_i = s_self[c] + (wp > 0.f ? 160 : 0);  // a positive normalised weight <=> a positive weight: block score
                        const float self_worthy = float(self_i) * wp, rival_anti = float(s_anti[c]) * wr;
                        prob[c] = 0.6f * self_worthy + 0.4f * rival_anti;
                    }
                }
            }
            pair_sync(bar);                                                  // (2) all 225 values stored
            float a2 = 0.f;                                                  // both warps: the sum in policy_heads_inc()'s order
#pragma unroll 1
            for (int c = lane; c < kCells; c += 32) { const float av = prob[c]; a2 += av * av; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, d);
            const float an = a2 > 0.f ? sqrtf(a2) : 1.f;
            pair_sync(bar);                                                  // (3) everybody has read the raw values
#pragma unroll 1
            for (int k = 4 * role; k < 4 * role + 4; ++k) {
                const int c = lane + 32 * k;
                if (c < kCells) prob[c] = n_stones == 0 ? (c == (kHeight / 2) * kWidth + kWidth / 2 ? 1.f : 0.f) : div_pos(prob[c], an);
            }
            pair_sync(bar);                                                  // (4) the probabilities are complete
            if (role == 0) {
                int cell = -1;
                if (n_stones < kCells && played < a.g_max_moves) {           // Evaluator::checkGameEnd, Pattern.cpp:343-353
                    uint32_t rnd = 0;
                    if (a.g_mode == 2)
                        rnd = philox_word(uint32_t(played) >> 2, 0u, uint32_t(a.g_game_base) + uint32_t(b), a.g_ctr_hi, a.g_key_lo,
                                          a.g_key_hi, uint32_t(played) & 3u);
                    cell = select_move(prob, lane, a.g_mode, rnd);
                }
                if (lane == 0) mail->cell = cell;
            }
            pair_sync(bar);                                                  // (5) the move is known
            const int cell = mail->cell;
            if (cell < 0) break;
            const int my = cell / kWidth, mx = cell - my * kWidth;
            if (lane == (p ? 15 : 0) + my) mine |= 1u << mx;
            if (lane == (cell & 31)) occ8 |= 1u << (cell >> 5);
            if (role == 0) {
                int wc[2];
                window_slots(cell, lane, wc);
                {   // the compounds this move can change are taken back while the lines still are as they were
                    const int cn = compound_candidates_window(ws, lists, wc, true, lane, lt);
                    compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, -600);
                }
                if (a.g_moves && lane == 0) a.g_moves[b * a.g_max_moves + played] = (int16_t)cell;
                if (lane == (cell >> 4)) { bw |= (p ? 1u : 2u) << ((cell & 15) * 2); ws.board[lane] = bw; }
                __syncwarp();
                win = guided_move_patterns(ws, lists, list_addr, T, cell, wc, lane);
                {   // ... and added again from the new counts and the new board
                    const int cn = compound_candidates_window(ws, lists, wc, false, lane, lt);
                    compounds_apply(ws, lists, cn, lane, T.next_addr, T.root_off, T.emit_thr, s_erec, s_patrec, 600);
                }
            } else {
                guided_density_add(dacc, s_lut, lane, mx, my, p);
            }
            ++played;
        }
        if (role == 0) {
            if (a.g_winner && lane == 0) a.g_winner[b] = (int8_t)result;
            if (a.g_length && lane == 0) a.g_length[b] = (int16_t)played;
            if (a.g_final && lane < kBoardWords) a.g_final[b * kBoardWords + lane] = lane == 14 ? (bw & 3u) : lane == 15 ? 0u : bw;
        }
    }
}

template <bool kHeads, bool kGuided>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
ac_eval_kernel(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_rows = a.n_clones + a.n_states;
    uint16_t* s_next = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* s_erec = reinterpret_cast<uint32_t*>(smem_raw + align16(size_t(n_rows) * 8));
    PatRec* s_patrec = reinterpret_cast<PatRec*>(reinterpret_cast<unsigned char*>(s_erec) + align16(size_t(a.n_clones) * 4));
    uint16_t* s_src = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_patrec) + align16(size_t(a.n_patterns) * sizeof(PatRec)));
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_src) + align16(size_t(a.tape_steps) * 64));
    unsigned char* s_warps = reinterpret_cast<unsigned char*>(s_lut) + (kHeads ? 4 * 128 * sizeof(uint16_t) : 0);

    for (int i = threadIdx.x; i < n_rows * 4; i += blockDim.x) s_next[i] = a.next16[i];
    for (int i = threadIdx.x; i < a.n_clones; i += blockDim.x) s_erec[i] = a.erec[i];
    for (int i = threadIdx.x; i < a.n_patterns; i += blockDim.x) s_patrec[i] = a.patrec[i];
    for (int i = threadIdx.x; i < a.tape_steps * 32; i += blockDim.x) s_src[i] = a.tape_src[i];
    if (kHeads) {
        // density of one window row: count | weight << 8 of the stones in a 7-bit row slice, by |dy|
        // (Evaluator::BlockWeights, Pattern.cpp:598-609; the matrix is symmetric in dx and dy)
        const int wts[4][7] = { { 1, 3, 4, 0, 4, 3, 1 }, { 0, 3, 5, 4, 5, 3, 0 }, { 0, 4, 3, 3, 3, 4, 0 }, { 2, 0, 0, 1, 0, 0, 2 } };
        for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) {
            int n = 0, w = 0;
            for (int bit = 0; bit < 7; ++bit)
                if ((i >> bit) & 1) { w += wts[i >> 7][bit]; n += wts[i >> 7][bit] > 0; }
            s_lut[i] = uint16_t(n | w << 8);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = a.list_cap;
    WarpSmem& ws = *reinterpret_cast<WarpSmem*>(s_warps + size_t(warp) * warp_bytes(cap, kHeads));
    uint16_t* lists = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(&ws) + sizeof(WarpSmem));   // [lane][cap]
    uint32_t* dflags = (kHeads && !kGuided) ? reinterpret_cast<uint32_t*>(lists + 32 * cap) : nullptr;
    uint16_t* dacc = kGuided ? reinterpret_cast<uint16_t*>(lists + 32 * cap) : nullptr;   // guided: density accumulators instead
    const uint32_t lt = lanemask_lt();
    const uint32_t emit_thr = uint32_t(a.n_clones) * 8u;
    // shared-window addresses, made opaque so the compiler keeps them in registers instead of
    // re-deriving them from the CTA's shared base inside the scan loop
    uint32_t next_addr = smem_addr(s_next), src_addr = smem_addr(s_src + lane);
    uint32_t list_addr = smem_addr(lists + lane * cap);
    asm volatile("" : "+r"(next_addr), "+r"(src_addr), "+r"(list_addr));

    // Deferred compounds (plain evaluator only).  A board yields ~5 compound candidates and ~10 updateAntis window
    // rescans, each a serial chain: run per board they keep 5..10 of 32 lanes busy.  Instead every lane parks at most
    // one candidate (board, cell/player index, its flag word) in registers; when the next board's candidates no
    // longer fit, all parked candidates run together -- Compound::locate, then their rescans spread over the lanes --
    // reading the boards from global memory and adding to the ALREADY STORED score blocks / totals with global atomics.
    const bool defer = !kHeads && a.scores != nullptr && (reinterpret_cast<uintptr_t>(a.cmp_totals) & 3u) == 0;
    // The plain evaluator hands its score blocks to the bulk-copy engine (phase 5) and scans the NEXT board while the block
    // leaves: the scan only touches the board registers and the emission lists, so the accumulators are cleared after it.
    constexpr bool kBulk = !kHeads;
    const uint32_t scores_addr = smem_addr(ws.scores);
    uint32_t pend_idx = 0xffffffffu, pend_flags = 0;
    long long pend_board = 0;
    auto flush_pending = [&]() {
        if (kBulk && lane == 0) bulk_wait_all();                            // the parked boards' score blocks have landed in global memory
        __threadfence();                                                    // the parked boards' stores precede the atomics
        __syncwarp();
        __threadfence();
        uint32_t t0 = 0, t1 = 0;
        if (pend_idx != 0xffffffffu)
            compound_at(a.scores + pend_board * kScoreWords,
                        a.cmp_totals ? reinterpret_cast<uint32_t*>(a.cmp_totals + pend_board * 6) : nullptr, 16, nullptr,
                        pend_flags, pend_idx, t0, t1);
        pend_idx = 0xffffffffu;
        const uint32_t owners = __ballot_sync(0xffffffffu, t0 != 0);
        const int ntask = 2 * __popc(owners);
        for (int s0 = 0; s0 < ntask; s0 += 32) {
            const int s = s0 + lane;
            const int owner = s < ntask ? int(__fns(owners, 0, (s >> 1) + 1)) : 0;
            const uint32_t ta = __shfl_sync(0xffffffffu, t0, owner), tb = __shfl_sync(0xffffffffu, t1, owner);
            const long long tboard = __shfl_sync(0xffffffffu, pend_board, owner);
            const uint32_t task = (s & 1) ? tb : ta;
            if (s < ntask) {
                const uint32_t black = (task >> 8) & 1u;
                anti_cells<true>(a.boards + tboard * kBoardWords, nullptr, 0u, next_addr, uint32_t(a.root_off), emit_thr, s_erec,
                                 s_patrec, int(task & 0xffu), (task >> 9) & 3u, (task >> 11) & 3u,
                                 a.scores + tboard * kScoreWords + (black + 1) * kCells);
            }
        }
        __syncwarp();
    };

    const int warps = blockDim.x >> 5;
    const long long b0 = (long long)blockIdx.x * warps + warp, bstride = (long long)gridDim.x * warps;
    // the warp's next board is requested one board ahead: its 64 bytes are in flight while the current board is evaluated
    uint32_t bw_next = (b0 < a.n && lane < kBoardWords) ? __ldg(a.boards + b0 * kBoardWords + lane) : 0u;
    for (long long b = b0; b < a.n; b += bstride) {
        // ---- phase 0 ---------------------------------------------------------------------------
        uint32_t bw = 0xffffffffu;
        if (lane < kBoardWords) bw = bw_next & ~((bw_next >> 1) & 0x55555555u);   // a cell holding the invalid value 3 reads as white
        if (lane == 14) bw |= 0xfffffffcu;                                  // cells 225.. are pads
        if (lane == 15) bw = 0xffffffffu;
        if (b + bstride < a.n && lane < kBoardWords) bw_next = __ldg(a.boards + (b + bstride) * kBoardWords + lane);
        int played = 0, result = 0;                                         // guided mode: moves played so far / final winner
      next_move:                                                            // guided mode re-enters here after every move
        uint32_t nx = a.start_off;
        uint32_t lp = list_addr;                                            // shared address of the lane's next free list slot
        auto scan = [&]() {
            // ---- phase 2: scan ---------------------------------------------------------------------
            // per step: tape entry -> board word (shuffle) -> cell value * 2 -> next row offset.  The only
            // state-dependent chain is  IADD (offset + value * 2), LDS.U16.
            uint32_t src = src_addr;
            for (int t = 0; t < a.tape_steps; t += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u, src += 64) {
                    const uint32_t e = lds_u16(src);
                    const uint32_t w = __shfl_sync(0xffffffffu, bw, e);     // source lane = e & 31 = board word index
                    const uint32_t v2 = __funnelshift_r(w, w, e >> 8) & 6u; // cell value * 2
                    nx = lds_u16(next_addr + nx + v2);
                    if (nx < emit_thr) {
                        sts_u16(lp, nx * 8u + uint32_t(t + u));             // clone id << 6 | step
                        lp += 2;
                    }
                }
            }
        };
        if (kBulk) {
            scan();                                                         // the previous board's block is still leaving
            if (lane == 0) bulk_wait_read();                                // ... now the copy engine has read it
            __syncwarp();
        }
        if (lane < kBoardSmem) ws.board[lane] = bw;
        {
            int4* z = reinterpret_cast<int4*>(ws.scores);
            for (int i = lane; i < (kScoreWords + kFlagWords) / 4; i += 32) z[i] = make_int4(0, 0, 0, 0);   // scores + flags are contiguous
            if (lane < kTotalWords) ws.totals[lane] = 0;
            if (kHeads && !kGuided) for (int i = lane; i < kDflagWords; i += 32) dflags[i] = 0;
        }
        __syncwarp();

        // ---- phase 1: block score ----------------------------------------------------------------
        uint32_t mine = 0;                                                  // this lane's 15-bit row of stones (also used by the heads)
        {
            const int pg = lane >= 15, y = lane - 15 * pg;                  // lanes 0..14 white rows, 15..29 black rows
            uint32_t occ = 0;
            if (lane < 30) {
                const int off = 30 * y;
                const uint32_t lo = ws.board[off >> 5], hi = ws.board[(off >> 5) + 1];
                const uint32_t v = __funnelshift_r(lo, hi, off & 31) & 0x3fffffffu;
                const uint32_t blk = squeeze_even(v), wht = squeeze_even(v >> 1);
                mine = pg ? blk : wht;
                occ = blk | wht;
            }
            uint32_t near = (mine << 1) | (mine >> 1) | (mine << 2) | (mine >> 2) | (mine << 3) | (mine >> 3);
#pragma unroll
            for (int k = 1; k <= 3; ++k) {
                const uint32_t up = __shfl_sync(0xffffffffu, mine, (lane + k) & 31);
                const uint32_t dn = __shfl_sync(0xffffffffu, mine, (lane - k) & 31);
                const uint32_t u = y + k < kHeight ? up : 0u, d = y - k >= 0 ? dn : 0u;
                const uint32_t both = u | d;
                if (k < 3) near |= both | (both << 1) | (both >> 1) | (both << 2) | (both >> 2);
                else near |= both | (both << 3) | (both >> 3);
            }
            near &= 0x7fffu & ~occ;
            if (lane >= 30) near = 0;
            int* dst = ws.scores + (pg ? 3 : 0) * kCells + y * kWidth;       // scores(P, P), Pattern.cpp:244,268
#pragma unroll
            for (int x = 0; x < kWidth; ++x)
                if (near & (1u << x)) dst[x] = 160;
        }
        __syncwarp();

        if (!kBulk) scan();
        // ---- phase 3: balanced scatter ----------------------------------------------------------------
        uint32_t win = 0;
        {
            uint32_t incl = (lp - list_addr) >> 1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const int total = int(__shfl_sync(0xffffffffu, incl, 31));
            const int excl = int(incl) - int((lp - list_addr) >> 1);            // emissions held by the lanes below this one
            __syncwarp();                                                       // the lanes' lists are complete
            for (int i0 = 0; i0 < total; i0 += 32) {                            // warp-uniform trip count: the search shuffles
                const int i = i0 + lane;
                int j = __shfl_sync(0xffffffffu, excl, 16) <= i ? 16 : 0;       // owner lane: excl[j] <= i < excl[j + 1], by binary
                if (__shfl_sync(0xffffffffu, excl, j + 8) <= i) j += 8;         // search over the prefix counts in the lanes' registers
                if (__shfl_sync(0xffffffffu, excl, j + 4) <= i) j += 4;
                if (__shfl_sync(0xffffffffu, excl, j + 2) <= i) j += 2;
                if (__shfl_sync(0xffffffffu, excl, j + 1) <= i) j += 1;
                const int first = __shfl_sync(0xffffffffu, excl, j);
                if (i >= total) continue;
                const uint32_t ent = lists[j * cap + (i - first)];
                const uint32_t er = s_erec[ent >> 6];
                const uint32_t inf = __ldg(a.tape_info + (ent & 63u) * 32u + uint32_t(j));
                const uint32_t dir = (inf >> 9) & 3u;
                const int vcell = inf & 0x1ff, stride = int(inf >> 11);
                win |= apply_emission(ws, dflags, s_patrec[er_pid(er, 0)], vcell - int(er_prev(er, 0)) * stride, dir, stride);
                const uint32_t p1 = er_pid(er, 1);
                if (p1 != kDevNoPid) win |= apply_emission(ws, dflags, s_patrec[p1], vcell - int(er_prev(er, 1)) * stride, dir, stride);
            }
        }
        __syncwarp();

        // ---- phase 4: compounds ----------------------------------------------------------------------
        {
            unsigned short* clist = lists;                                          // the emission lists are dead now (32 * cap >= 450)
            int cn = 0;
            const uint2* f2 = reinterpret_cast<const uint2*>(ws.flags);             // one cell: white word, black word
#pragma unroll 2
            for (int r = 0; r < (kCells + 31) / 32; ++r) {
                const int cell = r * 32 + lane;
                // cheap necessary condition first: a word with fewer than two raw bits cannot pass Compound::Test
                const uint2 f = f2[cell < kCells ? cell : kCells];                  // flags[450..451] stay zero
                const bool maybe = (((f.x & (f.x - 1)) | (f.y & (f.y - 1))) | ((f.x | f.y) & 0xaaaaaau)) != 0;   // two fields set, or a count of 2
                const uint32_t m = __ballot_sync(0xffffffffu, maybe);
                if (m) {                                                            // rare: a few cells per board
                    // Compound::Test, Pattern.cpp:424-433: per direction the classes' counts are OR-ed (01 | 10 = 11 reads as
                    // "two", like the reference's unary 11) and a binary 10 is widened to 11 before the two-bits test
                    uint32_t bw0 = (f.x | (f.x >> 8) | (f.x >> 16)) & 0xffu, bw1 = (f.y | (f.y >> 8) | (f.y >> 16)) & 0xffu;
                    bw0 |= (bw0 >> 1) & 0x55u;
                    bw1 |= (bw1 >> 1) & 0x55u;
                    const bool h0 = (bw0 & (bw0 - 1)) != 0, h1 = (bw1 & (bw1 - 1)) != 0;
                    const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                    if (h0) clist[cn + __popc(m0 & lt)] = (unsigned short)(cell * 2);
                    cn += __popc(m0);
                    if (h1) clist[cn + __popc(m1 & lt)] = (unsigned short)(cell * 2 + 1);
                    cn += __popc(m1);
                }
            }
            __syncwarp();
            for (int base = 0; base < cn; base += 32) {                             // warp-uniform trip count
                if (!kHeads && defer && cn <= 32) {                                 // park the candidates on free lanes (a board with more is done in place:
                    const int nnew = cn;                                            //  a flush must never meet candidates of a board not stored yet)
                    uint32_t free_lanes = __ballot_sync(0xffffffffu, pend_idx == 0xffffffffu);
                    if (__popc(free_lanes) < nnew) { flush_pending(); free_lanes = 0xffffffffu; }
                    const int r = __popc(free_lanes & lt);                          // this lane's rank among the free lanes
                    if (pend_idx == 0xffffffffu && r < nnew) {
                        pend_idx = clist[base + r];
                        pend_flags = ws.flags[pend_idx];
                        pend_board = b;
                    }
                    continue;
                }
                uint32_t t0 = 0, t1 = 0;
                if (base + lane < cn) {
                    const uint32_t idx = clist[base + lane];
                    compound_at(ws.scores, ws.totals + 16, 32, dflags, ws.flags[idx], idx, t0, t1);
                }
                // spread the window rescans: task s = 2 * (rank of the owning lane) + which
                const uint32_t owners = __ballot_sync(0xffffffffu, t0 != 0);
                const int ntask = 2 * __popc(owners);
                for (int s0 = 0; s0 < ntask; s0 += 32) {
                    const int s = s0 + lane;
                    const int owner = s < ntask ? int(__fns(owners, 0, (s >> 1) + 1)) : 0;
                    const uint32_t ta = __shfl_sync(0xffffffffu, t0, owner), tb = __shfl_sync(0xffffffffu, t1, owner);
                    const uint32_t task = (s & 1) ? tb : ta;
                    if (s < ntask) {
                        const uint32_t black = (task >> 8) & 1u;
                        anti_cells<false>(ws.board, dflags, 1u << (16 + ((task >> 14) & 3u) * 4 + black + 1), next_addr,
                                          uint32_t(a.root_off), emit_thr, s_erec, s_patrec, int(task & 0xffu), (task >> 9) & 3u,
                                          (task >> 11) & 3u, ws.scores + (black + 1) * kCells);
                    }
                }
            }
        }
        __syncwarp();

        if (kHeads) {
            float* prob = reinterpret_cast<float*>(lists);                   // 225 floats over the (dead) emission lists
            int n_stones, to_move;
            policy_heads(ws, prob, s_lut, mine, lane, (a.value && !kGuided) ? a.value + b : nullptr, n_stones, to_move, dacc, kGuided && played > 0);
            if (!kGuided) {
                if (a.decisive) decisive_filter(ws, dflags, prob, to_move, lane);  // TraditionalPolicy::hybridSimulate, Traditional.h:52-53
                if (a.probs)
                    for (int c = lane; c < kCells; c += 32) a.probs[b * kCells + c] = prob[c];
                if (a.dflags) for (int i = lane; i < kCells; i += 32) a.dflags[b * kCells + i] = dflags[i];
            }
            if (kGuided) {
                // Heuristic::EvaluatedRollout (Heuristic.hpp:61-72): while (!ev.checkGameEnd()) applyMove(probs_to_move(...))
                const uint32_t won = __reduce_or_sync(0xffffffffu, win);
                int cell = -1;
                if (won) result = (won & 1u) ? 1 : -1;                       // a Five emission ended the game, Pattern.cpp:140-145
                else if (n_stones < kCells && played < a.g_max_moves) {      // Evaluator::checkGameEnd, Pattern.cpp:343-353
                    uint32_t rnd = 0;
                    if (a.g_mode == 2)
                        rnd = philox_word(uint32_t(played) >> 2, 0u, uint32_t(a.g_game_base) + uint32_t(b), a.g_ctr_hi, a.g_key_lo,
                                          a.g_key_hi, uint32_t(played) & 3u);
                    cell = select_move(prob, lane, a.g_mode, rnd);
                }
                if (cell >= 0) {
                    if (a.g_moves && lane == 0) a.g_moves[b * a.g_max_moves + played] = (int16_t)cell;
                    if (lane == (cell >> 4)) bw |= (to_move ? 1u : 2u) << ((cell & 15) * 2);
                    {   // the new stone's contribution to the density accumulators of its colour: cells within 3 rows and columns
                        const int my = cell / kWidth, mx = cell - my * kWidth, dc = lane >= 15, dx = lane - 15 * dc, j = mx - dx + 3;
                        if (lane < 30 && dc == to_move && j >= 0 && j < 7) {
#pragma unroll 1
                            for (int d = -3; d <= 3; ++d) {
                                const int y = my + d;
                                if (y >= 0 && y < kHeight) dacc[dc * kCells + y * kWidth + dx] += s_lut[(d < 0 ? -d : d) * 128 + (1 << j)];
                            }
                        }
                    }
                    ++played;
                    __syncwarp();
                    goto next_move;
                }
                if (a.g_winner && lane == 0) a.g_winner[b] = (int8_t)result;
                if (a.g_length && lane == 0) a.g_length[b] = (int16_t)played;
                if (a.g_final && lane < kBoardWords) a.g_final[b * kBoardWords + lane] = lane == 14 ? (bw & 3u) : lane == 15 ? 0u : bw;
            }
        }
        // ---- phase 5: output --------------------------------------------------------------------------
        if (a.scores) {
            if (kBulk) {
                fence_async_shared();                                       // the lanes' shared-memory updates, made visible to the copy engine
                __syncwarp();
                if (lane == 0) bulk_store(a.scores + b * kScoreWords, scores_addr, kScoreWords * 4);
            } else {
                const int4* s4 = reinterpret_cast<const int4*>(ws.scores);
                int4* dst = reinterpret_cast<int4*>(a.scores + b * kScoreWords);
                for (int i = lane; i < kScoreWords / 4; i += 32) dst[i] = s4[i];
            }
        }
        if (a.pat_totals && lane < 16) a.pat_totals[b * 16 + lane] = (uint16_t)ws.totals[lane];
        if (a.cmp_totals && lane < 6) a.cmp_totals[b * 6 + lane] = (uint16_t)ws.totals[16 + lane];
        win = __reduce_or_sync(0xffffffffu, win);
        if (a.winner && lane == 0) a.winner[b] = (win & 1u) ? 1 : (win & 2u) ? -1 : 0;
        __syncwarp();
    }
    if (defer && __ballot_sync(0xffffffffu, pend_idx != 0xffffffffu)) flush_pending();
    if (kBulk && lane == 0) bulk_wait_read();                               // shared memory must outlive the copies that read it
}

// PatternSearch::matches for arbitrary symbol strings, one thread per string (test / tooling path).
__global__ void scan_strings_kernel(ScanArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_strings) return;
    const long long lo = a.starts[i], hi = a.starts[i + 1];
    uint32_t st = 0;
    int count = 0;
    int32_t* pids = a.pids + (long long)i * a.max_per_string;
    int32_t* offs = a.offsets + (long long)i * a.max_per_string;
    for (long long p = lo; p < hi; ++p) {
        const uint32_t tw = __ldg(a.trans + st * 4 + (a.codes[p] - 1u));
        st = tw_next(tw);
        const uint32_t ne = tw_nemit(tw);
        for (uint32_t k = 0; k < ne; ++k) {
            const uint32_t em = tw_emit(tw, k);
            if (count < a.max_per_string) { pids[count] = em_pid(em); offs[count] = int(p - lo) - int(em_prev(em)); }
            ++count;
        }
    }
    const int pending = a.flush[st];                                         // a run of xxxxx.. / ooooo.. reaching the end of input
    if (pending >= 0) {
        if (count < a.max_per_string) { pids[count] = pending; offs[count] = int(hi - lo) - 1; }
        ++count;
    }
    a.counts[i] = count;
}

}  // namespace

static size_t table_smem_bytes(const EvalArgs& a) {
    return align16(size_t(a.n_clones + a.n_states) * 8) + align16(size_t(a.n_clones) * 4) +
           align16(size_t(a.n_patterns) * sizeof(PatRec)) + align16(size_t(a.tape_steps) * 64);
}

// warps per CTA: as many as fit beside the tables (32 for the default table; bigger custom tables get fewer)
static bool wants_heads(const EvalArgs& a) {
    return a.probs != nullptr || a.value != nullptr || a.g_mode != 0 || a.dflags != nullptr;
}

// shared tables beside the automaton: the density LUT of the head variants, the quotient LUT of the incremental guided kernel
static bool incremental_guided(const EvalArgs& a) { return a.g_mode != 0 && !a.g_full_rescan && a.list_cap >= 13; }   // a lane's list holds one 13-symbol window
static size_t extra_table_bytes(const EvalArgs& a) {
    return (wants_heads(a) ? 4 * 128 * sizeof(uint16_t) : 0) + (incremental_guided(a) ? align16(size_t(kVlutN) * kVlutW * sizeof(float)) : 0);
}

static int eval_warps(const EvalArgs& a) {
    const size_t tables = table_smem_bytes(a) + extra_table_bytes(a),
                 per_warp = warp_bytes(a.list_cap, wants_heads(a));
    if (tables + per_warp > kSmemLimit) return 0;
    const size_t fit = (kSmemLimit - tables) / per_warp;
    return int(fit < size_t(kWarpsPerCta) ? fit : size_t(kWarpsPerCta));
}

size_t eval_smem_bytes(const EvalArgs& a) {
    return table_smem_bytes(a) + extra_table_bytes(a) + size_t(eval_warps(a)) * warp_bytes(a.list_cap, wants_heads(a));
}

cudaError_t launch_eval(const EvalArgs& a, int sm_count, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    int warps = eval_warps(a);
    if (warps == 0 || a.list_cap * 32 < 2 * kCells || a.tape_steps > kMaxTapeSteps)
        return cudaErrorInvalidConfiguration;                               // table too large for shared memory
    // A small batch is spread over all SMs (fewer warps per CTA, at least 4 to share the table load) instead of filling a
    // few: a warp is one serial chain, and seven of them per scheduler slow each other down (1 024 guided games: 0.99 ->
    // 0.88 ms, 256: 0.93 -> 0.80 ms; no difference from 4 096 games up).
    // Guided playouts may bound the games IN FLIGHT (one warp each): the n games are then a queue that the resident warps
    // work through (the kernels' grid-stride loop), i.e. self-play as a continuous stream with a fixed concurrency.
    const long long units = (a.g_mode != 0 && a.g_in_flight > 0 && a.g_in_flight < a.n) ? a.g_in_flight : a.n;
    const long long per_sm = (units + sm_count - 1) / sm_count;
    if (units >= sm_count && per_sm < warps) warps = (int)(per_sm > 4 ? per_sm : (warps < 4 ? warps : 4));   // (a handful of boards: one full CTA loads the tables fastest)
    const size_t smem = table_smem_bytes(a) + extra_table_bytes(a) + size_t(warps) * warp_bytes(a.list_cap, wants_heads(a));
    if (units < warps) warps = (int)(units > 0 ? units : 1);
    const bool bounded = units < a.n;                                        // a bound on the games in flight is not to be exceeded
    const long long want = bounded ? (units / warps > 0 ? units / warps : 1) : (units + warps - 1) / warps;
    const int grid = (int)(want < sm_count ? want : sm_count);               // persistent: one CTA per SM
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kernel<<<grid, warps * 32, smem, stream>>>(a);
        return cudaGetLastError();
    };
    // guided playouts: the incremental kernels, unless the caller asks for the full rescan.  Few games in flight (at most
    // kPairMax per SM) get two warps each -- the latency-bound regime, see guided_pair_kernel -- more get one.
    if (incremental_guided(a)) {
        const size_t tables = table_smem_bytes(a) + extra_table_bytes(a), per_pair = pair_bytes(a.list_cap);
        const long long fit = (long long)((kSmemLimit - tables) / per_pair) < kPairMax ? (long long)((kSmemLimit - tables) / per_pair) : kPairMax;
        if (!a.g_single_warp && fit >= 1 && units <= (long long)sm_count * fit) {
            const int pairs = (int)(per_sm < 1 ? 1 : per_sm > fit ? fit : per_sm);
            const long long ctas = bounded ? (units / pairs > 0 ? units / pairs : 1) : (units + pairs - 1) / pairs;
            const size_t psmem = tables + size_t(pairs) * per_pair;
            cudaError_t err = cudaFuncSetAttribute(guided_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
            if (err != cudaSuccess) return err;
            guided_pair_kernel<<<(int)(ctas < sm_count ? ctas : sm_count), pairs * 64, psmem, stream>>>(a);
            return cudaGetLastError();
        }
        return launch(guided_kernel);
    }
    if (a.g_mode != 0) return launch(ac_eval_kernel<true, true>);
    return wants_heads(a) ? launch(ac_eval_kernel<true, false>) : launch(ac_eval_kernel<false, false>);
}

cudaError_t launch_scan(const ScanArgs& a, cudaStream_t stream) {
    if (a.n_strings <= 0) return cudaSuccess;
    scan_strings_kernel<<<(a.n_strings + 127) / 128, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gk

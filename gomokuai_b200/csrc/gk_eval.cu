// gk_eval.cu -- K1 `ac_eval`: from-scratch evaluation of a batch of boards, one warp per board.
//
// Replaces, for a batch, Evaluator::applyMove replay + read-out of m_scores / totals
// (reference src/Pattern.cpp:128-302,418-550; semantics restated in SURVEY.md Appendix A).
//
//   phase 0  load the 64-byte board, pad cells = 3, clear the warp's shared-memory accumulators
//   phase 1  stone-density block score (+160 where a player has a stone on a weighted offset of
//            the 7x7 neighbourhood, Pattern.cpp:236-272,598-609) from 15-bit row masks
//   phase 2  the Aho-Corasick scan: 32 lanes walk their chains of whole lines in lock step, one
//            dependent shared-memory table lookup per symbol; emissions are compacted with
//            ballot/popc into a shared queue
//   phase 3  emission scatter: lanes take queue entries, add pattern scores to the '_' / '^' cells
//            (shared-memory atomics), bump totals, set the saturating per-cell flags
//   phase 4  compounds (double-three / four-three / double-four, Pattern.cpp:418-550) from the flags
//   phase 5  coalesced 128-bit store of the four 225-cell score maps + totals + winner
//
// Bound by integer issue and shared-memory latency, not HBM (64 B in, 3.6 KB out per board).
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_format.h"
#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kWarpsPerCta = 32;
constexpr int kQueueCap = 256;                 // words; also reused as the compound list (u16 x 450)
constexpr int kQueueFlush = kQueueCap - 64;    // one step can add at most 2 x 32 entries
constexpr int kScoreWords = 4 * kCells;        // 900
constexpr int kFlagWords = 2 * kCells + 2;     // 452 (16-byte multiple)
constexpr int kBoardSmem = 20;                 // 17 words used (cells up to 271 read as pad)
constexpr int kTotalWords = 24;                // 16 pattern + 6 compound + winner + spare

struct WarpSmem {
    int scores[kScoreWords];                   // [group][cell]
    uint32_t flags[kFlagWords];                // [cell][player grp]: 3 classes x 4 dirs x 2-bit unary count
    uint32_t queue[kQueueCap];
    uint32_t board[kBoardSmem];
    uint32_t totals[kTotalWords];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp block must keep 16-byte alignment");

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
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

// phase 3: apply queue[0 .. n) to the accumulators.  Returns winner bits (1 black, 2 white).
__device__ __forceinline__ uint32_t scatter_emissions(WarpSmem& ws, const PatRec* s_patrec, int n, int lane) {
    uint32_t win = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t ent = ws.queue[i];
        const PatRec rec = s_patrec[ent & 0x1ffu];
        const int vend = (ent >> 9) & 0x1ff;
        const uint32_t dir = (ent >> 18) & 3u;
        const uint32_t type = pr_type(rec.w0), black = pr_black(rec.w0);
        if (type == kTypeFive) { win |= black ? 1u : 2u; continue; }          // Pattern.cpp:140-145
        atomicAdd(&ws.totals[black * 8 + type], 1u);                          // :147
        const int score = dir >= 2 ? int(rec.w1 >> 16) : int(rec.w1 & 0xffffu);   // :151-152
        const int stride = dir_stride(dir);
        int* self = ws.scores + black * 3 * kCells;                          // Group(f, f)
        int* rival = ws.scores + (black + 1) * kCells;                       // Group(f, -f)
        const uint32_t cclass = pr_cclass(rec.w0);
        uint32_t kinds = pr_kinds(rec.w0);
        while (kinds) {
            const int j = (__ffs(kinds) - 1) >> 1;
            const uint32_t kind = (kinds >> (2 * j)) & 3u;
            kinds &= ~(3u << (2 * j));
            const int cell = vend - j * stride;
            atomicAdd(&rival[cell], score);                                  // '_' and '^', :158-161
            if (kind == 1u) {
                atomicAdd(&self[cell], score);
                if (cclass) {                                                // Record::set saturating 00 -> 01 -> 11, :395-400
                    uint32_t* word = &ws.flags[cell * 2 + black];
                    const uint32_t lo = 1u << ((cclass - 1) * 8 + dir * 2);
                    if (atomicOr(word, lo) & lo) atomicOr(word, lo << 1);
                }
            }
        }
    }
    return win;
}

// Compound::updateAntis (Pattern.cpp:520-543): rescan the 13-symbol window centred on `cell`
// from the root state and give +600 (rival's perspective) to the other '_' / '^' cells of the
// first emission of class `cclass` that has `cell` on a '_'.
__device__ __forceinline__ void anti_cells(WarpSmem& ws, const uint32_t* s_trans, const PatRec* s_patrec,
                                           int cell, uint32_t dir, uint32_t cclass, int* rival) {
    const int cx = cell % kWidth, cy = cell / kWidth;
    const int dx = dir == 1 ? 0 : dir == 3 ? -1 : 1, dy = dir == 0 ? 0 : 1;
    const int stride = dir_stride(dir);
    uint32_t st = 0;
    for (int i = 0; i < 13; ++i) {
        const int x = cx + dx * (i - 6), y = cy + dy * (i - 6);
        uint32_t sym = kSymPad;
        if (x >= 0 && x < kWidth && y >= 0 && y < kHeight) sym = (cell_value(ws.board, y * kWidth + x) - 1u) & 3u;
        const uint32_t tw = s_trans[st * 4 + sym];
        st = tw_next(tw);
        const uint32_t ne = tw_nemit(tw);
        for (uint32_t k = 0; k < ne; ++k) {
            const uint32_t em = tw_emit(tw, k);
            const PatRec rec = s_patrec[em_pid(em)];
            const int off = i - int(em_prev(em)) - 6;                       // position of `cell` counted from the pattern's end
            if (pr_cclass(rec.w0) != cclass || off < 0 || off >= int(pr_len(rec.w0))) continue;   // HasCovered, :22-25
            uint32_t kinds = pr_kinds(rec.w0);
            if (((kinds >> (2 * off)) & 3u) != 1u) continue;                 // `cell` must sit on a '_'
            kinds &= ~(3u << (2 * off));
            while (kinds) {
                const int j = (__ffs(kinds) - 1) >> 1;
                kinds &= ~(3u << (2 * j));
                atomicAdd(&rival[cell + (off - j) * stride], 600);
            }
            return;                                                          // only the first such pattern, :540
        }
    }
}

// phase 4 for one (cell, player) whose flags passed Compound::Test.
__device__ __forceinline__ void compound_at(WarpSmem& ws, const uint32_t* s_trans, const PatRec* s_patrec, uint32_t idx) {
    const uint32_t f = ws.flags[idx];
    const int cell = idx >> 1;
    const uint32_t black = idx & 1u;
    // Compound::locate, Pattern.cpp:440-486.  states: S0 0, L2 1, LD3 2, To33 3, To43 4, To44 5
    int state = 0, l3 = 0, ncomp = 0;
    bool triple = false;
    uint32_t adir[2] = { 0, 0 }, aclass[2] = { 0, 0 };
#pragma unroll
    for (uint32_t dir = 0; dir < 4; ++dir) {
        const uint32_t c1 = (f >> (dir * 2)) & 3u, c2 = (f >> (8 + dir * 2)) & 3u, c3 = (f >> (16 + dir * 2)) & 3u;
        const uint32_t cls = c1 ? 1u : c2 ? 2u : c3 ? 3u : 0u;              // LiveThree > DeadThree > LiveTwo
        if (!cls) continue;
        const uint32_t bits = cls == 1u ? c1 : cls == 2u ? c2 : c3;
        const int count = bits == 3u ? 2 : 1, cond = cls == 3u ? 1 : 2;
        l3 += cls == 1u;
        for (int i = 0; i < count; ++i) {
            if (ncomp < 2) { adir[ncomp] = dir; aclass[ncomp] = cls; }
            ++ncomp;
            int offset;
            if (state == 0) offset = 0;
            else if (state <= 2) offset = 1;
            else { triple = true; offset = state == 5 ? -cond : -1; }
            state += cond + offset;
        }
    }
    const int type = state - 3;
    if (type < 0) return;   // cannot happen for flags that passed Test unless one line holds an L3 and two lower-class patterns on the same cell
    int* self = ws.scores + black * 3 * kCells;
    int* rival = ws.scores + (black + 1) * kCells;
    atomicAdd(&self[cell], 600 * ncomp);                                     // updateCritical, :515-518
    atomicAdd(&rival[cell], 600 * ncomp);
    atomicAdd(&ws.totals[16 + black * 3 + type], 1u);                        // one compound, :505-508
    if (!triple && l3 == 0) {                                                // exactly two components here
        anti_cells(ws, s_trans, s_patrec, cell, adir[0], aclass[0], rival);
        anti_cells(ws, s_trans, s_patrec, cell, adir[1], aclass[1], rival);
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
ac_eval_kernel(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_trans = reinterpret_cast<uint32_t*>(smem_raw);
    PatRec* s_patrec = reinterpret_cast<PatRec*>(s_trans + ((a.n_states * 4 + 3) & ~3));
    uint32_t* s_tape = reinterpret_cast<uint32_t*>(s_patrec + ((a.n_patterns + 1) & ~1));
    WarpSmem* s_warps = reinterpret_cast<WarpSmem*>(s_tape + a.tape_steps * 32);

    for (int i = threadIdx.x; i < a.n_states * 4; i += blockDim.x) s_trans[i] = a.trans[i];
    for (int i = threadIdx.x; i < a.n_patterns; i += blockDim.x) s_patrec[i] = a.patrec[i];
    for (int i = threadIdx.x; i < a.tape_steps * 32; i += blockDim.x) s_tape[i] = a.tape[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem& ws = s_warps[warp];
    const uint32_t lt = lanemask_lt();

    for (long long b = (long long)blockIdx.x * kWarpsPerCta + warp; b < a.n; b += (long long)gridDim.x * kWarpsPerCta) {
        // ---- phase 0 ---------------------------------------------------------------------------
        uint32_t bw = 0xffffffffu;
        if (lane < kBoardWords) bw = __ldg(a.boards + b * kBoardWords + lane);
        if (lane == 14) bw |= 0xfffffffcu;                                  // cells 225.. are pads
        if (lane == 15) bw = 0xffffffffu;
        if (lane < kBoardSmem) ws.board[lane] = bw;
        {
            int4* z = reinterpret_cast<int4*>(ws.scores);
            for (int i = lane; i < (kScoreWords + kFlagWords) / 4; i += 32) z[i] = make_int4(0, 0, 0, 0);   // scores + flags are contiguous
            if (lane < kTotalWords) ws.totals[lane] = 0;
        }
        __syncwarp();

        // ---- phase 1: block score ----------------------------------------------------------------
        {
            const int pg = lane >= 15, y = lane - 15 * pg;                  // lanes 0..14 white rows, 15..29 black rows
            uint32_t mine = 0, occ = 0;
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
            while (near) {
                const int x = __ffs(near) - 1;
                near &= near - 1;
                dst[x] = 160;
            }
        }
        __syncwarp();

        // ---- phase 2: scan -------------------------------------------------------------------------
        uint32_t st = a.start_state, win = 0;
        int qn = 0;
        for (int t = 0; t < a.tape_steps; ++t) {
            const uint32_t e = s_tape[t * 32 + lane];
            const uint32_t sym = (cell_value(ws.board, tp_src(e)) - 1u) & 3u;
            if (e & kTapeStart) st = a.start_state;
            const uint32_t tw = s_trans[st * 4 + sym];
            st = tw_next(tw);
            const uint32_t ne = tw_nemit(tw);
            const uint32_t m1 = __ballot_sync(0xffffffffu, ne != 0);
            if (m1) {
                const uint32_t vcell = tp_vcell(e), dirbits = e & (3u << 18);
                const uint32_t stride = (e >> 21) & 31u;
                if (ne) {
                    const uint32_t em = tw_emit(tw, 0);
                    ws.queue[qn + __popc(m1 & lt)] = em_pid(em) | ((vcell - em_prev(em) * stride) << 9) | dirbits;
                }
                qn += __popc(m1);
                const uint32_t m2 = __ballot_sync(0xffffffffu, ne > 1);
                if (m2) {
                    if (ne > 1) {
                        const uint32_t em = tw_emit(tw, 1);
                        ws.queue[qn + __popc(m2 & lt)] = em_pid(em) | ((vcell - em_prev(em) * stride) << 9) | dirbits;
                    }
                    qn += __popc(m2);
                }
                if (qn > kQueueFlush) {
                    __syncwarp();
                    win |= scatter_emissions(ws, s_patrec, qn, lane);
                    __syncwarp();
                    qn = 0;
                }
            }
        }
        __syncwarp();
        win |= scatter_emissions(ws, s_patrec, qn, lane);
        __syncwarp();

        // ---- phase 4: compounds ----------------------------------------------------------------------
        {
            unsigned short* clist = reinterpret_cast<unsigned short*>(ws.queue);
            int cn = 0;
            for (int r = 0; r < (2 * kCells + 31) / 32; ++r) {
                const int idx = r * 32 + lane;
                bool hit = false;
                if (idx < 2 * kCells) {
                    const uint32_t f = ws.flags[idx];
                    const uint32_t bits = (f | (f >> 8) | (f >> 16)) & 0xffu;       // Compound::Test, Pattern.cpp:424-433
                    hit = (bits & (bits - 1)) != 0;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    if (hit) clist[cn + __popc(m & lt)] = (unsigned short)idx;
                    cn += __popc(m);
                }
            }
            __syncwarp();
            for (int i = lane; i < cn; i += 32) compound_at(ws, s_trans, s_patrec, clist[i]);
        }
        __syncwarp();

        // ---- phase 5: output --------------------------------------------------------------------------
        if (a.scores) {
            const int4* src = reinterpret_cast<const int4*>(ws.scores);
            int4* dst = reinterpret_cast<int4*>(a.scores + b * kScoreWords);
            for (int i = lane; i < kScoreWords / 4; i += 32) dst[i] = src[i];
        }
        if (a.pat_totals && lane < 16) a.pat_totals[b * 16 + lane] = (uint16_t)ws.totals[lane];
        if (a.cmp_totals && lane < 6) a.cmp_totals[b * 6 + lane] = (uint16_t)ws.totals[16 + lane];
        win = __reduce_or_sync(0xffffffffu, win);
        if (a.winner && lane == 0) a.winner[b] = (win & 1u) ? 1 : (win & 2u) ? -1 : 0;
        __syncwarp();
    }
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

size_t eval_smem_bytes(const EvalArgs& a) {
    return size_t((a.n_states * 4 + 3) & ~3) * 4 + size_t((a.n_patterns + 1) & ~1) * sizeof(PatRec) +
           size_t(a.tape_steps) * 32 * 4 + size_t(kWarpsPerCta) * sizeof(WarpSmem);
}

cudaError_t launch_eval(const EvalArgs& a, int sm_count, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    const size_t smem = eval_smem_bytes(a);
    cudaError_t err = cudaFuncSetAttribute(ac_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const long long want = (a.n + kWarpsPerCta - 1) / kWarpsPerCta;
    const int grid = (int)(want < sm_count ? want : sm_count);               // persistent: one 32-warp CTA per SM
    ac_eval_kernel<<<grid, kWarpsPerCta * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scan(const ScanArgs& a, cudaStream_t stream) {
    if (a.n_strings <= 0) return cudaSuccess;
    scan_strings_kernel<<<(a.n_strings + 127) / 128, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gk

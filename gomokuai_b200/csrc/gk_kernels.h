// gk_kernels.h -- launch interface between the C-ABI layer and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "gk_format.h"

namespace gk {

struct EvalArgs {
    const uint16_t* next16; const uint32_t* erec; int n_states, n_clones;   // device copies of the compiled table (gk_format.h)
    const PatRec* patrec; int n_patterns;
    const uint16_t* tape_src; const uint16_t* tape_info; int tape_steps;
    int root_off, start_off, list_cap, trail_pad;
    const uint32_t* boards; long long n;
    int32_t* scores; uint16_t* pat_totals; uint16_t* cmp_totals; int8_t* winner;   // any may be null
    float* probs; float* value;                 // policy heads for the side to move (null = not wanted)
    int decisive;                               // apply Heuristic::DecisiveFilter to probs
    uint32_t* dflags;                           // [n][225] per-cell decisive flag bits (gk_eval.cu), for inspection; may be null
    // guided playouts (g_mode != 0): every warp plays its board to the end inside the kernel
    int g_mode;                                 // 0 off, 1 most probable move, 2 move drawn from the probabilities
    int g_in_flight;                            // > 0: at most this many games are played at the same time (one warp each); the rest queue
    int g_single_warp;                          // never give a game two warps (guided_pair_kernel), whatever the batch size; for tests
    int g_full_rescan;                          // re-evaluate the whole board after every move (the first implementation; kept to test the incremental kernel against)
    uint32_t g_key_lo, g_key_hi, g_ctr_hi; int g_game_base, g_max_moves;
    int8_t* g_winner; int16_t* g_length; int16_t* g_moves; uint32_t* g_final;   // per game; any may be null
};
size_t eval_smem_bytes(const EvalArgs& a);
cudaError_t launch_eval(const EvalArgs& a, int sm_count, cudaStream_t stream);

struct ScanArgs {
    const uint32_t* trans; const int16_t* flush;   // host-format transition words (symbol-indexed)
    const uint8_t* codes; const long long* starts; int n_strings; int max_per_string;
    int32_t* pids; int32_t* offsets; int32_t* counts;
};
cudaError_t launch_scan(const ScanArgs& a, cudaStream_t stream);

struct RolloutArgs {
    const uint32_t* boards; int n; int rollouts_per_pos;
    uint32_t key_lo, key_hi, ctr_hi; int pos_base;
    const uint8_t* r_stream; int stream_stride;     // non-null: injected start indices instead of Philox
    int32_t* wdb; int8_t* winners; uint8_t* lengths; // any may be null
    uint8_t* moves;                                  // launch_rollout_small only: [n][rollouts_per_pos][225] cells played; may be null
};
// `images`: rollout_scratch_bytes(a.n) bytes of device scratch that stay untouched until the launch has finished
// (one buffer per stream in flight)
size_t rollout_scratch_bytes(int n);
cudaError_t launch_rollout(const RolloutArgs& a, int sm_count, uint32_t* images, cudaStream_t stream);
// a handful of positions with at most 256 rollouts each in ONE launch: a.boards, `out` (int32[n][3], may be null) and
// a.winners / a.lengths / a.moves may be page-locked host memory; same streams and counts as launch_rollout
cudaError_t launch_rollout_small(const RolloutArgs& a, int32_t* out, cudaStream_t stream);
// The latency form (gk_rollout_warp.cu): one WARP per rollout, line slots in registers, a.rollouts_per_pos <= 32.  Same
// arguments and results as launch_rollout_small; rollout_warp_fits says whether the batch is one resident wave of it.
bool rollout_warp_fits(int n, int rollouts_per_pos, int sm_count);
// h_board0: when a.n == 1, a HOST-readable copy of the one board (may be null): it is passed in the kernel parameters.
cudaError_t launch_rollout_warp(const RolloutArgs& a, int32_t* out, cudaStream_t stream, const uint32_t* h_board0 = nullptr);
// number of kernels one launch_rollout call enqueues (slot images + wdb clear, rollouts)
int rollout_launches(const RolloutArgs& a);

struct EncodeArgs {
    const uint32_t* boards; const int16_t* last_moves;   // last_moves: [n][2] = last, second-to-last (-1 = none); may be null
    long long n; int augment;                             // augment: 0 = one variant, 1 = the 8 rotations / reflections
    uint8_t* planes;                                      // [n][variants][6][225]
    const float* probs; float* probs_out;                 // optional: [n][225] -> [n][variants][225]
};
cudaError_t launch_encode(const EncodeArgs& a, int sm_count, cudaStream_t stream);

struct ExpandArgs {
    const uint32_t* boards0; const int16_t* moves; const int16_t* lengths; const int8_t* winners;   // winners may be null
    long long n; int max_moves; const long long* starts;                                            // starts: exclusive prefix sum of lengths
    uint32_t* out_boards; int16_t* out_last; int8_t* out_z;                                          // out_last / out_z may be null
};
cudaError_t launch_expand_games(const ExpandArgs& a, cudaStream_t stream);

// gk_peaks.cu: sustained warp instructions per second of a register-only stream (mode 0 LOP3, 1 IMAD, 2 both 1:1)
cudaError_t measure_issue_peak(int mode, int sm_count, int iters, double* warp_inst_per_s, cudaStream_t stream);

}  // namespace gk

/* gomoku_b200.h -- C-ABI of the B200-native hot path of Vigilans/GomokuAI.
 *
 * One shared library (gomokuai_b200/lib/libgomoku_b200.so, sm_100a only) replaces, for the
 * batched case, the reference's pattern evaluator and random-rollout simulator.  Every
 * entry point below names the reference interface it stands in for (paths relative to
 * /root/reference/core/lib/).  Signatures carry plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns a gk_status (0 = ok, < 0 = error) and never throws; the text
 *     of the last error on the calling thread is available from gk_last_error();
 *   - pointers named d_* are DEVICE pointers on the device given to gk_init(); pointers
 *     named h_* are HOST pointers (pinned memory makes the *_host calls faster, pageable
 *     memory is accepted); the caller owns every buffer, the library owns gk_table;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device entry
 *     points only enqueue work; *_host entry points return after the results are in h_*;
 *   - there is NO CPU fallback: without a usable sm_100 device every compute call fails
 *     with GK_ERR_NO_DEVICE.
 *
 * Data formats
 *   board      16 x uint32 per position (64 B): cell c = y*15 + x lives in word c/16,
 *              bits 2*(c%16)..+1; 0 = empty, 1 = black ('x'), 2 = white ('o'); the 62 unused
 *              high bits of word 14..15 are ignored.  Black moves first and the players
 *              alternate (include/Game.h:128, src/Game.cpp:37-47), so the side to move is
 *              black iff #black == #white.
 *   scores     int32[4][225] per position = Evaluator::m_scores (include/Pattern.h:219),
 *              group index g = 2*(favour==Black) + (perspective==Black) (Pattern.h:159-161)
 *   pat_totals uint16[2][8] per position: [0=White,1=Black][Pattern::Type DeadOne..LiveFour]
 *              = Evaluator::m_patternDist.back()[type].get(player) (Pattern.h:216, Pattern.cpp:413)
 *   cmp_totals uint16[2][3] per position: [player][DoubleThree,FourThree,DoubleFour]
 *              = Evaluator::m_compoundDist.back()[type].get(player) (Pattern.h:217)
 *   winner     int8 per position: +1 black / -1 white has five-or-more in a row (a Five
 *              emission, Pattern.cpp:140-145), 0 otherwise
 *   wdb        int32[3] per position: rollouts won by {white, nobody (draw), black}
 */
#ifndef GOMOKU_B200_H_
#define GOMOKU_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int gk_status;
enum {
    GK_OK = 0,
    GK_ERR_INVALID = -1,     /* bad argument */
    GK_ERR_NO_DEVICE = -2,   /* no sm_100 device / library built without a matching cubin */
    GK_ERR_CUDA = -3,        /* CUDA runtime error, see gk_last_error() */
    GK_ERR_TABLE = -4,       /* prototypes cannot be compiled to the flat table */
    GK_ERR_NOT_INIT = -5,    /* gk_init() has not been called on this process */
    GK_ERR_NCCL = -6
};

enum { GK_WIDTH = 15, GK_HEIGHT = 15, GK_CELLS = 225, GK_BOARD_WORDS = 16,
       GK_SCORE_GROUPS = 4, GK_PATTERN_TYPES = 8, GK_COMPOUND_TYPES = 3 };

typedef struct gk_table gk_table;

/* ---- lifecycle ---------------------------------------------------------------------- */
gk_status gk_init(int device);                      /* binds the process to one GPU (one process per GPU); every later
                                                        gk_* call rebinds its calling THREAD to that device */
gk_status gk_shutdown(void);
const char* gk_last_error(void);
gk_status gk_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor);
const char* gk_version(void);
/* Microbenchmark behind the integer-issue roofline: sustained warp instructions per second of a
 * register-only stream on this GPU at its current clocks -- mode 0: LOP3 only (the ALU pipe, which LOP3 / SHF /
 * IADD3 / PRMT / ISETP share), mode 1: IMAD only (the FMA pipe), mode 2: both 1 : 1 (the ceiling of a balanced
 * integer kernel).  Takes a few milliseconds. */
gk_status gk_measure_issue_peak(int mode, double* warp_inst_per_s);

/* ---- pattern automaton --------------------------------------------------------------
 * Replaces the static `PatternSearch Evaluator::Patterns` (src/Pattern.cpp:554-596) and
 * AhoCorasickBuilder::build (src/utils/ACAutomata.cpp:15-23).  The prototypes go through the
 * same reverse / colour-flip / boundary augmentation and ordering, and the resulting goto /
 * fail / output behaviour (including the emission-on-fail-landing and invariant-run rules
 * of PatternSearch::generator::operator++, src/Pattern.cpp:33-56) is compiled into one flat
 * (state, symbol) -> {next, up to 2 emissions} table that lives in shared memory on the GPU. */
gk_status gk_table_default(gk_table** out);         /* the reference's 41 prototypes; cached, do not free */
gk_status gk_table_build(const char* const* protos, const int* types, const int* scores, int n,
                         gk_table** out);           /* protos: "+xxxxx" / "-_oooo_" ... as Pattern.cpp:14-18 */
gk_status gk_table_free(gk_table* table);
gk_status gk_table_info(const gk_table* table, int* n_states, int* n_patterns, int* trail_pad, int* max_steps);
gk_status gk_table_pattern(const gk_table* table, int id, char str8[8], int* favour, int* type, int* score);
/* flat transition words (n_states*4 uint32, see gk_format.h) -- for inspection and tests */
gk_status gk_table_entries(const gk_table* table, uint32_t* h_entries, int capacity);
/* per state: pattern id still owed if the input ends in that state (a run of five-or-more that
 * reaches the end of the string, Pattern.cpp:40-45,54), else -1; n_states int16 */
gk_status gk_table_flush(const gk_table* table, int16_t* h_flush, int capacity);

/* The re-encoded automaton the eval kernel reads (csrc/gk_format.h): (n_clones + n_states) rows of four
 * uint16 row offsets and n_clones uint32 emission records.  info[6] = {n_rows, n_clones, root_off,
 * start_off, list_cap, tape_steps}.  h_next / h_erec may be NULL to query the sizes only.  For tests. */
gk_status gk_table_device_format(const gk_table* table, int info[6], uint16_t* h_next, int next_capacity,
                                 uint32_t* h_erec, int erec_capacity);

/* PatternSearch::execute / matches (src/Pattern.cpp:64-74) for a batch of symbol strings:
 * string i = d_codes[d_starts[i] .. d_starts[i+1]) with symbols 1..4 (EncodeCharset,
 * include/Mapping.h:40-48).  Emissions (pattern id, end offset) of string i are written to
 * d_pids/d_offsets[i*max_per_string ..]; d_counts[i] receives the number found (which may
 * exceed max_per_string; only that many are stored). */
gk_status gk_scan_batch(const gk_table* table, const uint8_t* d_codes, const int64_t* d_starts, int n_strings,
                        int max_per_string, int32_t* d_pids, int32_t* d_offsets, int32_t* d_counts, void* stream);

/* ---- board evaluation ---------------------------------------------------------------
 * Replaces, for a batch of positions, Evaluator::syncWithBoard / applyMove (src/Pattern.cpp:
 * 306-369) followed by reading m_scores, m_patternDist.back(), m_compoundDist.back() and the
 * winner.  Any output pointer may be NULL.  d_scores must be 16-byte aligned. */
gk_status gk_eval_batch(const gk_table* table, const uint32_t* d_boards, int n,
                        int32_t* d_scores, uint16_t* d_pat_totals, uint16_t* d_cmp_totals, int8_t* d_winner,
                        void* stream);
gk_status gk_eval_batch_host(const gk_table* table, const uint32_t* h_boards, int n,
                             int32_t* h_scores, uint16_t* h_pat_totals, uint16_t* h_cmp_totals, int8_t* h_winner);

/* ---- policy heads of the pattern evaluator ("next" row f1 of SURVEY.md section 8) ----------------
 * gk_eval_batch plus, fused into the same kernel, what TraditionalPolicy::hybridSimulate reads from
 * the evaluator for the side to move (include/policies/Traditional.h:49-69, include/algorithms/
 * Heuristic.hpp:16-45): d_probs float[n][225] = Heuristic::EvaluationProbs (density-weighted scores,
 * L2-normalised as Eigen's normalized(); a single 1.0 on the centre for an empty board) and d_value
 * float[n] = Heuristic::EvaluationValue (tanh of the weighted score balance).  Any output may be
 * NULL; with d_scores == NULL only 904 bytes per position leave the GPU instead of 3 645.
 * Floating point: equal to the reference up to summation order (see tests/test_heads.py). */
gk_status gk_eval_policy_batch(const gk_table* table, const uint32_t* d_boards, int n, float* d_probs, float* d_value,
                               int32_t* d_scores, uint16_t* d_pat_totals, uint16_t* d_cmp_totals, int8_t* d_winner,
                               void* stream);
gk_status gk_eval_policy_batch_host(const gk_table* table, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                                    int8_t* h_winner);
/* TraditionalPolicy::hybridSimulate (include/policies/Traditional.h:49-69) for a batch of leaf positions:
 * EvaluationProbs, then Heuristic::DecisiveFilter (Heuristic.hpp:93-161: if a four / live three / compound of
 * either side decides the position, only its key cells keep probability, re-normalised), and EvaluationValue.
 * d_dflags (nullable): uint32[n][225], which (pattern type >= DeadThree | compound type, favour, perspective)
 * flags are set per cell -- bit (type-4)*4 + g or 16 + compound*4 + g, g = 2*(favour==Black) + (perspective==Black). */
gk_status gk_hybrid_simulate_batch(const gk_table* table, const uint32_t* d_boards, int n, float* d_probs, float* d_value,
                                   int8_t* d_winner, uint32_t* d_dflags, void* stream);
gk_status gk_hybrid_simulate_batch_host(const gk_table* table, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                                        int8_t* h_winner);

/* ---- pattern-guided playouts (BASELINE config 5) -------------------------
 * Replaces Heuristic::EvaluatedRollout (include/algorithms/Heuristic.hpp:61-91) for n independent games:
 * until the evaluator reports a five or the board is full, the side to move plays the move chosen from
 * Heuristic::EvaluationProbs -- mode 1: the most probable cell, lowest index among equals
 * (MaxEvaluatedRollout); mode 2: a draw from the probabilities (RandomEvaluatedRollout).  Every game runs
 * to its end inside ONE kernel, one warp per game: the start position is evaluated from scratch, then every move
 * re-scans only the four lines through the new stone -- before and after, emissions taken back and added, the
 * reference's own incremental scheme (Updater::updateMove, src/Pattern.cpp:274-302).  mode | GK_GUIDED_FULL_RESCAN
 * re-evaluates the whole board after every move instead (the first implementation; both play identical games, which
 * the tests check).  With at most 15 games per SM in flight a game gets TWO warps (evaluator update and density weights of
 * a move run side by side; same games again); mode | GK_GUIDED_SINGLE_WARP keeps it on one.
 * Randomness (mode 2): weights w = round(p * 2^20), cells in index order, r = mulhi32(word, sum w) with word
 * (k & 3) of Philox4x32-10(counter = {k >> 2, 0, game_base + i, ctr_hi}, key) for move k of game i; the
 * first cell whose running sum exceeds r is played.  (The reference draws from std::discrete_distribution
 * over a process-global mt19937, Game.cpp:75-78, which is implementation defined.)
 * d_winner int8[n] (+1 black, -1 white, 0 draw or max_moves reached), d_length int16[n] moves played,
 * d_moves int16[n][max_moves] (nullable) the cells played, d_final_boards uint32[n][16] (nullable). */
enum { GK_GUIDED_FULL_RESCAN = 0x100, GK_GUIDED_SINGLE_WARP = 0x200 };
gk_status gk_guided_rollout_batch(const gk_table* table, const uint32_t* d_boards, int n, int mode, uint64_t philox_key,
                                  uint32_t ctr_hi, int game_base, int max_moves, int8_t* d_winner, int16_t* d_length,
                                  int16_t* d_moves, uint32_t* d_final_boards, void* stream);

/* The same with at most `max_in_flight` games being played at any moment (one warp each; <= 0: as many as the GPU holds):
 * the n games form a queue that the resident warps work through.  This is self-play as a continuous stream with a fixed
 * number of concurrent games (BASELINE config 5) -- one batch of that size lasts as long as its LONGEST game, a stream
 * runs at the rate of the mean.  Results are the same as gk_guided_rollout_batch's. */
gk_status gk_guided_rollout_queue(const gk_table* table, const uint32_t* d_boards, int n, int max_in_flight, int mode,
                                  uint64_t philox_key, uint32_t ctr_hi, int game_base, int max_moves, int8_t* d_winner,
                                  int16_t* d_length, int16_t* d_moves, uint32_t* d_final_boards, void* stream);

/* ---- random rollouts ----------------------------------------------------------------
 * Replaces Default::RandomRollout / Default::Simulate (include/algorithms/MonteCarlo.hpp:
 * 37-47,83-88) and RandomPolicy::averagedSimulate (include/policies/Random.h:22-35) for
 * `rollouts_per_pos` independent playouts from each of n positions: every move draws a start
 * index r in [0,225) and plays the first empty cell at or after r, cyclically
 * (Board::getRandomMove, src/Game.cpp:64-73); the game ends on five-or-more in a row through
 * the last stone or on a full board (Board::checkGameEnd, src/Game.cpp:88-136).
 * A position that is already decided (five on board) or full plays 0 moves.
 *
 * Randomness: Philox4x32-10, key = philox_key; move k of rollout j of position (pos_base+i)
 * uses word (k & 3) of philox(counter = {k >> 2, j, pos_base + i, ctr_hi}); r = mulhi32(word, 225).
 * Results therefore do not depend on how positions or rollouts are split across GPUs.
 * d_winners / d_lengths (nullable) receive per-rollout outcome (+1/-1/0) and move count. */
gk_status gk_rollout_batch(const uint32_t* d_boards, int n, int rollouts_per_pos,
                           uint64_t philox_key, uint32_t ctr_hi, int pos_base,
                           int32_t* d_wdb, int8_t* d_winners, uint8_t* d_lengths, void* stream);
gk_status gk_rollout_batch_host(const uint32_t* h_boards, int n, int rollouts_per_pos,
                                uint64_t philox_key, uint32_t ctr_hi, int pos_base, int32_t* h_wdb);
/* One position, `rollouts` (<= 256) playouts WITH their move lists, one fused launch -- what
 * PoolRAVEPolicy::defaultSimulate needs: it leaves the board at the END of its playout (include/policies/PoolRAVE.h:
 * 27-48) because RAVE::BackPropogate reads the final stones (include/algorithms/MonteCarlo.hpp:155-184).  Same Philox
 * stream as gk_rollout_batch for position index `pos`.  h_winners int8[rollouts] (+1 black, -1 white, 0 draw),
 * h_lengths uint8[rollouts], h_moves uint8[rollouts][225]: the cells in playing order (first h_lengths[r] valid). */
gk_status gk_rollout_trace_host(const uint32_t* h_board, int rollouts, uint64_t philox_key, uint32_t ctr_hi, int pos,
                                int8_t* h_winners, uint8_t* h_lengths, uint8_t* h_moves);
/* Asynchronous form of gk_rollout_batch_host for callers that keep several small batches in flight (the
 * root-parallel search: while one group of leaves is simulated, the host descends the trees of the next one).
 * `slot` in [0, 16) names an independent stream with its own device buffers.  submit enqueues copy-in, rollouts and
 * copy-out and returns; h_boards and h_wdb must stay valid and untouched until gk_rollout_wait(slot) has returned.
 * Page-locked buffers (gk_host_alloc) make the call truly asynchronous and let the kernel read the boards in place.
 * With both buffers page-locked and n <= 4096 the batch is one launch that stores a position's three counts into h_wdb
 * as that position's block retires; counts are never negative, so a caller that fills h_wdb with a negative value before
 * the submit may WATCH them arrive (volatile reads) instead of paying gk_rollout_wait's stream synchronisation -- the
 * root-parallel search and the MCTS mirror do.  One thread at a time per slot. */
gk_status gk_rollout_submit_host(int slot, const uint32_t* h_boards, int n, int rollouts_per_pos,
                                 uint64_t philox_key, uint32_t ctr_hi, int pos_base, int32_t* h_wdb);
gk_status gk_rollout_wait(int slot);
/* Same loop, but move k of rollout j of position i takes r = d_r_stream[(i*rollouts_per_pos + j)
 * * stream_stride + k] (values 0..224) -- the injected-stream protocol used to compare bit-exactly
 * with the reference's Board on any external stream (e.g. its own mt19937 draws).  A rollout
 * that needs more than stream_stride draws reports length 255 and winner 0. */
gk_status gk_rollout_injected(const uint32_t* d_boards, int n, int rollouts_per_pos,
                              const uint8_t* d_r_stream, int stream_stride,
                              int8_t* d_winners, uint8_t* d_lengths, void* stream);

/* ---- self-play feature planes ("next" row f3 of SURVEY.md section 8) ------------------
 * Replaces Board.encoded_states() of CorePyExt (core/py_ext/src/game_ext.hpp:87-104) for a batch,
 * and, with augment != 0, augment_game_data's 8 rotations / reflections of the planes
 * (network/data_helper.py:36-55; order: for i in 0..3: rot90(i), fliplr(rot90(i))).
 * d_last_moves: int16[n][2] = {last move, second-to-last move} as cell ids, -1 = none; NULL = none.
 * d_planes: uint8[n][V][6][15][15], V = 8 if augment else 1 (no alignment requirement).
 * Planes: stones of the side to move, stones of the opponent, empty cells, last move, second-to-last
 * move (one-hot), all-ones iff black is to move.
 * d_probs (nullable): float[n][225] move probabilities; d_probs_out: float[n][V][225] receives them
 * under the same rotations / reflections (augment_game_data's rot_probs / flip_probs). */
gk_status gk_encode_states_batch(const uint32_t* d_boards, const int16_t* d_last_moves, int n, int augment,
                                 uint8_t* d_planes, const float* d_probs, float* d_probs_out, void* stream);

/* The positions a batch of finished games went through, ready for gk_encode_states_batch (self-play samples; the
 * reference collects them move by move in dual_play, agents/utils.py:29-57).  Game g started from d_boards0[g] and
 * played d_moves[g][0 .. d_lengths[g]) (the outputs of gk_guided_rollout_batch); d_starts int64[n] is the exclusive
 * prefix sum of the lengths.  The position before ply k of game g is written to d_out_boards[d_starts[g] + k],
 * d_out_last_moves (nullable) receives {last move, second-to-last move} of that position (-1 = none, counted from the
 * start position) and d_out_z (nullable, needs d_winners) the game's outcome from the view of its side to move. */
gk_status gk_expand_games(const uint32_t* d_boards0, const int16_t* d_moves, const int16_t* d_lengths, const int8_t* d_winners,
                          int n, int max_moves, const int64_t* d_starts, uint32_t* d_out_boards, int16_t* d_out_last_moves,
                          int8_t* d_out_z, void* stream);

/* ---- root-parallel exchange (BASELINE config 4) -------------------------------------------------------
 * The only collective of the path: one allreduce(sum) of the int64[3][225] root statistics per move
 * ([0] visits, [1] black-won, [2] white-won rollouts per root child) over NCCL (NVLink / NVSwitch).
 * libnccl.so.2 is resolved at run time (dlopen), so the library has no link-time NCCL dependency and
 * shares the copy already loaded by the process (e.g. torch's).
 *   gk_nccl_unique_id   rank 0 creates the 128-byte id and hands it to the other ranks by any side channel
 *   gk_nccl_init        every rank (one process per GPU, after gk_init) joins the communicator
 *   gk_root_allreduce   in-place sum of d_stats over the ranks; `nccl_comm` = an ncclComm_t of the caller,
 *                       or NULL for the communicator made by gk_nccl_init */
gk_status gk_nccl_unique_id(uint8_t id[128]);
gk_status gk_nccl_init(const uint8_t id[128], int world_size, int rank);
gk_status gk_root_allreduce(void* nccl_comm, int64_t* d_stats, void* stream);
gk_status gk_nccl_shutdown(void);

/* Page-locked host buffers for the *_host entry points (cudaHostAlloc / cudaFreeHost). */
gk_status gk_host_alloc(void** out, size_t bytes);
gk_status gk_host_free(void* ptr);

/* ---- host utilities (no GPU needed) ------------------------------------------------- */
/* Write bandwidth of the host's memory as `threads` CPU threads see it (each streams `repeats` large memsets over its own
 * buffer; the slowest thread sets the clock): the ceiling of every *_host result copy when several GPUs of one box
 * deliver into the same DRAM (bench.py prints it next to the achieved end-to-end rate). */
gk_status gk_measure_host_write_bw(int threads, size_t bytes_per_thread, int repeats, double* gb_per_s);
/* move lists (black first, alternating; position i = moves[starts[i]..starts[i+1])) -> packed boards */
gk_status gk_pack_moves(const int16_t* moves, const int64_t* starts, int n, uint32_t* h_boards);
/* The synthetic "random mid-game" set of BASELINE.json / SURVEY.md section 8(d): position
 * (first + i) has 16 + (h mod 81) stones, black first, alternating, each on a uniformly random
 * empty cell, a stone that would complete five-or-more is redrawn; Philox4x32-10 keyed
 * 0x474F4D4F4B5531.  Writes packed boards and (optionally) the move lists: h_moves must hold
 * 96*n entries, h_starts n+1. */
gk_status gk_synth_positions(int64_t first, int n, uint32_t* h_boards, int16_t* h_moves, int64_t* h_starts);

#ifdef __cplusplus
}
#endif
#endif /* GOMOKU_B200_H_ */

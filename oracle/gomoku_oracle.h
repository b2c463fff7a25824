/* ORACLE -- test infrastructure, not product code.
 *
 * Plain-C CPU restatement of the reference's hot path (Vigilans/GomokuAI):
 * the Aho-Corasick pattern scan + incremental Evaluator, and the random rollout.
 * Every function in gomoku_oracle.c cites the reference file:line it follows.
 *
 * PARITY PINNED: the restatement is checked (tests/test_oracle_*.py) against
 *   (1) the reference's own golden vectors: augmentation list, fail identities, match
 *       KAT and invariant states of core/test/patternsearch_unittest.cpp, the line views
 *       of core/test/boardmap_unittest.cpp, the win/draw sequences of
 *       core/test/integration/board_integrationtest.cpp, committed under tests/golden/;
 *   (2) the reference itself, compiled unmodified into oracle/_ref/libgomoku_ref.so
 *       (oracle/Makefile), element-wise on base/check/fail arrays, emission sequences,
 *       m_scores / totals of replayed positions, and rollout outcomes.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
 * use this library.  gomokuai_b200/ never links, loads or calls it.
 */
#ifndef GOMOKU_ORACLE_H_
#define GOMOKU_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_W = 15, ORC_H = 15, ORC_CELLS = 225, ORC_MAX_PATTERN_LEN = 7, ORC_TARGET_LEN = 13,
       ORC_LINES = 88, ORC_MAX_PATTERNS = 1024 };

typedef struct orc_table orc_table;     /* augmented patterns + double-array trie + fail */
typedef struct orc_evaluator orc_evaluator;

/* ---- automaton ---- */
orc_table* orc_table_default(void);     /* the 41 prototypes of Pattern.cpp:554-596 (cached) */
orc_table* orc_table_build(const char* const* protos, const int* types, const int* scores, int n);
void orc_table_free(orc_table*);
int orc_table_sizes(const orc_table*, int* n_base, int* n_patterns);
int orc_table_arrays(const orc_table*, int32_t* base, int32_t* check, int32_t* fail, int32_t* invariants);
int orc_table_pattern(const orc_table*, int id, char* str8, int* favour, int* type, int* score);
int orc_augment(const char* const* protos, const int* types, const int* scores, int n, int stage,
                char* strs, int* favours, int* otypes, int* oscores, int cap);
int orc_encode(char ch);
int orc_scan(const orc_table*, const uint8_t* codes, int n, int32_t* pids, int32_t* offsets, int cap);
long orc_scan_many(const orc_table*, const uint8_t* codes, const int64_t* starts, int n_strings,
                   int32_t* pids, int32_t* offsets, int32_t* counts, long cap);

/* ---- line views ---- */
int orc_line_view(const int16_t* moves, int n_moves, int pose, int dir, uint8_t* out13);
int orc_line_map(const int16_t* moves, int n_moves, uint8_t* out, int* lens);

/* ---- evaluator (incremental, as the reference) ---- */
int orc_eval_moves(const int16_t* moves, int n_moves, int32_t* scores, uint16_t* pat_totals,
                   uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player);
int orc_eval_batch(const int16_t* moves, const int64_t* starts, int n_pos, int32_t* scores,
                   uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player);
int orc_eval_flags(uint32_t* pattern_flags, uint32_t* compound_flags, int32_t* density);
int orc_eval_apply_revert(const int16_t* moves, int n_moves, int n_revert, int32_t* scores,
                          uint16_t* pat_totals, uint16_t* cmp_totals);
long orc_degenerate_compounds(void);    /* how often Compound::locate produced type < 0 */
long orc_linescan_batch(const int16_t* moves, const int64_t* starts, int n_pos);

/* ---- from-scratch model of the kernel's algorithm ---- */
int orc_eval_scratch(const uint8_t* cells, int lead, int trail, int min_len, int32_t* scores,
                     uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner);
int orc_eval_scratch_batch(const int16_t* moves, const int64_t* starts, int n_pos, int lead, int trail, int min_len,
                           int32_t* scores, uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner);
long orc_scratch_gate_blocks(void);

/* ---- exhaustive single-line checks (see gomoku_oracle.c) ---- */
int orc_line_theorems(int max_len, long* out8);

/* ---- board / rollout ---- */
int orc_board_play(const int16_t* moves, int n_moves, int* out3);
int orc_rollout_injected(const int16_t* moves, int n_moves, const uint8_t* r_stream, int stream_len,
                         int* n_played);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* rollout whose start indices come from Philox4x32-10: move k uses word (k & 3) of
 * philox(ctr = {k >> 2, rollout, position, ctr_hi}, key), r = mulhi32(word, 225). */
int orc_rollout_philox(const int16_t* moves, int n_moves, uint64_t key, uint32_t position,
                       uint32_t rollout, uint32_t ctr_hi, int* n_played);
int orc_rollout_philox_batch(const int16_t* moves, const int64_t* starts, int n_pos, int rollouts_per_pos,
                             uint64_t key, uint32_t ctr_hi, int pos_base, int8_t* winners, uint8_t* lengths,
                             int32_t* wdb);

#ifdef __cplusplus
}
#endif
#endif

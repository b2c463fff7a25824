// gk_format.h -- bit layouts shared by the host table compiler and the sm_100a kernels.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GK_HD __host__ __device__
#else
#define GK_HD
#endif

namespace gk {

constexpr int kWidth = 15, kHeight = 15, kCells = 225, kBoardWords = 16;
constexpr int kLines = 72;          // rows + columns + the 2x21 diagonals of length >= 5
constexpr int kMaxStates = 1024;    // 10-bit state ids
constexpr int kMaxPatterns = 512;   // 9-bit pattern ids

// Symbol index used to address the transition table.  It is (cell value - 1) & 3 with the
// board's 2-bit cell values {0 empty, 1 black, 2 white, 3 off-board pad}:
//   0 = black 'x' (reference code 1), 1 = white 'o' (2), 2 = off-board '?' (3), 3 = empty (4)
constexpr int kSymBlack = 0, kSymWhite = 1, kSymPad = 2, kSymEmpty = 3;
GK_HD inline int sym_from_refcode(int code) { return code - 1; }

// ---- transition word: T[state * 4 + sym] ------------------------------------------------
//   [ 0,10) next state
//   [10,12) number of emissions (0..2)
//   [12,22) emission 0: [12,21) pattern id, bit 21 = "at previous symbol" (emitted on a fail
//           landing BEFORE the current symbol is consumed: end offset = i - 1)
//   [22,32) emission 1, same layout
constexpr uint32_t kNextMask = 0x3ffu;
GK_HD inline uint32_t tw_next(uint32_t w) { return w & kNextMask; }
GK_HD inline uint32_t tw_nemit(uint32_t w) { return (w >> 10) & 3u; }
GK_HD inline uint32_t tw_emit(uint32_t w, int k) { return (w >> (12 + 10 * k)) & 0x3ffu; }
GK_HD inline uint32_t em_pid(uint32_t e) { return e & 0x1ffu; }
GK_HD inline uint32_t em_prev(uint32_t e) { return (e >> 9) & 1u; }

// ======================= device-side encodings (what the kernels read) ==========================
// The host words above are the documented, inspectable form (gk_table_entries).  The copies
// uploaded to the GPU are re-encoded so that the scan loop needs as few integer ops as possible.
//
// ---- device automaton: "cloned arrival" next table -----------------------------------------------
// Every distinct (destination state, emission set) pair of an emitting transition becomes its own
// CLONE of the destination state: same outgoing row, but a low id.  Ids [0, n_clones) are clones,
// ids [n_clones, n_clones + n_states) are the plain states in host order.  Then
//   next16[id * 4 + v] (uint16) = BYTE offset (id' * 8) of the destination row, v = raw 2-bit cell
//                                 value (0 empty, 1 black, 2 white, 3 off-board pad),
//   "this step emitted"  <=>  offset < n_clones * 8   (one compare, no mask, no second load),
// and the state-dependent chain of one scan step is  IADD (offset + v * 2), LDS.U16.
//   erec[clone] (uint32): what an arrival at that clone emits
//     [ 0, 9) pattern id of emission 0      [ 9] emission 0 is "at previous symbol"
//     [10,19) pattern id of emission 1, 0x1ff = none      [19] same flag for emission 1
//     [20,22) compound class of emission 0's pattern (0 none, 1 LiveThree, 2 DeadThree, 3 LiveTwo)
// Emission list entry (uint16, kernel internal): clone id << 6 | scan step.
constexpr uint32_t kDevNoPid = 0x1ffu;
constexpr int kMaxClones = 1024;                  // clone id << 6 | step must fit 16 bits
constexpr int kMaxTapeSteps = 64;
GK_HD inline int sym_to_value(int sym) { return (sym + 1) & 3; }   // table symbol index -> raw cell value
GK_HD inline uint32_t er_pid(uint32_t e, int k) { return (e >> (10 * k)) & 0x1ffu; }
GK_HD inline uint32_t er_prev(uint32_t e, int k) { return (e >> (9 + 10 * k)) & 1u; }
GK_HD inline uint32_t er_cclass0(uint32_t e) { return (e >> 20) & 3u; }

// ---- device pattern record: 2 words per pattern ---------------------------------------------------
//   w0 [ 0,16) up to four scored cells, one nibble each: bits 0..2 = j (cell is the j-th char from
//              the END of the pattern), bit 3 = 1 for '_' (scored for both perspectives + flag),
//              0 for '^' (rival's perspective only)
//      [16,19) number of scored cells (0..4)
//      [19,23) Pattern::Type (0..8, 8 = Five)
//      [23]    favour is black
//      [24,27) length (5..7)
//      [27,29) compound class: 0 none, 1 LiveThree, 2 DeadThree, 3 LiveTwo
//   w1 [ 0,16) score on rows/columns, [16,32) score on diagonals (= int(1.2 * score))
struct PatRec { uint32_t w0, w1; };
GK_HD inline uint32_t pr_ncells(uint32_t w0) { return (w0 >> 16) & 7u; }
GK_HD inline uint32_t pr_type(uint32_t w0) { return (w0 >> 19) & 15u; }
GK_HD inline uint32_t pr_black(uint32_t w0) { return (w0 >> 23) & 1u; }
GK_HD inline uint32_t pr_len(uint32_t w0) { return (w0 >> 24) & 7u; }
GK_HD inline uint32_t pr_cclass(uint32_t w0) { return (w0 >> 27) & 3u; }
constexpr int kTypeFive = 8;

// ---- scan tape ---------------------------------------------------------------------------------------
// One warp evaluates one board; lane L walks a fixed chain of whole lines, one symbol per step.
// Every line is followed by trail_pad (>= 2) pad symbols; two pads take ANY state to the state
// "one '?' read from the root" (checked by the table compiler), which is also where a line has
// to start, so the chain needs no explicit restart between lines.
//   src[step * 32 + lane]  (uint16)  where this step's symbol lives in the warp's copy of the board
//        (lane w of the warp holds board word w; words 14 / 15 carry all-ones above cell 224):
//        [0,5)  board word index cell >> 4; pads read cell 225, which the kernel forces to 3
//        [8,13) rotate-right amount that brings the cell's 2 bits to bits 1..2: (2*(cell & 15) + 31) & 31
//   info[step * 32 + lane] (uint16)  only read when the step emitted:
//        [0,9)  virtual cell of this step on its line (cell0 + index * stride; runs past 224 on pads)
//        [9,11) direction: 0 row, 1 column, 2 diagonal (+1,+1), 3 anti-diagonal (-1,+1)
//        [11,16) cell stride of that direction (1, 15, 16, 14)
constexpr int kPadCell = 225;                     // any cell index in [225, 272) reads as pad
GK_HD inline int dir_stride(int dir) { return dir == 0 ? 1 : dir == 1 ? 15 : dir == 2 ? 16 : 14; }

}  // namespace gk

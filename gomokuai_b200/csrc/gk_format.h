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

// ---- pattern record: 2 words per pattern --------------------------------------------------
//   w0 [ 0,14) cell kinds, 2 bits per pattern char counted FROM THE END (j = 0 is the last
//              char): 0 = not scored, 1 = '_' (scored for both perspectives + flag),
//              2 = '^' (scored for the rival's perspective only)
//      [14,18) Pattern::Type (0..8, 8 = Five)
//      [18]    favour is black
//      [19,22) length (5..7)
//      [22,24) compound class: 0 none, 1 LiveThree, 2 DeadThree, 3 LiveTwo
//   w1 [ 0,16) score on rows/columns, [16,32) score on diagonals (= int(1.2 * score))
struct PatRec { uint32_t w0, w1; };
GK_HD inline uint32_t pr_kinds(uint32_t w0) { return w0 & 0x3fffu; }
GK_HD inline uint32_t pr_type(uint32_t w0) { return (w0 >> 14) & 15u; }
GK_HD inline uint32_t pr_black(uint32_t w0) { return (w0 >> 18) & 1u; }
GK_HD inline uint32_t pr_len(uint32_t w0) { return (w0 >> 19) & 7u; }
GK_HD inline uint32_t pr_cclass(uint32_t w0) { return (w0 >> 22) & 3u; }
constexpr int kTypeFive = 8;

// ---- scan tape: tape[step * 32 + lane] ------------------------------------------------------
// One warp evaluates one board; lane L walks a fixed chain of whole lines, one symbol per step.
//   [ 0, 9) source cell of this step's symbol: 0..224 = board cell, >= 225 = a pad cell (value 3)
//   [ 9,18) virtual cell of this step on its line (cell0 + index * stride, may run past 224 on
//           the trailing pads); an emission ending here covers cells vcell - j * stride
//   [18,20) direction of the line: 0 row, 1 column, 2 diagonal (+1,+1), 3 anti-diagonal (-1,+1)
//   [20]    first symbol of a line: the automaton restarts from the state reached after one
//           leading '?'
//   [21,26) cell stride of the line (1, 15, 16 or 14)
constexpr uint32_t kTapeStart = 1u << 20;
constexpr int kPadCell = 225;                     // any cell index in [225, 272) reads as pad
GK_HD inline uint32_t tp_src(uint32_t e) { return e & 0x1ffu; }
GK_HD inline uint32_t tp_vcell(uint32_t e) { return (e >> 9) & 0x1ffu; }
GK_HD inline uint32_t tp_dir(uint32_t e) { return (e >> 18) & 3u; }
GK_HD inline int dir_stride(int dir) { return dir == 0 ? 1 : dir == 1 ? 15 : dir == 2 ? 16 : 14; }

// ---- emission queue entry (kernel internal) ---------------------------------------------------
//   [0,9) pattern id, [9,18) virtual END cell, [18,20) direction
}  // namespace gk

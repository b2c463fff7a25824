// ORACLE SCAFFOLDING -- test infrastructure, not product code.
//
// Thin extern "C" harness over the reference's OWN object code.  The recipe in
// oracle/Makefile compiles /root/reference/core/lib/src/{Game,Mapping,Pattern}.cpp
// and utils/ACAutomata.cpp unmodified, where they lie, against oracle/eigen_shim and
// links them with this file into oracle/_ref/libgomoku_ref.so.  Everything the
// harness computes is computed by the reference's Board / BoardMap / PatternSearch /
// Evaluator; the only logic restated here is
//   * the 4-line loop of Default::RandomRollout (algorithms/MonteCarlo.hpp:37-47),
//   * the 4-line probe loop of Board::getRandomMove (Game.cpp:68-71) and
//   * the plane-filling lambda of Board.encoded_states (core/py_ext/src/game_ext.hpp:87-104; the
//     pybind module itself needs pybind11/eigen.h, i.e. real Eigen),
// because MonteCarlo.hpp needs float Eigen algebra the shim does not provide and the
// reference's RNG (`static mt19937 rnd_eng`, Game.cpp:11-12) cannot be seeded or fed
// from outside.  Private members are reached exactly the way the reference's own unit
// test does it (core/test/patternsearch_unittest.cpp:3-6).
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may
// load this library.
#include <Eigen/Dense>   // the shim, before the access hack below reaches its standard headers
#include <string>
#include <cstdint>
#include <cstring>
#define private public
#include "Pattern.h"
#include "utils/ACAutomata.h"
#undef private

using namespace Gomoku;

namespace {

PatternSearch g_custom;      // automaton built from caller-supplied prototypes
bool g_custom_ok = false;

PatternSearch& pick(int which) { return which == 0 ? Evaluator::Patterns : g_custom; }

Evaluator& evaluator() {     // one process-wide evaluator, reset per position
    static Evaluator ev;
    return ev;
}

void dump_eval(Evaluator& ev, int32_t* scores, uint16_t* pat_totals, uint16_t* cmp_totals,
               int8_t* winner, int8_t* cur_player) {
    if (scores) {
        for (int g = 0; g < 4; ++g)
            for (int i = 0; i < BOARD_SIZE; ++i) scores[g * BOARD_SIZE + i] = ev.m_scores[g][i];
    }
    // totals layout handed to the caller: [player: 0 = White, 1 = Black][type]
    // (Evaluator::Group(Player), Pattern.h:154-156; Record::get(Player), Pattern.cpp:413-416)
    if (pat_totals) {
        for (int t = 0; t < Pattern::Size - 1; ++t) {
            pat_totals[0 * 8 + t] = (uint16_t)ev.m_patternDist.back()[t].get(Player::White);
            pat_totals[1 * 8 + t] = (uint16_t)ev.m_patternDist.back()[t].get(Player::Black);
        }
    }
    if (cmp_totals) {
        for (int t = 0; t < Compound::Size; ++t) {
            cmp_totals[0 * 3 + t] = (uint16_t)ev.m_compoundDist.back()[t].get(Player::White);
            cmp_totals[1 * 3 + t] = (uint16_t)ev.m_compoundDist.back()[t].get(Player::Black);
        }
    }
    if (winner) *winner = (int8_t)ev.board().m_winner;
    if (cur_player) *cur_player = (int8_t)ev.board().m_curPlayer;
}

}  // namespace

extern "C" {

// ---- automaton introspection (golden values of SURVEY Appendix B) -------------------

int ref_table_sizes(int which, int* n_base, int* n_patterns) {
    auto& ps = pick(which);
    *n_base = (int)ps.m_base.size();
    *n_patterns = (int)ps.m_patterns.size();
    return 0;
}

int ref_table_arrays(int which, int32_t* base, int32_t* check, int32_t* fail, int32_t* invariants) {
    auto& ps = pick(which);
    for (size_t i = 0; i < ps.m_base.size(); ++i) {
        base[i] = ps.m_base[i];
        check[i] = ps.m_check[i];
        fail[i] = ps.m_fail[i];
    }
    for (int i = 0; i < 5; ++i) invariants[i] = ps.m_invariants[i];
    return 0;
}

// str must hold >= 8 bytes; favour: +1 black, -1 white.
int ref_table_pattern(int which, int id, char* str, int* favour, int* type, int* score) {
    auto& ps = pick(which);
    if (id < 0 || id >= (int)ps.m_patterns.size()) return -1;
    const Pattern& p = ps.m_patterns[id];
    std::memset(str, 0, 8);
    std::memcpy(str, p.str.data(), p.str.size());
    *favour = (int)p.favour;
    *type = (int)p.type;
    *score = p.score;
    return 0;
}

// Build the secondary automaton from caller prototypes ("+..."/"-..." strings as in
// Pattern.cpp:554-596).  Used for the 3-prototype KAT of patternsearch_unittest.cpp.
int ref_table_build_custom(const char* const* protos, const int* types, const int* scores, int n) {
    // PatternSearch only has an initializer_list constructor (Pattern.cpp:27-30); the
    // builder's vector is filled directly, exactly what that constructor ends up doing.
    AhoCorasickBuilder builder({});
    for (int i = 0; i < n; ++i) builder.m_patterns.emplace_back(protos[i], (Pattern::Type)types[i], scores[i]);
    g_custom = PatternSearch();
    builder.build(&g_custom);
    g_custom_ok = true;
    return 0;
}

// Augmentation stages only (for the golden list at patternsearch_unittest.cpp:40-73).
// stage: 1 = reverse, 2 = +flip, 3 = +boundary, 4 = +sort.  Returns pattern count; writes
// up to cap records of 8 chars (str) + favour/type/score.
int ref_augment(const char* const* protos, const int* types, const int* scores, int n, int stage,
                char* strs, int* favours, int* otypes, int* oscores, int cap) {
    AhoCorasickBuilder builder({});
    for (int i = 0; i < n; ++i) builder.m_patterns.emplace_back(protos[i], (Pattern::Type)types[i], scores[i]);
    if (stage >= 1) builder.reverseAugment();
    if (stage >= 2) builder.flipAugment();
    if (stage >= 3) builder.boundaryAugment();
    if (stage >= 4) builder.sortPatterns();
    int m = (int)builder.m_patterns.size();
    for (int i = 0; i < m && i < cap; ++i) {
        const Pattern& p = builder.m_patterns[i];
        std::memset(strs + 8 * i, 0, 8);
        std::memcpy(strs + 8 * i, p.str.data(), p.str.size());
        favours[i] = (int)p.favour;
        otypes[i] = (int)p.type;
        oscores[i] = p.score;
    }
    return m;
}

// ---- the scan itself: PatternSearch::execute over a code string ---------------------

// codes: symbols 1..4 (EncodeCharset, Mapping.h:40-48).  Returns the number of emissions
// (may exceed cap; only the first cap are stored).
int ref_scan(int which, const uint8_t* codes, int n, int32_t* pids, int32_t* offsets, int cap) {
    auto& ps = pick(which);
    std::string target((const char*)codes, (size_t)n);
    int count = 0;
    for (auto entry : ps.execute(target)) {
        const Pattern& p = std::get<0>(entry);
        if (count < cap) {
            pids[count] = (int32_t)(&p - ps.m_patterns.data());
            offsets[count] = std::get<1>(entry);
        }
        ++count;
    }
    return count;
}

// Many strings in one call (for the million-string transducer gate): strings are
// concatenated in `codes`, string i spans [starts[i], starts[i+1]).  Emissions are
// appended to pids/offsets; counts[i] receives the per-string count.
long ref_scan_many(int which, const uint8_t* codes, const int64_t* starts, int n_strings,
                   int32_t* pids, int32_t* offsets, int32_t* counts, long cap) {
    auto& ps = pick(which);
    long total = 0;
    for (int s = 0; s < n_strings; ++s) {
        std::string target((const char*)codes + starts[s], (size_t)(starts[s + 1] - starts[s]));
        int c = 0;
        for (auto entry : ps.execute(target)) {
            const Pattern& p = std::get<0>(entry);
            if (total < cap) {
                pids[total] = (int32_t)(&p - ps.m_patterns.data());
                offsets[total] = std::get<1>(entry);
            }
            ++total;
            ++c;
        }
        counts[s] = c;
    }
    return total;
}

// ---- BoardMap line views (golden windows of boardmap_unittest.cpp) ------------------

// Replays `moves` on a fresh BoardMap, then copies the 13-symbol window of
// lineView(pose, dir) (Mapping.cpp:31-34) into out13.
int ref_line_view(const int16_t* moves, int n_moves, int pose, int dir, uint8_t* out13) {
    BoardMap bm;
    for (int i = 0; i < n_moves; ++i) bm.applyMove(Position(moves[i]));
    auto v = bm.lineView(Position(pose), (Direction)dir);
    std::memcpy(out13, v.data(), TARGET_LEN);
    return 0;
}

// All 88 padded line strings of the position (Mapping.h:70), concatenated; lens[88].
int ref_line_map(const int16_t* moves, int n_moves, uint8_t* out, int* lens) {
    BoardMap bm;
    for (int i = 0; i < n_moves; ++i) bm.applyMove(Position(moves[i]));
    int k = 0;
    for (size_t l = 0; l < bm.m_lineMap.size(); ++l) {
        lens[l] = (int)bm.m_lineMap[l].size();
        std::memcpy(out + k, bm.m_lineMap[l].data(), bm.m_lineMap[l].size());
        k += lens[l];
    }
    return k;
}

// ---- Evaluator replay (the parity target of config 2) -------------------------------

// Replays a move list through Evaluator::applyMove (Pattern.cpp:310-335) and dumps
// m_scores[4][225], pattern totals [2][8], compound totals [2][3], winner, side to move.
// Returns 0, or 1 if the reference's always-on self-check threw (Pattern.cpp:314-333).
int ref_eval_moves(const int16_t* moves, int n_moves, int32_t* scores, uint16_t* pat_totals,
                   uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player) {
    Evaluator& ev = evaluator();
    ev.reset();
    try {
        for (int i = 0; i < n_moves; ++i) ev.applyMove(Position(moves[i]));
    } catch (...) {
        ev.reset();
        return 1;
    }
    dump_eval(ev, scores, pat_totals, cmp_totals, winner, cur_player);
    return 0;
}

// Batch form: position p spans moves[starts[p] .. starts[p+1]).  Any output pointer may
// be null (timing runs pass null for scores to measure the evaluator alone).
// Returns the number of positions whose self-check threw.
int ref_eval_batch(const int16_t* moves, const int64_t* starts, int n_pos, int32_t* scores,
                   uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player) {
    int bad = 0;
    for (int p = 0; p < n_pos; ++p) {
        bad += ref_eval_moves(moves + starts[p], (int)(starts[p + 1] - starts[p]),
                              scores ? scores + (size_t)p * 4 * BOARD_SIZE : nullptr,
                              pat_totals ? pat_totals + (size_t)p * 16 : nullptr,
                              cmp_totals ? cmp_totals + (size_t)p * 6 : nullptr,
                              winner ? winner + p : nullptr, cur_player ? cur_player + p : nullptr);
    }
    return bad;
}

// Per-cell flag words of the current evaluator state after ref_eval_moves (debug aid:
// m_patternDist[cell][type], m_compoundDist[cell][type]; Pattern.h:216-217).
int ref_eval_flags(uint32_t* pattern_flags /*225*8*/, uint32_t* compound_flags /*225*3*/, int32_t* density /*2*2*225*/) {
    Evaluator& ev = evaluator();
    for (int c = 0; c < BOARD_SIZE; ++c) {
        for (int t = 0; t < 8; ++t) pattern_flags[c * 8 + t] = ev.m_patternDist[c][t].field;
        for (int t = 0; t < 3; ++t) compound_flags[c * 3 + t] = ev.m_compoundDist[c][t].field;
    }
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < BOARD_SIZE; ++c) density[(a * 2 + b) * BOARD_SIZE + c] = ev.m_density[a][b][c];
    return 0;
}

// Full-line scan of the 88 padded lines of each position (the reference's scan, the
// reference's line strings); used for the lines/s CPU figure.  Returns total emissions.
long ref_linescan_batch(const int16_t* moves, const int64_t* starts, int n_pos) {
    long total = 0;
    BoardMap bm;
    for (int p = 0; p < n_pos; ++p) {
        bm.reset();
        for (int64_t i = starts[p]; i < starts[p + 1]; ++i) bm.applyMove(Position(moves[i]));
        for (auto& line : bm.m_lineMap)
            for (auto entry : Evaluator::Patterns.execute(line)) { (void)entry; ++total; }
    }
    return total;
}

// ---- Board: win / draw / rollout -----------------------------------------------------

// Replays moves through Board::applyMove (Game.cpp:37-47) with victory checks.
// out[0] = cur_player, out[1] = winner, out[2] = moves actually applied.
int ref_board_play(const int16_t* moves, int n_moves, int* out) {
    Board b;
    int applied = 0;
    for (int i = 0; i < n_moves; ++i) {
        Player before = b.m_curPlayer;
        Player after = b.applyMove(Position(moves[i]));
        if (after != before) ++applied;
    }
    out[0] = (int)b.m_curPlayer;
    out[1] = (int)b.m_winner;
    out[2] = applied;
    return 0;
}

// One rollout driven by an injected start-index stream.  Loop = Default::RandomRollout
// (MonteCarlo.hpp:39-41); move choice = the probe loop of Board::getRandomMove
// (Game.cpp:68-71) with `r_stream[k]` standing in for rnd(rnd_eng); everything else is
// the reference's Board::applyMove / checkGameEnd.  The board is set up by replaying
// `moves` WITHOUT victory checks (Policy::applyMove semantics, MCTS.cpp:48-50) followed by
// one checkGameEnd (MCTS.cpp:166), i.e. exactly how MCTS::playout reaches simulate().
// Returns the winner (-1/0/+1); *n_played = rollout length; -2 if the stream ran out.
int ref_rollout_injected(const int16_t* moves, int n_moves, const uint8_t* r_stream, int stream_len,
                         int* n_played) {
    Board b;
    for (int i = 0; i < n_moves; ++i) b.applyMove(Position(moves[i]), false);
    b.checkGameEnd();
    int total = 0;
    for (auto result = b.m_curPlayer; result != Player::None; ++total) {
        if (total >= stream_len) { *n_played = total; return -2; }
        int id = r_stream[total];
        while (!b.moveState(Player::None, id)) id = (id + 1) % (int)b.moveStates(Player::None).size();
        result = b.applyMove(Position(id));
    }
    *n_played = total;
    return (int)b.m_winner;
}

// Free-running rollouts with the reference's own RNG path: Board::getRandomMove()
// (global mt19937 seeded from random_device, Game.cpp:11-12,64-73).  wdb[0..2] +=
// {white wins, draws, black wins}; *total_moves += moves played.  Board restored with
// revertMove as RandomPolicy::averagedSimulate does (Random.h:27-32).
int ref_rollout_free(const int16_t* moves, int n_moves, int n_rollouts, int64_t* wdb, int64_t* total_moves) {
    Board b;
    for (int i = 0; i < n_moves; ++i) b.applyMove(Position(moves[i]), false);
    b.checkGameEnd();
    for (int k = 0; k < n_rollouts; ++k) {
        int total = 0;
        for (auto result = b.m_curPlayer; result != Player::None; ++total) {
            result = b.applyMove(b.getRandomMove());
        }
        wdb[(int)b.m_winner + 1] += 1;
        *total_moves += total;
        b.revertMove((size_t)total);
    }
    return 0;
}

// Board.encoded_states() (core/py_ext/src/game_ext.hpp:87-104) over the reference's Board: planes
// [stones of the side to move, opponent's, empty, last move, second-to-last move, black to move].
// Moves are replayed without victory checks so that any position can be encoded.
int ref_encoded_states(const int16_t* moves, int n_moves, uint8_t* out /*6*225*/) {
    Board b;
    for (int i = 0; i < n_moves; ++i) b.applyMove(Position(moves[i]), false);
    int index = 0;
    for (auto player : { b.m_curPlayer, -b.m_curPlayer, Player::None }) {
        for (int c = 0; c < BOARD_SIZE; ++c) out[index * BOARD_SIZE + c] = b.moveState(player, c) ? 1 : 0;
        ++index;
    }
    for (int i = 0; i <= 1; ++index, ++i) {
        std::memset(out + index * BOARD_SIZE, 0, BOARD_SIZE);
        if ((int)b.m_moveRecord.size() > i) {
            auto position = *(b.m_moveRecord.rbegin() + i);
            out[index * BOARD_SIZE + position.y() * WIDTH + position.x()] = 1;
        }
    }
    std::memset(out + index * BOARD_SIZE, b.m_curPlayer == Player::Black, BOARD_SIZE);
    return 0;
}

}  // extern "C"

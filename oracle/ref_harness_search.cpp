// ORACLE SCAFFOLDING -- test infrastructure, not product code.
//
// Second half of the extern "C" harness over the reference's OWN object code: everything ABOVE the
// evaluator.  oracle/Makefile compiles /root/reference/core/lib/src/MCTS.cpp unmodified, where it
// lies, and this file includes the reference's header-only layers unmodified:
//   include/algorithms/{MonteCarlo,Statistical,Heuristic}.hpp, include/policies/{Random,Traditional,PoolRAVE}.h
// against oracle/eigen_shim (whose header states which float semantics of Eigen it keeps).
// Everything computed here is computed by the reference's Heuristic / Policy / MCTS code; the only
// logic of this file's own is
//   * Philox4x32-10 (Salmon et al., SC'11; the public counter-based generator the product kernels
//     document in include/gomoku_b200.h) to feed the reference Board the SAME start-index stream
//     the rollout kernel draws, through the probe loop of Board::getRandomMove (Game.cpp:68-71),
//   * two injected `simulate` slots (Policy's own plugin mechanism, MCTS.h:86-92) that make a
//     search deterministic so that trees can be compared node for node, and
//   * a pre-order dump of the reference's tree.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load this.
#include <Eigen/Dense>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <string_view>
#define private public
#include "MCTS.h"
#include "Pattern.h"
#include "algorithms/Heuristic.hpp"
#include "algorithms/MonteCarlo.hpp"
#include "policies/PoolRAVE.h"
#include "policies/Random.h"
#include "policies/Traditional.h"
#undef private

using namespace Gomoku;
using Gomoku::Algorithms::Default;
using Gomoku::Algorithms::Heuristic;
using Gomoku::Algorithms::Stats;
using Gomoku::Policies::PoolRAVEPolicy;
using Gomoku::Policies::RandomPolicy;
using Gomoku::Policies::TraditionalPolicy;

namespace {

Evaluator& evaluator() {
    static Evaluator ev;
    return ev;
}

bool replay(Evaluator& ev, const int16_t* moves, int n) {
    ev.reset();
    try {
        for (int i = 0; i < n; ++i) ev.applyMove(Position(moves[i]));
    } catch (...) {
        ev.reset();
        return false;
    }
    return true;
}

void replay(Board& b, const int16_t* moves, int n) {       // as MCTS::playout walks: no victory checks, one at the end
    for (int i = 0; i < n; ++i) b.applyMove(Position(moves[i]), false);
    b.checkGameEnd();
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------
void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1), n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One rollout on the reference Board under the kernel's stream: word k of the rollout is word k & 3 of
// Philox(counter = (k / 4, rollout, position, ctr_hi), key); start index r = mulhi32(word, 225); move = first
// empty cell at or after r, cyclically (Game.cpp:68-71).  The loop is Default::RandomRollout (MonteCarlo.hpp:39-41).
Player philox_rollout(Board& b, uint64_t key, uint32_t ctr_hi, uint32_t position, uint32_t rollout, int* n_played) {
    const uint32_t k2[2] = { uint32_t(key), uint32_t(key >> 32) };
    uint32_t words[4];
    int total = 0;
    for (auto result = b.m_curPlayer; result != Player::None; ++total) {
        if ((total & 3) == 0) {
            const uint32_t ctr[4] = { uint32_t(total >> 2), rollout, position, ctr_hi };
            philox(ctr, k2, words);
        }
        int id = int((uint64_t(words[total & 3]) * 225u) >> 32);
        while (!b.moveState(Player::None, id)) id = (id + 1) % (int)b.moveStates(Player::None).size();
        result = b.applyMove(Position(id));
    }
    *n_played = total;
    return b.m_winner;
}

// ---- injected simulate slots ---------------------------------------------------------------------------
// kind 0: value = a fixed function of the position in {-1, -0.8, ..., 1}, uniform priors (Default::UniformProbs).
//         The same function is written in Python in tests/ (tests/search_util.py::hash_value).
float hash_value(const Board& b) {
    uint32_t h = 2166136261u;
    for (int c = 0; c < BOARD_SIZE; ++c) {
        const uint32_t s = b.moveState(Player::Black, c) ? 1u : b.moveState(Player::White, c) ? 2u : 0u;
        h = (h ^ s) * 16777619u;
    }
    return float(int(h % 11u) - 5) / 5.0f;
}

struct InjectedPolicy : public Policy {
    int kind; uint64_t key; uint32_t tree; int c_rollouts; uint32_t playout = 0;
    InjectedPolicy(int kind, uint64_t key, uint32_t tree, int c_rollouts, double c_puct)
        : Policy(nullptr, nullptr, [this](Board& board) { return run(board); }, nullptr, c_puct),
          kind(kind), key(key), tree(tree), c_rollouts(c_rollouts) {}
    EvalResult run(Board& board) {
        if (kind == 0) return { hash_value(board), Default::UniformProbs(board) };
        // kind 1: RandomPolicy::averagedSimulate (Random.h:22-35) with the rollouts driven by the stream the product's
        // root-parallel driver uses for (tree, playout): ctr_hi = playout index, position = tree index.
        auto init_player = board.m_curPlayer;
        double score = 0;
        for (int i = 0; i < c_rollouts; ++i) {
            int total = 0;
            const Player winner = philox_rollout(board, key, playout, tree, uint32_t(i), &total);
            score += CalcScore(init_player, winner);
            board.revertMove(total);
        }
        score /= c_rollouts;
        return { float(score), Default::UniformProbs(board) };
    }
};

struct Dump {
    int16_t* pos; int32_t* visits; float* value; float* prior; int16_t* depth; int32_t* n_children; long cap; long n = 0, visited = 0;
    void walk(const Node* node, int d) {
        if (n < cap) {
            pos[n] = node->position.id; visits[n] = (int32_t)node->node_visits; value[n] = node->state_value;
            prior[n] = node->action_prob; depth[n] = (int16_t)d; n_children[n] = (int32_t)node->children.size();
        }
        ++n;
        for (auto& ch : node->children)
            if (ch->node_visits > 0) walk(ch.get(), d + 1);      // never-visited children carry no information beyond their count
    }
};

std::shared_ptr<Policy> make_policy(int policy_kind, double c_puct, int c_rollouts, double c_bias) {
    switch (policy_kind) {
    case 0: return std::make_shared<RandomPolicy>(c_puct, size_t(c_rollouts));
    case 1: return std::make_shared<TraditionalPolicy>(c_puct);
    case 2: return std::make_shared<PoolRAVEPolicy>(c_puct, c_bias);
    default: return nullptr;
    }
}

}  // namespace

extern "C" {

// ---- Heuristic.hpp -------------------------------------------------------------------------------------

// Heuristic::EvaluationProbs / EvaluationValue / DensityWeight (Heuristic.hpp:16-45) for the side to move after
// replaying `moves` through the reference Evaluator.  dw = DensityWeight of [side to move, opponent], 2 x 225.
// Returns 0; 1 if the evaluator's self-check threw; 2 if the position is terminal (heads undefined: m_curPlayer = None).
int ref_heads(const int16_t* moves, int n_moves, float* probs, float* value, float* dw) {
    Evaluator& ev = evaluator();
    if (!replay(ev, moves, n_moves)) return 1;
    const Player cur = ev.board().m_curPlayer;
    if (cur == Player::None) return 2;
    auto p = Heuristic::EvaluationProbs(ev, cur);
    for (int i = 0; i < BOARD_SIZE; ++i) probs[i] = p[i];
    if (value) *value = Heuristic::EvaluationValue(ev, cur);
    if (dw) {
        auto a = Heuristic::DensityWeight(ev, cur), b = Heuristic::DensityWeight(ev, -cur);
        for (int i = 0; i < BOARD_SIZE; ++i) { dw[i] = a[i]; dw[BOARD_SIZE + i] = b[i]; }
    }
    return 0;
}

// Heuristic::DecisiveFilter (Heuristic.hpp:93-161) applied in place to caller-supplied probabilities.
int ref_decisive_filter(const int16_t* moves, int n_moves, float* probs) {
    Evaluator& ev = evaluator();
    if (!replay(ev, moves, n_moves)) return 1;
    if (ev.board().m_curPlayer == Player::None) return 2;
    Eigen::VectorXf p(BOARD_SIZE);
    for (int i = 0; i < BOARD_SIZE; ++i) p[i] = probs[i];
    Heuristic::DecisiveFilter(ev, p);
    for (int i = 0; i < BOARD_SIZE; ++i) probs[i] = p[i];
    return 0;
}

// TraditionalPolicy::hybridSimulate (Traditional.h:49-69) reached the way MCTS reaches it: prepare(board) synchronises
// the policy's evaluator with the board (Traditional.h:27-31), simulate(board) is the slot MCTS::playout calls.
int ref_hybrid_simulate(const int16_t* moves, int n_moves, float* probs, float* value) {
    static TraditionalPolicy policy;
    Board board;
    replay(board, moves, n_moves);
    if (board.m_curPlayer == Player::None) return 2;
    try {
        policy.prepare(board);
    } catch (...) {
        policy.m_evaluator.reset();
        return 1;
    }
    auto [v, p] = policy.simulate(board);
    *value = v;
    for (int i = 0; i < BOARD_SIZE; ++i) probs[i] = p[i];
    return 0;
}

// Heuristic::MaxEvaluatedRollout (Heuristic.hpp:61-85): returns the winner; played[] = the moves of the rollout.
int ref_guided_rollout_max(const int16_t* moves, int n_moves, int16_t* played, int cap, int* n_played) {
    Evaluator& ev = evaluator();
    if (!replay(ev, moves, n_moves)) return -3;
    int total = 0;
    try {
        auto [winner, count] = Heuristic::MaxEvaluatedRollout(ev, false);
        total = count;
    } catch (...) {
        ev.reset();
        return -3;
    }
    const auto& rec = ev.board().m_moveRecord;
    for (int i = 0; i < total && i < cap; ++i) played[i] = rec[n_moves + i].id;
    *n_played = total;
    return (int)ev.board().m_winner;
}

// First move of Heuristic::RandomEvaluatedRollout (Heuristic.hpp:88-91): Board::getRandomMove(probs) (Game.cpp:75-78,
// std::discrete_distribution over the reference's own global mt19937) drawn n_draws times from EvaluationProbs.
int ref_sampled_first_move(const int16_t* moves, int n_moves, int n_draws, int32_t* counts /*225*/) {
    Evaluator& ev = evaluator();
    if (!replay(ev, moves, n_moves)) return 1;
    const Player cur = ev.board().m_curPlayer;
    if (cur == Player::None) return 2;
    auto p = Heuristic::EvaluationProbs(ev, cur);
    std::memset(counts, 0, sizeof(int32_t) * BOARD_SIZE);
    for (int k = 0; k < n_draws; ++k) counts[ev.board().getRandomMove(p).id] += 1;
    return 0;
}

// n_games x Heuristic::RandomEvaluatedRollout from the position; wdb += {white wins, draws, black wins}.
int ref_guided_rollout_sampled(const int16_t* moves, int n_moves, int n_games, int64_t* wdb, int64_t* total_moves) {
    Evaluator& ev = evaluator();
    if (!replay(ev, moves, n_moves)) return 1;
    try {
        for (int g = 0; g < n_games; ++g) {
            auto [winner, count] = Heuristic::RandomEvaluatedRollout(ev, true);
            wdb[(int)winner + 1] += 1;
            *total_moves += count;
        }
    } catch (...) {
        ev.reset();
        return 1;
    }
    return 0;
}

// ---- rollouts under the kernel's Philox stream -------------------------------------------------------------

// wdb[3] += outcomes of `rollouts` playouts of position index `position` (reference Board, stream as documented above).
int ref_rollout_philox(const int16_t* moves, int n_moves, int rollouts, uint64_t key, uint32_t ctr_hi, uint32_t position,
                       int32_t* wdb, int64_t* total_moves) {
    Board b;
    replay(b, moves, n_moves);
    for (int k = 0; k < rollouts; ++k) {
        int total = 0;
        const Player w = philox_rollout(b, key, ctr_hi, position, uint32_t(k), &total);
        wdb[(int)w + 1] += 1;
        if (total_moves) *total_moves += total;
        b.revertMove((size_t)total);
    }
    return 0;
}

// ---- MCTS (MCTS.cpp, unmodified) ------------------------------------------------------------------------------

void ref_seed_noise(uint32_t seed) { Stats::RandomEngine().seed(seed); }       // Dirichlet noise engine (Statistical.hpp:23-26)

// A search with an injected deterministic `simulate` (sim_kind 0 / 1 above): the reference's MCTS::runPlayouts over
// `n_searches` consecutive moves -- search `iterations` playouts, step to the most visited child, play it on the board
// (MCTS::getAction, MCTS.cpp:99-102) -- then ONE more search whose tree is dumped in pre-order (visited nodes only).
// actions[n_searches] receives the moves chosen.  Returns the number of visited nodes of the final tree (may exceed cap).
long ref_mcts_injected(const int16_t* moves, int n_moves, int iterations, int n_searches, int sim_kind, uint64_t key,
                       uint32_t tree, int c_rollouts, double c_puct, uint32_t noise_seed, int16_t* actions,
                       int16_t* pos, int32_t* visits, float* value, float* prior, int16_t* depth, int32_t* n_children, long cap) {
    Board board;
    for (int i = 0; i < n_moves; ++i) board.applyMove(Position(moves[i]));
    auto policy = std::make_shared<InjectedPolicy>(sim_kind, key, tree, c_rollouts, c_puct);
    MCTS mcts(size_t(iterations), n_moves ? Position(moves[n_moves - 1]) : Position(-1),
              n_moves ? -board.m_curPlayer : Player::White, policy);
    Stats::RandomEngine().seed(noise_seed);
    // The stream's playout counter restarts with every search, like the driver's round index, and must advance once
    // per PLAYOUT -- also for playouts that end in a terminal leaf (no simulate call): wrap backPropogate, which
    // MCTS::playout calls exactly once per playout (MCTS.cpp:174).
    auto backprop = policy->backPropogate;
    policy->backPropogate = [policy_raw = policy.get(), backprop](Node* node, Board& b, double v) {
        backprop(node, b, v);
        policy_raw->playout += 1;
    };
    for (int s = 0; s < n_searches; ++s) {
        policy->playout = 0;
        const Position a = mcts.getAction(board);
        actions[s] = a.id;
        board.applyMove(a);
        if (board.m_curPlayer == Player::None) return -1;
    }
    policy->playout = 0;
    mcts.runPlayouts(board);
    Dump d{ pos, visits, value, prior, depth, n_children, cap };
    d.walk(mcts.m_root.get(), 0);
    return d.n;
}

// The reference's own search, its own policies and RNG (config 1 of BASELINE.json when policy_kind = 0, iterations =
// 10000, empty board): policy_kind 0 RandomPolicy(c_puct, c_rollouts), 1 TraditionalPolicy(c_puct), 2 PoolRAVEPolicy(c_puct, c_bias).
// Returns the chosen move; *seconds = wall clock of getAction; visits[225] = root child visits before the step.
int ref_mcts_get_action(const int16_t* moves, int n_moves, int iterations, int policy_kind, double c_puct, int c_rollouts,
                        double c_bias, double* seconds, int64_t* tree_size, int32_t* root_visits /*225, nullable*/) {
    Board board;
    for (int i = 0; i < n_moves; ++i) board.applyMove(Position(moves[i]));
    MCTS mcts(size_t(iterations), n_moves ? Position(moves[n_moves - 1]) : Position(-1),
              n_moves ? -board.m_curPlayer : Player::White, make_policy(policy_kind, c_puct, c_rollouts, c_bias));
    const auto t0 = std::chrono::steady_clock::now();
    mcts.runPlayouts(board);
    if (root_visits) {
        std::memset(root_visits, 0, sizeof(int32_t) * BOARD_SIZE);
        for (auto& ch : mcts.m_root->children) root_visits[ch->position.id] = (int32_t)ch->node_visits;
    }
    const Position a = mcts.stepForward()->position;
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (tree_size) *tree_size = (int64_t)mcts.m_size;
    return a.id;
}

// MCTS::evalState (MCTS.cpp:104-117) post-processing alone: visits -> normalized -> +1 -> TempBasedProbs
// (Statistical.hpp:37-42).  (evalState itself prints the visit matrix to stdout.)
int ref_temp_based_probs(const float* child_visits, int n_moves_played, float* out) {
    Eigen::VectorXf v(BOARD_SIZE);
    for (int i = 0; i < BOARD_SIZE; ++i) v[i] = child_visits[i];
    v = v.normalized().unaryExpr([](float x) { return x ? x + 1 : x; });
    auto p = Stats::TempBasedProbs(v, n_moves_played < 15 ? 1 : 1e-2);
    for (int i = 0; i < BOARD_SIZE; ++i) out[i] = p[i];
    return 0;
}

// Default::AddNoise (MonteCarlo.hpp:97-108) on a root whose children carry `priors` (0 = no child), engine seeded first.
int ref_add_noise(const float* priors, uint32_t seed, float* out) {
    Node root;
    root.player = Player::White;
    for (int i = 0; i < BOARD_SIZE; ++i)
        if (priors[i] != 0.0f) {
            auto ch = std::make_unique<Node>();
            ch->parent = &root; ch->position = Position(i); ch->player = Player::Black; ch->action_prob = priors[i];
            root.children.emplace_back(std::move(ch));
        }
    Stats::RandomEngine().seed(seed);
    Default::AddNoise(&root);
    std::memset(out, 0, sizeof(float) * BOARD_SIZE);
    for (auto& ch : root.children) out[ch->position.id] = ch->action_prob;
    return 0;
}

}  // extern "C"

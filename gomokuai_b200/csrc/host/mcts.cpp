// mcts.cpp -- see mcts.h.  Semantics follow SURVEY.md Appendix A.6 and the cited reference lines.
#include "mcts.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <random>
#include <stdexcept>
#include <string>

#include "../../../include/gomoku_b200.h"

namespace gomoku {

// ---- Policy (MCTS.cpp:18-58) ---------------------------------------------------------------------------
Policy::Policy(SelectFunc f1, ExpandFunc f2, EvalFunc f3, UpdateFunc f4, double c_puct)
    : select(f1 ? f1 : [this](const Node* node) { return Default::Select(this, node); }),
      expand(f2 ? f2 : [this](Node* node, Board& board, const Probs& probs) { return Default::Expand(this, node, board, probs); }),
      simulate(f3 ? f3 : [this](Board& board) { return Default::Simulate(this, board); }),
      backPropogate(f4 ? f4 : [this](Node* node, Board& board, double value) { Default::BackPropogate(this, node, board, value); }),
      c_puct(c_puct) {}

// ---- Node pool ---------------------------------------------------------------------------------------------
namespace {
struct NodePool {
    union Slot { Slot* next; alignas(Node) unsigned char bytes[sizeof(Node)]; };
    static constexpr std::size_t kBlock = 8192;          // nodes per slab (never returned: trees are rebuilt all the time)
    std::atomic_flag busy = ATOMIC_FLAG_INIT;            // the engine is single-threaded like the reference; the lock only
    Slot* free_list = nullptr;                           // makes a stray use from another thread safe
    std::vector<std::unique_ptr<Slot[]>> slabs;
    void lock() { while (busy.test_and_set(std::memory_order_acquire)) {} }
    void unlock() { busy.clear(std::memory_order_release); }
    void* take() {
        lock();
        if (!free_list) {
            slabs.emplace_back(new Slot[kBlock]);
            Slot* slab = slabs.back().get();
            for (std::size_t i = 0; i < kBlock; ++i) { slab[i].next = free_list; free_list = &slab[i]; }
        }
        Slot* s = free_list;
        free_list = s->next;
        unlock();
        return s;
    }
    void give(void* p) {
        Slot* s = static_cast<Slot*>(p);
        lock();
        s->next = free_list;
        free_list = s;
        unlock();
    }
};
NodePool& node_pool() { static NodePool* pool = new NodePool; return *pool; }   // leaked on purpose: nodes may outlive static destruction
}  // namespace

void* Node::operator new(std::size_t size) { return size == sizeof(Node) ? node_pool().take() : ::operator new(size); }
void Node::operator delete(void* p, std::size_t size) noexcept {
    if (!p) return;
    if (size == sizeof(Node)) node_pool().give(p); else ::operator delete(p);
}

std::unique_ptr<Node> Policy::createNode(Node* parent, Position pose, Player player, float value, float prob) {
    return std::make_unique<Node>(parent, pose, player, value, prob);
}
void Policy::prepare(Board& board) { m_initActs = board.m_moveRecord.size(); }
void Policy::cleanup(Board& board) { board.revertMove(board.m_moveRecord.size() - m_initActs); }
Player Policy::applyMove(Board& board, Position move) { return board.applyMove(move, false); }   // no victory check inside the tree
Player Policy::revertMove(Board& board, std::size_t count) { return board.revertMove(count); }
bool Policy::checkGameEnd(Board& board) { return board.checkGameEnd(); }

// ---- default algorithms (MonteCarlo.hpp) ---------------------------------------------------------------
double Default::PUCB(const Node* node, double c_puct) {                                 // :23-28
    const double P = node->action_prob, N = static_cast<double>(node->parent->node_visits), n = static_cast<double>(node->node_visits + 1);
    return c_puct * P * std::sqrt(N) / n;
}

Probs Default::UniformProbs(const Board& board) {                                       // :50-55
    Probs p(BOARD_SIZE, 0.0f);
    const float each = 1.0f / static_cast<float>(board.moveCounts(Player::None));
    for (int i = 0; i < BOARD_SIZE; ++i) if (board.cell(i) == 0) p[i] = each;
    return p;
}

Node* Default::Select(Policy* policy, const Node* node) {                               // :57-68, ties -> lowest index
    std::size_t best = 0;
    double best_score = -1.0;
    for (std::size_t i = 0; i < node->children.size(); ++i) {
        const Node* child = node->children[i].get();
        const double score = child->state_value + PUCB(child, policy->c_puct);
        if (score > best_score) { best_score = score; best = i; }
    }
    return node->children[best].get();
}

std::size_t Default::Expand(Policy* policy, Node* node, Board& board, const Probs& probs, bool extraCheck) {   // :71-80
    node->children.reserve(probs.size());
    for (int i = 0; i < BOARD_SIZE; ++i)
        if (probs[i] != 0.0f && (!extraCheck || board.checkMove(i)))
            node->children.emplace_back(policy->createNode(node, i, -node->player, 0.0f, probs[i]));
    return node->children.size();
}

// The reference needs no device set-up call; a drop-in user of Board / MCTS should not either:
// bind the process to a GPU (LOCAL_RANK, default 0) the first time a GPU slot runs.
void ensure_gpu() {
    int device = -1;
    if (gk_device_info(&device, nullptr, nullptr, nullptr) == GK_OK) return;
    const char* lr = std::getenv("LOCAL_RANK");
    if (gk_init(lr ? std::atoi(lr) : 0) != GK_OK) throw std::runtime_error(std::string("gk_init: ") + gk_last_error());
}

// ---- random sources ------------------------------------------------------------------------------------------
namespace {
std::atomic<std::uint64_t> g_rollout_key{ 0 };         // 0 = not drawn yet
std::atomic<std::uint64_t> g_rollout_calls{ 0 };       // every simulate call takes a fresh (ctr_hi, position) pair: independent streams
std::atomic<std::uint64_t> g_noise_epoch{ 0 }, g_noise_seed{ 0 };

std::uint64_t rollout_key() {
    std::uint64_t k = g_rollout_key.load(std::memory_order_relaxed);
    if (k == 0) {
        std::random_device rd;
        std::uint64_t fresh = (std::uint64_t(rd()) << 32 | rd()) | 1u;
        if (g_rollout_key.compare_exchange_strong(k, fresh)) k = fresh;
    }
    return k;
}

std::mt19937& noise_engine() {
    static thread_local std::mt19937 engine{ std::random_device{}() };
    static thread_local std::uint64_t epoch = 0;
    const std::uint64_t e = g_noise_epoch.load(std::memory_order_acquire);
    if (epoch != e) { engine.seed(static_cast<std::uint32_t>(g_noise_seed.load())); epoch = e; }
    return engine;
}
}  // namespace

void set_seed(std::uint64_t seed) {
    g_rollout_key.store(seed * 0x9E3779B97F4A7C15ull | 1u);
    g_rollout_calls.store(0);
    g_noise_seed.store(seed);
    g_noise_epoch.fetch_add(1, std::memory_order_release);     // every thread's noise engine reseeds at its next use
}

float Default::GpuRolloutValue(const Board& board, int rollouts) {
    ensure_gpu();
    std::uint32_t packed[16];
    std::int32_t wdb[3] = { 0, 0, 0 };
    board.pack(packed);
    const std::uint64_t call = g_rollout_calls.fetch_add(1);
    if (gk_rollout_batch_host(packed, 1, rollouts, rollout_key(), static_cast<std::uint32_t>(call >> 30),
                              static_cast<int>(call & 0x3fffffff), wdb) != GK_OK)
        throw std::runtime_error(std::string("gk_rollout_batch_host: ") + gk_last_error());
    const float black_value = static_cast<float>(wdb[2] - wdb[0]) / static_cast<float>(rollouts);
    return CalcScore(board.m_curPlayer, black_value);
}

Policy::EvalResult Default::Simulate(Policy*, Board& board) {                           // :83-88, one random game
    return { GpuRolloutValue(board, 1), UniformProbs(board) };
}

void Default::BackPropogate(Policy*, Node* node, Board&, double value) {                // :90-95
    float v = static_cast<float>(value);
    for (; node != nullptr; node = node->parent, v = -v) {
        node->node_visits += 1;
        node->state_value += (v - node->state_value) / static_cast<float>(node->node_visits);
    }
}

// MonteCarlo.hpp:97-108 + Stats::DirichletNoise, Statistical.hpp:29-34.  Operation for operation in float like the
// reference: the priors go into a 225-vector by cell, are scaled by (1 - epsilon), every cell whose scaled prior is
// non-zero draws gamma(alpha, 1) IN CELL ORDER, the draws are L2-normalised (Eigen's normalized(), not a sum-to-one
// Dirichlet) and epsilon times them is added.  A root without children draws nothing.
void Default::AddNoise(Node* node, float alpha, float epsilon) {
    float prior[BOARD_SIZE] = { 0.0f }, noise[BOARD_SIZE];
    for (auto& child : node->children) prior[child->position] = child->action_prob;
    const float keep = 1 - epsilon;
    for (float& p : prior) p *= keep;
    std::gamma_distribution<float> gamma(alpha, 1.0f);
    std::mt19937& engine = noise_engine();
    float norm2 = 0.0f;
    for (int i = 0; i < BOARD_SIZE; ++i) {
        noise[i] = prior[i] ? gamma(engine) : 0.0f;
        norm2 += noise[i] * noise[i];
    }
    if (norm2 > 0.0f) {
        const float norm = std::sqrt(norm2);
        for (float& x : noise) x /= norm;
    }
    for (int i = 0; i < BOARD_SIZE; ++i) prior[i] += epsilon * noise[i];
    for (auto& child : node->children) child->action_prob = prior[child->position];
}

namespace { constexpr std::int32_t kCountPending = INT32_MIN; }   // no count is negative

Probs Policy::simulateBegin(Board&) { throw std::logic_error("this policy has no split simulate"); }
float Policy::simulateEnd() { throw std::logic_error("this policy has no split simulate"); }

// ---- RandomPolicy (policies/Random.h) --------------------------------------------------------------------
RandomPolicy::RandomPolicy(double c_puct, std::size_t c_rollouts)
    : Policy(nullptr, nullptr, [this](Board& board) { return averagedSimulate(board); }, nullptr, c_puct), c_rollouts(c_rollouts) {
    static std::atomic<int> next_slot{ 0 };
    m_slot = 8 + next_slot.fetch_add(1) % 8;             // slots 8..15: 0..7 belong to the root-parallel search
    m_ownSimulate = &simulate.target_type();
}

RandomPolicy::~RandomPolicy() {
    if (m_pendingPlayer != Player::None) gk_rollout_wait(m_slot);
    if (m_pinned) gk_host_free(m_pinned);
}

Policy::EvalResult RandomPolicy::averagedSimulate(Board& board) {                      // Random.h:22-35; the board is left untouched
    return { Default::GpuRolloutValue(board, static_cast<int>(c_rollouts)), Default::UniformProbs(board) };
}

// averagedSimulate in two halves: same Philox stream per call as GpuRolloutValue, so the same value; the probabilities
// are uniform over the empty cells and need no device, which lets the search expand the leaf while the playouts run.
Probs RandomPolicy::simulateBegin(Board& board) {
    ensure_gpu();
    if (m_pendingPlayer != Player::None) { gk_rollout_wait(m_slot); m_pendingPlayer = Player::None; }   // a begin whose end never came
    if (!m_pinned) {
        void* p = nullptr;
        if (gk_host_alloc(&p, 128) != GK_OK) throw std::runtime_error(std::string("gk_host_alloc: ") + gk_last_error());
        m_pinned = static_cast<std::uint32_t*>(p);
    }
    board.pack(m_pinned);
    volatile std::int32_t* counts = reinterpret_cast<volatile std::int32_t*>(m_pinned + 16);
    counts[0] = counts[1] = counts[2] = kCountPending;
    const std::uint64_t call = g_rollout_calls.fetch_add(1);
    if (gk_rollout_submit_host(m_slot, m_pinned, 1, static_cast<int>(c_rollouts), rollout_key(), static_cast<std::uint32_t>(call >> 30),
                               static_cast<int>(call & 0x3fffffff), reinterpret_cast<std::int32_t*>(m_pinned + 16)) != GK_OK)
        throw std::runtime_error(std::string("gk_rollout_submit_host: ") + gk_last_error());
    m_pendingPlayer = board.m_curPlayer;
    return Default::UniformProbs(board);
}

float RandomPolicy::simulateEnd() {
    if (m_pendingPlayer == Player::None) throw std::logic_error("simulateEnd without simulateBegin");
    const Player player = m_pendingPlayer;
    m_pendingPlayer = Player::None;
    // The kernel stores the three counts into this page-locked block as its last act: watching them arrive saves the
    // stream synchronisation's own latency (2-3 us of a ~27 us playout).  Not there after a while (or a launch that
    // failed): ask the stream, which also reports the error.
    volatile std::int32_t* counts = reinterpret_cast<volatile std::int32_t*>(m_pinned + 16);
    bool arrived = false;
    for (int spins = 0; spins < (1 << 22) && !arrived; ++spins) {
        arrived = counts[0] != kCountPending && counts[1] != kCountPending && counts[2] != kCountPending;
#if defined(__x86_64__)
        if (!arrived) __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (!arrived && gk_rollout_wait(m_slot) != GK_OK) throw std::runtime_error(std::string("gk_rollout_wait: ") + gk_last_error());
    const std::int32_t* wdb = reinterpret_cast<const std::int32_t*>(m_pinned + 16);
    const float black_value = static_cast<float>(wdb[2] - wdb[0]) / static_cast<float>(c_rollouts);
    return CalcScore(player, black_value);
}

// ---- RAVE (MonteCarlo.hpp:112-186) ------------------------------------------------------------------------------
double RAVE::HandSelect(const AMAFNode* node, std::size_t eqv_param) {                     // :126-130
    const double n = static_cast<double>(node->node_visits), k = static_cast<double>(eqv_param);
    return std::sqrt(k / (3 * n + k));
}
double RAVE::MinMSE(const AMAFNode* node, double c_bias) {                                 // :132-137
    const double n = static_cast<double>(node->node_visits + 1), n_ = static_cast<double>(node->amaf_visits);
    return n_ / (n + n_ + n * n_ * c_bias * c_bias);
}
double RAVE::WeightedValue(const AMAFNode* node, double) {                                 // :139-143: the hand-tuned schedule is the live one
    const double weight = HandSelect(node);
    return (1 - weight) * node->state_value + weight * node->amaf_value;
}

Node* RAVE::Select(Policy*, const Node* node) { return node->children[0].get(); }          // :150-153

void RAVE::BackPropogate(Policy* policy, Node* node, Board& board, double value, bool use_rave, double c_bias) {   // :155-186
    float v = static_cast<float>(value);
    for (; node != nullptr; node = node->parent, v = -v) {
        std::size_t max_index = 0;
        double max_score = -std::numeric_limits<double>::infinity();
        for (std::size_t i = 0; i < node->children.size(); ++i) {
            Node* child = node->children[i].get();
            double score = Default::PUCB(child, policy->c_puct);
            if (use_rave) {
                AMAFNode* rave = static_cast<AMAFNode*>(child);
                if (board.moveState(rave->player, rave->position)) {                      // the same player's stone stands there at the end
                    rave->amaf_visits += 1;
                    rave->amaf_value += (-v - rave->amaf_value) / static_cast<float>(rave->amaf_visits);
                }
                score += WeightedValue(rave, c_bias);
            } else {
                score += child->state_value;
            }
            if (score > max_score) { max_score = score; max_index = i; }
        }
        if (!node->children.empty()) node->children[0].swap(node->children[max_index]);       // best child to the front
        node->node_visits += 1;
        node->state_value += (v - node->state_value) / static_cast<float>(node->node_visits);
    }
}

// ---- TraditionalPolicy (policies/Traditional.h) ---------------------------------------------------------------------
TraditionalPolicy::TraditionalPolicy(double c_puct, double c_bias, bool use_rave)
    : Policy([this](const Node* node) { return RAVE::Select(this, node); },
             [this](Node* node, Board& board, const Probs& probs) { return Default::Expand(this, node, board, probs, false); },
             [this](Board& board) { return hybridSimulate(board); },
             [this](Node* node, Board& board, double value) { RAVE::BackPropogate(this, node, board, value, c_useRave, this->c_bias); }, c_puct),
      c_bias(c_bias), c_useRave(use_rave) {}

std::unique_ptr<Node> TraditionalPolicy::createNode(Node* parent, Position pose, Player player, float value, float prob) {
    if (c_useRave) return std::unique_ptr<Node>(new RAVE::AMAFNode(parent, pose, player, value, prob));
    return Policy::createNode(parent, pose, player, value, prob);
}

// ---- PoolRAVEPolicy (policies/PoolRAVE.h) ----------------------------------------------------------------------------
PoolRAVEPolicy::PoolRAVEPolicy(double c_puct, double c_bias)
    : Policy([this](const Node* node) { return RAVE::Select(this, node); },
             [this](Node* node, Board& board, const Probs& probs) { return Default::Expand(this, node, board, probs, false); },
             [this](Board& board) { return defaultSimulate(board); },
             [this](Node* node, Board& board, double value) { RAVE::BackPropogate(this, node, board, value, true, this->c_bias); }, c_puct),
      c_bias(c_bias) {}

std::unique_ptr<Node> PoolRAVEPolicy::createNode(Node* parent, Position pose, Player player, float value, float prob) {   // PoolRAVE.h:23-25
    return std::unique_ptr<Node>(new RAVE::AMAFNode(parent, pose, player, value, prob));
}

Policy::EvalResult PoolRAVEPolicy::defaultSimulate(Board& board) {                         // PoolRAVE.h:27-48
    Probs probs = Default::UniformProbs(board);              // before the playout: the board is NOT restored afterwards
    const Player init_player = board.m_curPlayer;
    ensure_gpu();
    std::uint32_t packed[16];
    board.pack(packed);
    std::int8_t winner = 0;
    std::uint8_t length = 0, moves[BOARD_SIZE];
    const std::uint64_t call = g_rollout_calls.fetch_add(1);
    if (gk_rollout_trace_host(packed, 1, rollout_key(), static_cast<std::uint32_t>(call >> 30), static_cast<int>(call & 0x3fffffff),
                              &winner, &length, moves) != GK_OK)
        throw std::runtime_error(std::string("gk_rollout_trace_host: ") + gk_last_error());
    for (int k = 0; k < length; ++k) board.applyMove(Position(moves[k]));                    // Default::RandomRollout's applyMove, victory checks on
    if (static_cast<int>(board.m_winner) != winner || board.m_curPlayer != Player::None)
        throw std::runtime_error("PoolRAVEPolicy: the replayed playout does not end where the kernel's did");
    return { CalcScore(init_player, board.m_winner), probs };
}

Policy::EvalResult TraditionalPolicy::hybridSimulate(Board& board) {                       // Traditional.h:49-69
    ensure_gpu();
    gk_table* table = nullptr;
    if (gk_table_default(&table) != GK_OK) throw std::runtime_error(std::string("gk_table_default: ") + gk_last_error());
    std::uint32_t packed[16];
    board.pack(packed);
    Probs probs(BOARD_SIZE, 0.0f);
    float value = 0.0f;
    if (gk_hybrid_simulate_batch_host(table, packed, 1, probs.data(), &value, nullptr) != GK_OK)
        throw std::runtime_error(std::string("gk_hybrid_simulate_batch_host: ") + gk_last_error());
    return { value, probs };                                                                // DecisiveFilter never sets report.level
}

// ---- MCTS (MCTS.cpp:60-198) ---------------------------------------------------------------------------------
static Node* updateRoot(MCTS& mcts, std::unique_ptr<Node>&& next) {
    mcts.m_root = std::move(next);
    mcts.m_root->parent = nullptr;
    return mcts.m_root.get();
}

MCTS::MCTS(milliseconds c_duration, Position last_move, Player last_player, std::shared_ptr<Policy> policy)
    : m_policy(policy ? policy : std::make_shared<RandomPolicy>()),
      m_root(m_policy->createNode(nullptr, last_move, last_player, 0.0f, 1.0f)),
      m_size(1), m_iterations(0), m_duration(c_duration), c_constraint(Constraint::Duration) {}

MCTS::MCTS(std::size_t c_iterations, Position last_move, Player last_player, std::shared_ptr<Policy> policy)
    : m_policy(policy ? policy : std::make_shared<RandomPolicy>()),
      m_root(m_policy->createNode(nullptr, last_move, last_player, 0.0f, 1.0f)),
      m_size(1), m_iterations(c_iterations), m_duration(0), c_constraint(Constraint::Iterations) {}

Position MCTS::getAction(Board& board) {
    runPlayouts(board);
    return stepForward()->position;
}

// Statistical.hpp:37-42, operation for operation: log and the division by the temperature in float, softmax in double
// (exp, left-to-right sum, division), back to float, entries <= epsilon zeroed
Probs TempBasedProbs(const Probs& logits, float temperature) {
    const float eps = std::numeric_limits<float>::epsilon();
    std::vector<double> t(logits.size());
    double sum = 0.0;
    for (std::size_t i = 0; i < t.size(); ++i) { t[i] = std::exp(static_cast<double>(std::log(logits[i] + eps) / temperature)); sum += t[i]; }
    Probs out(logits.size());
    for (std::size_t i = 0; i < t.size(); ++i) { const float p = static_cast<float>(t[i] / sum); out[i] = p > eps ? p : 0.0f; }
    return out;
}

Policy::EvalResult MCTS::evalState(Board& board) {                                      // MCTS.cpp:104-117 (the debug print is dropped)
    runPlayouts(board);
    Probs visits(BOARD_SIZE, 0.0f);
    for (auto& node : m_root->children) visits[node->position] = static_cast<float>(node->node_visits);
    float norm2 = 0.0f;                                                                 // normalized(): float, unchanged when the norm is 0
    for (float v : visits) norm2 += v * v;
    if (norm2 > 0.0f) { const float norm = std::sqrt(norm2); for (float& v : visits) v /= norm; }
    for (float& v : visits) if (v != 0.0f) v += 1.0f;
    return { m_root->state_value, TempBasedProbs(visits, board.m_moveRecord.size() < 15 ? 1.0f : 1e-2f) };
}

void MCTS::syncWithBoard(Board& board) {                                                // MCTS.cpp:119-125
    auto iter = std::find_if(board.m_moveRecord.begin(), board.m_moveRecord.end(),
                             [this](Position p) { return p.id == m_root->position.id; });
    for (iter = (iter == board.m_moveRecord.end() ? board.m_moveRecord.begin() : iter + 1); iter != board.m_moveRecord.end(); ++iter)
        stepForward(*iter);
}

Node* MCTS::stepForward() {                                                             // most visited child becomes the root
    auto iter = std::max_element(m_root->children.begin(), m_root->children.end(),
                                 [](auto&& l, auto&& r) { return l->node_visits < r->node_visits; });
    return iter != m_root->children.end() ? updateRoot(*this, std::move(*iter)) : m_root.get();
}

Node* MCTS::stepForward(Position next_move) {                                           // MCTS.cpp:136-147
    auto iter = std::find_if(m_root->children.begin(), m_root->children.end(),
                             [next_move](auto&& node) { return node->position.id == next_move.id; });
    if (iter == m_root->children.end())
        iter = m_root->children.emplace(m_root->children.end(), m_policy->createNode(nullptr, next_move, -m_root->player, 0.0f, 1.0f));
    return updateRoot(*this, std::move(*iter));
}

void MCTS::reset() {                                                                    // MCTS.cpp:149-156
    m_root = m_policy->createNode(nullptr, Position(-1), Player::White, 0.0f, 1.0f);
    m_size = 1;
}

std::size_t MCTS::playout(Board& board) {                                               // MCTS.cpp:158-177
    Node* node = m_root.get();
    while (!node->isLeaf()) {
        node = m_policy->select(node);
        m_policy->applyMove(board, node->position);
    }
    double node_value;
    std::size_t expand_size;
    if (!m_policy->checkGameEnd(board)) {
        if (m_policy->splitSimulate()) {                                 // expand while the GPU plays the leaf's rollouts
            const Probs action_probs = m_policy->simulateBegin(board);
            expand_size = m_policy->expand(node, board, action_probs);
            node_value = -m_policy->simulateEnd();
        } else {
            auto [state_value, action_probs] = m_policy->simulate(board);   // value for the player to move
            expand_size = m_policy->expand(node, board, action_probs);
            node_value = -state_value;                                   // the node stores the view of who moved INTO it
        }
    } else {
        expand_size = 0;
        node_value = CalcScore(node->player, board.m_winner);
    }
    m_policy->backPropogate(node, board, node_value);
    m_policy->revertMove(board, board.m_moveRecord.size() - m_policy->m_initActs);
    return expand_size;
}

void MCTS::runPlayouts(Board& board) {                                                  // MCTS.cpp:179-198
    const auto start = std::chrono::system_clock::now();
    syncWithBoard(board);
    Default::AddNoise(m_root.get());
    m_policy->prepare(board);
    if (c_constraint == Constraint::Duration) {
        m_iterations = 0;
        for (auto end = start; end - start < m_duration; end = std::chrono::system_clock::now(), ++m_iterations) m_size += playout(board);
    } else {
        for (std::size_t i = 0; i < m_iterations; ++i) m_size += playout(board);
        m_duration = std::chrono::duration_cast<milliseconds>(std::chrono::system_clock::now() - start);
    }
    m_policy->cleanup(board);
}

}  // namespace gomoku

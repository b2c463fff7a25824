// mcts.h -- host-side mirror of the reference's search layer (include/MCTS.h, src/MCTS.cpp,
// include/algorithms/MonteCarlo.hpp, include/policies/Random.h): Node, the Policy plugin (four
// std::function slots + virtuals), MCTS.  The tree stays on the host; what a slot computes may run
// on the GPU (RandomPolicy::simulate calls gk_rollout_batch_host).  Probability vectors cross as
// std::vector<float>(225) where the reference uses Eigen::VectorXf.
#pragma once
#include <chrono>
#include <cstdint>
#include <functional>
#include <memory>
#include <tuple>
#include <typeinfo>
#include <vector>

#include "game.h"

namespace gomoku {

using std::chrono::milliseconds;
using Probs = std::vector<float>;

constexpr double C_PUCT = 5.0;                       // MCTS.h:17-19
constexpr std::size_t C_ITERATIONS = 10000;
constexpr milliseconds C_DURATION{ 1000 };

struct Node {                                        // MCTS.h:25-66
    Node* parent = nullptr;
    Position position = Position(-1);                // the move that led here ...
    Player player = Player::None;                    // ... and who played it
    float state_value = 0.0f;                        // running mean, from `player`'s point of view
    float action_prob = 0.0f;
    std::size_t node_visits = 0;
    std::vector<std::unique_ptr<Node>> children;

    Node() = default;
    Node(const Node&) = delete;                      // children are uniquely owned (MCTS.h:56-61)
    Node(Node&&) = default;
    Node& operator=(Node&&) = default;
    Node(Node* parent, Position position, Player player, float value, float prob)
        : parent(parent), position(position), player(player), state_value(value), action_prob(prob) {}
    // Policy::createNode may return a derived node (RAVE::AMAFNode) behind unique_ptr<Node>.  The reference deletes it
    // through the base pointer without a virtual destructor (MonteCarlo.hpp:114-124, PoolRAVE.h:23-25); here the
    // destructor is virtual so that the pooled allocator below sees the true size.
    virtual ~Node() = default;
    bool isLeaf() const { return children.empty(); }
    bool isFull(const Board& board) const { return children.size() == board.moveCounts(Player::None); }

    // Nodes come from a pooled free list (row f2 of SURVEY 8f: the reference spends 40-50 % of a playout in operator new,
    // one make_unique per child, MonteCarlo.hpp:71-80).  Same ownership model -- unique_ptr<Node> children created
    // through Policy::createNode -- only the allocation behind it changes.
    static void* operator new(std::size_t size);
    static void operator delete(void* p, std::size_t size) noexcept;
};

class Policy {                                       // MCTS.h:69-132
public:
    using SelectFunc = std::function<Node*(const Node*)>;
    using ExpandFunc = std::function<std::size_t(Node*, Board&, const Probs&)>;
    using EvalResult = std::tuple<float, Probs>;     // value for the player to move, 225 move probabilities (0 on occupied cells)
    using EvalFunc = std::function<EvalResult(Board&)>;
    using UpdateFunc = std::function<void(Node*, Board&, double)>;

    SelectFunc select;
    ExpandFunc expand;
    EvalFunc simulate;
    UpdateFunc backPropogate;

    // a null slot takes the default algorithm (MCTS.cpp:18-33)
    explicit Policy(SelectFunc = nullptr, ExpandFunc = nullptr, EvalFunc = nullptr, UpdateFunc = nullptr, double c_puct = C_PUCT);
    virtual ~Policy() = default;

    virtual std::unique_ptr<Node> createNode(Node* parent, Position pose, Player player, float value, float prob);
    virtual void prepare(Board& board);
    virtual void cleanup(Board& board);
    virtual Player applyMove(Board& board, Position move);
    virtual Player revertMove(Board& board, std::size_t count = 1);
    virtual bool checkGameEnd(Board& board);
    virtual void reset() {}

    // Split form of the policy's OWN simulate, for policies whose move probabilities do not need the device (new; the
    // reference's simulate is synchronous CPU code): simulateBegin sends the leaf to the GPU and returns the
    // probabilities, simulateEnd waits for the value, and MCTS::playout expands the leaf in between.  Same numbers as
    // `simulate`.  Only used while the `simulate` slot still holds what the policy's constructor put there.
    bool splitSimulate() const { return m_ownSimulate != nullptr && simulate.target_type() == *m_ownSimulate; }
    virtual Probs simulateBegin(Board& board);
    virtual float simulateEnd();

    double c_puct;
    std::size_t m_initActs = 0;

protected:
    const std::type_info* m_ownSimulate = nullptr;   // type of the callable the constructor stored in `simulate`, if it has a split form
};

// The default algorithms (MonteCarlo.hpp:13-110).  Simulate plays its random game on the GPU.
struct Default {
    static double PUCB(const Node* node, double c_puct);
    static Probs UniformProbs(const Board& board);
    static Node* Select(Policy* policy, const Node* node);
    static std::size_t Expand(Policy* policy, Node* node, Board& board, const Probs& probs, bool extraCheck = true);
    static Policy::EvalResult Simulate(Policy* policy, Board& board);
    static void BackPropogate(Policy* policy, Node* node, Board& board, double value);
    static void AddNoise(Node* node, float alpha = 0.05f, float epsilon = 0.25f);
    // mean value for the player to move of `rollouts` random playouts from `board`, run by the
    // rollout kernel (include/gomoku_b200.h: gk_rollout_batch_host); throws std::runtime_error on failure
    static float GpuRolloutValue(const Board& board, int rollouts);
};

// RandomPolicy (policies/Random.h:15-35): `c_rollouts` playouts averaged per leaf
class RandomPolicy : public Policy {
public:
    explicit RandomPolicy(double c_puct = C_PUCT, std::size_t c_rollouts = 5);
    ~RandomPolicy() override;
    RandomPolicy(const RandomPolicy&) = delete;      // the slots capture `this`, and the page-locked block has one owner
    RandomPolicy& operator=(const RandomPolicy&) = delete;
    EvalResult averagedSimulate(Board& board);
    Probs simulateBegin(Board& board) override;      // the c_rollouts playouts of averagedSimulate, in flight on the GPU
    float simulateEnd() override;
    std::size_t c_rollouts;

private:
    std::uint32_t* m_pinned = nullptr;               // page-locked: 16 board words + 3 counts
    Player m_pendingPlayer = Player::None;           // side to move of the leaf in flight (None: nothing in flight)
    int m_slot = 8;
};

// RAVE (algorithms/MonteCarlo.hpp:112-186): back-propagation keeps the best-scoring child at index 0 of every node on
// the path, select takes children[0].  UseRave: the score is PUCB + WeightedValue(state value, AMAF value) and a child
// whose (player, position) stone stands on `board` -- the END of the playout -- gets an AMAF update; otherwise
// PUCB + state value (what TraditionalPolicy uses).
struct RAVE {
    struct AMAFNode : public Node {                  // MonteCarlo.hpp:114-124
        float amaf_value = 0.0f;
        std::size_t amaf_visits = 0;
        AMAFNode(Node* parent, Position pose, Player player, float Q, float P) : Node(parent, pose, player, Q, P) {}
    };
    static double HandSelect(const AMAFNode* node, std::size_t eqv_param = 800);
    static double MinMSE(const AMAFNode* node, double c_bias);
    static double WeightedValue(const AMAFNode* node, double c_bias);
    static Node* Select(Policy* policy, const Node* node);
    static void BackPropogate(Policy* policy, Node* node, Board& board, double value, bool use_rave = false, double c_bias = 0.0);
};

// PoolRAVEPolicy (policies/PoolRAVE.h:13-52): AMAF nodes, RAVE select / update, ONE random playout per leaf whose
// final position stays on the board until MCTS::playout reverts it (the AMAF update reads it).  The playout runs on the
// GPU (gk_rollout_trace_host) and its moves are replayed on the host board.
class PoolRAVEPolicy : public Policy {
public:
    explicit PoolRAVEPolicy(double c_puct = 1e-4, double c_bias = 1e-1);
    std::unique_ptr<Node> createNode(Node* parent, Position pose, Player player, float value, float prob) override;
    EvalResult defaultSimulate(Board& board);
    double c_bias;
};

// TraditionalPolicy (policies/Traditional.h:13-73): pattern evaluator instead of random playouts.
// simulate = hybridSimulate: Heuristic::EvaluationProbs filtered by Heuristic::DecisiveFilter, and
// Heuristic::EvaluationValue, computed by ac_eval_kernel (gk_hybrid_simulate_batch_host).  The reference keeps an
// incremental Evaluator synchronised with the tree walk (CachedApplyMove / CachedRevertMove); the GPU evaluates
// the leaf position from scratch, so the board itself is walked (Policy::applyMove / revertMove).
// The 3-argument constructor exists only in the reference's Python binding (core/py_ext/src/policy_ext.hpp:39-45; the
// library header has TraditionalPolicy(puct) alone): use_rave = true here selects AMAF nodes and RAVE's UseRave update
// with c_bias, fed -- as the binding's argument order implies -- by the stones of the leaf position.
class TraditionalPolicy : public Policy {
public:
    explicit TraditionalPolicy(double c_puct = C_PUCT, double c_bias = 0.0, bool use_rave = false);
    std::unique_ptr<Node> createNode(Node* parent, Position pose, Player player, float value, float prob) override;
    EvalResult hybridSimulate(Board& board);
    double c_bias;
    bool c_useRave;
};

class MCTS {                                         // MCTS.h:135-180
public:
    explicit MCTS(milliseconds c_duration = C_DURATION, Position last_move = -1, Player last_player = Player::White,
                  std::shared_ptr<Policy> policy = nullptr);
    explicit MCTS(std::size_t c_iterations, Position last_move = -1, Player last_player = Player::White,
                  std::shared_ptr<Policy> policy = nullptr);

    Position getAction(Board& board);
    Policy::EvalResult evalState(Board& board);
    Node* stepForward();
    Node* stepForward(Position next_move);
    void syncWithBoard(Board& board);
    void reset();

    std::shared_ptr<Policy> m_policy;
    std::unique_ptr<Node> m_root;
    std::size_t m_size;
    std::size_t m_iterations;
    milliseconds m_duration;

private:
    std::size_t playout(Board& board);
    void runPlayouts(Board& board);
    enum class Constraint { Iterations, Duration } c_constraint;
};

void ensure_gpu();   // gk_init(LOCAL_RANK or 0) on first use; throws std::runtime_error without a usable GPU

// The mirror's random sources -- the Philox key of the GPU playouts behind RandomPolicy / PoolRAVEPolicy / Default::Simulate
// and the calling thread's Dirichlet-noise engine -- are seeded from std::random_device once per process / thread, as the
// reference seeds its engines (Game.cpp:11-12, Statistical.hpp:23-26).  set_seed makes a run reproducible.
void set_seed(std::uint64_t seed);

// Statistical helpers used by evalState (algorithms/Statistical.hpp:37-42)
Probs TempBasedProbs(const Probs& logits, float temperature);

}  // namespace gomoku

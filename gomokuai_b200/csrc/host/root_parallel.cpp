// root_parallel.cpp -- see root_parallel.h
#include "root_parallel.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../../include/gomoku_b200.h"
#include "mcts.h"

namespace gomoku {

namespace {

// Arena node.  The reference expands a leaf into one heap node per empty cell (MonteCarlo.hpp:71-80; 40-50 % of its
// time is operator new, Final Report.md:102).  With RandomPolicy every child of a node has the same prior and starts
// at value 0 / visits 0, so all NEVER-VISITED children of a node tie on the PUCB score and Default::Select takes the
// lowest cell among them (strict `>`, MonteCarlo.hpp:57-68).  Children are therefore materialised only when they are
// first selected -- in increasing cell order, kept in a sibling list -- and the untouched ones are represented by a
// cursor.  Selection visits exactly the nodes the eager tree would (tests/test_gpu_mcts.py compares the two).
// The root keeps a materialised block of children when Dirichlet noise makes their priors differ.
struct ANode {
    std::int32_t parent, first_child, last_child, next_sibling;
    std::int16_t n_children;        // children that exist as nodes
    std::int16_t n_moves;           // legal moves here (empty cells); > 0 once the node has been expanded
    std::int16_t position, cursor;  // the move into this node; highest cell materialised so far (-1 none)
    float value, prior;             // running mean from the view of who moved into the node; prior of the move
    std::int32_t visits;
    std::int8_t player;             // who played `position`
    std::int8_t eager;              // children are one contiguous block [first_child, first_child + n_children)
};

struct Tree {
    std::vector<ANode> nodes;
    Board board;
    std::int32_t leaf = 0;          // selected this round
    std::int16_t root_child = -1;   // root move of the current path (-1: the leaf is the root itself)
    bool terminal = false;
    std::mt19937 rng;
    std::array<std::int32_t, BOARD_SIZE> black_wins{}, white_wins{};   // rollouts below each root child
};

// minimal fork-join pool: run(f) calls f(worker) on every worker and returns when all are done
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int i = 1; i < n_; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int)>& f) {
        { std::lock_guard<std::mutex> l(mu_); job_ = &f; pending_ = n_ - 1; ++epoch_; }
        cv_.notify_all();
        f(0);
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
    }
private:
    void loop(int id) {
        std::uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                job = job_;
            }
            (*job)(id);
            { std::lock_guard<std::mutex> l(mu_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    int pending_ = 0;
    std::uint64_t epoch_ = 0;
    bool stop_ = false;
};

}  // namespace

struct RootParallelSearch::Impl {
    std::vector<Tree> trees;
    std::uint32_t* packed = nullptr;        // trees x 16, page-locked (allocated on first run, when the GPU is bound)
    std::int32_t* wdb = nullptr;            // trees x 3, page-locked
    std::unique_ptr<Pool> pool;
    ~Impl() { gk_host_free(packed); gk_host_free(wdb); }
};

RootParallelSearch::RootParallelSearch(const RootParallelConfig& cfg) : m(new Impl), m_cfg(cfg) {
    if (cfg.trees <= 0 || cfg.c_rollouts <= 0) throw std::invalid_argument("trees and c_rollouts must be positive");
    int threads = cfg.threads > 0 ? cfg.threads : static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(threads, cfg.trees));
    m->pool.reset(new Pool(threads));
    m->trees.resize(cfg.trees);
}

RootParallelSearch::~RootParallelSearch() { delete m; }

Position RootParallelSearch::bestMove(const Stats& stats) {
    int best = -1;
    std::int64_t most = -1;
    for (int c = 0; c < BOARD_SIZE; ++c)
        if (stats[c] > most) { most = stats[c]; best = c; }
    return most > 0 ? Position(best) : Position(-1);
}

static ANode make_node(std::int32_t parent, int position, float prior, int player) {
    ANode n{};
    n.parent = parent; n.first_child = n.last_child = n.next_sibling = -1;
    n.position = static_cast<std::int16_t>(position); n.cursor = -1;
    n.prior = prior; n.player = static_cast<std::int8_t>(player);
    return n;
}

// expand `node` over the empty cells of `board` with uniform priors (MonteCarlo.hpp:50-55,71-80); `eager`
// materialises the children at once (the reference's layout), otherwise they appear when first selected
static void expand(Tree& t, std::int32_t node, const Board& board, bool eager) {
    const int empties = static_cast<int>(board.moveCounts(Player::None));
    if (empties == 0) return;
    t.nodes[node].n_moves = static_cast<std::int16_t>(empties);
    if (!eager) return;
    const float prior = 1.0f / static_cast<float>(empties);
    const std::int32_t first = static_cast<std::int32_t>(t.nodes.size());
    const int player = -t.nodes[node].player;
    for (int c = 0; c < BOARD_SIZE; ++c)
        if (board.cell(c) == 0) t.nodes.push_back(make_node(node, c, prior, player));
    t.nodes[node].first_child = first;
    t.nodes[node].n_children = static_cast<std::int16_t>(empties);
    t.nodes[node].eager = 1;
}

// Default::Select (MonteCarlo.hpp:57-68) on a lazily expanded node: the best materialised child, or -- when the
// common score of the never-visited children is strictly higher -- the lowest never-visited cell, created now.
static std::int32_t select_lazy(Tree& t, std::int32_t node, double c_puct) {
    const ANode parent = t.nodes[node];
    const double sq = std::sqrt(static_cast<double>(parent.visits));
    const float prior = 1.0f / static_cast<float>(parent.n_moves);
    std::int32_t best = parent.first_child;                          // as the reference: the first child unless one scores above -1
    double best_score = -1.0;
    for (std::int32_t c = parent.first_child; c >= 0; c = t.nodes[c].next_sibling) {   // increasing cell order
        const ANode& ch = t.nodes[c];
        const double score = ch.value + c_puct * ch.prior * sq / static_cast<double>(ch.visits + 1);
        if (score > best_score) { best_score = score; best = c; }
    }
    if (parent.n_children < parent.n_moves) {                       // someone has never been visited: value 0, visits 0
        const double fresh = 0.0 + c_puct * prior * sq / 1.0;
        if (best < 0 || fresh > best_score) {
            int cell = parent.cursor + 1;
            while (t.board.cell(cell) != 0) ++cell;                 // the lowest empty cell above the cursor
            const std::int32_t created = static_cast<std::int32_t>(t.nodes.size());
            t.nodes.push_back(make_node(node, cell, prior, -parent.player));
            ANode& p = t.nodes[node];
            if (p.last_child >= 0) t.nodes[p.last_child].next_sibling = created; else p.first_child = created;
            p.last_child = created;
            p.n_children += 1;
            p.cursor = static_cast<std::int16_t>(cell);
            return created;
        }
    }
    return best;
}

void RootParallelSearch::run(const Board& root, int playouts_per_tree) {
    const auto t_start = std::chrono::steady_clock::now();
    m_stats.fill(0);
    seconds_gpu = 0;
    leaves = 0;
    const int n_trees = m_cfg.trees;
    const Player root_last = root.m_moveRecord.empty() ? Player::White : -root.m_curPlayer;   // MCTS.h:138-151
    for (int i = 0; i < n_trees; ++i) {
        Tree& t = m->trees[i];
        t.nodes.clear();
        t.nodes.reserve(m_cfg.eager ? static_cast<std::size_t>(playouts_per_tree) * 64 + 256 : static_cast<std::size_t>(playouts_per_tree) * 2 + 256);
        t.nodes.push_back(make_node(-1, -1, 1.0f, static_cast<int>(root_last)));
        t.board = root;
        t.black_wins.fill(0);
        t.white_wins.fill(0);
        t.rng.seed(static_cast<std::uint32_t>(m_cfg.seed * 2654435761u + static_cast<std::uint32_t>(m_cfg.replica_base + i)));
    }
    ensure_gpu();
    if (!m->packed) {
        void *a = nullptr, *b = nullptr;
        if (gk_host_alloc(&a, static_cast<std::size_t>(n_trees) * 64) != GK_OK || gk_host_alloc(&b, static_cast<std::size_t>(n_trees) * 12) != GK_OK)
            throw std::runtime_error(std::string("gk_host_alloc: ") + gk_last_error());
        m->packed = static_cast<std::uint32_t*>(a);
        m->wdb = static_cast<std::int32_t*>(b);
    }
    Board probe = root;
    if (probe.checkGameEnd() || playouts_per_tree <= 0) { seconds_total = 0; return; }   // nothing to search from a decided position
    const double c_puct = m_cfg.c_puct;
    const std::size_t root_depth = root.m_moveRecord.size();
    const int workers = m->pool->size();

    for (int round = 0; round < playouts_per_tree; ++round) {
        // ---- selection: every tree walks to a leaf and packs its position -----------------------------
        m->pool->run([&](int w) {
            for (int i = w; i < n_trees; i += workers) {
                Tree& t = m->trees[i];
                std::int32_t node = 0;
                t.root_child = -1;
                while (t.nodes[node].n_moves > 0) {                  // expanded: descend
                    if (!t.nodes[node].eager) {
                        node = select_lazy(t, node, c_puct);
                    } else {
                        const ANode& parent = t.nodes[node];
                        const double sq = std::sqrt(static_cast<double>(parent.visits));
                        std::int32_t best = parent.first_child;
                        double best_score = -1.0;
                        for (std::int32_t c = parent.first_child, e = c + parent.n_children; c < e; ++c) {
                            const ANode& ch = t.nodes[c];
                            const double score = ch.value + c_puct * ch.prior * sq / static_cast<double>(ch.visits + 1);   // PUCB, :23-28,57-68
                            if (score > best_score) { best_score = score; best = c; }
                        }
                        node = best;
                    }
                    if (t.root_child < 0) t.root_child = t.nodes[node].position;
                    t.board.applyMove(Position(t.nodes[node].position), false);
                }
                t.leaf = node;
                t.terminal = t.board.checkGameEnd();                  // MCTS.cpp:166
                t.board.pack(&m->packed[static_cast<std::size_t>(i) * 16]);
            }
        });
        // ---- simulation: all leaves in one launch ------------------------------------------------------------
        const auto g0 = std::chrono::steady_clock::now();
        if (gk_rollout_batch_host(m->packed, n_trees, m_cfg.c_rollouts, m_cfg.seed, static_cast<std::uint32_t>(round),
                                  m_cfg.replica_base, m->wdb) != GK_OK)
            throw std::runtime_error(std::string("gk_rollout_batch_host: ") + gk_last_error());
        seconds_gpu += std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
        leaves += n_trees;
        // ---- expansion + backup -----------------------------------------------------------------------------------
        m->pool->run([&](int w) {
            for (int i = w; i < n_trees; i += workers) {
                Tree& t = m->trees[i];
                const std::int32_t* r = &m->wdb[static_cast<std::size_t>(i) * 3];
                const float black_value = static_cast<float>(r[2] - r[0]) / static_cast<float>(m_cfg.c_rollouts);
                if (!t.terminal) {
                    expand(t, t.leaf, t.board, m_cfg.eager || (t.leaf == 0 && m_cfg.noise));
                    if (t.leaf == 0 && m_cfg.noise) {                // Default::AddNoise on the root's fresh children
                        ANode& rt = t.nodes[0];
                        std::gamma_distribution<float> gamma(0.05f, 1.0f);
                        std::vector<float> g(rt.n_children);
                        double n2 = 0;
                        for (float& x : g) { x = gamma(t.rng); n2 += double(x) * x; }
                        const float inv = n2 > 0 ? static_cast<float>(1.0 / std::sqrt(n2)) : 0.0f;
                        for (int k = 0; k < rt.n_children; ++k) {
                            ANode& ch = t.nodes[rt.first_child + k];
                            ch.prior = ch.prior * 0.75f + 0.25f * g[k] * inv;
                        }
                    }
                }
                if (t.root_child >= 0) { t.black_wins[t.root_child] += r[2]; t.white_wins[t.root_child] += r[0]; }
                float v = static_cast<float>(t.nodes[t.leaf].player) * black_value;   // value for who moved into the leaf
                for (std::int32_t n = t.leaf; n >= 0; n = t.nodes[n].parent, v = -v) {
                    ANode& nd = t.nodes[n];
                    nd.visits += 1;
                    nd.value += (v - nd.value) / static_cast<float>(nd.visits);       // MonteCarlo.hpp:90-95
                }
                t.board.revertMove(t.board.m_moveRecord.size() - root_depth);
            }
        });
    }
    // ---- root statistics: integers, summed over trees ---------------------------------------------------------------
    nodes = 0;
    for (Tree& t : m->trees) {
        nodes += static_cast<std::int64_t>(t.nodes.size());
        const ANode& rt = t.nodes[0];
        if (rt.eager) {
            for (std::int32_t c = rt.first_child, e = c + rt.n_children; c < e; ++c) m_stats[t.nodes[c].position] += t.nodes[c].visits;
        } else {
            for (std::int32_t c = rt.first_child; c >= 0; c = t.nodes[c].next_sibling) m_stats[t.nodes[c].position] += t.nodes[c].visits;
        }
        for (int c = 0; c < BOARD_SIZE; ++c) {
            m_stats[BOARD_SIZE + c] += t.black_wins[c];
            m_stats[2 * BOARD_SIZE + c] += t.white_wins[c];
        }
    }
    seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
}

}  // namespace gomoku

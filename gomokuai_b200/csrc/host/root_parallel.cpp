// root_parallel.cpp -- see root_parallel.h
#include "root_parallel.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sys/mman.h>
#endif

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
// A node carries its statistics inline (one cache line per child when a sibling list is walked: the search is bound
// by memory latency, ~80 MB of trees per rank).  Only the root's materialised block of 200+ children keeps them in
// three parallel arrays instead, so that its scan streams 2.6 KB and vectorises.
struct ANode {                      // 32 bytes: two per cache line, none across a line boundary
    std::int32_t parent, first_child, last_child, next_sibling;
    float value;                    // running mean from the view of who moved into the node
    std::int32_t visits;
    std::int16_t position, cursor;  // the move into this node; highest cell materialised so far (-1 none)
    std::uint8_t n_children;        // children that exist as nodes (<= 225)
    std::uint8_t n_moves;           // legal moves here (empty cells); > 0 once the node has been expanded
    std::int8_t player;             // who played `position`
    std::int8_t eager;              // children are one contiguous block [first_child, first_child + n_children)
    // (no prior: every child of a node has the prior 1 / n_moves of its parent -- Default::UniformProbs -- except the
    //  root's block, whose priors Dirichlet noise may change and which live in Tree::rprior; the root itself has 1)
};
static_assert(sizeof(ANode) == 32, "two nodes per cache line");

// A tree's nodes: the little of std::vector the search uses, over storage that normally is a slice of one slab shared by
// all trees of the searcher (huge pages: ~70 MB of nodes touched at random otherwise miss the TLB on every access) and
// only falls back to its own heap block when a tree outgrows its slice.
class NodeArray {
public:
    NodeArray() = default;
    NodeArray(const NodeArray&) = delete;
    NodeArray& operator=(const NodeArray&) = delete;
    NodeArray(NodeArray&& o) noexcept { *this = std::move(o); }
    NodeArray& operator=(NodeArray&& o) noexcept {
        std::swap(m_data, o.m_data); std::swap(m_size, o.m_size); std::swap(m_cap, o.m_cap); std::swap(m_own, o.m_own);
        return *this;
    }
    ~NodeArray() { if (m_own) std::free(m_data); }
    ANode& operator[](std::size_t i) { return m_data[i]; }
    const ANode& operator[](std::size_t i) const { return m_data[i]; }
    std::size_t size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    void clear() { m_size = 0; }
    void adopt(ANode* storage, std::size_t capacity) {            // empty array over caller-owned storage
        if (m_own) std::free(m_data);
        m_data = storage; m_size = 0; m_cap = capacity; m_own = false;
    }
    void reserve(std::size_t n) { if (n > m_cap) grow(n); }
    void push_back(const ANode& n) {
        if (m_size == m_cap) grow(m_cap ? 2 * m_cap : 256);
        m_data[m_size++] = n;
    }

private:
    void grow(std::size_t n) {
        ANode* fresh = static_cast<ANode*>(std::malloc(n * sizeof(ANode)));
        if (!fresh) throw std::bad_alloc();
        if (m_size) std::memcpy(fresh, m_data, m_size * sizeof(ANode));
        if (m_own) std::free(m_data);
        m_data = fresh; m_cap = n; m_own = true;
    }
    ANode* m_data = nullptr;
    std::size_t m_size = 0, m_cap = 0;
    bool m_own = false;
};

struct Tree {
    NodeArray nodes;
    std::array<float, BOARD_SIZE> rvalue, rprior;   // statistics of the root's eager block, nodes [rfirst, rfirst + rn)
    std::array<std::int32_t, BOARD_SIZE> rvisits;
    std::int32_t rfirst = 0, rn = 0;
    Board board;
    std::int32_t leaf = 0;          // selected this round
    std::int16_t root_child = -1;   // root move of the current path (-1: the leaf is the root itself)
    bool terminal = false;
    std::mt19937 rng;
    struct Wins { std::int32_t black, white; };
    std::array<Wins, BOARD_SIZE> wins{};                             // rollouts won below each root child (one cache line per update)

    std::int32_t add(std::int32_t parent, int position, int player) {
        ANode n{};
        n.parent = parent; n.first_child = n.last_child = n.next_sibling = -1;
        n.position = static_cast<std::int16_t>(position); n.cursor = -1;
        n.player = static_cast<std::int8_t>(player);
        nodes.push_back(n);
        return static_cast<std::int32_t>(nodes.size()) - 1;
    }
    void clear(std::size_t reserve, ANode* slice) {               // slice: `reserve` nodes of the searcher's slab, or null
        rfirst = rn = 0;
        if (slice) nodes.adopt(slice, reserve);
        else { nodes.clear(); nodes.reserve(reserve); }
    }
    bool in_root_block(std::int32_t i) const { return static_cast<std::uint32_t>(i - rfirst) < static_cast<std::uint32_t>(rn); }
    std::int32_t visits_of(std::int32_t i) const { return in_root_block(i) ? rvisits[i - rfirst] : nodes[i].visits; }
};

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

// spin, then yield: the workers of a search own their cores, but must survive being oversubscribed
struct Backoff {
    int spins = 0;
    void operator()() { if (++spins < 4096) cpu_relax(); else std::this_thread::yield(); }
};

// Thread team: run(f) calls f(id) on every member and returns when all are done (one fork-join per search).
class Team {
public:
    explicit Team(int n) : n_(n) {
        for (int i = 1; i < n_; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~Team() {
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

constexpr std::int32_t kCountPending = INT32_MIN;   // no count is negative
constexpr int kMaxGroups = 8;       // leaf batches in flight at most (gk_rollout_submit_host slots 0..7)
constexpr int kAutoGroups = 8;      // ... and when the caller leaves the choice open (measured: scripts/ab_root_parallel.py,
                                    // profiles/r03q_root_parallel_groups.json)

}  // namespace

struct RootParallelSearch::Impl {
    std::vector<Tree> trees;                // declared before the slab's owner below: the trees' slices die first
    std::uint32_t* packed = nullptr;        // trees x 16, page-locked (allocated on first run, when the GPU is bound)
    std::int32_t* wdb = nullptr;            // trees x 3, page-locked
    std::unique_ptr<Team> team;
    void* slab = nullptr;                   // node storage of all trees, 2 MB aligned, MADV_HUGEPAGE
    std::size_t slab_bytes = 0;
    // `nodes_per_tree` nodes for each of `n` trees, or null when that is too much to hold in one piece
    ANode* node_slab(std::size_t n, std::size_t nodes_per_tree) {
        constexpr std::size_t kHuge = std::size_t(2) << 20, kMost = std::size_t(1) << 30;
        const std::size_t want = (n * nodes_per_tree * sizeof(ANode) + kHuge - 1) / kHuge * kHuge;
        if (want > kMost) return nullptr;
        if (want > slab_bytes) {
            for (Tree& t : trees) t.nodes.adopt(nullptr, 0);
            std::free(slab);
            slab = std::aligned_alloc(kHuge, want);
            slab_bytes = slab ? want : 0;
#if defined(__linux__)
            if (slab) madvise(slab, want, MADV_HUGEPAGE);            // advisory: 4 KB pages work too, just slower
#endif
        }
        return static_cast<ANode*>(slab);
    }
    ~Impl() {
        for (Tree& t : trees) t.nodes.adopt(nullptr, 0);
        std::free(slab);
        gk_host_free(packed); gk_host_free(wdb);
    }
};

RootParallelSearch::RootParallelSearch(const RootParallelConfig& cfg) : m(new Impl), m_cfg(cfg) {
    if (cfg.trees <= 0 || cfg.c_rollouts <= 0) throw std::invalid_argument("trees and c_rollouts must be positive");
    int threads = cfg.threads > 0 ? cfg.threads : static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(threads, cfg.trees));             // every thread owns trees; none is set aside to drive the GPU
    m->team.reset(new Team(threads));
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

// expand `node` over the empty cells of `board` with uniform priors (MonteCarlo.hpp:50-55,71-80); `eager`
// materialises the children at once (the reference's layout), otherwise they appear when first selected
static void expand(Tree& t, std::int32_t node, const Board& board, bool eager) {
    const int empties = static_cast<int>(board.moveCounts(Player::None));
    if (empties == 0) return;
    t.nodes[node].n_moves = static_cast<std::uint8_t>(empties);
    if (!eager) return;
    const float prior = 1.0f / static_cast<float>(empties);
    const std::int32_t first = static_cast<std::int32_t>(t.nodes.size());
    const int player = -t.nodes[node].player;
    for (int c = 0; c < BOARD_SIZE; ++c)
        if (board.cell(c) == 0) t.add(node, c, player);
    t.nodes[node].first_child = first;
    t.nodes[node].n_children = static_cast<std::uint8_t>(empties);
    t.nodes[node].eager = 1;
    if (node == 0) {                                                 // the root's block: statistics in parallel arrays
        t.rfirst = first; t.rn = empties;
        std::fill_n(t.rvalue.begin(), empties, 0.0f); std::fill_n(t.rprior.begin(), empties, prior); std::fill_n(t.rvisits.begin(), empties, 0);
    }
}

// PUCB score of node c under a parent with sqrt(visits) = sq: state_value + c_puct * P * sqrt(N) / (n + 1), the
// product evaluated left to right in double like the reference (MonteCarlo.hpp:23-28,62)
static inline double pucb(const ANode& ch, float prior, double c_puct, double sq) {
    return ch.value + c_puct * prior * sq / static_cast<double>(ch.visits + 1);
}

// Default::Select (MonteCarlo.hpp:57-68) over a contiguous block of children: the first child with the highest score.
// "The first child whose score is above -1 and above all before it" is the first occurrence of the maximum (every
// score is >= -1), so the scan is a max reduction plus an equality search: no unpredictable branches, and four
// children per step with AVX2 (same IEEE operations in the same order as the scalar form, hence the same choice).
static int select_block_scalar(const float* v, const float* p, const std::int32_t* k, int n, double c_puct, double sq) {
    double score[BOARD_SIZE];
    double top = -1.0;
    for (int i = 0; i < n; ++i) {
        score[i] = static_cast<double>(v[i]) + c_puct * static_cast<double>(p[i]) * sq / static_cast<double>(k[i] + 1);
        top = score[i] > top ? score[i] : top;
    }
    for (int i = 0; i < n; ++i)
        if (score[i] == top) return i;
    return 0;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static int select_block_avx2(const float* v, const float* p, const std::int32_t* k, int n,
                                                             double c_puct, double sq) {
    alignas(32) double score[BOARD_SIZE + 4];
    const __m256d vc = _mm256_set1_pd(c_puct), vsq = _mm256_set1_pd(sq);
    __m256d vtop = _mm256_set1_pd(-1.0);
    const int n4 = n & ~3;
    for (int i = 0; i < n4; i += 4) {
        const __m256d val = _mm256_cvtps_pd(_mm_loadu_ps(v + i)), pr = _mm256_cvtps_pd(_mm_loadu_ps(p + i));
        const __m256d cnt = _mm256_cvtepi32_pd(_mm_add_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(k + i)), _mm_set1_epi32(1)));
        const __m256d sc = _mm256_add_pd(val, _mm256_div_pd(_mm256_mul_pd(_mm256_mul_pd(vc, pr), vsq), cnt));
        _mm256_store_pd(score + i, sc);
        vtop = _mm256_max_pd(vtop, sc);
    }
    alignas(32) double lanes[4];
    _mm256_store_pd(lanes, vtop);
    double top = lanes[0];
    for (int j = 1; j < 4; ++j) top = lanes[j] > top ? lanes[j] : top;
    for (int i = n4; i < n; ++i) {
        score[i] = static_cast<double>(v[i]) + c_puct * static_cast<double>(p[i]) * sq / static_cast<double>(k[i] + 1);
        top = score[i] > top ? score[i] : top;
    }
    const __m256d vt = _mm256_set1_pd(top);
    for (int i = 0; i < n4; i += 4) {
        const int hit = _mm256_movemask_pd(_mm256_cmp_pd(_mm256_load_pd(score + i), vt, _CMP_EQ_OQ));
        if (hit) return i + __builtin_ctz(static_cast<unsigned>(hit));
    }
    for (int i = n4; i < n; ++i)
        if (score[i] == top) return i;
    return 0;
}
#endif

static std::int32_t select_block(const Tree& t, std::int32_t first, int n, double c_puct, double sq) {
    if (first != t.rfirst || n != t.rn) {                            // a block below the root (eager mode only): plain loop
        std::int32_t best = first;
        double best_score = -1.0;
        const float prior = 1.0f / static_cast<float>(n);              // an eager block holds every legal move: n = n_moves
        for (std::int32_t c = first; c < first + n; ++c) {
            const double score = pucb(t.nodes[c], prior, c_puct, sq);
            if (score > best_score) { best_score = score; best = c; }
        }
        return best;
    }
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return first + select_block_avx2(t.rvalue.data(), t.rprior.data(), t.rvisits.data(), n, c_puct, sq);
#endif
    return first + select_block_scalar(t.rvalue.data(), t.rprior.data(), t.rvisits.data(), n, c_puct, sq);
}

// Default::Select on a lazily expanded node: the best materialised child, or -- when the common score of the
// never-visited children is strictly higher -- the lowest never-visited cell, created now.
static std::int32_t select_lazy(Tree& t, std::int32_t node, double c_puct) {
    const ANode parent = t.nodes[node];
    const double sq = std::sqrt(static_cast<double>(t.visits_of(node)));
    const float prior = 1.0f / static_cast<float>(parent.n_moves);
    std::int32_t best = parent.first_child;                          // as the reference: the first child unless one scores above -1
    double best_score = -1.0;
    for (std::int32_t c = parent.first_child; c >= 0; c = t.nodes[c].next_sibling) {   // increasing cell order
        const double score = pucb(t.nodes[c], prior, c_puct, sq);
        if (score > best_score) { best_score = score; best = c; }
    }
    if (parent.n_children < parent.n_moves) {                       // someone has never been visited: value 0, visits 0
        const double fresh = 0.0 + c_puct * prior * sq / 1.0;
        if (best < 0 || fresh > best_score) {
            int cell = parent.cursor + 1;
            while (t.board.cell(cell) != 0) ++cell;                 // the lowest empty cell above the cursor
            const std::int32_t created = t.add(node, cell, -parent.player);
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

// One playout's descent (MCTS::playout, MCTS.cpp:158-166): walk to a leaf, leave the board there, pack it.
static void select_leaf(Tree& t, double c_puct, std::uint32_t* packed) {
    std::int32_t node = 0;
    t.root_child = -1;
    while (t.nodes[node].n_moves > 0) {                              // expanded: descend
        const ANode& parent = t.nodes[node];
        node = parent.eager ? select_block(t, parent.first_child, parent.n_children, c_puct, std::sqrt(static_cast<double>(t.visits_of(node))))
                            : select_lazy(t, node, c_puct);
        if (t.root_child < 0) t.root_child = t.nodes[node].position;
        t.board.applyMove(Position(t.nodes[node].position), false);
    }
    t.leaf = node;
    t.terminal = t.board.checkGameEnd();                             // MCTS.cpp:166
    t.board.pack(packed);
}

// Expansion of the leaf + back-propagation of its simulated value (MCTS.cpp:167-176, MonteCarlo.hpp:90-95).
static void backup(Tree& t, const std::int32_t* r, const RootParallelConfig& cfg, std::size_t root_depth) {
    const float black_value = static_cast<float>(r[2] - r[0]) / static_cast<float>(cfg.c_rollouts);
    if (!t.terminal) {
        expand(t, t.leaf, t.board, cfg.eager || t.leaf == 0);          // the root always gets its block: its scan vectorises
        if (t.leaf == 0 && cfg.noise) {                              // Default::AddNoise on the root's fresh children
            const ANode& rt = t.nodes[0];
            std::gamma_distribution<float> gamma(0.05f, 1.0f);
            std::vector<float> g(rt.n_children);
            double n2 = 0;
            for (float& x : g) { x = gamma(t.rng); n2 += double(x) * x; }
            const float inv = n2 > 0 ? static_cast<float>(1.0 / std::sqrt(n2)) : 0.0f;
            for (int k = 0; k < rt.n_children; ++k) {
                float& pr = t.rprior[k];
                pr = pr * 0.75f + 0.25f * g[k] * inv;
            }
        }
    }
    if (t.root_child >= 0) { t.wins[t.root_child].black += r[2]; t.wins[t.root_child].white += r[0]; }
    float v = static_cast<float>(t.nodes[t.leaf].player) * black_value;   // value for who moved into the leaf
    for (std::int32_t n = t.leaf; n >= 0; n = t.nodes[n].parent, v = -v) {
        if (t.in_root_block(n)) {
            const std::int32_t k = n - t.rfirst;
            t.rvisits[k] += 1;
            t.rvalue[k] += (v - t.rvalue[k]) / static_cast<float>(t.rvisits[k]);
        } else {
            ANode& nd = t.nodes[n];
            nd.visits += 1;
            nd.value += (v - nd.value) / static_cast<float>(nd.visits);       // MonteCarlo.hpp:90-95
        }
    }
    t.board.revertMove(t.board.m_moveRecord.size() - root_depth);
}

std::vector<RootParallelSearch::DumpNode> RootParallelSearch::dumpTree(int tree) const {
    std::vector<DumpNode> out;
    if (tree < 0 || tree >= static_cast<int>(m->trees.size())) return out;
    const Tree& t = m->trees[tree];
    if (t.nodes.empty()) return out;
    std::vector<std::pair<std::int32_t, int>> stack{ { 0, 0 } };
    std::vector<std::int32_t> kids;
    while (!stack.empty()) {
        const auto [n, depth] = stack.back();
        stack.pop_back();
        const ANode& nd = t.nodes[n];
        const bool in_block = t.in_root_block(n);
        const float prior = in_block ? t.rprior[n - t.rfirst] : nd.parent < 0 ? 1.0f : 1.0f / static_cast<float>(t.nodes[nd.parent].n_moves);
        out.push_back({ nd.position, t.visits_of(n), in_block ? t.rvalue[n - t.rfirst] : nd.value, prior,
                        static_cast<std::int16_t>(depth), static_cast<std::int16_t>(nd.n_moves) });
        kids.clear();
        if (nd.eager) for (std::int32_t c = nd.first_child; c < nd.first_child + nd.n_children; ++c) kids.push_back(c);
        else for (std::int32_t c = nd.first_child; c >= 0; c = t.nodes[c].next_sibling) kids.push_back(c);
        for (auto it = kids.rbegin(); it != kids.rend(); ++it)
            if (t.visits_of(*it) > 0) stack.push_back({ *it, depth + 1 });
    }
    return out;
}

void RootParallelSearch::run(const Board& root, int playouts_per_tree, std::uint64_t seed) {
    m_cfg.seed = seed;
    run(root, playouts_per_tree);
}

void RootParallelSearch::run(const Board& root, int playouts_per_tree) {
    const auto t_start = std::chrono::steady_clock::now();
    m_stats.fill(0);
    seconds_gpu = 0;
    leaves = 0;
    const int n_trees = m_cfg.trees;
    const Player root_last = root.m_moveRecord.empty() ? Player::White : -root.m_curPlayer;   // MCTS.h:138-151
    const std::size_t reserve = m_cfg.eager ? static_cast<std::size_t>(std::max(playouts_per_tree, 0)) * 64 + 256
                                            : static_cast<std::size_t>(std::max(playouts_per_tree, 0)) * 2 + 256;
    {
        ANode* const slab = m->node_slab(static_cast<std::size_t>(n_trees), reserve);
        std::atomic<int> next{ 0 };                                  // fresh trees, set up by the whole team (first-touch allocation)
        m->team->run([&](int) {
            for (int i; (i = next.fetch_add(1, std::memory_order_relaxed)) < n_trees;) {
                Tree& t = m->trees[i];
                t.clear(reserve, slab ? slab + static_cast<std::size_t>(i) * reserve : nullptr);
                t.add(-1, root.m_moveRecord.empty() ? -1 : static_cast<int>(root.m_moveRecord.back()), static_cast<int>(root_last));   // MCTS.h:138-151
                t.board = root;
                t.wins.fill({ 0, 0 });
                t.rng.seed(static_cast<std::uint32_t>(m_cfg.seed * 2654435761u + static_cast<std::uint32_t>(m_cfg.replica_base + i)));
            }
        });
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

    // The trees are cut into G groups whose leaf batches are in flight at the same time: while the GPU simulates the
    // leaves of one group, the threads back up and descend the trees of the others.  "Visit" v handles round v / G of
    // group v % G: back up the group's previous batch, select the next leaves.  There is no driver thread: every thread
    // walks the visits in order and takes trees in chunks; whoever finishes a group's LAST tree submits its batch
    // (one launch, a few microseconds, while the others are already at the next visit), and whoever first needs a
    // batch that is in flight watches its counts arrive on behalf of all.  A tree's Philox stream is (round, global
    // tree index), whatever G and the thread count are.
    // (a group needs few trees to be worth a launch since a leaf batch is ONE fused launch: with 128 trees per rank -- the
    // benchmark's 1 024 trees over 8 GPUs -- eight groups of 16 keep eight round trips in flight)
    const int groups = m_cfg.groups > 0 ? std::min({ m_cfg.groups, kMaxGroups, n_trees })
                                        : std::max(1, std::min(kAutoGroups, n_trees / 16));
    std::array<int, kMaxGroups + 1> gs{};
    for (int g = 0; g <= groups; ++g) gs[g] = static_cast<int>(static_cast<long long>(n_trees) * g / groups);
    const long long visits_total = static_cast<long long>(playouts_per_tree + 1) * groups;
    // Trees are handed out in chunks from a per-group counter, so a thread that loses its core for a time slice
    // delays one chunk, not the whole group.  All counters only grow: round r of a group owns the chunk numbers
    // [r * chunks, (r + 1) * chunks), so a thread that arrives late can never claim work of a finished round.
    // A visit ends when its LAST chunk does, so a small group is cut into one chunk per thread (170 trees over 16 threads
    // in chunks of 8 would be 16 + 6 chunks: two waves for the work of 1.3); measured, finer chunks than that lose more
    // to the shared counter than they gain in balance.
    const int n_threads = m->team->size();
    const int share = (n_trees / groups + n_threads - 1) / n_threads;
    const int kChunk = share <= 16 ? std::max(1, share) : 8;
    struct alignas(64) GroupState {
        std::atomic<long long> claimed{ 0 }, done{ 0 };   // chunks handed out / trees finished, over all rounds
        std::atomic<int> submitted{ 0 }, arrived{ 0 };    // batches launched / batches whose results are in m->wdb
        std::atomic_flag waiting = ATOMIC_FLAG_INIT;      // someone is watching / waiting for this group's batch
    };
    std::array<GroupState, kMaxGroups> gst;
    std::atomic<bool> failed{ false };
    std::mutex error_mutex;
    std::string error;
    auto fail = [&](const char* what) {
        std::lock_guard<std::mutex> lock(error_mutex);
        if (error.empty()) error = std::string(what) + ": " + gk_last_error();
        failed.store(true);
    };
    std::vector<std::array<double, 8>> clock(n_threads);   // per thread (padded): seconds waiting for results / of it, as the group's watcher / inside gk_rollout_submit_host
    for (auto& c : clock) c.fill(0.0);
    auto seconds_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };

    // take chunks of visit v (whose results have arrived) until none is left: back up, select, pack; the thread that
    // finishes the group's last tree sends the leaves off
    auto process = [&](long long v, int id) {
        const int g = static_cast<int>(v % groups), round = static_cast<int>(v / groups);
        GroupState& G = gst[g];
        const int count = gs[g + 1] - gs[g], chunks = (count + kChunk - 1) / kChunk;
        const long long first = static_cast<long long>(round) * chunks, last = first + chunks;
        for (;;) {
            long long c = G.claimed.load(std::memory_order_relaxed);
            if (c >= last) break;
            if (!G.claimed.compare_exchange_weak(c, c + 1, std::memory_order_acq_rel)) continue;
            const int lo = gs[g] + static_cast<int>(c - first) * kChunk, hi = std::min(lo + kChunk, gs[g + 1]);
            for (int i = lo; i < hi; ++i) {
                Tree& t = m->trees[i];
                if (round > 0) backup(t, &m->wdb[static_cast<std::size_t>(i) * 3], m_cfg, root_depth);
                if (round < playouts_per_tree) select_leaf(t, c_puct, &m->packed[static_cast<std::size_t>(i) * 16]);
            }
            const long long finished = G.done.fetch_add(hi - lo, std::memory_order_acq_rel) + (hi - lo);
            if (finished == static_cast<long long>(round + 1) * count && round < playouts_per_tree) {
                const auto t0 = std::chrono::steady_clock::now();
                volatile std::int32_t* counts = m->wdb + static_cast<std::size_t>(gs[g]) * 3;      // every backup of the round has read them
                for (int k = 0; k < 3 * count; ++k) counts[k] = kCountPending;
                if (gk_rollout_submit_host(g, m->packed + static_cast<std::size_t>(gs[g]) * 16, count, m_cfg.c_rollouts, m_cfg.seed,
                                           static_cast<std::uint32_t>(round), m_cfg.replica_base + gs[g],
                                           m->wdb + static_cast<std::size_t>(gs[g]) * 3) != GK_OK)
                    fail("gk_rollout_submit_host");
                G.submitted.store(round + 1, std::memory_order_release);
                clock[id][2] += seconds_since(t0);
            }
        }
    };
    auto worker = [&](int id) {
        for (long long v = 0; v < visits_total && !failed.load(std::memory_order_relaxed); ++v) {
            const int g = static_cast<int>(v % groups), round = static_cast<int>(v / groups);
            GroupState& G = gst[g];
            if (round > 0 && G.arrived.load(std::memory_order_acquire) < round) {      // the batch of round - 1 is still out
                const auto t0 = std::chrono::steady_clock::now();
                Backoff wait;
                while (G.arrived.load(std::memory_order_acquire) < round && !failed.load(std::memory_order_relaxed)) {
                    if (G.submitted.load(std::memory_order_acquire) >= round && !G.waiting.test_and_set(std::memory_order_acquire)) {
                        if (G.arrived.load(std::memory_order_acquire) < round) {
                            // The kernel stores a leaf's three counts into the page-locked block as that leaf's CTA retires:
                            // watching them arrive saves the stream synchronisation's own latency (~3 us of a ~33 us round
                            // trip).  No progress for a while (or a failed launch): ask the stream, which reports the error.
                            const auto t1 = std::chrono::steady_clock::now();
                            volatile const std::int32_t* counts = m->wdb + static_cast<std::size_t>(gs[g]) * 3;
                            const int total = 3 * (gs[g + 1] - gs[g]);
                            int seen = 0;
                            for (int spins = 0; m_cfg.watch && spins < (1 << 22); ++spins) {
                                while (seen < total && counts[seen] != kCountPending) ++seen;
                                if (seen == total) break;
                                cpu_relax();
                            }
                            std::atomic_thread_fence(std::memory_order_acquire);
                            if (seen < total && gk_rollout_wait(g) != GK_OK) fail("gk_rollout_wait");
                            G.arrived.store(round, std::memory_order_release);
                            clock[id][1] += seconds_since(t1);
                        }
                        G.waiting.clear(std::memory_order_release);
                    } else {
                        wait();
                    }
                }
                clock[id][0] += seconds_since(t0);
            }
            process(v, id);
        }
    };
    m->team->run([&](int id) { worker(id); });
    for (int g = 0; g < groups; ++g)                              // the streams are idle by now; a launch error surfaces here at the latest
        if (gk_rollout_wait(g) != GK_OK && !failed.load()) fail("gk_rollout_wait");
    if (failed.load()) throw std::runtime_error(error);
    double idle = 0, t_sync = 0, t_submit = 0;
    for (const auto& c : clock) { idle += c[0]; t_sync += c[1]; t_submit += c[2]; }
    idle /= n_threads;                           // mean over the threads: time spent waiting for results
    const double t_workers = 0;
    seconds_gpu = idle;
    driver_seconds = { t_sync, t_workers, t_submit };
    leaves = static_cast<std::int64_t>(n_trees) * playouts_per_tree;
    // ---- root statistics: integers, summed over trees ---------------------------------------------------------------
    nodes = 0;
    for (Tree& t : m->trees) {
        nodes += static_cast<std::int64_t>(t.nodes.size());
        const ANode& rt = t.nodes[0];
        if (rt.eager) {
            for (std::int32_t c = rt.first_child, e = c + rt.n_children; c < e; ++c) m_stats[t.nodes[c].position] += t.visits_of(c);
        } else {
            for (std::int32_t c = rt.first_child; c >= 0; c = t.nodes[c].next_sibling) m_stats[t.nodes[c].position] += t.nodes[c].visits;
        }
        for (int c = 0; c < BOARD_SIZE; ++c) {
            m_stats[BOARD_SIZE + c] += t.wins[c].black;
            m_stats[2 * BOARD_SIZE + c] += t.wins[c].white;
        }
    }
    seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
}

}  // namespace gomoku

// root_parallel.h -- root-parallel MCTS driver (BASELINE config 4; new, the reference has no
// parallel search).  `trees` independent search trees start from the same root position; every
// round each tree descends to one leaf (PUCB), the leaves of all trees are simulated in
// ONE gk_rollout_batch call (c_rollouts playouts each, disjoint Philox counters), and the results
// are backed up.  Trees live in arenas (one contiguous child block per expansion) instead of the
// reference's make_unique per child (MonteCarlo.hpp:71-80).  The per-root-child statistics are
// integers, so summing them over trees, ranks and GPUs is order-independent: that sum is the only
// thing the multi-GPU path exchanges (one allreduce of int64[3][225] per move).
#pragma once
#include <array>
#include <cstdint>
#include <vector>

#include "game.h"

namespace gomoku {

struct RootParallelConfig {
    int trees = 256;            // independent trees on this rank
    int c_rollouts = 5;         // RandomPolicy::c_rollouts (policies/Random.h:15)
    double c_puct = 5.0;
    std::uint64_t seed = 1;     // Philox key
    int replica_base = 0;       // first global tree index of this rank (keeps streams disjoint across ranks)
    int threads = 0;            // host threads, all doing tree work; whoever finishes a group's last tree launches its batch (0 = hardware concurrency)
    // Dirichlet noise on the priors of the root's fresh children.  An EXTENSION, off by default: the reference's
    // Default::AddNoise runs before the playouts and only touches children that already exist (MonteCarlo.hpp:97-108,
    // MCTS.cpp:182), so a search from a fresh tree -- which is what every root-parallel move is -- draws none.
    bool noise = false;
    int groups = 0;             // leaf batches in flight, 1..8 (0 = chosen from the tree count); never changes a result
    bool watch = true;          // see a batch's counts arrive in the page-locked block instead of synchronising its stream
    bool eager = false;         // materialise every child at expansion like the reference (slow; kept to test the lazy tree against)
};

class RootParallelSearch {
public:
    using Stats = std::array<std::int64_t, 3 * BOARD_SIZE>;   // [0] visits, [1] black-won rollouts, [2] white-won rollouts, per root child

    explicit RootParallelSearch(const RootParallelConfig& cfg);
    ~RootParallelSearch();

    // fresh trees from `root`, `playouts_per_tree` playouts on each; fills stats()
    void run(const Board& root, int playouts_per_tree);
    void run(const Board& root, int playouts_per_tree, std::uint64_t seed);   // same, with a new Philox key / noise seed (the object and its arenas are reused)

    const Stats& stats() const { return m_stats; }
    // Tree `tree` of the last run in pre-order, visited nodes only (children in increasing cell order, as the reference
    // expands them): the move into the node, its visits, value, prior, depth and number of legal moves once expanded.
    // For tests that compare a tree node for node with the reference's MCTS under the same playout stream.
    struct DumpNode { std::int16_t position; std::int32_t visits; float value, prior; std::int16_t depth; std::int32_t n_moves; };
    std::vector<DumpNode> dumpTree(int tree) const;
    static Position bestMove(const Stats& stats);             // most visited root child, ties -> lowest cell (MCTS.cpp:129-134)

    double seconds_total = 0, seconds_gpu = 0;                // wall clock of the last run / of it, the time a thread waited for GPU results (mean over threads)
    std::array<double, 3> driver_seconds{};                   // summed over threads: watching / waiting for a batch as its group's watcher, (unused, 0), inside gk_rollout_submit_host
    std::int64_t leaves = 0, nodes = 0;

private:
    struct Impl;
    Impl* m;
    RootParallelConfig m_cfg;
    Stats m_stats{};
};

}  // namespace gomoku

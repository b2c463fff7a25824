"""The reference-facing plugin surface with its simulate slot on the GPU: MCTS + RandomPolicy
(config 1 semantics) and the root-parallel driver (config 4).  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core(gpu):
    from gomokuai_b200 import build
    build.build_pyext()
    from gomokuai_b200 import core
    return core


def test_random_policy_mcts_plays_a_legal_move_and_restores_the_board(core):
    # core/test/unit/mcts_unittest.cpp:14-42 (the reference's disabled invariance test)
    b = core.Board()
    for c in (112, 113, 97):
        b.apply_move(c)
    snapshot = list(map(int, b.move_record))
    m = core.MCTS(c_iterations=400, policy=core.RandomPolicy(5.0, 5))
    q, pi = m.policy.eval_state(b)
    assert list(map(int, b.move_record)) == snapshot and -1.0 <= q <= 1.0 and abs(sum(pi) - 1.0) < 1e-4
    move = m.get_action(b)
    assert list(map(int, b.move_record)) == snapshot and b.check_move(move) and m.iterations == 400
    assert m.size > 400


def test_split_simulate_builds_the_same_tree_as_the_synchronous_slot(core):
    """MCTS::playout expands the leaf while RandomPolicy's rollouts are in flight (simulateBegin / simulateEnd); a policy
    whose eval_state slot is a user function takes the synchronous path.  Same seed -> same Philox stream per call ->
    the same tree."""
    b = core.Board()
    for c in (112, 113, 97, 98):
        b.apply_move(c)

    def search(policy):
        core.seed(123)
        m = core.MCTS(c_iterations=600, policy=policy)
        move = m.get_action(b)
        kids = sorted((int(ch.position), int(ch.node_visits), float(ch.state_value)) for ch in m.root.children)
        return int(move), int(m.size), kids

    rp = core.RandomPolicy(5.0, 5)
    split = search(rp)
    inner = core.RandomPolicy(5.0, 5)
    sync = search(core.Policy(eval_state=lambda board: inner.eval_state(board), c_puct=5.0))
    assert split == sync
    assert split == search(rp)                          # and the split path is repeatable under the seed


def test_random_policy_value_tracks_the_rollout_kernel(core, gpu):
    b = core.Board()
    for c in (112, 0, 113, 1, 114, 2, 115, 30):        # black to move with an open four: black wins most rollouts
        b.apply_move(c)
    vals = [core.RandomPolicy(5.0, 64).eval_state(b)[0] for _ in range(8)]
    wdb = gpu.rollout_batch(b.packed().reshape(1, 16), 4096)["wdb"].cpu().numpy()[0]
    expect = (wdb[2] - wdb[0]) / 4096.0                 # side to move is black
    assert abs(np.mean(vals) - expect) < 0.15 and np.mean(vals) > 0.2


def test_root_parallel_search_statistics(core):
    b = core.Board()
    for c in (112, 113, 97, 98):
        b.apply_move(c)
    s = core.RootParallelSearch(trees=64, c_rollouts=5, seed=7, threads=4)
    stats = s.run(b, 40)
    assert stats.shape == (3, 225) and stats.dtype == np.int64
    assert stats[0].sum() == 64 * (40 - 1)              # the first playout of every tree evaluates the root itself
    occupied = [112, 113, 97, 98]
    assert (stats[:, occupied] == 0).all()
    assert ((stats[1] + stats[2]) <= stats[0] * 5).all()
    again = core.RootParallelSearch(trees=64, c_rollouts=5, seed=7, threads=2).run(b, 40)
    assert np.array_equal(stats, again)                 # seeded: independent of the host thread count
    other = core.RootParallelSearch(trees=64, c_rollouts=5, seed=8).run(b, 40)
    assert not np.array_equal(stats, other)
    assert s.leaves == 64 * 40 and s.nodes > 64 * 200


def test_lazy_tree_equals_the_eager_tree(core):
    """children materialised on first selection (the arena design) visit exactly the nodes of the reference's
    expand-everything tree: same seeds -> identical root statistics, with and without root noise"""
    b = core.Board()
    for c in (112, 113, 97, 98, 128):
        b.apply_move(c)
    for noise in (True, False):
        lazy = core.RootParallelSearch(trees=48, c_rollouts=5, seed=3, threads=3, noise=noise)
        eager = core.RootParallelSearch(trees=48, c_rollouts=5, seed=3, threads=3, noise=noise, eager=True)
        a, e = lazy.run(b, 300), eager.run(b, 300)
        assert np.array_equal(a, e)
        assert lazy.nodes * 20 < eager.nodes                # ~1 node per playout instead of ~220


def test_root_parallel_search_finds_the_winning_move(core):
    from gomokuai_b200 import root_parallel as rp
    b = core.Board()
    for c in (112, 0, 113, 1, 114, 2, 115, 30):        # black completes five at 111 or 116
        b.apply_move(c)
    # With uniform priors and no noise the reference's own search (fresh root: Default::AddNoise draws nothing) keeps
    # exploiting the first cells it tried -- black wins most random playouts from here, so a visited child scores far
    # above every unvisited one and the search never reaches cell 111.  The root-noise extension spreads the trees.
    move, stats, _ = rp.search(b, playouts_total=32 * 1500, trees_per_rank=32, c_puct=1.0, noise=True)
    assert move in (111, 116)
    # white to move must block: after black's four is open on both sides every move loses, so just check legality
    b.apply_move(30 + 15)
    move, stats, _ = rp.search(b, playouts_total=128 * 20, trees_per_rank=128)
    assert b.check_move(move)


def test_decided_root_returns_empty_statistics(core, kats):
    b = core.Board()
    for x, y in kats["black_win"]:
        b.apply_move(core.Position(x, y))
    stats = core.RootParallelSearch(trees=8).run(b, 10)
    assert stats.sum() == 0


def test_traditional_policy_simulate_is_hybrid_simulate(core, gpu):
    """TraditionalPolicy.eval_state == TraditionalPolicy::hybridSimulate (Traditional.h:49-69) of the oracle"""
    from oracle import pyoracle
    port = pyoracle.port()
    moves = [112, 113, 97, 98, 127, 128, 82]
    b = core.Board()
    for c in moves:
        b.apply_move(c)
    value, probs = core.TraditionalPolicy().eval_state(b)
    rv, rp = pyoracle.hybrid_simulate(port, moves)
    assert abs(value - float(rv)) <= 2e-5 and np.allclose(np.asarray(probs), rp, rtol=2e-5, atol=2e-6)
    assert list(map(int, b.move_record)) == moves


def test_traditional_policy_mcts_blocks_an_open_four(core):
    b = core.Board()
    for c in (112, 0, 113, 1, 114, 30, 115, 31):        # black has four in a row (112..115), black to move: 111 or 116 wins
        b.apply_move(c)
    m = core.MCTS(c_iterations=200, policy=core.TraditionalPolicy())
    move = int(m.get_action(b))
    assert move in (111, 116)
    b2 = core.Board()
    for c in (112, 0, 113, 1, 114, 30, 45):             # white to move must answer the open three / four threat on row 7
        b2.apply_move(c)
    m2 = core.MCTS(c_iterations=200, policy=core.TraditionalPolicy())
    assert int(m2.get_action(b2)) in (110, 111, 115, 116)


def test_root_allreduce_through_the_c_abi(gpu):
    """gk_root_allreduce on the library's own NCCL communicator (libnccl resolved at run time).  One rank:
    the sum over a world of one is the identity; the 2-GPU run is scripts/bench_root_parallel.py."""
    import torch
    uid = gpu.nccl_unique_id()
    assert len(uid) == 128
    gpu.nccl_init(uid, 1, 0)
    try:
        with pytest.raises(gpu.GomokuB200Error):
            gpu.nccl_init(uid, 1, 0)                     # a second communicator is refused
        rng = np.random.default_rng(3)
        stats = rng.integers(0, 1 << 40, size=(3, 225)).astype(np.int64)
        t = gpu.root_allreduce(torch.from_numpy(stats).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), stats)
        with pytest.raises(gpu.GomokuB200Error):
            gpu.root_allreduce(torch.zeros(10, dtype=torch.int64, device="cuda"))
    finally:
        gpu.nccl_shutdown()


def test_root_parallel_pipeline_is_invisible_in_the_statistics(core):
    """Several tree groups in flight, chunked work distribution, whichever thread launching and awaiting the batches,
    counts watched as they arrive or streams synchronised: none of it may show in the result -- a tree's Philox stream
    is (round, global tree index), its noise stream is seeded per tree."""
    b = core.Board()
    for c in (112, 113, 97):
        b.apply_move(c)
    ref = core.RootParallelSearch(trees=300, c_rollouts=5, seed=21, threads=2).run(b, 60)      # 8 groups of 37 / 38 trees
    assert ref[0].sum() == 300 * 59
    for threads, groups, watch in ((3, 0, True), (7, 0, False), (1, 1, True), (5, 3, False), (16, 8, True)):
        s = core.RootParallelSearch(trees=300, c_rollouts=5, seed=21, threads=threads, groups=groups, watch=watch)
        assert np.array_equal(s.run(b, 60), ref), (threads, groups, watch)
    s = core.RootParallelSearch(trees=300, c_rollouts=5, seed=5, threads=4)
    first = s.run(b, 30)
    assert np.array_equal(s.run(b, 60, 21), ref)                    # the searcher is reusable and reseedable
    assert not np.array_equal(first[0], ref[0])
    # replica_base shifts the tree indices: two ranks' halves are disjoint streams of one 600-tree search
    lo = core.RootParallelSearch(trees=300, c_rollouts=5, seed=21, threads=4, replica_base=0).run(b, 20)
    hi = core.RootParallelSearch(trees=300, c_rollouts=5, seed=21, threads=4, replica_base=300).run(b, 20)
    assert not np.array_equal(lo, hi) and (lo + hi)[0].sum() == 600 * 19
    whole = core.RootParallelSearch(trees=600, c_rollouts=5, seed=21, threads=4).run(b, 20)
    assert np.array_equal(lo + hi, whole)                           # what the allreduce of two ranks yields = one rank with all the trees

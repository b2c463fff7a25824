"""Rows f2 / b of SURVEY.md section 8: the host tree engines against the reference's OWN search.

oracle/_ref holds the reference's MCTS.cpp, MonteCarlo.hpp, Statistical.hpp and policies/*.h compiled unmodified
(oracle/Makefile, oracle/ref_harness_search.cpp).  A search is made deterministic through the reference's own plugin
mechanism -- an injected `simulate` slot (MCTS.h:86-92) -- and the trees are then compared NODE FOR NODE (move, visits,
running-mean value, prior, depth, number of children, in pre-order):

  * the CorePyExt mirror (csrc/host/mcts.cpp) with the same evaluator filled in from Python, over several consecutive
    moves: tree reuse (stepForward), Dirichlet root noise (same engine type, same seed), select / expand / back-propagate;
  * (GPU) the lazy-arena root-parallel driver (csrc/host/root_parallel.cpp) whose leaves are simulated by the rollout
    kernel, against the reference's MCTS whose simulate slot plays the SAME Philox-driven playouts on the reference Board.
"""
import numpy as np
import pytest

from search_util import assert_same_tree, dump_tree, injected_eval_state


@pytest.fixture(scope="module")
def core(gk):
    from gomokuai_b200 import build
    build.build_pyext()
    from gomokuai_b200 import core
    return core


def _board(core, moves):
    b = core.Board()
    for c in moves:
        b.apply_move(c)
    return b


@pytest.mark.parametrize("moves,iterations", [([], 1500), ([112, 113, 97], 1200), ([112, 0, 113, 1, 114, 2, 115, 30], 900)])
def test_mirror_tree_equals_reference_tree(core, ref, moves, iterations):
    """one search from a fresh tree (no noise: the root has no children yet, MonteCarlo.hpp:97-108)"""
    want = ref.mcts_injected(moves, iterations, n_searches=0, sim_kind=0)
    b = _board(core, moves)
    m = core.MCTS(c_iterations=iterations, policy=core.Policy(eval_state=injected_eval_state),
                  **({"last_move": core.Position(moves[-1]), "last_player": -b.status["cur_player"]} if moves else {}))
    m.eval_state(b)                                     # runPlayouts without the step
    assert_same_tree(dump_tree(m.root), want, f"from {moves}")
    assert list(map(int, b.move_record)) == moves


@pytest.mark.parametrize("seed", [1, 12345])
def test_mirror_equals_reference_over_consecutive_moves(core, ref, seed):
    """get_action three times (tree reuse + seeded Dirichlet noise on the reused root), then the fourth tree"""
    moves, iterations = [112, 98], 700
    want = ref.mcts_injected(moves, iterations, n_searches=3, sim_kind=0, noise_seed=seed)
    core.seed(seed)
    b = _board(core, moves)
    m = core.MCTS(c_iterations=iterations, last_move=core.Position(moves[-1]), last_player=-b.status["cur_player"],
                  policy=core.Policy(eval_state=injected_eval_state))
    actions = []
    for _ in range(3):
        a = m.get_action(b)
        actions.append(int(a))
        b.apply_move(a)
    assert actions == want["actions"]
    m.eval_state(b)
    assert_same_tree(dump_tree(m.root), want, f"after {actions}")


def test_eval_state_probabilities_equal_reference(core, ref):
    """MCTS::evalState's post-processing (MCTS.cpp:104-117, Statistical.hpp:37-42): visits -> normalized -> +1 ->
    TempBasedProbs, for both temperature regimes (T = 1 below 15 moves, 1e-2 after)"""
    for moves in ([112, 113], [112, 113, 97, 98, 127, 128, 83, 82, 67, 68, 52, 54, 37, 39, 23, 24]):
        b = _board(core, moves)
        m = core.MCTS(c_iterations=600, last_move=core.Position(moves[-1]), last_player=-b.status["cur_player"],
                      policy=core.Policy(eval_state=injected_eval_state))
        q, pi = m.eval_state(b)
        visits = np.zeros(225, np.float32)
        for ch in m.root.children:
            visits[int(ch.position)] = ch.node_visits
        want = ref.temp_based_probs(visits, len(moves))
        assert np.array_equal(np.asarray(pi, np.float32), want), np.abs(np.asarray(pi) - want).max()
        assert q == m.root.state_value


# ---- GPU: the root-parallel arena driver and the GPU-backed policies ----------------------------------------------
@pytest.fixture(scope="module")
def gcore(gpu, core):
    return core


@pytest.mark.gpu
@pytest.mark.parametrize("moves", [[], [112, 113, 97, 98, 128]])
def test_root_parallel_trees_equal_reference_mcts(gcore, ref, moves):
    """Every arena tree == the reference's MCTS::playout loop whose simulate slot is RandomPolicy::averagedSimulate
    (Random.h:22-35) driven by the kernel's own Philox stream on the reference Board: (tree index, playout index) ->
    (position, ctr_hi) of include/gomoku_b200.h."""
    key, playouts, trees = 77, 900, 6
    b = _board(gcore, moves)
    s = gcore.RootParallelSearch(trees=trees, c_rollouts=5, seed=key, threads=3, noise=False, replica_base=40)
    stats = s.run(b, playouts)
    total_visits = np.zeros(225, np.int64)
    for t in range(trees):
        want = ref.mcts_injected(moves, playouts, n_searches=0, sim_kind=1, key=key, tree=40 + t, c_rollouts=5)
        assert_same_tree(s.tree_dump(t), want, f"tree {t} from {moves}")
        d1 = want["depth"] == 1
        np.add.at(total_visits, want["pos"][d1].astype(int), want["visits"][d1])
    assert np.array_equal(stats[0], total_visits)           # what the allreduce sums = the reference trees' root visits


@pytest.mark.gpu
def test_rollout_trace_equals_reference_board_playout(gpu, ref):
    """gk_rollout_trace_host: winners, lengths and move lists of the kernel's playouts == the reference Board driven by the
    same Philox stream; the moves replay legally to the same result"""
    port = __import__("oracle.pyoracle", fromlist=["x"]).port()
    from conftest import random_positions
    for i, m in enumerate([[], [112, 113]] + random_positions(5, 6, lo=4, hi=80)):
        if port.eval_moves(m)["winner"] != 0:
            continue
        mv, st = __import__("oracle.pyoracle", fromlist=["x"]).pack_moves([m])
        tr = gpu.rollout_trace_host(gpu.pack_moves(mv, st)[0], 24, key=9, ctr_hi=3, pos=100 + i)
        wn, ln, wdb = port.rollout_philox_batch(mv, st, 24, 9, ctr_hi=3, pos_base=100 + i)
        assert np.array_equal(tr["winners"], wn[0]) and np.array_equal(tr["lengths"], ln[0])
        rw, _ = ref.rollout_philox(m, 24, 9, ctr_hi=3, position=100 + i)
        assert np.array_equal(rw, wdb[0])
        for r in range(24):
            played = tr["moves"][r, :tr["lengths"][r]].tolist()
            res = port.board_play(list(m) + played)
            assert res["applied"] == len(m) + len(played) and res["winner"] == tr["winners"][r] and res["cur_player"] == 0


@pytest.mark.gpu
def test_poolrave_policy_search(gcore):
    """PoolRAVEPolicy (policies/PoolRAVE.h): AMAF nodes, the playout's final position on the board during back-propagation,
    the board restored afterwards; finds the immediate win"""
    gcore.seed(5)
    b = _board(gcore, [112, 0, 113, 1, 114, 2, 115, 30])        # black to move: 111 or 116 completes five
    snapshot = list(map(int, b.move_record))
    policy = gcore.PoolRAVEPolicy(2.0, 0.0)
    q, pi = policy.eval_state(b)                                 # defaultSimulate leaves the playout on the board (PoolRAVE.h:27-48)
    assert b.status["is_end"] and len(b.move_record) > len(snapshot) and q in (-1.0, 0.0, 1.0)
    assert abs(float(np.sum(pi)) - 1.0) < 1e-4
    b.revert_move(len(b.move_record) - len(snapshot))
    m = gcore.MCTS(c_iterations=3000, last_move=gcore.Position(30), last_player=gcore.Player.white, policy=policy)
    move = int(m.get_action(b))
    assert list(map(int, b.move_record)) == snapshot and not b.status["is_end"]
    assert move in (111, 116), move
    assert "PoolRAVEPolicy" in repr(policy)


@pytest.mark.gpu
def test_traditional_policy_with_rave_constructs_and_searches(gcore):
    b = _board(gcore, [112, 0, 113, 1, 114, 30, 115, 31])
    m = gcore.MCTS(c_iterations=150, policy=gcore.TraditionalPolicy(5.0, 0.1, True))   # policy_ext.hpp:39-45
    assert int(m.get_action(b)) in (111, 116)

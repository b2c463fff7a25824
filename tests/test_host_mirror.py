"""Host-side mirror of the reference's Board / Policy / MCTS surface (CorePyExt).  CPU only: the
search runs with Python-filled Policy slots, exactly how agents/alphazero.py plugs into the
reference (core/py_ext/src/mcts_ext.hpp:43-61), so no GPU slot is exercised here."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def core(gk):
    from gomokuai_b200 import build
    build.build_pyext()
    from gomokuai_b200 import core
    return core


def test_surface_names(core):
    # core/__init__.py:2-4 and game_ext.hpp / mcts_ext.hpp / policy_ext.hpp
    assert core.GameConfig == {"width": 15, "height": 15, "board_size": 225, "max_renju": 5}
    for cls, names in {
        core.Board: ["apply_move", "revert_move", "random_move", "check_move", "check_end", "reset", "move_record",
                     "last_move", "move_counts", "move_states", "status", "encoded_states"],
        core.Node: ["parent", "position", "player", "state_value", "action_prob", "node_visits", "children", "is_leaf", "is_full"],
        core.Policy: ["prepare", "clean_up", "apply_move", "revert_move", "check_game_end", "create_node", "select", "expand",
                      "eval_state", "back_prop"],
        core.MCTS: ["size", "iterations", "duration", "root", "policy", "get_action", "eval_state", "step_forward",
                    "sync_with_board", "reset"],
    }.items():
        for n in names:
            assert hasattr(cls, n), (cls, n)
    assert float(core.Player.white) == -1.0 and -core.Player.black == core.Player.white
    assert core.Player.calc_score(core.Player.black, core.Player.black) == 1.0
    p = core.Position(3, 4)
    assert (p.id, p.x, p.y, int(p), list(p), len(p)) == (63, 3, 4, 63, [3, 4], 2)


def test_board_matches_oracle_sequences(core, port, kats):
    # core/test/integration/board_integrationtest.cpp:66-155
    for key, want in (("black_win", core.Player.black), ("white_win", core.Player.white)):
        b = core.Board()
        cur = core.Player.black
        for x, y in kats[key]:
            assert b.apply_move(core.Position(x, y)) != cur
            cur = -cur
        assert b.status["is_end"] and b.status["winner"] == want
        o = port.board_play([y * 15 + x for x, y in kats[key]])
        assert int(b.status["winner"]) == o["winner"]
        for _ in kats[key]:
            b.revert_move()
        assert b.move_counts[core.Player.none] == 225 and b.status["cur_player"] == core.Player.black
    b = core.Board()
    for y in kats["tie_row_order"]:
        for x in range(15):
            r = b.apply_move(core.Position(x, y))
    assert r == core.Player.none and b.status["winner"] == core.Player.none
    with pytest.raises(Exception):
        b.random_move()


def test_board_random_game_and_invalid_moves(core, port):
    rng = np.random.default_rng(0)
    for _ in range(20):
        b = core.Board()
        moves = []
        while not b.status["is_end"]:
            mv = b.random_move()
            before = b.status["cur_player"]
            assert b.apply_move(mv) != before
            moves.append(int(mv))
            if not b.status["is_end"]:
                assert b.apply_move(mv) == b.status["cur_player"]      # replaying the cell is a no-op
        o = port.board_play(moves)
        assert (int(b.status["winner"]), len(b.move_record)) == (o["winner"], o["applied"])
        st = b.move_states
        assert st[core.Player.black].sum() + st[core.Player.white].sum() == len(moves)
        enc = b.encoded_states()
        assert enc.shape == (6, 15, 15) and enc[3].sum() == 1 and enc[4].sum() == 1


def test_encoded_states_and_packed(core, gk):
    b = core.Board()
    for c in (112, 113, 97, 98, 82):
        b.apply_move(c)
    enc = b.encoded_states()                                          # game_ext.hpp:87-104
    cur_is_black = b.status["cur_player"] == core.Player.black
    assert enc[5].all() == cur_is_black
    mine = core.Player.black if cur_is_black else core.Player.white
    assert np.array_equal(enc[0], b.move_states[mine]) and np.array_equal(enc[2], b.move_states[core.Player.none])
    assert enc[3].reshape(-1)[82] == 1 and enc[4].reshape(-1)[98] == 1
    from oracle import pyoracle as po
    mv, st = po.pack_moves([[112, 113, 97, 98, 82]])
    assert np.array_equal(b.packed(), gk.pack_moves(mv, st)[0])


def test_mcts_with_python_slots(core):
    """All four slots filled from Python: playout / select / expand / backprop semantics of SURVEY A.6."""
    calls = {"n": 0}

    def eval_state(board):
        calls["n"] += 1
        probs = np.zeros(225, np.float32)
        empties = [i for i in range(225) if board.check_move(i)]
        probs[empties] = 1.0 / len(empties)
        return 0.0, probs

    policy = core.Policy(eval_state=eval_state, c_puct=5.0)
    b = core.Board()
    b.apply_move(112)
    m = core.MCTS(c_iterations=300, policy=policy)
    snapshot = (list(map(int, b.move_record)), b.status["cur_player"])
    move = m.get_action(b)
    assert (list(map(int, b.move_record)), b.status["cur_player"]) == snapshot      # board restored (MCTS.cpp:175,197)
    assert b.check_move(move) and calls["n"] > 0 and m.iterations == 300
    root = m.root                                                                    # after get_action the chosen child is the root
    assert int(root.position) == int(move) and root.player == core.Player.white and root.parent is None
    # value 0 everywhere + uniform priors => PUCB visits every root child once before any twice
    m2 = core.MCTS(c_iterations=1 + 224 + 10, policy=core.Policy(eval_state=eval_state))
    q, pi = m2.eval_state(b)
    assert pi.shape == (225,) and abs(pi.sum() - 1.0) < 1e-4 and pi[112] == 0
    visits = sorted(c.node_visits for c in m2.root.children)
    assert len(visits) == 224 and visits[0] >= 1 and sum(visits) == 1 + 224 + 10 - 1 and visits[-1] <= 2


def test_mcts_finds_forced_win_with_exact_evaluator(core):
    """A leaf evaluator that knows five-in-a-row makes the search pick the winning move."""
    def eval_state(board):
        probs = np.zeros(225, np.float32)
        empties = [i for i in range(225) if board.check_move(i)]
        probs[empties] = 1.0 / len(empties)
        return 0.0, probs
    b = core.Board()
    for c in (112, 0, 113, 1, 114, 2, 115, 30):        # black: four in a row on row 7, (4,7) and (8,7) open
        b.apply_move(c)
    m = core.MCTS(c_iterations=3000, policy=core.Policy(eval_state=eval_state, c_puct=1.0))
    move = int(m.get_action(b))
    assert move in (111, 116)                          # terminal leaves are scored by CalcScore (MCTS.cpp:170-173)


def test_root_stats_allreduce_single_process(gk):
    from gomokuai_b200 import root_parallel as rp
    s = np.arange(675, dtype=np.int64).reshape(3, 225)
    assert np.array_equal(rp.allreduce_root_stats(s), s)
    assert rp.best_move(s) == 224 and rp.best_move(np.zeros((3, 225), np.int64)) == -1

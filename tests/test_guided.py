"""Pattern-guided playouts (BASELINE config 5): Heuristic::EvaluatedRollout semantics
(include/algorithms/Heuristic.hpp:61-91) with every game played to its end inside one kernel.

Parity statement.  The moves are chosen from floating-point probabilities (tests/test_heads.py states their
tolerance), so a trajectory can legitimately differ from the reference where two candidate cells are closer
than that tolerance.  The tests therefore require:
  * every game is LEGAL and correctly adjudicated (replayed through the oracle's Board);
  * arg-max mode: >= 97 % of the games are move-for-move identical to Heuristic::MaxEvaluatedRollout of the
    COMPILED reference (oracle/_ref, `ref.guided_rollout_max`; the numpy restatement stands in only where
    oracle/_ref was not built), and to the committed games in tests/golden/reference_heads.json;
  * sampled mode: the reference draws with std::discrete_distribution over its process-global mt19937
    (Board::getRandomMove(probs), Game.cpp:75-78) -- an implementation-defined stream -- so parity is statistical
    (SURVEY 8d): per position, the first-move distribution and the black-win rate of whole games must agree with
    the reference's own draws within 4.5 sigma of the binomial difference, |p1 - p2| <= 4.5 sqrt(p(1-p)(1/n1+1/n2)).
    Against the documented Philox draw (include/gomoku_b200.h) the restatement is exact, move for move."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_positions
from oracle import pyoracle

KEY = 0x474F4D4F4B5531


def _starts(seed, n):
    port = pyoracle.port()
    lists = [[], [112], [0], [224, 210]] + random_positions(seed, n, lo=2, hi=60)
    return [m for m in lists if port.eval_moves(m)["winner"] == 0]


def test_oracle_guided_rollout_is_a_legal_game(port):
    for mode in ("max", "sample"):
        for g, m in enumerate(_starts(4, 12)):
            winner, played = pyoracle.guided_rollout(port, port, m, mode, KEY, g)
            full = list(m) + played
            assert len(set(full)) == len(full) and all(0 <= c < 225 for c in full)
            r = port.board_play(full)
            assert r["applied"] == len(full) and r["winner"] == winner
            assert winner != 0 or len(full) == 225 or len(played) == 0 or port.eval_moves(full)["winner"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["max", "sample"])
def test_gpu_guided_rollouts_vs_oracle(gpu, mode):
    port, ref = pyoracle.port(), pyoracle.ref()
    lists = _starts(31, 160)
    mv, st = pyoracle.pack_moves(lists)
    out = gpu.guided_rollout_batch(gpu.pack_moves(mv, st), mode=mode, key=KEY, game_base=1000)
    winner, length, moves = out["winner"].cpu().numpy(), out["length"].cpu().numpy(), out["moves"].cpu().numpy()
    final = out["final_boards"].cpu().numpy().view(np.uint32)
    same = 0
    for i, m in enumerate(lists):
        played = moves[i, :length[i]].tolist()
        full = list(m) + played
        assert (moves[i, length[i]:] == -1).all()
        assert len(set(full)) == len(full)                              # legal: no cell twice
        r = port.board_play(full)                                       # adjudication by the oracle's Board
        assert r["applied"] == len(full) and r["winner"] == winner[i], (i, full)
        if winner[i] == 0:
            assert len(full) == 225 or port.eval_moves(full)["winner"] == 0
        fm, fs = pyoracle.pack_moves([full])
        assert np.array_equal(gpu.pack_moves(fm, fs)[0], final[i])
        if mode == "max" and ref is not None:
            w, p = ref.guided_rollout_max(m)                            # the reference's own MaxEvaluatedRollout
        else:
            w, p = pyoracle.guided_rollout(port, port, m, mode, KEY, 1000 + i)
        same += int(w == winner[i] and p == played)
    print(f"guided[{mode}]: {same} of {len(lists)} games identical to", "the compiled reference" if mode == "max" and ref is not None else "the restatement")
    assert same >= 0.97 * len(lists), f"{same} of {len(lists)} games identical to the oracle"


def test_restated_max_rollout_equals_the_compiled_reference(port, ref):
    same = total = 0
    for m in _starts(9, 40):
        total += 1
        same += ref.guided_rollout_max(m) == pyoracle.guided_rollout(port, port, m, "max", KEY, 0)
    assert same >= 0.97 * total, (same, total)


@pytest.mark.gpu
def test_gpu_guided_max_golden(gpu):
    """the committed games of the compiled reference's MaxEvaluatedRollout"""
    with open(os.path.join(GOLDEN, "reference_heads.json")) as f:
        games = json.load(f)["guided_max"]
    mv, st = pyoracle.pack_moves([g["moves"] for g in games])
    out = gpu.guided_rollout_batch(gpu.pack_moves(mv, st), mode="max", key=KEY)
    winner, length, moves = out["winner"].cpu().numpy(), out["length"].cpu().numpy(), out["moves"].cpu().numpy()
    same = sum(int(winner[i]) == g["winner"] and moves[i, :length[i]].tolist() == g["played"] for i, g in enumerate(games))
    assert same >= len(games) - 1, f"{same} of {len(games)} games identical to the reference's"


def _binomial_gate(k1, n1, k2, n2, sigmas=4.5):
    p = (k1 + k2) / float(n1 + n2)
    return abs(k1 / n1 - k2 / n2) <= sigmas * np.sqrt(max(p * (1 - p), 1e-12) * (1.0 / n1 + 1.0 / n2)) + 1e-9


@pytest.mark.gpu
def test_gpu_sampled_first_move_distribution_vs_reference(gpu, ref):
    """RandomEvaluatedRollout's move draw: GPU first moves of N games from one position vs N draws of the reference's
    Board::getRandomMove(EvaluationProbs) -- every cell within the 4.5 sigma binomial tolerance."""
    n = 20000
    lists = _starts(77, 24)[:20]
    for i, m in enumerate(lists):
        mv, st = pyoracle.pack_moves([m])
        boards = np.repeat(gpu.pack_moves(mv, st), n, axis=0)
        out = gpu.guided_rollout_batch(boards, mode="sample", key=KEY + i, max_moves=1)
        first = out["moves"].cpu().numpy()[:, 0]
        assert (out["length"].cpu().numpy() == 1).all()
        g = np.bincount(first, minlength=225)
        r = ref.sampled_first_move(m, n)
        assert g[np.array(m, int)].sum() == 0 and r[np.array(m, int)].sum() == 0 if m else True
        bad = [c for c in range(225) if not _binomial_gate(int(g[c]), n, int(r[c]), n)]
        assert not bad, (i, m, bad, g[bad], r[bad])


@pytest.mark.gpu
def test_gpu_sampled_games_win_rate_vs_reference(gpu, ref):
    """whole RandomEvaluatedRollout games: black-win rate and mean length per position vs the compiled reference"""
    n_g, n_r = 4096, 600
    for i, m in enumerate(_starts(78, 10)[4:10]):
        mv, st = pyoracle.pack_moves([m])
        boards = np.repeat(gpu.pack_moves(mv, st), n_g, axis=0)
        out = gpu.guided_rollout_batch(boards, mode="sample", key=KEY + 100 + i, want_moves=False)
        w = out["winner"].cpu().numpy()
        wdb, total = ref.guided_rollout_sampled(m, n_r)
        assert _binomial_gate(int((w == 1).sum()), n_g, int(wdb[2]), n_r), (m, (w == 1).mean(), wdb)
        assert _binomial_gate(int((w == -1).sum()), n_g, int(wdb[0]), n_r), (m, (w == -1).mean(), wdb)
        assert abs(out["length"].cpu().numpy().mean() - total / n_r) <= 0.15 * max(total / n_r, 4.0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["max", "sample"])
def test_gpu_incremental_guided_kernel_plays_the_games_of_the_full_rescan(gpu, mode):
    """guided_kernel (per move: only the four lines through the new stone, before and after) against
    ac_eval_kernel<true, true> (whole board after every move): same winners, lengths, moves and final boards, game for
    game -- from the empty board, from the synthetic mid-game set, from clustered / nearly full / decided positions."""
    import torch
    lists = _starts(55, 300) + random_positions(56, 200, lo=60, hi=200)
    mv, st = pyoracle.pack_moves(lists)
    boards = np.concatenate([np.zeros((40, 16), np.uint32), gpu.synth_positions(7, 1500, want_moves=False)[0], gpu.pack_moves(mv, st)])
    assert len(boards) <= 148 * 15                          # few enough for the two-warps-per-game kernel to be the default
    for max_moves in (225, 7):
        full = gpu.guided_rollout_batch(boards, mode=mode, key=KEY + 3, game_base=17, max_moves=max_moves, full_rescan=True)
        variants = {"two warps per game": gpu.guided_rollout_batch(boards, mode=mode, key=KEY + 3, game_base=17, max_moves=max_moves),
                    "one warp per game": gpu.guided_rollout_batch(boards, mode=mode, key=KEY + 3, game_base=17, max_moves=max_moves, single_warp=True),
                    "two warps, 64 in flight": gpu.guided_rollout_batch(boards, mode=mode, key=KEY + 3, game_base=17, max_moves=max_moves, max_in_flight=64)}
        torch.cuda.synchronize()
        for name, inc in variants.items():
            for k in ("winner", "length", "moves", "final_boards"):
                a, b = inc[k].cpu().numpy(), full[k].cpu().numpy()
                assert np.array_equal(a, b), (name, mode, max_moves, k, int(np.flatnonzero((a != b).reshape(len(boards), -1).any(axis=1))[0]))
    assert int(full["length"].sum()) > 0


@pytest.mark.gpu
def test_gpu_guided_max_moves_and_determinism(gpu):
    boards, _, _ = gpu.synth_positions(0, 64)
    a = gpu.guided_rollout_batch(boards, mode="sample", key=KEY, max_moves=3)
    assert int(a["length"].max()) <= 3
    b = gpu.guided_rollout_batch(boards, mode="sample", key=KEY)
    q = gpu.guided_rollout_batch(boards, mode="sample", key=KEY, max_in_flight=8)     # a queue worked through by 8 warps
    assert all(np.array_equal(b[k].cpu().numpy(), q[k].cpu().numpy()) for k in ("winner", "length", "moves", "final_boards"))
    c = gpu.guided_rollout_batch(boards[32:], mode="sample", key=KEY, game_base=32)
    assert np.array_equal(b["moves"].cpu().numpy()[32:], c["moves"].cpu().numpy())   # independent of batch split
    assert np.array_equal(b["winner"].cpu().numpy()[32:], c["winner"].cpu().numpy())


@pytest.mark.gpu
def test_gpu_expand_games_replays_every_ply(gpu):
    """gk_expand_games: the position before every ply of every game, its two last moves and the outcome for the side
    to move -- against a plain replay of the move lists; the planes of those positions are Board.encoded_states()."""
    import torch
    from oracle import pyoracle as po
    boards = gpu.synth_positions(40, 48, want_moves=False)[0]
    boards[:8] = 0                                              # some games from the empty board
    g = gpu.guided_rollout_batch(boards, mode="sample", key=5)
    ex = gpu.expand_games(boards, g)
    torch.cuda.synchronize()
    lengths = g["length"].cpu().numpy().astype(int); moves = g["moves"].cpu().numpy(); winner = g["winner"].cpu().numpy()
    out_b = ex["boards"].cpu().numpy().view(np.uint32); out_l = ex["last_moves"].cpu().numpy(); out_z = ex["z"].cpu().numpy()
    assert len(out_b) == lengths.sum() and ex["game"].cpu().numpy().tolist() == np.repeat(np.arange(48), lengths).tolist()
    row = 0
    for i in range(48):
        cells = gpu.unpack_boards(boards[i:i + 1])[0].astype(int).copy()
        black = int((cells == 1).sum() == (cells == 2).sum())
        last = [-1, -1]
        for k in range(lengths[i]):
            assert np.array_equal(gpu.unpack_boards(out_b[row:row + 1])[0], cells)
            assert out_l[row].tolist() == last and out_z[row] == (winner[i] if black else -winner[i])
            c = int(moves[i, k])
            assert cells[c] == 0
            cells[c] = 1 if black else 2
            last = [c, last[0]]
            black ^= 1
            row += 1
        assert np.array_equal(gpu.unpack_boards(g["final_boards"][i:i + 1].cpu().numpy().view(np.uint32))[0], cells)
    # the planes of one game from the empty board equal the restated Board.encoded_states() of its prefixes
    planes = gpu.encode_states_batch(ex["boards"], ex["last_moves"]).cpu().numpy()
    s0 = int(ex["starts"][0])
    for k in range(min(lengths[0], 12)):
        assert np.array_equal(planes[s0 + k].reshape(6, 15, 15), po.encoded_states([int(m) for m in moves[0, :k]]))

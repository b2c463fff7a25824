"""K1 `ac_eval` through the C-ABI against the oracle: bit-exact m_scores, pattern totals, compound
totals and winner.  Needs a B200 (-m gpu)."""
import numpy as np
import pytest

from conftest import fnv, random_positions
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _gpu_eval(gk, boards, **kw):
    import torch
    out = gk.eval_batch(boards, **kw)
    torch.cuda.synchronize()
    return {
        "scores": None if out.get("scores") is None else out["scores"].cpu().numpy(),
        "pat_totals": out["pat_totals"].cpu().numpy().view(np.uint16),
        "cmp_totals": out["cmp_totals"].cpu().numpy().view(np.uint16),
        "winner": out["winner"].cpu().numpy(),
    }


def _assert_same(got, want, keys=("scores", "pat_totals", "cmp_totals", "winner")):
    for k in keys:
        if not np.array_equal(got[k], want[k]):
            bad = np.nonzero((got[k] != want[k]).reshape(len(got[k]), -1).any(axis=1))[0]
            raise AssertionError(f"{k}: {len(bad)} positions differ, first {bad[:8].tolist()}")


def test_synthetic_midgame_positions_bit_exact(gpu, port):
    boards, moves, starts = gpu.synth_positions(0, 8192)
    want = port.eval_batch(moves, starts)
    assert want["bad"] == 0
    _assert_same(_gpu_eval(gpu, boards), want)
    assert (want["cmp_totals"].sum(axis=(1, 2)) > 0).mean() > 0.3      # compounds are exercised


def test_random_clustered_and_terminal_positions_bit_exact(gpu, port):
    lists = random_positions(101, 6000, lo=0, hi=160)
    mv, st = po.pack_moves(lists)
    want = port.eval_batch(mv, st)
    assert want["bad"] == 0 and (want["winner"] != 0).sum() > 50
    _assert_same(_gpu_eval(gpu, gpu.pack_moves(mv, st)), want)


def test_against_compiled_reference(gpu, ref):
    lists = random_positions(202, 3000, lo=0, hi=130)
    mv, st = po.pack_moves(lists)
    want = ref.eval_batch(mv, st)
    assert want["bad"] == 0
    _assert_same(_gpu_eval(gpu, gpu.pack_moves(mv, st)), want)


def test_committed_reference_outputs(gpu, ref_outputs):
    names = list(ref_outputs["eval"])
    lists = [ref_outputs["eval"][n]["moves"][:ref_outputs["eval"][n]["applied"]] for n in names]
    mv, st = po.pack_moves(lists)
    got = _gpu_eval(gpu, gpu.pack_moves(mv, st))
    for i, n in enumerate(names):
        g = ref_outputs["eval"][n]
        assert got["scores"][i].sum(axis=1).tolist() == g["score_sums"], n
        assert [fnv(got["scores"][i][k]) for k in range(4)] == g["score_fnv"], n
        assert got["pat_totals"][i].tolist() == g["pat_totals"], n
        assert got["cmp_totals"][i].tolist() == g["cmp_totals"], n
        assert int(got["winner"][i]) == g["winner"], n
    r1 = names.index("R1")
    assert np.array_equal(got["scores"][r1], np.array(ref_outputs["eval"]["R1"]["scores"], np.int32))


def test_edge_positions(gpu, port, kats):
    tie = [y * 15 + x for y in kats["tie_row_order"] for x in range(15)]
    lists = [
        [],                                                   # empty board
        [112],                                                # one stone
        [0], [14], [210], [224], [7], [105],                  # corners and edges
        [y * 15 + x for x, y in kats["black_win"]],           # black five (diagonal)
        [y * 15 + x for x, y in kats["white_win"]],           # white five (column)
        tie,                                                  # full board, draw
        tie[:-1], tie[:200],                                  # almost full
        [0, 15, 1, 16, 2, 17, 3, 18, 4],                      # five on the top edge
        [210, 0, 211, 1, 212, 2, 213, 3, 214],                # five on the bottom edge
        [14, 0, 28, 1, 42, 2, 56, 3, 70],                     # five on an anti-diagonal from the corner
    ]
    mv, st = po.pack_moves(lists)
    want = port.eval_batch(mv, st)
    assert want["bad"] == 0
    got = _gpu_eval(gpu, gpu.pack_moves(mv, st))
    _assert_same(got, want)
    assert got["winner"].tolist()[8:11] == [1, -1, 0]
    assert (got["scores"][0] == 0).all() and (got["scores"][10] == 0).all()


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 4735, 4737])
def test_ragged_batch_sizes_and_null_outputs(gpu, port, n):
    import torch
    boards, moves, starts = gpu.synth_positions(7, max(n, 1))
    boards = boards[:n]
    out = gpu.eval_batch(boards) if n else gpu.eval_batch(np.zeros((0, 16), np.uint32))
    torch.cuda.synchronize()
    if n == 0:
        return
    want = port.eval_batch(moves[:starts[n]], starts[:n + 1])
    assert np.array_equal(out["scores"].cpu().numpy(), want["scores"])
    only_totals = _gpu_eval(gpu, boards, want_scores=False)
    assert only_totals["scores"] is None
    _assert_same(only_totals, want, keys=("pat_totals", "cmp_totals", "winner"))


def test_host_buffer_entry_point(gpu, port):
    boards, moves, starts = gpu.synth_positions(50000, 40000)      # > 2 pipeline chunks
    out = gpu.eval_batch_host(boards)
    sub = np.arange(0, 40000, 13)
    mv, st = po.pack_moves([moves[starts[i]:starts[i + 1]] for i in sub])
    want = port.eval_batch(mv, st)
    assert np.array_equal(out["scores"][sub], want["scores"])
    assert np.array_equal(out["pat_totals"][sub], want["pat_totals"])
    assert np.array_equal(out["cmp_totals"][sub], want["cmp_totals"])
    dev = _gpu_eval(gpu, boards)
    for k in ("scores", "pat_totals", "cmp_totals", "winner"):
        assert np.array_equal(out[k], dev[k])


def test_full_size_properties_1m_positions(gpu, port):
    """BASELINE config 2 size: 1M positions.  Size-independent properties + a sampled oracle check."""
    import torch
    n = 1 << 20
    boards, moves, starts = gpu.synth_positions(0, n)
    bt = torch.from_numpy(boards.view(np.int32)).cuda()
    a = gpu.eval_batch(bt)
    torch.cuda.synchronize()
    # (1) the reference's own always-on self-check (Pattern.cpp:314-333): occupied cells score 0, no negative score
    cells = torch.from_numpy(gpu.unpack_boards(boards)).cuda()
    occ = (cells != 0).unsqueeze(1)
    assert int((a["scores"] * occ).abs().sum()) == 0
    assert int((a["scores"] < 0).sum()) == 0
    assert int(a["winner"].abs().sum()) == 0                   # the synthetic set is non-terminal
    # (2) deterministic and independent of how the batch is cut
    b = gpu.eval_batch(bt)
    c0 = gpu.eval_batch(bt[:300001])
    c1 = gpu.eval_batch(bt[300001:])
    torch.cuda.synchronize()
    for k in ("scores", "pat_totals", "cmp_totals", "winner"):
        assert torch.equal(a[k], b[k])
        assert torch.equal(a[k][:300001], c0[k]) and torch.equal(a[k][300001:], c1[k])
    # (3) block score: exactly 160 per empty cell that has an own stone on a weighted offset, so
    #     scores(P,P) >= 160 there; total score mass equals the oracle's on a sample
    sub = np.arange(0, n, 257)
    mv, st = po.pack_moves([moves[starts[i]:starts[i + 1]] for i in sub])
    want = port.eval_batch(mv, st)
    idx = torch.from_numpy(sub).cuda()
    assert np.array_equal(a["scores"][idx].cpu().numpy(), want["scores"])
    assert np.array_equal(a["pat_totals"][idx].cpu().numpy().view(np.uint16), want["pat_totals"])
    assert np.array_equal(a["cmp_totals"][idx].cpu().numpy().view(np.uint16), want["cmp_totals"])

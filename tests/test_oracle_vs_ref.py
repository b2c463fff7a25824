"""The C restatement against the reference itself (oracle/_ref, compiled from /root/reference by
oracle/Makefile).  Skipped where oracle/_ref has not been built.  CPU only."""
import numpy as np

from conftest import random_positions
from oracle import pyoracle as po


def test_automaton_arrays_identical(port, ref):
    a, b = port.table(), ref.table()
    for k in ("base", "check", "fail", "invariants"):
        assert np.array_equal(a[k], b[k]), k
    assert a["patterns"] == b["patterns"]


def test_custom_automaton_identical(port, ref, kats):
    k = kats["kat_protos"]
    port.build_custom(k["protos"], k["types"], k["scores"])
    ref.build_custom(k["protos"], k["types"], k["scores"])
    a, b = port.table(True), ref.table(True)
    for key in ("base", "check", "fail", "invariants"):
        assert np.array_equal(a[key], b[key]), key
    assert a["patterns"] == b["patterns"]


def _strings(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        kind = i % 3
        if kind == 0:                                   # anything over the 4 symbols
            s = rng.integers(1, 5, size=int(rng.integers(1, 30)))
        elif kind == 1:                                 # a padded board line
            body = rng.choice([1, 2, 4], size=int(rng.integers(1, 16)), p=[.3, .3, .4])
            s = np.concatenate([[3] * 6, body, [3] * 6])
        else:                                           # a 13-symbol window cut from a padded line
            body = rng.choice([1, 2, 4], size=int(rng.integers(5, 16)), p=[.35, .35, .3])
            full = np.concatenate([[3] * 6, body, [3] * 6])
            a = int(rng.integers(0, len(full) - 12))
            s = full[a:a + 13]
        out.append(np.asarray(s, np.uint8))
    return out


def test_scan_emissions_identical(port, ref):
    strings = _strings(5, 60000)
    starts = np.zeros(len(strings) + 1, np.int64)
    starts[1:] = np.cumsum([len(s) for s in strings])
    codes = np.concatenate(strings)
    a = port.scan_many(codes, starts)
    b = ref.scan_many(codes, starts)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[2].sum() > 20000


def test_line_maps_identical(port, ref):
    for mv in random_positions(3, 20):
        for la, lb in zip(port.line_map(mv), ref.line_map(mv)):
            assert np.array_equal(la, lb)


def test_evaluator_identical_on_random_positions(port, ref):
    mv, st = po.pack_moves(random_positions(11, 1500))
    a, b = port.eval_batch(mv, st), ref.eval_batch(mv, st)
    assert a["bad"] == 0 and b["bad"] == 0
    for k in ("scores", "pat_totals", "cmp_totals", "winner", "cur_player"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["cmp_totals"].sum(axis=(1, 2)) > 0).sum() > 300
    assert (a["winner"] != 0).sum() > 20
    assert port.degenerate_compounds() == 0


def test_per_cell_flag_words_identical(port, ref):
    for mv in random_positions(12, 40):
        port.eval_moves(mv)
        ref.eval_moves(mv)
        for x, y in zip(port.eval_flags(), ref.eval_flags()):
            assert np.array_equal(x, y)


def test_injected_rollouts_identical(port, ref):
    rng = np.random.default_rng(9)
    for mv in random_positions(13, 60, lo=0, hi=90):
        for _ in range(4):
            rs = rng.integers(0, 225, size=232).astype(np.uint8)
            assert port.rollout_injected(mv, rs) == ref.rollout_injected(mv, rs)


def test_from_scratch_model_equals_incremental_reference(port, ref):
    """The position-only recomputation the CUDA kernel uses against the
    reference's incremental evaluator, with the reference's 6/6 padding and with the kernel's 1/2."""
    mv, st = po.pack_moves(random_positions(21, 1500))
    want = ref.eval_batch(mv, st)
    for lead, trail, min_len in ((6, 6, 1), (1, 2, 5)):
        got = port.eval_scratch_batch(mv, st, lead, trail, min_len)
        assert got["bad"] == 0
        for k in ("scores", "pat_totals", "cmp_totals", "winner"):
            assert np.array_equal(got[k], want[k]), (k, lead, trail)
    assert port.scratch_gate_blocks() == 0

"""Row f1: the evaluator's policy heads -- Heuristic::DensityWeight / EvaluationProbs /
EvaluationValue (include/algorithms/Heuristic.hpp:16-45) fused into ac_eval_kernel.

Floating point.  The reference computes these with Eigen float vectors; the kernel sums in a
different order (lane-strided, then warp shuffles).  Tolerances, stated once:
    probs: |gpu - oracle| <= 2e-6 + 2e-5 * |oracle|   (unit-norm vectors, entries <= 1)
    value: |gpu - oracle| <= 2e-5                      (tanh output in [-1, 1])
No reference test pins these functions and Heuristic.hpp itself needs real Eigen, which
is absent here: the formulas are restated in oracle/pyoracle.py::policy_heads and evaluated on the
density / score arrays of the COMPILED reference evaluator (oracle/_ref) for the golden fixture."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_positions
from oracle import pyoracle

PROBS_ATOL, PROBS_RTOL, VALUE_ATOL = 2e-6, 2e-5, 2e-5


def _positions(seed, n):
    lists = [[], [112], [0], [224, 0], [112, 113, 97]] + random_positions(seed, n, lo=2, hi=110)
    # heads are read for non-terminal positions only (TraditionalPolicy::hybridSimulate runs after checkGameEnd)
    out = []
    for m in lists:
        out.append(m)
    return out


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "reference_heads.json")) as f:
        return json.load(f)


def test_heads_restatement_on_reference_fixture(port, golden):
    """numpy restatement over the C oracle's evaluator state == committed values computed over the compiled reference"""
    for item in golden["positions"]:
        probs, value = pyoracle.policy_heads(port, item["moves"])
        assert abs(float(value) - item["value"]) <= 1e-6
        assert np.allclose(probs[item["top_cells"]], np.array(item["top_probs"], np.float32), rtol=1e-6, atol=1e-7)
        assert abs(float(np.abs(probs).sum()) - item["l1"]) <= 1e-4


def test_heads_port_equals_reference(port, ref):
    for m in _positions(3, 120):
        if port.eval_moves(m)["winner"] != 0:
            continue
        pp, pv = pyoracle.policy_heads(port, m)
        rp, rv = pyoracle.policy_heads(ref, m)
        assert np.array_equal(pp, rp) and pv == rv          # identical inputs (scores, density) => identical floats


def test_heads_properties(port):
    probs, value = pyoracle.policy_heads(port, [])
    assert probs[112] == 1.0 and probs.sum() == 1.0 and value == 0.0
    for m in _positions(8, 40)[1:]:
        probs, value = pyoracle.policy_heads(port, m)
        assert np.all(probs[np.array(m)] == 0)               # occupied cells carry no probability (MCTS.h:86-92)
        assert np.all(probs >= 0) and -1.0 <= value <= 1.0
        if probs.any():
            assert abs(float(np.sqrt((probs.astype(np.float64) ** 2).sum())) - 1.0) < 1e-5


def _check_gpu(gk, lists, out):
    port = pyoracle.port()
    probs, value = out["probs"].cpu().numpy(), out["value"].cpu().numpy()
    worst_p = worst_v = 0.0
    for i, m in enumerate(lists):
        rp, rv = pyoracle.policy_heads(port, m)
        err = np.abs(probs[i] - rp) - (PROBS_ATOL + PROBS_RTOL * np.abs(rp))
        assert err.max() <= 0, (i, m, float(err.max()))
        assert abs(float(value[i]) - float(rv)) <= VALUE_ATOL, (i, m, float(value[i]), float(rv))
        worst_p = max(worst_p, float(np.abs(probs[i] - rp).max()))
        worst_v = max(worst_v, abs(float(value[i]) - float(rv)))
    return worst_p, worst_v


@pytest.mark.gpu
def test_gpu_policy_heads_vs_oracle(gpu):
    port = pyoracle.port()
    lists = [m for m in _positions(21, 700) if port.eval_moves(m)["winner"] == 0]
    mv, st = pyoracle.pack_moves(lists)
    out = gpu.eval_policy_batch(gpu.pack_moves(mv, st), want_scores=True)
    _check_gpu(gpu, lists, out)
    # the integer outputs of the fused kernel are the same as the plain evaluator's
    ref = port.eval_batch(mv, st)
    assert np.array_equal(out["scores"].cpu().numpy(), ref["scores"])
    assert np.array_equal(out["pat_totals"].cpu().numpy().view(np.uint16), ref["pat_totals"])
    assert np.array_equal(out["cmp_totals"].cpu().numpy().view(np.uint16), ref["cmp_totals"])


@pytest.mark.gpu
def test_gpu_policy_heads_synthetic_set_and_host_path(gpu):
    boards, moves, starts = gpu.synth_positions(0, 512)
    lists = [moves[starts[i]:starts[i + 1]].tolist() for i in range(512)]
    out = gpu.eval_policy_batch(boards)
    _check_gpu(gpu, lists, out)
    probs, value, winner = gpu.eval_policy_batch_host(boards)
    assert np.array_equal(probs, out["probs"].cpu().numpy()) and np.array_equal(value, out["value"].cpu().numpy())
    assert np.array_equal(winner, out["winner"].cpu().numpy())


@pytest.mark.gpu
def test_gpu_policy_heads_golden(gpu, golden):
    lists = [it["moves"] for it in golden["positions"]]
    mv, st = pyoracle.pack_moves(lists)
    out = gpu.eval_policy_batch(gpu.pack_moves(mv, st))
    probs, value = out["probs"].cpu().numpy(), out["value"].cpu().numpy()
    for i, it in enumerate(golden["positions"]):
        want = np.array(it["top_probs"], np.float32)
        assert np.all(np.abs(probs[i][it["top_cells"]] - want) <= PROBS_ATOL + PROBS_RTOL * np.abs(want))
        assert abs(float(value[i]) - it["value"]) <= VALUE_ATOL


# ---- DecisiveFilter / hybridSimulate ------------------------------------------------------------------------
def _dflag_bits(pf_cell, cf_cell):
    """the kernel's per-cell word from the oracle's Record fields: any-direction flag per (type, group)"""
    w = 0
    for t in range(4, 8):
        for g in range(4):
            if (int(pf_cell[t]) >> (8 * g)) & 0xff:
                w |= 1 << ((t - 4) * 4 + g)
    for ct in range(3):
        for g in range(4):
            if (int(cf_cell[ct]) >> (8 * g)) & 0xff:
                w |= 1 << (16 + ct * 4 + g)
    return w


def test_decisive_filter_restatement_properties(port):
    fired = 0
    for m in _positions(12, 150):
        if port.eval_moves(m)["winner"] != 0:
            continue
        probs, _ = pyoracle.policy_heads(port, m)
        out, cand = pyoracle.decisive_filter(port, m, probs)
        if cand:
            fired += 1
            assert np.all((out == 0) | (probs > 0))                          # the filter only removes cells
            if out.any():
                assert abs(float(np.sqrt((out.astype(np.float64) ** 2).sum())) - 1.0) < 1e-5
        else:
            assert np.array_equal(out, probs)
    assert fired > 30


def test_hybrid_simulate_port_equals_reference(port, ref):
    for m in _positions(14, 80):
        if port.eval_moves(m)["winner"] != 0:
            continue
        pv, pp = pyoracle.hybrid_simulate(port, m)
        rv, rp = pyoracle.hybrid_simulate(ref, m)
        assert pv == rv and np.array_equal(pp, rp)


@pytest.mark.gpu
def test_gpu_hybrid_simulate_vs_oracle(gpu):
    """probs after DecisiveFilter within the stated tolerance, and the per-cell flag words the filter reads compared
    bit for bit with the incremental evaluator's.  Pattern-flag bits (0..15) must be IDENTICAL.  Compound-flag bits
    (16..27) may only differ one way: the reference keeps a 2-bit saturating shift register per (cell, type, group,
    direction) (Record::set, Pattern.cpp:395-400), so a slot that once held >= 3 contributions (a compound's own key
    cell that is also an anti cell of other compounds on the same line) forgets the excess when those are removed
    and can read 0 while contributions remain -- it depends on the move ORDER, which a from-scratch evaluation of
    the position cannot know.  Hence: kernel bits are a superset of the oracle's, on at most 2 % of the positions."""
    port = pyoracle.port()
    lists = [m for m in _positions(41, 900) if port.eval_moves(m)["winner"] == 0]
    mv, st = pyoracle.pack_moves(lists)
    out = gpu.hybrid_simulate_batch(gpu.pack_moves(mv, st), want_flags=True)
    probs, value = out["probs"].cpu().numpy(), out["value"].cpu().numpy()
    dflags = out["dflags"].cpu().numpy().view(np.uint32)
    flag_mismatch = prob_mismatch = 0
    for i, m in enumerate(lists):
        rv, rp = pyoracle.hybrid_simulate(port, m)                              # leaves the evaluator at position m
        pf, cf, _ = port.eval_flags()
        want = np.array([_dflag_bits(pf[c], cf[c]) for c in range(225)], np.uint32)
        same_flags = np.array_equal(dflags[i], want)
        flag_mismatch += not same_flags
        assert np.array_equal(dflags[i] & 0xffff, want & 0xffff), (i, m)        # pattern flags: exact
        assert np.all((dflags[i] & want) == want), (i, m)                       # compound flags: only extra bits
        ok = np.all(np.abs(probs[i] - rp) <= PROBS_ATOL + PROBS_RTOL * np.abs(rp))
        prob_mismatch += not ok
        assert ok or not same_flags, (i, m)                                     # a probs difference must come from a flag difference
        assert abs(float(value[i]) - float(rv)) <= VALUE_ATOL
    assert flag_mismatch <= len(lists) // 50, (flag_mismatch, len(lists))
    assert prob_mismatch <= len(lists) // 100, (prob_mismatch, len(lists))
    print("hybridSimulate: flag-word mismatches", flag_mismatch, "prob mismatches", prob_mismatch, "of", len(lists))

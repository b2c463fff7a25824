"""Row f1: the evaluator's policy heads -- Heuristic::DensityWeight / EvaluationProbs /
EvaluationValue (include/algorithms/Heuristic.hpp:16-45), Heuristic::DecisiveFilter (:93-161) and
TraditionalPolicy::hybridSimulate (include/policies/Traditional.h:49-69) fused into ac_eval_kernel.

The checker is the REFERENCE ITSELF: Heuristic.hpp / Traditional.h compiled unmodified into oracle/_ref
(oracle/ref_harness_search.cpp; `ref.heads`, `ref.decisive_filter`, `ref.hybrid_simulate`), plus the committed
outputs of that build (tests/golden/reference_heads.json, made by tests/golden/make_golden_heads.py).  The numpy
restatement in oracle/pyoracle.py is itself pinned against the compiled reference here and only stands in where
oracle/_ref has not been built.

Floating point.  The reference computes these with Eigen float vectors; Eigen, the shim the oracle is built
against, and the kernel sum the same products in three different orders (SIMD packets / left to right /
lane-strided then warp shuffles), so equality is up to rounding.  Tolerances in units of float32 ulp(1) = 2^-23:
    probs: |gpu - ref| <= 16 ulp(1) + 168 ulp(1) * |ref|   (1.9e-6 + 2.0e-5 |ref|; unit-norm vectors, entries <= 1)
    value: |gpu - ref| <= 256 ulp(1)                        (3.1e-5; tanh output in [-1, 1]; its argument is the
                                                             difference of two float dot products of 225 terms with
                                                             sums of 1e3..1e4, each good to ~1e-3 under reordering,
                                                             divided by 500; measured worst case: 136 ulp(1))"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_positions
from oracle import pyoracle

ULP = 2.0 ** -23
PROBS_ATOL, PROBS_RTOL, VALUE_ATOL = 16 * ULP, 168 * ULP, 256 * ULP


def _checker():
    """the compiled reference when it was built, else the numpy restatement over the C port"""
    r = pyoracle.ref()
    if r is not None:
        return "reference", r.heads, r.hybrid_simulate
    port = pyoracle.port()
    return "restatement", (lambda m: pyoracle.policy_heads(port, m)), (lambda m: pyoracle.hybrid_simulate(port, m))


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
    """numpy restatement over the C oracle's evaluator state vs the committed outputs of the COMPILED reference
    (Heuristic.hpp itself): same formulas, different summation order"""
    for item in golden["positions"]:
        probs, value = pyoracle.policy_heads(port, item["moves"])
        assert abs(float(value) - item["value"]) <= VALUE_ATOL
        want = np.array(item["top_probs"], np.float32)
        assert np.all(np.abs(probs[item["top_cells"]] - want) <= 4 * ULP + 8 * ULP * np.abs(want))
        assert abs(float(np.abs(probs).sum()) - item["l1"]) <= 1e-4
        hv, hp = pyoracle.hybrid_simulate(port, item["moves"])
        want = np.zeros(225, np.float32)
        want[item["hybrid_cells"]] = item["hybrid_probs"]
        assert abs(float(hv) - item["hybrid_value"]) <= VALUE_ATOL
        assert np.all(np.abs(hp - want) <= 4 * ULP + 8 * ULP * np.abs(want)), item["moves"]


def test_compiled_reference_reproduces_its_fixture(ref, golden):
    for item in golden["positions"]:
        probs, value = ref.heads(item["moves"])
        assert float(value) == item["value"] and probs[item["top_cells"]].tolist() == item["top_probs"]
        hv, hp = ref.hybrid_simulate(item["moves"])
        assert float(hv) == item["hybrid_value"] and np.flatnonzero(hp).tolist() == item["hybrid_cells"]
    for g in golden["guided_max"]:
        assert ref.guided_rollout_max(g["moves"]) == (g["winner"], g["played"])


def test_restatement_equals_compiled_heuristic_hpp(port, ref):
    """oracle/pyoracle.py::policy_heads / decisive_filter / hybrid_simulate against Heuristic::EvaluationProbs /
    EvaluationValue / DensityWeight / DecisiveFilter and TraditionalPolicy::hybridSimulate of the compiled reference"""
    fired = 0
    for m in _positions(3, 160):
        h = ref.heads(m, want_dw=True)
        if h is None:
            assert port.eval_moves(m)["winner"] != 0 or len(m) == 225
            continue
        for orc in (port, ref):                             # the restatement over either evaluator state
            pp, pv = pyoracle.policy_heads(orc, m)
            assert np.all(np.abs(pp - h[0]) <= 4 * ULP + 8 * ULP * np.abs(h[0])) and abs(float(pv) - float(h[1])) <= VALUE_ATOL
        assert np.all(h[2] >= 0) and all(abs(float(np.sqrt((d.astype(np.float64) ** 2).sum())) - 1.0) < 1e-5 for d in h[2] if d.any())
        filtered = ref.decisive_filter(m, h[0])
        mine, cand = pyoracle.decisive_filter(port, m, h[0])
        assert np.array_equal(filtered != 0, mine != 0), m  # the same cells survive
        assert np.all(np.abs(filtered - mine) <= 4 * ULP + 8 * ULP * np.abs(filtered))
        fired += bool(cand)
        hv, hp = ref.hybrid_simulate(m)
        assert hv == h[1] and np.array_equal(hp, filtered)  # hybridSimulate = EvaluationValue + filtered EvaluationProbs
    assert fired > 30


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
    kind, heads, _ = _checker()
    probs, value = out["probs"].cpu().numpy(), out["value"].cpu().numpy()
    worst_p = worst_v = 0.0
    for i, m in enumerate(lists):
        rp, rv = heads(m)
        err = np.abs(probs[i] - rp) - (PROBS_ATOL + PROBS_RTOL * np.abs(rp))
        assert err.max() <= 0, (i, m, float(err.max()))
        assert abs(float(value[i]) - float(rv)) <= VALUE_ATOL, (i, m, float(value[i]), float(rv))
        worst_p = max(worst_p, float(np.abs(probs[i] - rp).max()))
        worst_v = max(worst_v, abs(float(value[i]) - float(rv)))
    print(f"policy heads vs {kind}: worst |dprobs| = {worst_p / ULP:.1f} ulp(1), worst |dvalue| = {worst_v / ULP:.1f} ulp(1) over {len(lists)} positions")
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
    hyb = gpu.hybrid_simulate_batch(gpu.pack_moves(mv, st))
    hprobs, hvalue = hyb["probs"].cpu().numpy(), hyb["value"].cpu().numpy()
    for i, it in enumerate(golden["positions"]):
        want = np.array(it["top_probs"], np.float32)
        assert np.all(np.abs(probs[i][it["top_cells"]] - want) <= PROBS_ATOL + PROBS_RTOL * np.abs(want))
        assert abs(float(value[i]) - it["value"]) <= VALUE_ATOL
        want = np.zeros(225, np.float32)
        want[it["hybrid_cells"]] = it["hybrid_probs"]
        assert np.all(np.abs(hprobs[i] - want) <= PROBS_ATOL + PROBS_RTOL * np.abs(want)), it["moves"]
        assert abs(float(hvalue[i]) - it["hybrid_value"]) <= VALUE_ATOL


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


def test_reference_compound_flags_depend_on_the_move_order(ref):
    """Why the compound-flag bits of a from-scratch evaluation can only be compared one way (see the GPU test below): the
    REFERENCE's own per-cell compound flags are not a function of the position.  Replaying the same stones in another
    order (black stones permuted among the black moves, white among the white) leaves m_scores and every pattern-flag
    presence bit unchanged, but changes compound-flag bits -- Record::set's 2-bit saturating counters (Pattern.cpp:395-400)
    forget contributions that are removed after a third was added, and which ones depends on the history."""
    rng = np.random.default_rng(0)

    def state(m):
        r = ref.eval_moves(m)
        if r["bad"] or r["winner"] != 0:
            return None
        pf, cf, _ = ref.eval_flags()
        return r["scores"].copy(), np.array([_dflag_bits(pf[c], cf[c]) for c in range(225)], np.uint32)

    tried = differ = 0
    for m in _positions(41, 260):
        a = state(m) if len(m) >= 6 else None
        if a is None:
            continue
        for _ in range(3):
            mm = [0] * len(m)
            mm[0::2] = [int(c) for c in rng.permutation(m[0::2])]
            mm[1::2] = [int(c) for c in rng.permutation(m[1::2])]
            b = state(mm)                                    # a permuted order can complete a five early: skipped
            if b is None:
                continue
            tried += 1
            assert np.array_equal(a[0], b[0])                # scores: order-independent
            assert np.array_equal(a[1] & 0xffff, b[1] & 0xffff)   # pattern-flag presence bits: order-independent
            differ += not np.array_equal(a[1] >> 16, b[1] >> 16)
    assert tried > 400 and differ >= 3, (tried, differ)     # compound-flag bits: the reference disagrees with itself


def test_hybrid_simulate_port_equals_reference(port, ref):
    for m in _positions(14, 80):
        if port.eval_moves(m)["winner"] != 0:
            continue
        pv, pp = pyoracle.hybrid_simulate(port, m)
        rv, rp = pyoracle.hybrid_simulate(ref, m)
        assert pv == rv and np.array_equal(pp, rp)          # identical inputs (scores, density, flags) => identical floats


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
    kind, _, hybrid = _checker()
    lists = [m for m in _positions(41, 900) if port.eval_moves(m)["winner"] == 0]
    mv, st = pyoracle.pack_moves(lists)
    out = gpu.hybrid_simulate_batch(gpu.pack_moves(mv, st), want_flags=True)
    probs, value = out["probs"].cpu().numpy(), out["value"].cpu().numpy()
    dflags = out["dflags"].cpu().numpy().view(np.uint32)
    flag_mismatch = prob_mismatch = 0
    for i, m in enumerate(lists):
        rv, rp = hybrid(m)                                                      # TraditionalPolicy::hybridSimulate of the compiled reference
        port.eval_moves(m)                                                      # leaves the port's evaluator at position m
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
    print("hybridSimulate vs", kind, ": flag-word mismatches", flag_mismatch, "prob mismatches", prob_mismatch, "of", len(lists))

"""Row f3: Board.encoded_states() planes and augment_game_data's 8 variants.
CPU: the numpy restatement (oracle/pyoracle.py) against the committed reference outputs
(tests/golden/reference_features.json) and, where oracle/_ref exists, against the reference Board.
GPU: encode_states_kernel through the C-ABI against the restatement."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_positions
from oracle import pyoracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def feats():
    with open(os.path.join(GOLDEN, "reference_features.json")) as f:
        return json.load(f)


def test_encoded_states_matches_reference_outputs(feats):
    for item in feats["encoded_states"]:
        planes = pyoracle.encoded_states(item["moves"])
        assert planes.shape == (6, 15, 15) and planes.dtype == np.uint8
        assert sha(planes) == item["sha256"], item["moves"]
        assert [int(x) for x in planes.reshape(6, -1).sum(1)] == item["plane_sums"]
        if "planes" in item:
            assert planes.reshape(6, 225).tolist() == item["planes"]


def test_augment_matches_reference_outputs(feats):
    rng = np.random.default_rng(7)
    for item in feats["augment"]:
        probs = rng.random(225).astype(np.float32)
        st, pr = pyoracle.augment_planes(pyoracle.encoded_states(item["moves"]), probs)
        assert st.shape == (8, 6, 15, 15) and pr.shape == (8, 225)
        assert sha(st.astype(np.uint8)) == item["states_sha256"]
        assert sha(pr.astype(np.float32)) == item["probs_sha256"]
        assert np.allclose(pr[:, :8], np.array(item["probs_first8"], np.float32))


def test_encoded_states_vs_compiled_reference(ref):
    for mv in [[], [0], [224, 0]] + random_positions(11, 200, lo=1, hi=224):
        assert np.array_equal(pyoracle.encoded_states(mv), ref.encoded_states(mv))


def _case(gk, seed, n):
    lists = [[], [112], [112, 113], list(range(0, 224))] + random_positions(seed, n - 4, lo=1, hi=224)
    mv, st = pyoracle.pack_moves(lists)
    boards = gk.pack_moves(mv, st)
    last = np.full((len(lists), 2), -1, np.int16)
    for i, m in enumerate(lists):
        for k in (0, 1):
            if len(m) > k:
                last[i, k] = m[-1 - k]
    return lists, boards, last


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 513])
def test_gpu_encoded_states(gpu, n):
    lists, boards, last = _case(gpu, 100 + n, max(n, 4))
    lists, boards, last = lists[:n], boards[:n], last[:n]
    planes = gpu.encode_states_batch(boards, last).cpu().numpy()
    assert planes.shape == (n, 1, 6, 15, 15)
    for i, m in enumerate(lists):
        assert np.array_equal(planes[i, 0], pyoracle.encoded_states(m)), i


@pytest.mark.gpu
def test_gpu_augmented_states_and_probs(gpu):
    lists, boards, last = _case(gpu, 5, 300)
    rng = np.random.default_rng(3)
    probs = rng.random((len(lists), 225)).astype(np.float32)
    planes, pr = gpu.encode_states_batch(boards, last, augment=True, probs=probs)
    planes, pr = planes.cpu().numpy(), pr.cpu().numpy()
    for i, m in enumerate(lists):
        st_ref, pr_ref = pyoracle.augment_planes(pyoracle.encoded_states(m), probs[i])
        assert np.array_equal(planes[i], st_ref), i
        assert np.array_equal(pr[i], pr_ref), i


@pytest.mark.gpu
def test_gpu_encoded_states_golden(gpu, feats):
    items = feats["augment"]
    lists = [it["moves"] for it in items]
    mv, st = pyoracle.pack_moves(lists)
    boards = gpu.pack_moves(mv, st)
    last = np.array([[m[-1] if len(m) > 0 else -1, m[-2] if len(m) > 1 else -1] for m in lists], np.int16)
    planes = gpu.encode_states_batch(boards, last, augment=True).cpu().numpy()
    for i, it in enumerate(items):
        assert sha(planes[i]) == it["states_sha256"]


@pytest.mark.gpu
def test_gpu_encode_without_last_moves_and_empty_batch(gpu):
    _, boards, _ = _case(gpu, 9, 16)
    planes = gpu.encode_states_batch(boards).cpu().numpy()
    assert planes[:, 0, 3:5].sum() == 0
    assert gpu.encode_states_batch(boards[:0]).shape[0] == 0

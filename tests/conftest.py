import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ENC = {"x": 1, "o": 2, "?": 3, "-": 4, "_": 4, "^": 4, "~": 4}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_itemcollected(item):
    """A test that asks for the `gpu` fixture IS a gpu test, marked or not: `-m "not gpu"` must never reach it
    (collection hooks run before the -m filter)."""
    names = getattr(item, "fixturenames", ())
    if "gpu" in names and item.get_closest_marker("gpu") is None:
        item.add_marker(pytest.mark.gpu)


def encode(s):
    return np.array([ENC[c] for c in s], np.uint8)


def codes_of(s):
    return [ENC[c] for c in s]


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_outputs():
    with open(os.path.join(GOLDEN, "reference_outputs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    pyoracle.build(want_ref=False)
    return pyoracle.port()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref); tests that need it are skipped where it was not built."""
    from oracle import pyoracle
    r = pyoracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libgomoku_ref.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def gk():
    """The product package.  On a GPU box the library and CorePyExt are REBUILT from the sources of the snapshot once per
    session, before anything loads them (nvcc is in the image), so the kernels under test are the ones in the tree and
    not a stale prebuilt file; GK_NO_REBUILD=1 skips that.  Without a GPU the in-tree build is used (built if missing)."""
    import gomokuai_b200
    from gomokuai_b200 import build
    on_gpu_box = False
    try:
        import torch
        on_gpu_box = torch.cuda.is_available()
    except Exception:
        pass
    if on_gpu_box and not os.environ.get("GK_NO_REBUILD"):
        assert gomokuai_b200._lib is None and "gomokuai_b200.CorePyExt" not in sys.modules, "rebuild must precede the first load"
        build.build(force=True)
        build.build_pyext(force=True)
    elif not os.path.exists(gomokuai_b200.LIB_PATH):
        build.build()
    gomokuai_b200.lib()
    return gomokuai_b200


@pytest.fixture(scope="session")
def gpu(gk):
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    gk.init(0)
    return gk


def fnv(vec):
    h = 2166136261
    for v in vec:
        h = ((h ^ (int(v) & 0xffffffff)) * 16777619) & 0xffffffff
    return h


def _makes_five(cells, c, col):
    x0, y0 = c % 15, c // 15
    for dx, dy in ((1, 0), (0, 1), (1, 1), (1, -1)):
        n = 1
        for sgn in (1, -1):
            x, y = x0, y0
            for _ in range(5):
                x += sgn * dx
                y += sgn * dy
                if 0 <= x < 15 and 0 <= y < 15 and cells[y * 15 + x] == col:
                    n += 1
                else:
                    break
        if n >= 5:
            return True
    return False


def random_positions(seed, n, lo=1, hi=120, clustered_every=3):
    """Legal move lists: uniformly random and centre-clustered, alternating colours.  A list ends
    with the move that completes five-or-more (the reference accepts no move after that)."""
    rng = np.random.default_rng(seed)
    lists = []
    for i in range(n):
        k = int(rng.integers(lo, hi))
        if clustered_every and i % clustered_every == clustered_every - 1:
            cells = [y * 15 + x for y in range(3, 12) for x in range(3, 12)]
            order = [int(c) for c in rng.permutation(cells)[:min(k, 75)]]
        else:
            order = [int(c) for c in rng.permutation(225)[:k]]
        board, moves = [0] * 225, []
        for j, c in enumerate(order):
            col = 1 if j % 2 == 0 else 2
            board[c] = col
            moves.append(c)
            if _makes_five(board, c, col):
                break
        lists.append(moves)
    return lists

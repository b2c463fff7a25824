"""Generates tests/golden/reference_features.json (run in the build container only).

* ``encoded_states``: planes produced by the reference's own Board (compiled unmodified into
  oracle/_ref) through oracle/ref_harness.cpp::ref_encoded_states, which restates only the
  plane-filling lambda of core/py_ext/src/game_ext.hpp:87-104;
* ``augment``: outputs of the reference's own Python ``augment_game_data``
  (network/data_helper.py:36-55).  The module itself cannot be imported here (it imports the
  compiled `core` package and TensorFlow), so the function's source is taken from the file with
  ``ast`` and executed unmodified with numpy and the reference's Game = {15, 15} config.

    python tests/golden/make_golden_features.py
"""
import ast
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle  # noqa: E402
from conftest import random_positions  # noqa: E402

DATA_HELPER = "/root/reference/network/data_helper.py"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_augment():
    tree = ast.parse(open(DATA_HELPER).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "augment_game_data")
    ns = {"np": np, "Game": {"width": 15, "height": 15}}
    exec(compile(ast.Module([fn], []), DATA_HELPER, "exec"), ns)
    return ns["augment_game_data"]


def main():
    ref = pyoracle.ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    lists = [[], [112], [112, 113]] + random_positions(20261018, 29, lo=1, hi=120)
    enc = []
    for mv in lists:
        planes = ref.encoded_states(mv)
        enc.append({"moves": [int(m) for m in mv], "sha256": sha(planes),
                    "plane_sums": [int(x) for x in planes.reshape(6, -1).sum(1)]})
    # one full sample, spelled out
    enc[5]["planes"] = ref.encoded_states(lists[5]).reshape(6, 225).tolist()

    augment = reference_augment()
    rng = np.random.default_rng(7)
    aug = []
    for mv in lists[:12]:
        planes = ref.encoded_states(mv)
        probs = rng.random(225).astype(np.float32)
        out = augment([(planes, np.array(1.0), probs)])
        assert len(out) == 8
        st = np.stack([o[0] for o in out])
        pr = np.stack([o[2] for o in out])
        aug.append({"moves": [int(m) for m in mv], "probs_seed": 7, "states_sha256": sha(st.astype(np.uint8)),
                    "probs_sha256": sha(pr.astype(np.float32)),
                    "probs_first8": [[float(x) for x in row[:8]] for row in pr]})
    json.dump({"_source": "oracle/_ref (reference Board) + network/data_helper.py::augment_game_data executed unmodified",
               "encoded_states": enc, "augment": aug},
              open(os.path.join(HERE, "reference_features.json"), "w"), indent=0)
    print("wrote reference_features.json:", len(enc), "positions,", len(aug), "augmented samples")


if __name__ == "__main__":
    main()

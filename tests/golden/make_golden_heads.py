"""Generates tests/golden/reference_heads.json (run in the build container only): the policy heads of
Heuristic.hpp:16-45 evaluated on the density / score arrays of the COMPILED reference evaluator
(oracle/_ref), through the numpy restatement oracle/pyoracle.py::policy_heads (Heuristic.hpp needs real
Eigen, absent here, so the three formulas cannot be compiled from the reference).

    python tests/golden/make_golden_heads.py
"""
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


def main():
    ref = pyoracle.ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    lists = [[], [112], [0], [112, 113, 97]] + random_positions(77, 60, lo=2, hi=100)
    items = []
    for m in lists:
        if ref.eval_moves(m)["winner"] != 0:
            continue
        probs, value = pyoracle.policy_heads(ref, m)
        top = np.argsort(-probs, kind="stable")[:12]
        items.append({"moves": [int(x) for x in m], "value": float(value), "l1": float(np.abs(probs).sum()),
                      "top_cells": [int(c) for c in top], "top_probs": [float(probs[c]) for c in top]})
    json.dump({"_source": "oracle/_ref evaluator state + oracle/pyoracle.py::policy_heads (Heuristic.hpp:16-45)",
               "positions": items}, open(os.path.join(HERE, "reference_heads.json"), "w"), indent=0)
    print("wrote reference_heads.json:", len(items), "positions")


if __name__ == "__main__":
    main()

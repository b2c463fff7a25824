"""Generates tests/golden/reference_heads.json (run in the build container only, where /root/reference exists):
outputs of the reference's OWN Heuristic.hpp / Traditional.h, compiled unmodified into oracle/_ref
(oracle/Makefile, oracle/ref_harness_search.cpp) --

  positions[i]: Heuristic::EvaluationProbs / EvaluationValue (Heuristic.hpp:16-36) and
                TraditionalPolicy::hybridSimulate (Traditional.h:49-69: the same probabilities after
                Heuristic::DecisiveFilter, Heuristic.hpp:93-161) for the side to move;
  guided_max[i]: the game Heuristic::MaxEvaluatedRollout (Heuristic.hpp:61-85) plays from a position.

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
        h = ref.heads(m)
        if h is None:
            continue
        probs, value = h
        hv, hp = ref.hybrid_simulate(m)
        top = np.argsort(-probs, kind="stable")[:12]
        keep = np.flatnonzero(hp)
        items.append({"moves": [int(x) for x in m], "value": float(value), "l1": float(np.abs(probs).sum()),
                      "top_cells": [int(c) for c in top], "top_probs": [float(probs[c]) for c in top],
                      "hybrid_value": float(hv), "hybrid_cells": [int(c) for c in keep],
                      "hybrid_probs": [float(hp[c]) for c in keep]})
    games = []
    for m in [[], [112], [224, 210]] + random_positions(78, 21, lo=2, hi=60):
        if ref.heads(m) is None:
            continue
        winner, played = ref.guided_rollout_max(m)
        games.append({"moves": [int(x) for x in m], "winner": winner, "played": played})
    json.dump({"_source": "oracle/_ref: the reference's Heuristic.hpp / Traditional.h compiled unmodified (oracle/ref_harness_search.cpp)",
               "positions": items, "guided_max": games}, open(os.path.join(HERE, "reference_heads.json"), "w"), indent=0)
    print("wrote reference_heads.json:", len(items), "positions,", len(games), "guided games")


if __name__ == "__main__":
    main()

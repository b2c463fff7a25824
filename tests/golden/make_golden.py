"""Generates the committed golden fixtures in tests/golden/ (run in the build container only).

Two kinds of fixture:

* ``reference_kats.json`` -- the golden vectors the reference's OWN tests hold for this path,
  transcribed by hand below with the file:line they come from (no code is executed for these);
* ``reference_outputs.json`` -- outputs of the reference itself, compiled unmodified into
  ``oracle/_ref/libgomoku_ref.so`` (oracle/Makefile), for inputs no reference test pins:
  automaton arrays, evaluator scores / totals of fixed positions, rollout outcomes under
  injected start-index streams.

/root/reference is read here (to parse the four regression boards that exist only as comments
in core/test/evaluator_integrationtest.cpp) and nowhere else; the tests only read the JSON.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402

REF_TESTS = "/root/reference/core/test"


def fnv(vec):
    h = 2166136261
    for v in vec:
        h = ((h ^ (int(v) & 0xffffffff)) * 16777619) & 0xffffffff
    return h


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


# ---- hand-transcribed golden vectors of the reference's tests -----------------------------------------
KATS = {
    "_source": "Vigilans/GomokuAI core/test/*.cpp (transcribed)",
    # patternsearch_unittest.cpp:26-30 (fixture prototypes)
    "kat_protos": {"protos": ["-~_ooo_~", "-x^ooo_~", "-x_ooo_x"], "types": [5, 5, 4], "scores": [0, 0, 1]},
    # patternsearch_unittest.cpp:40-73 (AugmentPattern): [proto, type, score] after each stage
    "augment_reverse": [["-~_ooo_~", 5, 0], ["-x^ooo_~", 5, 0], ["-x_ooo_x", 4, 1], ["-~_ooo^x", 5, 0]],
    "augment_flip_added": [["+~_xxx_~", 5, 0], ["+o^xxx_~", 5, 0], ["+o_xxx_o", 4, 1], ["+~_xxx^o", 5, 0]],
    "augment_boundary_added": [
        ["-?^ooo_~", 5, 0], ["-?_ooo_x", 4, 1], ["-?_ooo_?", 4, 1], ["-x_ooo_?", 4, 1], ["-~_ooo^?", 5, 0],
        ["+?^xxx_~", 5, 0], ["+?_xxx_o", 4, 1], ["+?_xxx_?", 4, 1], ["+o_xxx_?", 4, 1], ["+~_xxx^?", 5, 0]],
    # patternsearch_unittest.cpp:144-149 (DoubleArrayTrie): strings that are / are not root-to-leaf paths
    "dat_positive": ["x_ooo_x", "?_ooo_x", "?_xxx_?", "_~xxx^o"],
    "dat_negative": ["xoooo_o", "x_oxo_x", "?_oooox", "x_oo"],
    # patternsearch_unittest.cpp:193-200 (ACFailPointers): fail[travel(b)] == travel(a)
    "fail_identities": [["", ""], ["", "o"], ["-", "o_"], ["x", "o_x"], ["x", "o_xx"], ["x", "o_xxx"], ["x-", "o_xxx_"], ["x-o", "o_xxx_o"]],
    # patternsearch_unittest.cpp:206-222 (PatternMatch)
    "match_target": "??-xxx-ooo-xxx-o-xxx--xxx-?",
    "match_expected": [["?-xxx-o", 7], ["x-ooo-x", 11], ["o-xxx-o", 15], ["o-xxx--", 21], ["--xxx-?", 26]],
    # patternsearch_unittest.cpp:225-252 (InvariantState) on the production table
    "invariant_paths": {"x": "xxxxx", "o": "ooooo", "?": "?", "-": "----"},
    # boardmap_unittest.cpp:21-33 (InitialLineView): [x, y, dir, window]
    "initial_views": [
        [7, 7, 0, "-------------"], [7, 7, 1, "-------------"], [7, 7, 2, "-------------"], [7, 7, 3, "-------------"],
        [0, 0, 0, "??????-------"], [0, 0, 1, "??????-------"], [0, 0, 2, "??????-------"], [0, 0, 3, "??????-??????"],
        [1, 2, 0, "?????--------"], [1, 2, 1, "????---------"], [1, 2, 2, "?????--------"], [1, 2, 3, "????----?????"]],
    # boardmap_unittest.cpp:35-72 (UpdateMove): kifu and [moves applied, x, y, dir, window]
    "kifu": [[7, 7], [8, 7], [7, 6], [7, 8], [6, 9]],
    "kifu_views": [
        [1, 7, 7, 0, "------x------"], [1, 7, 7, 1, "------x------"], [1, 7, 7, 2, "------x------"], [1, 7, 7, 3, "------x------"],
        [2, 7, 7, 0, "------xo-----"], [2, 7, 7, 1, "------x------"], [2, 7, 7, 2, "------x------"], [2, 7, 7, 3, "------x------"],
        [3, 7, 7, 0, "------xo-----"], [3, 7, 7, 1, "-----xx------"], [3, 8, 7, 2, "-----xo------"], [3, 7, 6, 2, "------xo-----"],
        [3, 7, 7, 2, "------x------"], [3, 7, 7, 3, "------x------"],
        [4, 7, 7, 1, "-----xxo-----"], [4, 7, 8, 3, "-----oo------"], [4, 7, 7, 2, "------x------"], [4, 7, 7, 3, "------x------"],
        [5, 7, 8, 3, "-----oox-----"]],
    # board_integrationtest.cpp:70,83 (CheckVictory): [(x, y)...] -> winner
    "black_win": [[3, 3], [3, 4], [4, 4], [3, 5], [5, 5], [3, 6], [6, 6], [3, 7], [7, 7]],
    "white_win": [[3, 3], [3, 4], [4, 4], [3, 5], [5, 5], [3, 6], [6, 6], [3, 7], [8, 8], [3, 8]],
    # board_integrationtest.cpp:98-123 (CheckTie): row order j -> y = 2j (j <= 7) else 2(j-7)-1, x ascending; ends in a draw
    "tie_row_order": [0, 2, 4, 6, 8, 10, 12, 14, 1, 3, 5, 7, 9, 11, 13],
}


def parse_regression_boards():
    """The four ASCII boards in the comments of core/test/evaluator_integrationtest.cpp:14-107."""
    text = open(os.path.join(REF_TESTS, "evaluator_integrationtest.cpp"), "rb").read().decode("latin-1")
    boards, cur = [], []
    for line in text.splitlines():
        m = re.match(r"^([0-9a-e]) ((?:[_xo] ?){15})\s*$", line)
        if m:
            cur.append(m.group(2).replace(" ", ""))
            if len(cur) == 15:
                boards.append(cur)
                cur = []
    assert len(boards) == 4, len(boards)
    out = []
    for rows in boards:
        black = [y * 15 + x for y in range(15) for x in range(15) if rows[y][x] == "x"]
        white = [y * 15 + x for y in range(15) for x in range(15) if rows[y][x] == "o"]
        assert len(black) - len(white) in (0, 1)
        moves = []
        for i in range(len(black)):
            moves.append(black[i])
            if i < len(white):
                moves.append(white[i])
        out.append(moves)
    return out


def g1_moves():
    """SURVEY.md Appendix B position G1: std::mt19937 rng(12345); 60x { do id = rng() % 225 while occupied }."""
    mt = np.zeros(624, np.uint32)
    mt[0] = 12345
    for i in range(1, 624):
        mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xffffffff
    bg = np.random.MT19937()
    st = bg.state
    st["state"]["key"] = mt
    st["state"]["pos"] = 624
    bg.state = st
    occ, moves = set(), []
    for _ in range(60):
        while True:
            i = int(bg.random_raw()) % 225
            if i not in occ:
                break
        occ.add(i)
        moves.append(i)
    return moves


def philox_r_stream(port, key, position, rollout, ctr_hi, n):
    out = []
    for k in range(n):
        if k % 4 == 0:
            w = port.philox([k >> 2, rollout, position, ctr_hi], [key & 0xffffffff, key >> 32])
        out.append((w[k & 3] * 225) >> 32)
    return out


def main():
    ref = pyoracle.ref()
    port = pyoracle.port()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)

    out = {"_source": "oracle/_ref/libgomoku_ref.so = reference sources compiled unmodified (g++ 13.3, libstdc++)"}
    t = ref.table()
    used = np.nonzero(t["check"] >= 0)[0]
    hist = [0] * 9
    for p in t["patterns"]:
        hist[p[2]] += 1
    out["automaton"] = {
        "size": int(len(t["base"])), "used_max_slot": int(used.max()), "n_patterns": len(t["patterns"]),
        "invariants": t["invariants"].tolist(), "base_sha256": sha(t["base"]), "check_sha256": sha(t["check"]),
        "fail_sha256": sha(t["fail"]), "type_histogram": hist,
        "patterns_sha256": hashlib.sha256(json.dumps(t["patterns"]).encode()).hexdigest(),
        "five_ids": [i for i, p in enumerate(t["patterns"]) if p[2] == 8],
        # the three 7-symbol patterns the libstdc++-ordered trie cannot reach, and what their prefix reports
        "aliased": [[i, t["patterns"][i][0]] for i in (217, 218, 238, 239, 250, 251)],
    }

    positions = {"G1": g1_moves()}
    for i, mv in enumerate(parse_regression_boards()):
        positions[f"R{i + 1}"] = mv
    rng = np.random.default_rng(20261018)
    for i in range(24):                                  # random and clustered positions, some terminal
        n = int(rng.integers(6, 110))
        perm = rng.permutation(225)[:n].tolist()
        if i % 3 == 2:                                   # clustered around the centre
            perm = [int(c) for c in rng.permutation([y * 15 + x for y in range(3, 12) for x in range(3, 12)])[:min(n, 70)]]
        positions[f"P{i}"] = perm
    evals = {}
    for name, mv in positions.items():
        r = ref.eval_moves(mv)
        assert r["bad"] == 0
        # a terminal position stops accepting moves: keep only the moves the reference applied
        applied = ref.board_play(mv)["applied"] if r["winner"] != 0 else len(mv)
        evals[name] = {
            "moves": [int(m) for m in mv], "applied": int(applied),
            "score_sums": r["scores"].sum(axis=1).tolist(), "score_fnv": [fnv(r["scores"][g]) for g in range(4)],
            "pat_totals": r["pat_totals"].tolist(), "cmp_totals": r["cmp_totals"].tolist(),
            "winner": int(r["winner"]), "cur_player": int(r["cur_player"]),
        }
    evals["R1"]["scores"] = ref.eval_moves(positions["R1"])["scores"].tolist()
    out["eval"] = evals

    # rollouts: Philox-injected streams (the protocol of include/gomoku_b200.h) through the reference's Board
    key = 0x474F4D4F4B5531
    roll = []
    for pi, name in enumerate(["G1", "R1", "R2", "R3", "R4", "P0", "P1", "P5"]):
        mv = positions[name]
        if evals[name]["winner"] != 0:
            continue
        res = []
        for j in range(16):
            rs = philox_r_stream(port, key, pi, j, 0, 232)
            w, n = ref.rollout_injected(mv, rs)
            res.append([int(w), int(n)])
        roll.append({"name": name, "position_index": pi, "outcomes": res})
    out["rollout_philox"] = {"key": key, "ctr_hi": 0, "cases": roll}
    # explicit streams (any bytes 0..224): outcome under the reference's Board
    ex = []
    for j in range(12):
        mv = positions["R2"] if j % 2 else []
        rs = rng.integers(0, 225, size=232).astype(np.uint8)
        w, n = ref.rollout_injected(mv, rs)
        ex.append({"moves": [int(m) for m in mv], "r": rs.tolist(), "winner": int(w), "length": int(n)})
    out["rollout_explicit"] = ex
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(out, f)
    print("G1 sums", evals["G1"]["score_sums"], "fnv", evals["G1"]["score_fnv"])
    for k in ("R1", "R2", "R3", "R4"):
        print(k, evals[k]["score_sums"], evals[k]["score_fnv"], evals[k]["pat_totals"], evals[k]["cmp_totals"])


if __name__ == "__main__":
    main()

"""The C restatement (oracle/gomoku_oracle.c) against the golden vectors the reference's own tests
hold (tests/golden/reference_kats.json) and against committed outputs of the compiled reference
(tests/golden/reference_outputs.json).  CPU only."""
import hashlib
import json

import numpy as np

from conftest import codes_of, encode, fnv

TYPES = {"DeadOne": 0, "LiveOne": 1, "DeadTwo": 2, "LiveTwo": 3, "DeadThree": 4, "LiveThree": 5, "DeadFour": 6, "LiveFour": 7, "Five": 8}


def _pat(entry):
    proto, typ, score = entry
    return (proto[1:], 1 if proto[0] == "+" else -1, typ, score)


def test_augmentation_golden_list(port, kats):
    # core/test/patternsearch_unittest.cpp:40-73
    k = kats["kat_protos"]
    stage1 = [_pat(e) for e in kats["augment_reverse"]]
    assert port.augment(k["protos"], k["types"], k["scores"], 1) == stage1
    stage2 = stage1 + [_pat(e) for e in kats["augment_flip_added"]]
    assert port.augment(k["protos"], k["types"], k["scores"], 2) == stage2
    stage3 = stage2 + [_pat(e) for e in kats["augment_boundary_added"]]
    assert port.augment(k["protos"], k["types"], k["scores"], 3) == stage3


def test_sort_is_lexicographic_on_codes_for_kat_set(port, kats):
    # core/test/patternsearch_unittest.cpp:75-87
    k = kats["kat_protos"]
    got = port.augment(k["protos"], k["types"], k["scores"], 4)
    want = sorted(port.augment(k["protos"], k["types"], k["scores"], 3), key=lambda p: codes_of(p[0]))
    assert got == want


def _travel(t, s):
    st = 0
    for ch in s:
        if t["check"][t["base"][st]] == st:      # reached a leaf-bearing node: the walk of the unit test stops
            break
        nx = t["base"][st] + codes_of(ch)[0]
        if t["check"][nx] != st:
            return st, False
        st = nx
    return st, True


def test_double_array_trie_paths(port, kats):
    # core/test/patternsearch_unittest.cpp:136-169
    k = kats["kat_protos"]
    port.build_custom(k["protos"], k["types"], k["scores"])
    t = port.table(custom=True)

    def match(s):
        st = 0
        for ch in s:
            if t["check"][t["base"][st]] == st:
                break
            nx = t["base"][st] + codes_of(ch)[0]
            if t["check"][nx] != st:
                return False
            st = nx
        return t["check"][t["base"][st]] == st
    for s in kats["dat_positive"]:
        assert match(s), s
    for s in kats["dat_negative"]:
        assert not match(s), s


def test_fail_pointer_identities(port, kats):
    # core/test/patternsearch_unittest.cpp:171-201
    k = kats["kat_protos"]
    port.build_custom(k["protos"], k["types"], k["scores"])
    t = port.table(custom=True)
    for a, b in kats["fail_identities"]:
        assert _travel(t, a)[0] == t["fail"][_travel(t, b)[0]], (a, b)


def test_pattern_match_kat(port, kats):
    # core/test/patternsearch_unittest.cpp:204-223
    k = kats["kat_protos"]
    port.build_custom(k["protos"], k["types"], k["scores"])
    pats = port.table(custom=True)["patterns"]
    got = [(codes_of(pats[pid][0]), off) for pid, off in port.scan(encode(kats["match_target"]), custom=True)]
    want = [(codes_of(s), off) for s, off in kats["match_expected"]]
    assert got == want


def test_invariant_states_of_production_table(port, kats):
    # core/test/patternsearch_unittest.cpp:225-252
    t = port.table()
    for sym, path in kats["invariant_paths"].items():
        st = 0
        for ch in path:
            st = t["base"][st] + codes_of(ch)[0]
        assert st == t["invariants"][codes_of(sym)[0]]


def test_line_views(port, kats):
    # core/test/boardmap_unittest.cpp:21-72
    for x, y, d, window in kats["initial_views"]:
        assert port.line_view([], y * 15 + x, d).tolist() == codes_of(window)
    kifu = [y * 15 + x for x, y in kats["kifu"]]
    for n, x, y, d, window in kats["kifu_views"]:
        assert port.line_view(kifu[:n], y * 15 + x, d).tolist() == codes_of(window), (n, x, y, d)


def test_win_and_draw_sequences(port, kats):
    # core/test/integration/board_integrationtest.cpp:66-123
    b = port.board_play([y * 15 + x for x, y in kats["black_win"]])
    assert (b["cur_player"], b["winner"], b["applied"]) == (0, 1, 9)
    w = port.board_play([y * 15 + x for x, y in kats["white_win"]])
    assert (w["cur_player"], w["winner"], w["applied"]) == (0, -1, 10)
    tie = [y * 15 + x for y in kats["tie_row_order"] for x in range(15)]
    t = port.board_play(tie)
    assert (t["cur_player"], t["winner"], t["applied"]) == (0, 0, 225)
    short = port.board_play(tie[:-1])
    assert short["cur_player"] != 0 and short["applied"] == 224


def test_invalid_move_is_a_no_op(port):
    # core/lib/src/Game.cpp:37-47 and board_integrationtest.cpp:147-153
    r = port.board_play([112, 112, 113, -1, 225, 113])
    assert r["applied"] == 2 and r["cur_player"] == 1


def test_automaton_matches_compiled_reference_fingerprint(port, ref_outputs):
    a = ref_outputs["automaton"]
    t = port.table()
    sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    assert len(t["base"]) == a["size"] and len(t["patterns"]) == a["n_patterns"]
    assert int(np.nonzero(t["check"] >= 0)[0].max()) == a["used_max_slot"]
    assert t["invariants"].tolist() == a["invariants"]
    assert sha(t["base"]) == a["base_sha256"] and sha(t["check"]) == a["check_sha256"] and sha(t["fail"]) == a["fail_sha256"]
    assert hashlib.sha256(json.dumps([list(p) for p in t["patterns"]]).encode()).hexdigest() == a["patterns_sha256"]
    hist = [0] * 9
    for p in t["patterns"]:
        hist[p[2]] += 1
    assert hist == a["type_histogram"]
    assert [[i, t["patterns"][i][0]] for i, _ in a["aliased"]] == a["aliased"]


def test_evaluator_matches_compiled_reference_outputs(port, ref_outputs):
    for name, g in ref_outputs["eval"].items():
        r = port.eval_moves(g["moves"])
        assert r["bad"] == 0, name
        assert r["scores"].sum(axis=1).tolist() == g["score_sums"], name
        assert [fnv(r["scores"][k]) for k in range(4)] == g["score_fnv"], name
        assert r["pat_totals"].tolist() == g["pat_totals"], name
        assert r["cmp_totals"].tolist() == g["cmp_totals"], name
        assert int(r["winner"]) == g["winner"] and int(r["cur_player"]) == g["cur_player"], name
    full = np.array(ref_outputs["eval"]["R1"]["scores"], np.int32)
    assert np.array_equal(port.eval_moves(ref_outputs["eval"]["R1"]["moves"])["scores"], full)


def test_evaluator_self_check_invariants(port, ref_outputs):
    # the reference's always-on self-check (Pattern.cpp:314-333): occupied cells score 0, empties >= 0
    for name, g in ref_outputs["eval"].items():
        r = port.eval_moves(g["moves"])
        occ = np.zeros(225, bool)
        occ[g["moves"][:g["applied"]]] = True
        assert (r["scores"][:, occ] == 0).all(), name
        assert (r["scores"] >= 0).all(), name


def test_apply_revert_symmetry(port, ref_outputs):
    # board_integrationtest.cpp:46-64 style, for the evaluator: apply n, revert k == apply n-k
    mv = ref_outputs["eval"]["G1"]["moves"]
    for k in (1, 5, 17, 60):
        a = port.eval_apply_revert(mv, k)
        b = port.eval_moves(mv[:len(mv) - k])
        assert a["bad"] == 0
        assert np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["pat_totals"], b["pat_totals"])
        assert np.array_equal(a["cmp_totals"], b["cmp_totals"])


def test_rollout_outcomes_match_compiled_reference(port, ref_outputs):
    g = ref_outputs["rollout_philox"]
    for case in g["cases"]:
        mv = ref_outputs["eval"][case["name"]]["moves"]
        for j, (w, n) in enumerate(case["outcomes"]):
            rs = []
            for k in range(232):
                if k % 4 == 0:
                    words = port.philox([k >> 2, j, case["position_index"], g["ctr_hi"]], [g["key"] & 0xffffffff, g["key"] >> 32])
                rs.append((words[k & 3] * 225) >> 32)
            assert port.rollout_injected(mv, rs) == (w, n), (case["name"], j)
        import oracle.pyoracle as po
        mvs, st = po.pack_moves([mv])
        wn, ln, wdb = port.rollout_philox_batch(mvs, st, 16, g["key"], g["ctr_hi"], case["position_index"])
        assert wn[0].tolist() == [o[0] for o in case["outcomes"]] and ln[0].tolist() == [o[1] for o in case["outcomes"]]
    for ex in ref_outputs["rollout_explicit"]:
        assert port.rollout_injected(ex["moves"], ex["r"]) == (ex["winner"], ex["length"])


def test_philox_known_answer(port):
    # Random123 kat_vectors: philox4x32-10, counter 0 / key 0 and the all-ones vector
    assert port.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert port.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_exhaustive_single_line_theorems(port):
    """Every line of length 5..10 over {x, o, blank}: the facts that make a from-scratch evaluation of
    the final position equal to the reference's incremental replay (oracle/gomoku_oracle.c,
    orc_line_theorems).  The same enumeration up to length 15 (all lines a board can hold) was run
    once while writing DESIGN.md section 2.2."""
    import ctypes
    out = (ctypes.c_long * 8)()
    port.lib.orc_line_theorems(10, out)
    t1, t2, t4, t5, lines, emissions, t3, five_moves = list(out)
    assert lines == sum(3 ** n for n in range(5, 11)) and emissions > 50000
    assert (t1, t2, t3, t4, t5) == (0, 0, 0, 0, 0)
    assert five_moves > 0          # Five emissions do move with the run length (they only set the winner)

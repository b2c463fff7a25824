"""The product's table compiler (gomokuai_b200/csrc/gk_table.cpp) against the oracle's generator:
the flat transducer must produce the same (pattern, offset) stream on every string.  The table is
walked here in Python (test code); no product compute path runs on the CPU."""
import numpy as np

from test_oracle_vs_ref import _strings


def _walk(entries, flush, codes):
    st, out = 0, []
    for i, c in enumerate(codes):
        w = int(entries[st, c - 1])
        st = w & 1023
        for k in range((w >> 10) & 3):
            em = (w >> (12 + 10 * k)) & 1023
            out.append((em & 511, i - ((em >> 9) & 1)))
    if flush[st] >= 0:
        out.append((int(flush[st]), len(codes) - 1))
    return out


def test_default_table_shape(gk, port):
    t = gk.default_table()
    info = t.info()
    assert info == {"n_states": 558, "n_patterns": 294, "trail_pad": 2, "tape_steps": 34}
    assert t.patterns() == port.table()["patterns"]        # same ids as PatternSearch::m_patterns
    e = t.entries()
    hist = np.bincount(((e >> 10) & 3).ravel(), minlength=4)
    assert hist[3] == 0 and hist[2] == 39                  # at most two emissions per (state, symbol)
    assert (t.flush() >= 0).sum() == 2                     # only the xxxxx / ooooo run states owe an emission at end of input


def test_emissions_equal_oracle_generator(gk, port):
    t = gk.default_table()
    entries, flush = t.entries(), t.flush()
    strings = _strings(17, 30000)
    # long runs of five-or-more, also reaching the end of the input (the invariant-state rule)
    for tail in ([1] * 5, [1] * 9, [2] * 6 + [4], [4, 1, 1, 1, 1, 1, 1, 2], [2] * 12):
        strings.append(np.array([3, 4] + tail, np.uint8))
    starts = np.zeros(len(strings) + 1, np.int64)
    starts[1:] = np.cumsum([len(s) for s in strings])
    pids, offs, counts = port.scan_many(np.concatenate(strings), starts)
    at = 0
    for i, s in enumerate(strings):
        want = list(zip(pids[at:at + counts[i]].tolist(), offs[at:at + counts[i]].tolist()))
        at += counts[i]
        assert _walk(entries, flush, s.tolist()) == want, s.tolist()


def test_custom_table_equals_oracle(gk, port, kats):
    k = kats["kat_protos"]
    t = gk.Table(k["protos"], k["types"], k["scores"])
    port.build_custom(k["protos"], k["types"], k["scores"])
    assert t.patterns() == port.table(custom=True)["patterns"]
    entries, flush = t.entries(), t.flush()
    from conftest import encode
    pats = t.patterns()
    got = [(pats[p][0], o) for p, o in _walk(entries, flush, encode(kats["match_target"]).tolist())]
    from conftest import codes_of
    assert [(codes_of(s), o) for s, o in got] == [(codes_of(s), o) for s, o in kats["match_expected"]]
    rng = np.random.default_rng(2)
    for _ in range(3000):
        s = rng.integers(1, 5, size=int(rng.integers(1, 40))).astype(np.uint8)
        assert _walk(entries, flush, s.tolist()) == port.scan(s, custom=True)


def test_bad_prototypes_are_rejected(gk):
    import pytest
    for protos in (["xxxxx"], ["+xx#xx"], ["+"], ["+xxxxxxxxx"]):
        with pytest.raises(gk.GomokuB200Error):
            gk.Table(protos, [0] * len(protos), [1] * len(protos))


def test_pack_and_synth(gk, port):
    boards, moves, starts = gk.synth_positions(0, 300)
    n = np.diff(starts)
    assert n.min() >= 16 and n.max() <= 96
    assert np.array_equal(gk.pack_moves(moves, starts), boards)
    again, m2, s2 = gk.synth_positions(100, 50)
    assert np.array_equal(again, boards[100:150])            # position i depends on i only
    cells = gk.unpack_boards(boards)
    assert ((cells == 1).sum(1) - (cells == 2).sum(1) >= 0).all() and ((cells == 1).sum(1) - (cells == 2).sum(1) <= 1).all()
    # non-terminal by construction: the reference's evaluator reports no winner
    r = port.eval_batch(moves, starts, want_scores=False)
    assert r["bad"] == 0 and (r["winner"] == 0).all() and (r["cur_player"] != 0).all()


def _walk_device(dev, codes, start=None):
    """The eval kernel's scan loop in Python: one 16-bit load per symbol, emission iff offset < n_clones * 8."""
    nxt, erec, thr = dev["next"], dev["erec"], dev["n_clones"] * 8
    off, out, emitting_steps = (dev["root_off"] if start is None else start), [], 0
    for i, c in enumerate(codes):
        off = int(nxt[off >> 3, c % 4])                    # reference code 1..4 -> raw cell value 1, 2, 3, 0
        if off < thr:
            emitting_steps += 1
            e = int(erec[off >> 3])
            for k in range(2):
                pid = (e >> (10 * k)) & 511
                if pid != 511:
                    out.append((pid, i - ((e >> (9 + 10 * k)) & 1)))
    return out, emitting_steps


def test_device_format_equals_host_format(gk):
    """cloned-arrival table (what the kernel reads) == flat transducer (what the oracle pins), on random strings"""
    t = gk.default_table()
    entries, dev = t.entries(), t.device_format()
    assert dev["next"].shape[0] == dev["n_clones"] + 558 and dev["n_clones"] <= 1024
    assert dev["tape_steps"] == 34 and dev["list_cap"] == 18 and dev["root_off"] == dev["n_clones"] * 8
    no_flush = np.full(558, -1)
    for s in _strings(23, 20000):
        codes = s.tolist()
        assert _walk_device(dev, codes)[0] == _walk(entries, no_flush, codes), codes


def test_list_capacity_bounds_emitting_steps(gk):
    """no lane chain of the tape can emit on more steps than the per-lane list holds: adversarial +
    random lines through the device table, from the line-start state, with the two trailing pads"""
    dev = gk.default_table().device_format()
    rng = np.random.default_rng(5)
    worst = 0
    dense = [[4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4], [1, 4] * 7 + [1], [4, 4, 1, 4, 4, 1, 4, 4, 1, 4, 4, 1, 4, 4, 1],
             [2, 1, 4, 4, 4, 1, 4, 4, 4, 1, 4, 4, 4, 1, 2]]
    lines = dense + [rng.choice([1, 2, 4, 4], size=15).tolist() for _ in range(20000)]
    for line in lines:
        _, steps = _walk_device(dev, line + [3, 3], start=dev["start_off"])
        worst = max(worst, steps)
    assert worst <= 9                                       # 15-cell line: at most 9 emitting steps (table compiler's DP)
    assert 2 * worst <= dev["list_cap"]


def test_custom_table_device_format(gk, kats):
    k = kats["kat_protos"]
    t = gk.Table(k["protos"], k["types"], k["scores"])
    entries, dev = t.entries(), t.device_format()
    no_flush = np.full(entries.shape[0], -1)
    rng = np.random.default_rng(2)
    for _ in range(3000):
        codes = rng.integers(1, 5, size=int(rng.integers(1, 30))).tolist()
        assert _walk_device(dev, codes)[0] == _walk(entries, no_flush, codes)

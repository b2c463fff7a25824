"""ORACLE -- test infrastructure, not product code.

ctypes front-end shared by the two CPU checkers:

* ``port()``  -> oracle/libgomoku_oracle.so, the plain-C restatement (gomoku_oracle.c)
* ``ref()``   -> oracle/_ref/libgomoku_ref.so, the reference's own sources compiled by
                 oracle/Makefile (None when it has not been built)

Both expose the same Python methods so a test can run the same check against either.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libgomoku_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgomoku_ref.so")
REFERENCE_ROOT = "/root/reference/core/lib"
PYREF_ZIP = os.path.join(_HERE, "_ref", "pyref.zip")     # the reference's agents/ + config.py as sourceless byte code (make pyref)


def build(want_ref=True):
    """Compile the C restatement, and oracle/_ref when /root/reference is present."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    if want_ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-s", "-C", _HERE, "ref", "pyref"], check=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pack_moves(move_lists):
    """list of move lists -> (int16 moves, int64 starts)"""
    starts = np.zeros(len(move_lists) + 1, np.int64)
    for i, m in enumerate(move_lists):
        starts[i + 1] = starts[i] + len(m)
    moves = np.zeros(max(int(starts[-1]), 1), np.int16)
    for i, m in enumerate(move_lists):
        moves[starts[i]:starts[i + 1]] = m
    return moves, starts


class Oracle:
    def __init__(self, path, prefix, kind):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.kind = kind          # "port" | "reference"
        self._table = None
        if prefix == "orc_":
            self.lib.orc_table_default.restype = ctypes.c_void_p
            self.lib.orc_table_build.restype = ctypes.c_void_p
            self._table = ctypes.c_void_p(self.lib.orc_table_default())
            self._default = self._table
        for name in ("scan_many", "linescan_batch"):
            getattr(self.lib, prefix + name).restype = ctypes.c_long

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _which(self, custom):
        # the port passes a table pointer, the reference harness a selector
        if self.prefix == "orc_":
            return self._table if custom else self._default
        return 1 if custom else 0

    # ---- automaton -------------------------------------------------------------------
    def build_custom(self, protos, types, scores):
        n = len(protos)
        arr = (ctypes.c_char_p * n)(*[p.encode() for p in protos])
        t = (ctypes.c_int * n)(*types)
        s = (ctypes.c_int * n)(*scores)
        if self.prefix == "orc_":
            self._table = ctypes.c_void_p(self.lib.orc_table_build(arr, t, s, n))
        else:
            self.lib.ref_table_build_custom(arr, t, s, n)

    def table(self, custom=False):
        nb, npat = ctypes.c_int(), ctypes.c_int()
        self._f("table_sizes")(self._which(custom), ctypes.byref(nb), ctypes.byref(npat))
        base = np.zeros(nb.value, np.int32)
        check, fail = base.copy(), base.copy()
        inv = np.zeros(5, np.int32)
        self._f("table_arrays")(self._which(custom), _p(base), _p(check), _p(fail), _p(inv))
        pats = []
        for i in range(npat.value):
            s = ctypes.create_string_buffer(8)
            fav, typ, sc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            self._f("table_pattern")(self._which(custom), i, s, ctypes.byref(fav), ctypes.byref(typ), ctypes.byref(sc))
            pats.append((s.value.decode(), fav.value, typ.value, sc.value))
        return {"base": base, "check": check, "fail": fail, "invariants": inv, "patterns": pats}

    def augment(self, protos, types, scores, stage):
        n = len(protos)
        arr = (ctypes.c_char_p * n)(*[p.encode() for p in protos])
        t = (ctypes.c_int * n)(*types)
        s = (ctypes.c_int * n)(*scores)
        cap = 12 * n + 1
        strs = ctypes.create_string_buffer(8 * cap)
        fav, typ, sc = (np.zeros(cap, np.int32) for _ in range(3))
        m = self._f("augment")(arr, t, s, n, stage, strs, _p(fav), _p(typ), _p(sc), cap)
        raw = strs.raw
        return [(raw[8 * i:8 * i + 8].split(b"\0")[0].decode(), int(fav[i]), int(typ[i]), int(sc[i])) for i in range(m)]

    def scan(self, codes, custom=False, cap=256):
        codes = np.ascontiguousarray(codes, np.uint8)
        pids, offs = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n = self._f("scan")(self._which(custom), _p(codes), len(codes), _p(pids), _p(offs), cap)
        assert n <= cap
        return list(zip(pids[:n].tolist(), offs[:n].tolist()))

    def scan_many(self, codes, starts, custom=False, cap=None):
        codes = np.ascontiguousarray(codes, np.uint8)
        starts = np.ascontiguousarray(starts, np.int64)
        ns = len(starts) - 1
        cap = cap or 4 * len(codes) + 16
        pids, offs = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        counts = np.zeros(ns, np.int32)
        total = self._f("scan_many")(self._which(custom), _p(codes), _p(starts), ns, _p(pids), _p(offs), _p(counts), ctypes.c_long(cap))
        assert total <= cap
        return pids[:total], offs[:total], counts

    # ---- line views -----------------------------------------------------------------
    def line_view(self, moves, pose, direction):
        mv = np.ascontiguousarray(moves, np.int16)
        out = np.zeros(13, np.uint8)
        self._f("line_view")(_p(mv), len(mv), int(pose), int(direction), _p(out))
        return out

    def line_map(self, moves):
        mv = np.ascontiguousarray(moves, np.int16)
        out = np.zeros(88 * 28, np.uint8)
        lens = np.zeros(88, np.int32)
        k = self._f("line_map")(_p(mv), len(mv), _p(out), _p(lens))
        lines, at = [], 0
        for ln in lens:
            lines.append(out[at:at + ln].copy())
            at += ln
        assert at == k
        return lines

    # ---- evaluator ------------------------------------------------------------------
    def eval_batch(self, moves, starts, want_scores=True):
        """-> dict(scores[N,4,225] i32, pat_totals[N,2,8] u16, cmp_totals[N,2,3] u16, winner, cur_player, bad)"""
        moves = np.ascontiguousarray(moves, np.int16)
        starts = np.ascontiguousarray(starts, np.int64)
        n = len(starts) - 1
        scores = np.zeros((n, 4, 225), np.int32) if want_scores else None
        pt = np.zeros((n, 2, 8), np.uint16)
        ct = np.zeros((n, 2, 3), np.uint16)
        win = np.zeros(n, np.int8)
        cur = np.zeros(n, np.int8)
        bad = self._f("eval_batch")(_p(moves), _p(starts), n, _p(scores), _p(pt), _p(ct), _p(win), _p(cur))
        return {"scores": scores, "pat_totals": pt, "cmp_totals": ct, "winner": win, "cur_player": cur, "bad": bad}

    def eval_moves(self, move_list):
        mv, st = pack_moves([move_list])
        r = self.eval_batch(mv, st)
        return {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in r.items()}

    def eval_flags(self):
        pf = np.zeros((225, 8), np.uint32)
        cf = np.zeros((225, 3), np.uint32)
        den = np.zeros((2, 2, 225), np.int32)
        self._f("eval_flags")(_p(pf), _p(cf), _p(den))
        return pf, cf, den

    def linescan_batch(self, moves, starts):
        moves = np.ascontiguousarray(moves, np.int16)
        starts = np.ascontiguousarray(starts, np.int64)
        return self._f("linescan_batch")(_p(moves), _p(starts), len(starts) - 1)

    # ---- board / rollout ------------------------------------------------------------
    def board_play(self, moves):
        mv = np.ascontiguousarray(moves, np.int16)
        out = np.zeros(3, np.int32)
        self._f("board_play")(_p(mv), len(mv), _p(out))
        return {"cur_player": int(out[0]), "winner": int(out[1]), "applied": int(out[2])}

    def rollout_injected(self, moves, r_stream):
        mv = np.ascontiguousarray(moves, np.int16)
        rs = np.ascontiguousarray(r_stream, np.uint8)
        n = ctypes.c_int()
        w = self._f("rollout_injected")(_p(mv), len(mv), _p(rs), len(rs), ctypes.byref(n))
        return w, n.value


# ---- self-play feature planes (row f3): numpy restatements --------------------------------------------
def encoded_states(move_list):
    """Board.encoded_states() (core/py_ext/src/game_ext.hpp:87-104): uint8[6,15,15] =
    [stones of the side to move, opponent's stones, empty cells, last move, second-to-last move,
    all-ones iff black is to move].  Black moves first and players alternate (Game.cpp:37-47)."""
    cells = np.zeros(225, np.int8)
    for k, c in enumerate(move_list):
        cells[c] = 1 if k % 2 == 0 else -1                    # Player::Black = 1, White = -1 (Game.h)
    cur = 1 if len(move_list) % 2 == 0 else -1
    out = np.zeros((6, 225), np.uint8)
    out[0] = cells == cur                                     # :91-93 moveStates(m_curPlayer)
    out[1] = cells == -cur                                    #        moveStates(-m_curPlayer)
    out[2] = cells == 0                                       #        moveStates(Player::None)
    for i in (0, 1):                                          # :94-100 one-hot of the last two moves
        if len(move_list) > i:
            out[3 + i, move_list[-1 - i]] = 1
    out[5] = cur == 1                                         # :101
    return out.reshape(6, 15, 15)


def augment_planes(states, probs):
    """augment_game_data (network/data_helper.py:36-55) for one sample: the 8 variants in the
    reference's order -- for i in 0..3: rot90(i), fliplr(rot90(i)).  -> (uint8[8,6,15,15], f32[8,225])"""
    st, pr = [], []
    for i in range(4):
        rs = np.array([np.rot90(plane, i) for plane in states])
        rp = np.rot90(probs.reshape(15, 15), i)
        st.append(rs)
        pr.append(rp.flatten())
        st.append(np.array([np.fliplr(plane) for plane in rs]))
        pr.append(np.fliplr(rp).flatten())
    return np.stack(st), np.stack(pr)


# ---- policy heads (row f1): numpy restatement over an oracle's evaluator state ---------------------------
def policy_heads(orc, move_list):
    """Heuristic::EvaluationProbs / EvaluationValue / DensityWeight (include/algorithms/Heuristic.hpp:16-45)
    for the side to move, computed from the scores and density arrays of `orc` (the C restatement or the
    compiled reference) after replaying `move_list`.  float32 throughout, as the reference's VectorXf;
    only the summation order is numpy's instead of Eigen's.  -> (probs f32[225], value f32)"""
    f32 = np.float32
    r = orc.eval_moves(move_list)
    assert r["bad"] == 0
    _, _, den = orc.eval_flags()
    den = den.reshape(2, 2, 225)                                  # [Group(player)][count | weight][cell], Pattern.h:218
    scores = r["scores"].astype(f32)                              # [Group(favour, perspective)][cell]
    p = 1 if len(move_list) % 2 == 0 else 0                       # Group(player to move); black moves first

    def density_weight(g):                                        # :40-45  normalize(3W / (1 + 2N)), max(x, 0) filter
        n = np.maximum(den[g, 0], 0).astype(f32)
        w = np.maximum(den[g, 1], 0).astype(f32)
        v = (f32(3) * w) / (f32(1) + f32(2) * n)
        z = f32(np.sum(v * v, dtype=f32))
        return v / f32(np.sqrt(z)) if z > 0 else v                # Eigen normalized(): unchanged when the norm is 0

    dw_self, dw_rival = density_weight(p), density_weight(1 - p)
    if len(move_list) > 0:                                        # :18-23
        self_worthy = scores[3 * p] * dw_self                     # scores(player, player)
        rival_anti = scores[2 * (1 - p) + p] * dw_rival           # scores(-player, player)
        a = f32(0.6) * self_worthy + f32(0.4) * rival_anti
        z = f32(np.sum(a * a, dtype=f32))
        probs = a / f32(np.sqrt(z)) if z > 0 else a
    else:                                                         # :24-27 empty board: the centre
        probs = np.zeros(225, f32)
        probs[7 * 15 + 7] = 1.0
    sw = f32(np.dot(scores[3 * p], dw_self))                      # :33-35
    rw = f32(np.dot(scores[3 * (1 - p)], dw_rival))
    value = f32(np.tanh((1.2 * float(sw) - float(rw)) / 500.0))
    return probs.astype(f32), value


def decisive_filter(orc, move_list, probs):
    """Heuristic::DecisiveFilter (include/algorithms/Heuristic.hpp:93-161) on the evaluator state of `orc` after
    replaying `move_list`: walk the priority automaton +4 > -4 > +L3 == +To44 > -L3 == -To44 >= +To43 > -To43 >
    +To33 > -To33 over the pattern / compound totals; at the first class with a non-zero count keep only the cells
    whose per-cell flag Record::get(player, cur_player) is set for one of the remaining candidates, re-normalise.
    Returns (filtered probs f32[225], candidates [(pattern index, player)]) -- candidates is empty when nothing fired."""
    f32 = np.float32
    r = orc.eval_moves(move_list)
    pf, cf, _ = orc.eval_flags()                                   # per-cell Record fields: [225][8], [225][3]
    cur = 1 if len(move_list) % 2 == 0 else -1                     # Player::Black = 1, White = -1
    SIZE, L4, D4, L3, D3 = 9, 7, 6, 5, 4                            # Pattern::Type (include/Pattern.h:19-24)
    S4, SL3, TO44, TO43, TO33, END = range(6)                      # Heuristic.hpp:98
    table = [[(S4, 1), (TO44, 0), (SL3, 1), (TO43, 1), (TO33, 1), (END, 0)],      # :101-105
             [(SL3, 0), (TO44, 1), (TO43, 0), (TO33, 0), (END, 0), (END, 1)]]

    def count(pattern, player):                                     # :125-130 totals, [0 = White, 1 = Black]
        g = 1 if player == 1 else 0
        return int(r["pat_totals"][g][pattern]) if pattern < SIZE else int(r["cmp_totals"][g][pattern - SIZE])

    def flag(cell, pattern, player):                                # :147-151 Record::get(favour, perspective), Pattern.cpp:408-411
        group = ((player == 1) << 1) | (cur == 1)
        field = int(pf[cell][pattern]) if pattern < SIZE else int(cf[cell][pattern % SIZE])
        return (field >> (8 * group)) & 0xff

    probs = np.array(probs, f32)
    state, anti, cand = S4, 0, []
    while state != END:
        player = -cur if anti else cur
        if state == S4:
            cand += [(L4, player), (D4, player)]
        elif state == SL3:
            cand += [(L3, player)]
        else:
            cand += [(SIZE + (TO33 - state), player)]
        while cand:
            pattern, pl = cand[0]
            if count(pattern, pl) != 0:
                if anti and state != S4:
                    cand.append((D3, -pl))
                break
            cand.pop(0)
        if cand:
            keep = np.array([any(flag(i, pt, pl) for pt, pl in cand) for i in range(225)])
            probs = np.where(keep, probs, f32(0)).astype(f32)
            z = f32(np.sum(probs * probs, dtype=f32))
            if z > 0:
                probs = probs / f32(np.sqrt(z))
            return probs, cand
        state, anti = table[anti][state]
    return probs, []


def hybrid_simulate(orc, move_list):
    """TraditionalPolicy::hybridSimulate (include/policies/Traditional.h:49-69): DecisiveFilter never sets
    report.level, so the result is always {EvaluationValue, filtered EvaluationProbs}."""
    probs, value = policy_heads(orc, move_list)
    probs, _ = decisive_filter(orc, move_list, probs)
    return value, probs


def guided_rollout(orc, port_oracle, move_list, mode, key, game, ctr_hi=0, max_moves=225):
    """Heuristic::EvaluatedRollout (include/algorithms/Heuristic.hpp:61-91) from the position after `move_list`:
    while the evaluator has no winner and the board is not full, play the move chosen from policy_heads --
    mode "max": first maximum (MaxEvaluatedRollout); mode "sample": the documented quantised draw of
    include/gomoku_b200.h (weights round(p * 2^20), r = mulhi32(Philox word, sum)).  -> (winner, moves played)"""
    moves, played = list(move_list), []
    f32 = np.float32
    while True:
        r = orc.eval_moves(moves)
        if r["winner"] != 0:
            return int(r["winner"]), played
        if len(moves) == 225 or len(played) >= max_moves:
            return 0, played
        probs, _ = policy_heads(orc, moves)
        if mode == "max":
            if not (probs.max() > 0):
                return 0, played
            cell = int(np.argmax(probs))                            # first maximum, as Eigen's maxCoeff(&index)
        else:
            w = (probs * f32(1048576.0) + f32(0.5)).astype(np.uint32)
            total = int(w.sum(dtype=np.uint64))
            if total == 0:
                return 0, played
            k = len(played)
            word = port_oracle.philox([k >> 2, 0, game, ctr_hi], [key & 0xffffffff, key >> 32])[k & 3]
            rr = (word * total) >> 32
            cell = int(np.searchsorted(np.cumsum(w.astype(np.uint64)), rr, side="right"))
        moves.append(cell)
        played.append(cell)


class PortOracle(Oracle):
    """extras only the C restatement has (Philox stream, apply/revert, degenerate counter)"""

    def __init__(self):
        super().__init__(PORT_SO, "orc_", "port")
        self.lib.orc_degenerate_compounds.restype = ctypes.c_long

    def philox(self, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        o = (ctypes.c_uint32 * 4)()
        self.lib.orc_philox4x32_10(c, k, o)
        return list(o)

    def rollout_philox_batch(self, moves, starts, rollouts_per_pos, key, ctr_hi=0, pos_base=0):
        moves = np.ascontiguousarray(moves, np.int16)
        starts = np.ascontiguousarray(starts, np.int64)
        n = len(starts) - 1
        winners = np.zeros((n, rollouts_per_pos), np.int8)
        lengths = np.zeros((n, rollouts_per_pos), np.uint8)
        wdb = np.zeros((n, 3), np.int32)
        self.lib.orc_rollout_philox_batch(_p(moves), _p(starts), n, rollouts_per_pos, ctypes.c_uint64(key),
                                          ctypes.c_uint32(ctr_hi), pos_base, _p(winners), _p(lengths), _p(wdb))
        return winners, lengths, wdb

    def eval_apply_revert(self, moves, n_revert):
        mv = np.ascontiguousarray(moves, np.int16)
        scores = np.zeros((4, 225), np.int32)
        pt = np.zeros((2, 8), np.uint16)
        ct = np.zeros((2, 3), np.uint16)
        bad = self.lib.orc_eval_apply_revert(_p(mv), len(mv), n_revert, _p(scores), _p(pt), _p(ct))
        return {"scores": scores, "pat_totals": pt, "cmp_totals": ct, "bad": bad}

    def eval_scratch_batch(self, moves, starts, lead=1, trail=2, min_len=5):
        """from-scratch model of the kernel's algorithm (same output dict as eval_batch)"""
        moves = np.ascontiguousarray(moves, np.int16)
        starts = np.ascontiguousarray(starts, np.int64)
        n = len(starts) - 1
        scores = np.zeros((n, 4, 225), np.int32)
        pt = np.zeros((n, 2, 8), np.uint16)
        ct = np.zeros((n, 2, 3), np.uint16)
        win = np.zeros(n, np.int8)
        bad = self.lib.orc_eval_scratch_batch(_p(moves), _p(starts), n, lead, trail, min_len, _p(scores), _p(pt), _p(ct), _p(win))
        return {"scores": scores, "pat_totals": pt, "cmp_totals": ct, "winner": win, "bad": bad}

    def scratch_gate_blocks(self):
        self.lib.orc_scratch_gate_blocks.restype = ctypes.c_long
        return self.lib.orc_scratch_gate_blocks()

    def degenerate_compounds(self):
        return self.lib.orc_degenerate_compounds()


class RefOracle(Oracle):
    def __init__(self):
        super().__init__(REF_SO, "ref_", "reference")

    def encoded_states(self, moves):
        mv = np.ascontiguousarray(moves, np.int16)
        out = np.zeros((6, 15, 15), np.uint8)
        self.lib.ref_encoded_states(_p(mv), len(mv), _p(out))
        return out

    def rollout_free(self, moves, n_rollouts):
        mv = np.ascontiguousarray(moves, np.int16)
        wdb = np.zeros(3, np.int64)
        total = np.zeros(1, np.int64)
        self.lib.ref_rollout_free(_p(mv), len(mv), n_rollouts, _p(wdb), _p(total))
        return wdb, int(total[0])

    # ---- everything above the evaluator: the reference's Heuristic.hpp / policies / MCTS.cpp, compiled unmodified
    # ---- (oracle/ref_harness_search.cpp) --------------------------------------------------------------------------
    def heads(self, moves, want_dw=False):
        """Heuristic::EvaluationProbs / EvaluationValue (Heuristic.hpp:16-36) for the side to move.
        -> (probs f32[225], value f32[, DensityWeight f32[2,225] = side to move, opponent]); None for terminal positions"""
        mv = np.ascontiguousarray(moves, np.int16)
        probs = np.zeros(225, np.float32)
        value = ctypes.c_float()
        dw = np.zeros((2, 225), np.float32) if want_dw else None
        rc = self.lib.ref_heads(_p(mv), len(mv), _p(probs), ctypes.byref(value), _p(dw))
        if rc == 2:
            return None
        assert rc == 0, "the reference evaluator's self-check threw"
        return (probs, np.float32(value.value), dw) if want_dw else (probs, np.float32(value.value))

    def decisive_filter(self, moves, probs):
        """Heuristic::DecisiveFilter (Heuristic.hpp:93-161) on a copy of `probs`"""
        mv = np.ascontiguousarray(moves, np.int16)
        out = np.array(probs, np.float32)
        rc = self.lib.ref_decisive_filter(_p(mv), len(mv), _p(out))
        assert rc == 0, rc
        return out

    def hybrid_simulate(self, moves):
        """TraditionalPolicy::prepare + simulate (Traditional.h:27-31,49-69) -> (value, probs); None for terminal positions"""
        mv = np.ascontiguousarray(moves, np.int16)
        probs = np.zeros(225, np.float32)
        value = ctypes.c_float()
        rc = self.lib.ref_hybrid_simulate(_p(mv), len(mv), _p(probs), ctypes.byref(value))
        if rc == 2:
            return None
        assert rc == 0, rc
        return np.float32(value.value), probs

    def guided_rollout_max(self, moves):
        """Heuristic::MaxEvaluatedRollout (Heuristic.hpp:61-85) -> (winner, moves played)"""
        mv = np.ascontiguousarray(moves, np.int16)
        played = np.zeros(226, np.int16)
        n = ctypes.c_int()
        w = self.lib.ref_guided_rollout_max(_p(mv), len(mv), _p(played), 226, ctypes.byref(n))
        assert w != -3, "the reference evaluator threw"
        return int(w), played[:n.value].tolist()

    def sampled_first_move(self, moves, n_draws):
        """Board::getRandomMove(EvaluationProbs) (Game.cpp:75-78) drawn n_draws times -> counts i32[225]"""
        mv = np.ascontiguousarray(moves, np.int16)
        counts = np.zeros(225, np.int32)
        rc = self.lib.ref_sampled_first_move(_p(mv), len(mv), int(n_draws), _p(counts))
        assert rc == 0, rc
        return counts

    def guided_rollout_sampled(self, moves, n_games):
        """n_games x Heuristic::RandomEvaluatedRollout -> (wdb i64[3] = white, draw, black; total moves)"""
        mv = np.ascontiguousarray(moves, np.int16)
        wdb = np.zeros(3, np.int64)
        total = np.zeros(1, np.int64)
        rc = self.lib.ref_guided_rollout_sampled(_p(mv), len(mv), int(n_games), _p(wdb), _p(total))
        assert rc == 0, rc
        return wdb, int(total[0])

    def rollout_philox(self, moves, rollouts, key, ctr_hi=0, position=0):
        """the reference Board under the rollout kernel's Philox stream -> (wdb i32[3], total moves)"""
        mv = np.ascontiguousarray(moves, np.int16)
        wdb = np.zeros(3, np.int32)
        total = np.zeros(1, np.int64)
        self.lib.ref_rollout_philox(_p(mv), len(mv), int(rollouts), ctypes.c_uint64(key), ctypes.c_uint32(ctr_hi),
                                    ctypes.c_uint32(position), _p(wdb), _p(total))
        return wdb, int(total[0])

    def mcts_injected(self, moves, iterations, n_searches=0, sim_kind=0, key=1, tree=0, c_rollouts=5, c_puct=5.0,
                      noise_seed=1, cap=1 << 20):
        """The reference's MCTS (MCTS.cpp) with a deterministic injected `simulate` slot; see ref_harness_search.cpp.
        -> dict(actions, pos, visits, value, prior, depth, n_children): the visited nodes of the last tree in pre-order"""
        mv = np.ascontiguousarray(moves, np.int16)
        actions = np.zeros(max(n_searches, 1), np.int16)
        pos, depth = np.zeros(cap, np.int16), np.zeros(cap, np.int16)
        visits, nch = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        value, prior = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        self.lib.ref_mcts_injected.restype = ctypes.c_long
        n = self.lib.ref_mcts_injected(_p(mv), len(mv), int(iterations), int(n_searches), int(sim_kind), ctypes.c_uint64(key),
                                       ctypes.c_uint32(tree), int(c_rollouts), ctypes.c_double(c_puct), ctypes.c_uint32(noise_seed),
                                       _p(actions), _p(pos), _p(visits), _p(value), _p(prior), _p(depth), _p(nch), ctypes.c_long(cap))
        assert 0 <= n <= cap, n
        return {"actions": actions[:n_searches].tolist(), "pos": pos[:n], "visits": visits[:n], "value": value[:n],
                "prior": prior[:n], "depth": depth[:n], "n_children": nch[:n]}

    def mcts_get_action(self, moves, iterations, policy="random", c_puct=5.0, c_rollouts=5, c_bias=0.0):
        """The reference's own search with its own policies and RNG -> dict(action, seconds, size, root_visits)"""
        mv = np.ascontiguousarray(moves, np.int16)
        kind = {"random": 0, "traditional": 1, "poolrave": 2}[policy]
        sec, size = ctypes.c_double(), ctypes.c_int64()
        rv = np.zeros(225, np.int32)
        a = self.lib.ref_mcts_get_action(_p(mv), len(mv), int(iterations), kind, ctypes.c_double(c_puct), int(c_rollouts),
                                         ctypes.c_double(c_bias), ctypes.byref(sec), ctypes.byref(size), _p(rv))
        return {"action": int(a), "seconds": sec.value, "size": int(size.value), "root_visits": rv}

    def temp_based_probs(self, child_visits, n_moves_played):
        v = np.ascontiguousarray(child_visits, np.float32)
        out = np.zeros(225, np.float32)
        self.lib.ref_temp_based_probs(_p(v), int(n_moves_played), _p(out))
        return out

    def add_noise(self, priors, seed):
        v = np.ascontiguousarray(priors, np.float32)
        out = np.zeros(225, np.float32)
        self.lib.ref_add_noise(_p(v), ctypes.c_uint32(seed), _p(out))
        return out


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build(want_ref=False)
        _port = PortOracle()
    return _port


def ref():
    """The compiled reference, or None if oracle/_ref has not been built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = RefOracle()
    return _ref

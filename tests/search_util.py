"""Helpers shared by the search-parity tests: the deterministic leaf evaluator that oracle/ref_harness_search.cpp
injects into the reference's MCTS (sim_kind 0), written in Python for the mirror's Policy(eval_state=...) slot, and
pre-order dumps of the mirror's tree in the format of `ref.mcts_injected`."""
import numpy as np


def hash_value(board):
    """ref_harness_search.cpp::hash_value: FNV-1a over the 225 cell states (0 empty, 1 black, 2 white) -> {-1, -0.8, .., 1}"""
    packed = board.packed()
    h = 2166136261
    for c in range(225):
        s = (int(packed[c >> 4]) >> ((c & 15) * 2)) & 3
        h = ((h ^ s) * 16777619) & 0xffffffff
    return np.float32(np.float32(int(h % 11) - 5) / np.float32(5.0))


def uniform_probs(board):
    """Default::UniformProbs (MonteCarlo.hpp:50-55): 1 / #empty on every empty cell, in float"""
    packed = board.packed()
    empty = np.array([((int(packed[c >> 4]) >> ((c & 15) * 2)) & 3) == 0 for c in range(225)])
    probs = np.zeros(225, np.float32)
    probs[empty] = np.float32(1.0) / np.float32(empty.sum())
    return probs


def injected_eval_state(board):
    return float(hash_value(board)), uniform_probs(board)


def dump_tree(root):
    """visited nodes of a CorePyExt tree in pre-order -> dict of arrays like ref.mcts_injected"""
    pos, visits, value, prior, depth, nch = [], [], [], [], [], []
    stack = [(root, 0)]
    while stack:
        node, d = stack.pop()
        pos.append(int(node.position)); visits.append(node.node_visits); value.append(node.state_value)
        prior.append(node.action_prob); depth.append(d)
        children = node.children
        nch.append(len(children))
        for ch in reversed(children):
            if ch.node_visits > 0:
                stack.append((ch, d + 1))
    return {"pos": np.array(pos, np.int16), "visits": np.array(visits, np.int32), "value": np.array(value, np.float32),
            "prior": np.array(prior, np.float32), "depth": np.array(depth, np.int16), "n_children": np.array(nch, np.int32)}


def assert_same_tree(a, b, what=""):
    assert len(a["pos"]) == len(b["pos"]), (what, len(a["pos"]), len(b["pos"]))
    for k in ("pos", "depth", "visits", "n_children", "value", "prior"):
        same = np.array_equal(np.asarray(a[k]), np.asarray(b[k]))
        if not same:
            i = int(np.flatnonzero(np.asarray(a[k]) != np.asarray(b[k]))[0])
            raise AssertionError(f"{what}: trees differ in '{k}' at pre-order node {i}: {a[k][i]} vs {b[k][i]} "
                                 f"(pos {a['pos'][i]}, depth {a['depth'][i]})")

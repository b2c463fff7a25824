"""One-off differential run of the root-parallel search (arena trees, lazy children, no driver thread, leaf batches through
gk_rollout_submit_host -> rollout_warp_kernel / rollout_small_kernel) against the COMPILED reference's MCTS::playout loop
whose simulate slot is RandomPolicy::averagedSimulate driven by the kernels' Philox streams on the reference Board
(oracle/ref_harness_search.cpp, sim_kind 1): every tree node for node, from random positions of 0 .. 215 stones, with
varying tree counts, thread counts and numbers of batches in flight.
    python tests/tools/fuzz_search.py [positions] [seed]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from conftest import random_positions
from oracle import pyoracle as po
from search_util import assert_same_tree

n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4100
import gomokuai_b200 as gk
from gomokuai_b200 import build, core
build.build_pyext()
gk.init(0)
ref = po.ref()
rng = np.random.default_rng(seed)
t0 = time.time()
trees_checked = nodes = skipped = 0
lists = random_positions(seed, n_pos, lo=0, hi=216, clustered_every=3)
for i, moves in enumerate(lists):
    moves = [int(m) for m in moves]
    b = core.Board()
    for m in moves:
        b.apply_move(m)
    if b.status["is_end"]:                              # a list ends with the move that decides the game: search the position before it
        b.revert_move(1)
        moves = moves[:-1]
    if b.status["is_end"] or not moves and i % 7:       # (a full board without a five stays decided; the empty board now and then)
        skipped += 1
        continue
    trees = int(rng.choice([1, 3, 8, 40, 150, 400]))
    playouts = int(rng.choice([30, 120, 400])) if trees <= 40 else int(rng.choice([20, 60]))
    threads = int(rng.choice([1, 2, 5, 16]))
    groups = int(rng.choice([0, 1, 3, 8]))
    base = int(rng.integers(0, 1 << 20))
    key = int(rng.integers(1, 1 << 40))
    s = core.RootParallelSearch(trees=trees, c_rollouts=5, seed=key, threads=threads, noise=False, replica_base=base, groups=groups,
                                watch=bool(rng.integers(0, 2)))
    s.run(b, playouts)
    for t in sorted(set(rng.integers(0, trees, size=min(trees, 4)).tolist())):
        want = ref.mcts_injected(moves, playouts, n_searches=0, sim_kind=1, key=key, tree=base + t, c_rollouts=5)
        assert_same_tree(s.tree_dump(t), want, f"position {i} ({len(moves)} stones), tree {t} of {trees}, {threads} threads, groups {groups}")
        trees_checked += 1
        nodes += len(want["pos"])
    del s
print(json.dumps({"positions": n_pos - skipped, "decided_positions_skipped": skipped, "trees_compared": trees_checked, "nodes_compared": nodes,
                  "differing_trees": 0, "against": "compiled reference MCTS::playout + RandomPolicy::averagedSimulate on the reference Board, same Philox streams",
                  "seconds": round(time.time() - t0, 1)}))

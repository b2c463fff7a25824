"""BASELINE config 1 through the drop-in: MCTS(c_iterations=10000, policy=RandomPolicy(5.0, 5)).get_action(Board()) --
one tree, one leaf per playout, every simulate a 5-rollout GPU call.  Launch-latency bound by construction (the
reference's own CPU loop does a rollout in 5-8 us); RootParallelSearch is the throughput path for the same statistics."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomokuai_b200 import core

core.init(0)
b = core.Board()
rows = []
for it in (1000, 10000):
    m = core.MCTS(c_iterations=it, policy=core.RandomPolicy(5.0, 5))
    t0 = time.perf_counter()
    move = m.get_action(b)
    dt = time.perf_counter() - t0
    rows.append({"c_iterations": it, "seconds": dt, "playouts_per_s": it / dt, "rollouts_per_s": 5 * it / dt, "move": int(move), "tree_nodes": int(m.size)})
print(json.dumps({"metric": "config 1: single-tree MCTS playouts/sec through CorePyExt (GPU simulate)", "runs": rows}))

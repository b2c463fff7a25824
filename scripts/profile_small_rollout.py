"""One-leaf rollout launches for ncu: 12 launches of rollout_warp_kernel (5 playouts) and 12 of rollout_small_kernel
(40 playouts: more than a warp per playout allows) from the same position, each with the lengths of its games, so that
kernel cycles can be regressed on the longest game of the launch (scripts/small_rollout_summary.py).
    ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none --csv \
        --log-file launches.csv python scripts/profile_small_rollout.py > lengths.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gomokuai_b200 as gk

gk.init(0)
b = np.zeros((1, 16), np.uint32)
for c, v in zip((112, 113, 97, 98), (1, 2, 1, 2)):
    b[0, c >> 4] |= np.uint32(v << ((c & 15) * 2))
rows = []
for rep in range(12):
    for rollouts in (5, 40):
        tr = gk.rollout_trace_host(b[0], rollouts, key=3, ctr_hi=rep, pos=0)
        rows.append({"launch": len(rows), "rollouts": rollouts, "longest_game": int(tr["lengths"].max()),
                     "moves": int(tr["lengths"].astype(np.int64).sum())})
print(json.dumps(rows))

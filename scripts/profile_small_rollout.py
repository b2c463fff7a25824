"""Kernel time of one-leaf rollout batches against the longest game in the batch (run under ncu --metrics
gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum): prints, per launch, the game lengths."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gomokuai_b200 as gk

gk.init(0)
b = np.zeros((1, 16), np.uint32)
for c, v in zip((112, 113, 97, 98), (1, 2, 1, 2)):
    b[0, c >> 4] |= np.uint32(v << ((c & 15) * 2))
rows = []
for rep in range(8):
    tr = gk.rollout_trace_host(b[0], 5, key=3, ctr_hi=rep, pos=0)          # launch 2*rep: the trace variant
    w = gk.rollout_batch_host(b, 5, key=3, ctr_hi=rep, pos_base=0)         # launch 2*rep+1: counts only, same games
    rows.append({"rep": rep, "lengths": [int(x) for x in tr["lengths"]], "max": int(tr["lengths"].max()), "wdb": w[0].tolist()})
many = np.repeat(b, 128, 0)
for rep in range(4):
    gk.rollout_batch_host(many[:16], 5, key=3, ctr_hi=rep, pos_base=0)
print(json.dumps(rows))

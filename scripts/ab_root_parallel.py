"""Root-parallel search, one GPU: seconds per 10^6-playout move against the number of leaf batches in flight.
    python scripts/ab_root_parallel.py [trees[:threads] ...]       (prints one JSON line per tree count)"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gomokuai_b200 as gk
from gomokuai_b200 import core

gk.init(0)
all_threads = os.cpu_count() or 8
b = core.Board()
for c in (112, 113, 97, 98):
    b.apply_move(c)
for spec in sys.argv[1:] or ["1024", "128"]:
    trees, threads = (int(x) for x in spec.split(":")) if ":" in spec else (int(spec), all_threads)
    per_tree = max(1, 1_000_000 // trees)
    row = {"trees": trees, "playouts_per_tree": per_tree, "threads": threads, "groups": {}}
    ref = None                                                # the statistics do not depend on the group count
    for groups in (1, 2, 3, 4, 6, 8):
        if groups > trees:
            continue
        s = core.RootParallelSearch(trees=trees, c_rollouts=5, seed=5, threads=threads, groups=groups, watch=os.environ.get("RP_WATCH", "1") != "0")
        secs, drv, idle = [], [], []
        for rep in range(4):
            st = s.run(b, per_tree)
            ref = st if ref is None else ref
            assert (st == ref).all()
            if rep:
                secs.append(s.seconds_total); drv.append(s.driver_seconds); idle.append(s.seconds_gpu)
        k = secs.index(statistics.median_low(secs))
        row["groups"][groups] = {"seconds": round(secs[k], 5), "playouts_per_s": round(trees * per_tree / secs[k]),
                                 "driver_wait_submit_workers_s": [round(drv[k][0], 4), round(drv[k][2], 4), round(drv[k][1], 4)],
                                 "worker_idle_frac": round(idle[k] / secs[k], 3)}
        del s
    print(json.dumps(row), flush=True)

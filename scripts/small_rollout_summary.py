"""cycles per move of the one-leaf rollout kernels: least-squares line of sm__cycles_elapsed.max against the longest game
of each launch.    python scripts/small_rollout_summary.py launches.csv lengths.json > profiles/<tag>_small_rollout.json"""
import csv, json, sys
import numpy as np

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
col = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]]})
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
lengths = json.load(open(sys.argv[2]))
out = {"source": "ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none "
                 "python scripts/profile_small_rollout.py", "kernels": {}, "launches": []}
by_kernel = {}
for i, L in enumerate(lengths):
    d = launches[i]
    name = "rollout_warp_kernel" if "rollout_warp" in d["kernel"] else "rollout_small_kernel" if "rollout_small" in d["kernel"] else d["kernel"]
    row = {"kernel": name, "rollouts": L["rollouts"], "longest_game": L["longest_game"], "moves": L["moves"],
           "cycles": d["sm__cycles_elapsed.max"], "ns": d["gpu__time_duration.sum"], "warp_instructions": d["smsp__inst_executed.sum"]}
    out["launches"].append(row)
    by_kernel.setdefault(name, []).append(row)
for name, rs in by_kernel.items():
    x = np.array([r["longest_game"] for r in rs], float); y = np.array([r["cycles"] for r in rs], float)
    slope, icpt = np.polyfit(x, y, 1)
    out["kernels"][name] = {"launches": len(rs), "rollouts_per_launch": rs[0]["rollouts"], "cycles_per_move": round(float(slope), 1),
                            "cycles_fixed": round(float(icpt)), "us_median": round(float(np.median([r["ns"] for r in rs])) / 1e3, 2),
                            "warp_instructions_per_move_played": round(float(sum(r["warp_instructions"] for r in rs) / sum(r["moves"] for r in rs)), 1)}
print(json.dumps(out, indent=1))

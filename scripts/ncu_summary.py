"""Turns an `ncu --set full` report of the bench-size launches into profiles/ncu_summary.json
(per-launch dram bytes, warp instructions, issue utilisation, stalls) -- the figures bench.py quotes.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv
The first launch of every distinct kernel is kept; ac_eval_kernel<0, 0> and rollout_kernel<0, 12> are
stored under the plain names bench.py looks up.
"""
import csv, json, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
col = {x: i for i, x in enumerate(h)}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def num(r, name, unit_scaled=False):
    try:
        v = float(r[col[name]].replace(",", ""))
        return v * scale.get(u[col[name]], 1.0) if unit_scaled else v
    except Exception:
        return None


out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"([A-Za-z_]\w*(?:<[^<>()]*>)?)\s*\(", name)          # function name (+ template arguments) before the argument list
    key = m.group(1) if m else name
    key = "ac_eval_kernel" if key.startswith("ac_eval_kernel<0, 0>") else "rollout_kernel" if key.startswith("rollout_kernel") else key
    if key in out:
        continue
    rd, wr = num(r, "dram__bytes_read.sum", True), num(r, "dram__bytes_write.sum", True)
    out[key] = {
        "kernel": name, "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]],
        "ncu_duration": r[col["gpu__time_duration.sum"]] + " " + u[col["gpu__time_duration.sum"]],
        "dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_read_bytes": rd, "dram_write_bytes": wr,
        "warp_inst_per_launch": num(r, "smsp__inst_executed.sum"),
        "threads_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": num(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "shared_bank_conflicts": num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "shared_wavefronts": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": num(r, "launch__registers_per_thread"),
        "stalls_per_issue": {x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): num(r, x)
                             for x in h if "issue_stalled" in x and "per_issue_active" in x and (num(r, x) or 0) > 0.1},
    }
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from bench import git_blob_sha  # noqa: E402
# stamp: the kernel sources this capture was taken from (bench.py drops the figures once a source has changed)
out["_sources"] = {f: git_blob_sha(os.path.join(root, "gomokuai_b200", "csrc", f)) for f in ("gk_eval.cu", "gk_rollout.cu")}
dst = os.path.join(root, "profiles", "ncu_summary.json")
json.dump(out, open(dst, "w"), indent=1)
for k, v in out.items():
    if k.startswith("_"):
        continue
    print(k, v["ncu_duration"], "warp-inst", v["warp_inst_per_launch"], "issue%", v["issue_active_pct"], "dram bytes", v["dram_bytes_per_launch"])

"""Turns an `ncu --set full` report of the bench-size launches into profiles/ncu_summary.json
(per-launch dram bytes, warp instructions, issue utilisation) -- the figures bench.py quotes.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv
"""
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name, scale=1.0):
    try: return float(r[col[name]].replace(",", "")) * scale
    except Exception: return None
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = "ac_eval_kernel" if "ac_eval" in name else "rollout_kernel" if "rollout_kernel" in name else None
    if key is None: continue
    unit_scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = num(r, "dram__bytes_read.sum", unit_scale.get(units[col["dram__bytes_read.sum"]], 1.0))
    wr = num(r, "dram__bytes_write.sum", unit_scale.get(units[col["dram__bytes_write.sum"]], 1.0))
    t_unit = units[col["gpu__time_duration.sum"]]
    out[key] = {
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
        "ncu_duration": r[col["gpu__time_duration.sum"]] + " " + t_unit,
        "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]],
    }
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))

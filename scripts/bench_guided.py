"""Pattern-guided playouts (configs[4]): games/s of the incremental kernel and of the full-rescan variant at several batch
sizes, kernel alone (CUDA events, best of 5 after 2 warm-ups).    python scripts/bench_guided.py [1024 8192 ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk

gk.init(0)
sizes = [int(x) for x in sys.argv[1:]] or [256, 1024, 8192, 65536]
rows = []
for n in sizes:
    boards = torch.zeros((n, 16), dtype=torch.int32, device="cuda")
    row = {"games": n}
    for name, full, single in (("incremental", False, False), ("single_warp", False, True), ("full_rescan", True, False)):
        for mode in ("sample", "max"):
            best = None
            for it in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                r = gk.guided_rollout_batch(boards, mode=mode, key=gk.SYNTH_KEY, game_base=0, want_moves=True, full_rescan=full, single_warp=single)
                b.record()
                torch.cuda.synchronize()
                if it >= 2:
                    best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
            moves = float(r["length"].float().sum().item())
            row[f"{name}_{mode}"] = {"ms": best, "games_per_s": n / (best * 1e-3), "moves_per_s": moves / (best * 1e-3), "mean_length": moves / n}
    # the same concurrency as a STREAM: 16 batches' worth of games, at most n in flight
    big = torch.zeros((16 * n, 16), dtype=torch.int32, device="cuda")
    for name, full in (("incremental", False), ("full_rescan", True)):
        best = None
        for it in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = gk.guided_rollout_batch(big, mode="sample", key=gk.SYNTH_KEY, game_base=0, want_moves=True, full_rescan=full, max_in_flight=n)
            b.record()
            torch.cuda.synchronize()
            if it >= 1:
                best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
        row[f"{name}_sample_stream"] = {"ms": best, "games": 16 * n, "in_flight": n, "games_per_s": 16 * n / (best * 1e-3)}
    row["speedup_sample"] = row["full_rescan_sample"]["ms"] / row["incremental_sample"]["ms"]
    row["speedup_max"] = row["full_rescan_max"]["ms"] / row["incremental_max"]["ms"]
    rows.append(row)
    print(json.dumps(row), file=sys.stderr)
print(json.dumps({"metric": "guided playouts: incremental kernel vs full rescan, kernel alone", "rows": rows}))

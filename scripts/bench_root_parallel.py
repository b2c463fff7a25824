"""BASELINE config 4: root-parallel MCTS, playouts/move across the GPUs of one box with ONE
allreduce of the root statistics per move.

    python scripts/bench_root_parallel.py --playouts 1000000 --trees 2048
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_root_parallel.py ...
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--playouts", type=int, default=1_000_000)
ap.add_argument("--trees", type=int, default=4096, help="trees per GPU")
ap.add_argument("--rollouts", type=int, default=5)
ap.add_argument("--moves", type=int, default=3)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--check", action="store_true", help="rank 0 repeats every search alone with world x trees trees and compares the statistics")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gomokuai_b200 as gk
from gomokuai_b200 import core, root_parallel as rp
gk.init(local)
threads = args.threads or max(1, (os.cpu_count() or 8) // world)
b = core.Board()
for c in (112, 113, 97, 98):
    b.apply_move(c)
rows = []
for mv in range(args.moves):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    move, stats, s = rp.search(b, args.playouts, trees_per_rank=args.trees, c_rollouts=args.rollouts, seed=11 + mv, threads=threads)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    played = int(stats[0].sum()) + world * args.trees
    rows.append({"move": int(move), "seconds": float(t.item()), "playouts": played, "playouts_per_s": played / float(t.item()),
                 "gpu_fraction": s.seconds_gpu / max(s.seconds_total, 1e-9), "nodes_rank0": int(s.nodes),
                 "playouts_per_tree": int(s.leaves // args.trees),
                 "search_seconds": s.seconds_total, "driver_seconds": [round(x, 4) for x in s.driver_seconds]})
    if args.check and rank == 0:
        # the same search on ONE GPU: world x trees trees with tree indices 0 .. world*trees-1 -- identical by construction
        alone = core.RootParallelSearch(trees=world * args.trees, c_rollouts=args.rollouts, seed=11 + mv, threads=threads)
        rows[-1]["equals_single_gpu_search"] = bool(np.array_equal(alone.run(b, rows[-1]["playouts_per_tree"]), stats))
        rows[-1]["single_gpu_move"] = int(core.RootParallelSearch.best_move(alone.run(b, rows[-1]["playouts_per_tree"])))
        del alone
    if world > 1:
        dist.barrier()
    b.apply_move(move)
if rank == 0:
    best = max(rows, key=lambda r: r["playouts_per_s"])
    print(json.dumps({"metric": "root-parallel MCTS playouts/sec", "n_gpus": world, "trees_per_gpu": args.trees, "c_rollouts": args.rollouts,
                      "host_threads_per_rank": threads, "value": best["playouts_per_s"], "rollouts_per_s": best["playouts_per_s"] * args.rollouts,
                      "moves": rows}))
if world > 1:
    dist.destroy_process_group()

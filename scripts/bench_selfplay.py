"""BASELINE config 5: self-play data generation with pattern-guided rollouts -- G concurrent games per GPU,
every game played to its end inside one kernel (gk_guided_rollout_batch, mode "sample"), then every visited
position with its outcome (gk_expand_games) and its feature planes (gk_encode_states_batch).  Reports games/s and moves/s; weak scaling, no
collective on the data path.

    python scripts/bench_selfplay.py --games 8192
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_selfplay.py --games 1024
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192, help="concurrent games per GPU")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--mode", default="sample")
ap.add_argument("--augment", action="store_true", help="emit the 8 rotations / reflections of every sample")
ap.add_argument("--from-synth", action="store_true", help="start from the synthetic mid-game set instead of empty boards")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gomokuai_b200 as gk
gk.init(local)
if args.from_synth:
    boards = gk.synth_positions(rank * args.games, args.games, want_moves=False)[0]
else:
    boards = np.zeros((args.games, 16), np.uint32)
d_boards = torch.from_numpy(boards.view(np.int32)).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
res = None
for it in range(args.steps + 2):
    if it == 2:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); ev[0].record()
    res = gk.guided_rollout_batch(d_boards, mode=args.mode, key=gk.SYNTH_KEY + it, game_base=rank * args.games)
    ex = gk.expand_games(d_boards, res)                                      # every visited position, its last moves, its z
    planes = gk.encode_states_batch(ex["boards"], ex["last_moves"], augment=args.augment)   # the training samples' planes
ev[1].record(); torch.cuda.synchronize()
ms = torch.tensor([ev[0].elapsed_time(ev[1]) / args.steps], dtype=torch.float64, device="cuda")
moves = torch.tensor([float(res["length"].float().sum().item())], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(moves, op=dist.ReduceOp.SUM)
if rank == 0:
    w = res["winner"].cpu().numpy()
    print(json.dumps({"metric": "self-play games/sec (pattern-guided, 15x15)", "n_gpus": world, "games_per_gpu": args.games, "mode": args.mode,
                      "start": "synthetic mid-game" if args.from_synth else "empty board", "ms_per_batch": float(ms.item()),
                      "value": world * args.games / float(ms.item()) * 1e3, "unit": "games/s",
                      "moves_per_s": float(moves.item()) / float(ms.item()) * 1e3,
                      "samples_per_s": float(moves.item()) * (8 if args.augment else 1) / float(ms.item()) * 1e3,
                      "pipeline": "guided_rollout_batch -> expand_games -> encode_states_batch (planes of every visited position)", "mean_game_length": float(moves.item()) / (world * args.games),
                      "black_win_rate_rank0": float((w == 1).mean()), "white_win_rate_rank0": float((w == -1).mean())}))
if world > 1:
    dist.destroy_process_group()

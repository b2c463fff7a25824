"""Config 4, the honest picture: how many trees should share 1M playouts per move?

Root parallelism trades tree DEPTH for throughput: T trees of 1M / T playouts each keep the GPU batch full (T leaves per
launch) but every tree is shallow; one tree of 1M playouts is deep but simulates one leaf per launch (latency bound).
This script sweeps T at a fixed 1M playouts per move on ONE GPU and reports, per T, the throughput and a strength proxy:
the share of tactical positions whose forced move the merged arg-max finds --

  win    the side to move has a four with an empty completion cell: every completion cell is a solution;
  block  the side to move has no four, the opponent has exactly one completion cell: that cell is the only move that
         does not lose at once.

Positions come from the synthetic mid-game set (gk_synth_positions), filtered by the evaluator's own pattern totals
(LiveFour / DeadFour counts of gk_eval_batch) and solved cells are found by a plain five-in-a-row check.

    python scripts/sweep_root_parallel.py [--positions 100] [--playouts 1000000] > profiles/r02_root_parallel_sweep.json
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gomokuai_b200 as gk
from gomokuai_b200 import core


def completion_cells(cells, colour):
    out = []
    for c in np.flatnonzero(cells == 0):
        x0, y0 = c % 15, c // 15
        for dx, dy in ((1, 0), (0, 1), (1, 1), (1, -1)):
            n = 1
            for sgn in (1, -1):
                x, y = x0 + sgn * dx, y0 + sgn * dy
                while 0 <= x < 15 and 0 <= y < 15 and cells[y * 15 + x] == colour:
                    n += 1
                    x += sgn * dx
                    y += sgn * dy
            if n >= 5:
                out.append(int(c))
                break
    return out


def tactical_suite(n_each):
    wins, blocks, first = [], [], 0
    while len(wins) < n_each or len(blocks) < n_each:
        boards, moves, starts = gk.synth_positions(first, 8192)
        first += 8192
        tot = gk.eval_batch(boards)["pat_totals"].cpu().numpy().view(np.uint16).reshape(-1, 2, 8)
        fours = tot[:, :, 6] + tot[:, :, 7]                       # [position][0 = white, 1 = black]
        for i in np.flatnonzero(fours.sum(axis=1) > 0):
            mv = moves[starts[i]:starts[i + 1]].tolist()
            cells = gk.unpack_boards(boards[i:i + 1])[0].astype(int)
            side = 1 if len(mv) % 2 == 0 else 2
            mine, theirs = completion_cells(cells, side), completion_cells(cells, 3 - side)
            if mine and len(wins) < n_each:
                wins.append({"kind": "win", "moves": mv, "answers": mine})
            elif not mine and len(theirs) == 1 and len(blocks) < n_each:
                blocks.append({"kind": "block", "moves": mv, "answers": theirs})
    return wins + blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=100)
    ap.add_argument("--playouts", type=int, default=1_000_000)
    ap.add_argument("--trees", type=str, default="8192,1024,64,8,1")
    ap.add_argument("--budget-s", type=float, default=75.0, help="wall-clock budget per tree count: fewer positions where a search is slow")
    args = ap.parse_args()
    gk.init(0)
    suite = tactical_suite(args.positions // 2)
    threads = len(os.sched_getaffinity(0))
    rows = []
    for trees in [int(t) for t in args.trees.split(",")]:
        per_tree = math.ceil(args.playouts / trees)
        s = core.RootParallelSearch(trees=trees, c_rollouts=5, seed=1, threads=threads)
        solved = {"win": 0, "block": 0}
        tried = {"win": 0, "block": 0}
        secs, t_begin = [], time.perf_counter()
        order = [suite[(i // 2) + (len(suite) // 2) * (i % 2)] for i in range(len(suite))]     # alternate win / block
        for k, pos in enumerate(order):
            if k >= 2 and time.perf_counter() - t_begin > args.budget_s:
                break
            b = core.Board()
            for c in pos["moves"]:
                b.apply_move(c, False)
            t0 = time.perf_counter()
            stats = s.run(b, per_tree, 100 + k)
            secs.append(time.perf_counter() - t0)
            move = int(core.RootParallelSearch.best_move(stats))
            tried[pos["kind"]] += 1
            solved[pos["kind"]] += move in pos["answers"]
        n = sum(tried.values())
        rows.append({"trees": trees, "playouts_per_tree": per_tree, "positions": n, "tried": tried, "solved": solved,
                     "solved_frac": sum(solved.values()) / max(n, 1), "seconds_per_move_median": float(np.median(secs)),
                     "playouts_per_s": trees * per_tree / float(np.median(secs))})
        print(json.dumps(rows[-1]), file=sys.stderr)
        del s
    print(json.dumps({"metric": "root-parallel MCTS on one B200: trees vs throughput and tactical strength at a fixed playout budget per move",
                      "playouts_per_move": args.playouts, "host_threads": threads, "suite": {"win": args.positions // 2, "block": args.positions // 2},
                      "rows": rows}))


if __name__ == "__main__":
    main()

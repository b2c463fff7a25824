"""Full-size parity, every element (too slow for pytest; run once per round on the GPU box):
  config 2: ALL 1,048,576 synthetic positions -- GPU scores / totals / winner against the compiled REFERENCE
            evaluator (oracle/_ref, Evaluator::applyMove replay) when it is present, else the C restatement;
  config 3: ALL 4096 x 4096 rollouts -- per-position win/draw/loss counts against the C restatement of
            Board::getRandomMove / applyMove / checkGameEnd consuming the same Philox stream.
CPU work is spread over the host cores (workers are forked before CUDA is initialised).
    python tests/tools/full_parity.py [--positions N] [--rollout-positions P] [--rollouts R] > profiles/rXX_full_parity.json
"""
import argparse, json, multiprocessing as mp, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np


def _eval_worker(job):
    kind, moves, starts = job
    from oracle import pyoracle
    o = pyoracle.ref() if kind == "reference" else pyoracle.port()
    r = o.eval_batch(moves, starts)
    return r["scores"], r["pat_totals"], r["cmp_totals"], r["winner"], int(r.get("bad", 0))


def _roll_worker(job):
    moves, starts, R, key, base = job
    from oracle import pyoracle
    _, _, wdb = pyoracle.port().rollout_philox_batch(moves, starts, R, key, pos_base=base)
    return wdb


def _slices(moves, starts, lo, hi, parts):
    per = (hi - lo + parts - 1) // parts
    for a in range(lo, hi, per):
        b = min(hi, a + per)
        yield a, b, moves[starts[a]:starts[b]], starts[a:b + 1] - starts[a]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=1 << 20)
    ap.add_argument("--rollout-positions", type=int, default=4096)
    ap.add_argument("--rollouts", type=int, default=4096)
    args = ap.parse_args()
    from oracle import pyoracle
    import gomokuai_b200 as gk
    kind = "reference" if pyoracle.ref() is not None else "port"
    pyoracle.port()
    cores = len(os.sched_getaffinity(0))
    pool = mp.get_context("fork").Pool(cores)
    n = args.positions
    boards, moves, starts = gk.synth_positions(0, n)
    out = {"host_cores": cores}

    # ---- config 2 -------------------------------------------------------------------------------------------
    t0 = time.perf_counter()
    jobs = [(kind, m, s) for _, _, m, s in _slices(moves, starts, 0, n, max(cores * 8, n // 8192))]
    stream = pool.imap(_eval_worker, jobs)                              # ordered, a few thousand boards per message
    import torch
    gk.init(0)
    dev = gk.eval_batch(torch.from_numpy(boards.view(np.int32)).cuda())
    torch.cuda.synchronize()
    g_scores = dev["scores"].cpu().numpy(); g_pat = dev["pat_totals"].cpu().numpy().view(np.uint16)
    g_cmp = dev["cmp_totals"].cpu().numpy().view(np.uint16); g_win = dev["winner"].cpu().numpy()
    bad_boards, lo, crc = 0, 0, 0
    for sc, pt, ct, wn, bad in stream:
        hi = lo + len(wn)
        same = ((g_scores[lo:hi] == sc).all(axis=(1, 2)) & (g_pat[lo:hi] == pt).all(axis=(1, 2)) &
                (g_cmp[lo:hi] == ct).all(axis=(1, 2)) & (g_win[lo:hi] == wn))
        bad_boards += int((~same).sum()) + bad
        crc = zlib.crc32(sc.tobytes(), crc)
        lo = hi
    cpu_s = time.perf_counter() - t0
    assert lo == n
    out["config2"] = {"positions": n, "checker": kind + (" (oracle/_ref: the reference's Evaluator compiled unmodified)" if kind == "reference" else " (C restatement)"),
                      "positions_differing": bad_boards, "score_values_compared": int(g_scores.size), "crc32_of_all_scores": f"{crc:08x}",
                      "crc32_gpu": f"{zlib.crc32(g_scores.tobytes()):08x}", "compound_total": int(g_cmp.sum()), "pattern_total": int(g_pat.sum()),
                      "checker_seconds": round(cpu_s, 1)}
    print("config 2:", out["config2"], file=sys.stderr, flush=True)

    # ---- config 3 -------------------------------------------------------------------------------------------
    P, R = args.rollout_positions, args.rollouts
    t0 = time.perf_counter()
    jobs = [(m, s, R, gk.SYNTH_KEY, a) for a, _, m, s in _slices(moves, starts, 0, P, cores * 8)]
    res_async = pool.map_async(_roll_worker, jobs)
    g_wdb = gk.rollout_batch(torch.from_numpy(boards[:P].view(np.int32)).cuda(), R)["wdb"].cpu().numpy()
    want = np.concatenate(res_async.get())
    out["config3"] = {"positions": P, "rollouts_per_position": R, "rollouts": P * R, "checker": "port (C restatement of Game.cpp:37-136 on the same Philox stream)",
                      "positions_differing": int((g_wdb != want).any(axis=1).sum()), "black_wins": int(g_wdb[:, 2].sum()),
                      "white_wins": int(g_wdb[:, 0].sum()), "draws": int(g_wdb[:, 1].sum()), "checker_seconds": round(time.perf_counter() - t0, 1)}
    print("config 3:", out["config3"], file=sys.stderr, flush=True)
    pool.close()
    print(json.dumps(out))
    return 0 if out["config2"]["positions_differing"] == 0 and out["config3"]["positions_differing"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())

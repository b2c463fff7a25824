"""One-off differential run: GPU evaluator against the CPU restatement on many dense / clustered / terminal positions
(the synthetic benchmark set is uniformly random; these stress the compound and pattern-count paths).
    python tests/tools/fuzz_eval.py [n_lists] [seed]"""
import multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def _work(job):
    seed, n = job
    from conftest import random_positions
    from oracle import pyoracle as po
    lists = random_positions(seed, n, lo=20, hi=160, clustered_every=2)
    mv, st = po.pack_moves(lists)
    r = po.port().eval_batch(mv, st)
    return mv, st, r["scores"], r["pat_totals"], r["cmp_totals"], r["winner"]


if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    cores = len(os.sched_getaffinity(0))
    per = 2000
    pool = mp.get_context("fork").Pool(cores)
    jobs = [(seed0 + i, per) for i in range(total // per)]
    it = pool.imap(_work, jobs)
    import torch
    import gomokuai_b200 as gk
    gk.init(0)
    bad = done = comp = 0
    t0 = time.time()
    for mv, st, sc, pt, ct, wn in it:
        out = gk.eval_batch(gk.pack_moves(mv, st))
        ok = (np.array_equal(out["scores"].cpu().numpy(), sc) and np.array_equal(out["pat_totals"].cpu().numpy().view(np.uint16), pt)
              and np.array_equal(out["cmp_totals"].cpu().numpy().view(np.uint16), ct) and np.array_equal(out["winner"].cpu().numpy(), wn))
        pol = gk.eval_policy_batch(gk.pack_moves(mv, st), want_scores=True)
        ok = ok and np.array_equal(pol["scores"].cpu().numpy(), sc) and np.array_equal(pol["cmp_totals"].cpu().numpy().view(np.uint16), ct)
        bad += 0 if ok else 1
        done += len(st) - 1
        comp += int(ct.sum())
    pool.close(); pool.join()
    print(f"{done} positions in {time.time() - t0:.1f} s, {comp} compounds, {bad} differing batches of {per}")
    sys.exit(1 if bad else 0)

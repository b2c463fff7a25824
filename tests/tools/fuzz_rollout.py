"""One-off differential run: GPU rollouts (batched kernels and the fused single-launch path) against the CPU restatement
on dense / clustered / terminal / nearly full positions under the same Philox streams.
    python tests/tools/fuzz_rollout.py [n_lists] [seed]"""
import multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

R = 48


def _work(job):
    seed, n = job
    from conftest import random_positions
    from oracle import pyoracle as po
    lists = random_positions(seed, n, lo=1, hi=226, clustered_every=3)
    mv, st = po.pack_moves(lists)
    _, _, wdb = po.port().rollout_philox_batch(mv, st, R, 4242, ctr_hi=seed & 0xffff, pos_base=seed)
    return seed, mv, st, wdb


if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
    cores = len(os.sched_getaffinity(0))
    per = 1000
    pool = mp.get_context("fork").Pool(cores)
    it = pool.imap(_work, [(seed0 + i, per) for i in range(total // per)])
    import torch
    import gomokuai_b200 as gk
    gk.init(0)
    bad = done = 0
    for seed, mv, st, want in it:
        boards = gk.pack_moves(mv, st)
        got = gk.rollout_batch(boards, R, key=4242, ctr_hi=seed & 0xffff, pos_base=seed)["wdb"].cpu().numpy()
        fused = gk.rollout_batch_host(boards[:16], R, key=4242, ctr_hi=seed & 0xffff, pos_base=seed)      # the single-launch path
        ok = np.array_equal(got, want) and np.array_equal(fused, want[:16])
        bad += 0 if ok else 1
        done += len(st) - 1
    pool.close(); pool.join()
    print(f"{done} positions x {R} rollouts, {bad} differing batches of {per}")
    sys.exit(1 if bad else 0)

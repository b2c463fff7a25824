"""One-off differential run: GPU rollouts -- the batched kernels, the one-launch thread-per-rollout kernel (48 rollouts per
position) and the one-launch warp-per-rollout kernel (5 and 32 rollouts per position, EVERY position, 16 per launch) --
against the CPU restatement on dense / clustered / terminal / nearly full positions under the same Philox streams.
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
    winners, _, wdb = po.port().rollout_philox_batch(mv, st, R, 4242, ctr_hi=seed & 0xffff, pos_base=seed)
    return seed, mv, st, wdb, winners


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
    bad = done = warp_bad = warp_launches = 0
    for seed, mv, st, want, winners in it:
        boards = gk.pack_moves(mv, st)
        got = gk.rollout_batch(boards, R, key=4242, ctr_hi=seed & 0xffff, pos_base=seed)["wdb"].cpu().numpy()
        fused = gk.rollout_batch_host(boards[:16], R, key=4242, ctr_hi=seed & 0xffff, pos_base=seed)      # the single-launch path
        ok = np.array_equal(got, want) and np.array_equal(fused, want[:16])
        bad += 0 if ok else 1
        # a rollout's stream is (rollout index, position), whatever the count: the first r of the 48 are the r-rollout call's
        for r in (5, 32):
            w = winners[:, :r]
            want_r = np.stack([(w == -1).sum(1), (w == 0).sum(1), (w == 1).sum(1)], 1).astype(np.int32)
            for lo in range(0, len(st) - 1, 16):
                got_r = gk.rollout_batch_host(boards[lo:lo + 16], r, key=4242, ctr_hi=seed & 0xffff, pos_base=seed + lo)
                warp_bad += 0 if np.array_equal(got_r, want_r[lo:lo + 16]) else 1
                warp_launches += 1
        done += len(st) - 1
    pool.close(); pool.join()
    print(f"{done} positions x {R} rollouts, {bad} differing batches of {per}; warp-per-rollout kernel: {warp_launches} launches "
          f"of 16 positions x 5 / 32 rollouts, {warp_bad} differ")
    sys.exit(1 if bad or warp_bad else 0)

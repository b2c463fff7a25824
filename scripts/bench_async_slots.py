"""Latency of small rollout batches through gk_rollout_submit_host / gk_rollout_wait: one batch alone, and K batches
in flight on K slots (do they overlap on the GPU?).  Development tool for the root-parallel driver."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gomokuai_b200 as gk

gk.init(0)
L = gk.lib()
L.gk_rollout_submit_host.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
L.gk_rollout_wait.argtypes = [ctypes.c_int]


def pinned(nbytes):
    p = ctypes.c_void_p()
    assert L.gk_host_alloc(ctypes.byref(p), ctypes.c_size_t(nbytes)) == 0
    return p


for n in (512, 2048):
    boards = pinned(8 * n * 64)
    wdb = pinned(8 * n * 12)
    ctypes.memset(boards, 0, 8 * n * 64)
    # a near-empty position: 4 stones
    arr = np.ctypeslib.as_array(ctypes.cast(boards, ctypes.POINTER(ctypes.c_uint32)), shape=(8 * n, 16))
    b = np.zeros(16, np.uint32)
    for i, c in enumerate((112, 113, 97, 98)):
        b[c >> 4] |= (1 + (i & 1)) << ((c & 15) * 2)
    arr[:] = b
    for k in (1, 2, 4, 8):
        best = 1e9
        sub = 0.0
        for rep in range(30):
            t0 = time.perf_counter()
            for s in range(k):
                assert L.gk_rollout_submit_host(s, boards.value + s * n * 64, n, 5, 7, rep, s * n, wdb.value + s * n * 12) == 0
            t1 = time.perf_counter()
            for s in range(k):
                assert L.gk_rollout_wait(s) == 0
            t2 = time.perf_counter()
            if rep >= 5 and t2 - t0 < best:
                best, sub = t2 - t0, t1 - t0
        print(f"n={n} leaves x 5 rollouts, {k} slots in flight: {best * 1e6:7.1f} us total, submits {sub * 1e6:6.1f} us  -> {best * 1e6 / k:6.1f} us per batch")

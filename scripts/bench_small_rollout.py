"""Latency of one small rollout batch through gk_rollout_submit_host / gk_rollout_wait (page-locked buffers, one launch):
microseconds per round trip against the batch size.    python scripts/bench_small_rollout.py"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gomokuai_b200 as gk

gk.init(0)
L = gk.lib()
L.gk_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
L.gk_host_free.argtypes = [ctypes.c_void_p]


def host_alloc(nbytes):
    p = ctypes.c_void_p()
    assert L.gk_host_alloc(ctypes.byref(p), nbytes) == 0
    return p.value


rows = []
for stones in (0, 4, 60):
    for n in (1, 16, 64, 128, 148, 256, 296, 512, 1024):
        boards, _, _ = gk.synth_positions(7, n, want_moves=False) if stones == 60 else (None, None, None)
        hb = host_alloc(n * 64); hw = host_alloc(n * 12)
        b = np.frombuffer((ctypes.c_char * (n * 64)).from_address(hb), np.uint32).reshape(n, 16)
        w = np.frombuffer((ctypes.c_char * (n * 12)).from_address(hw), np.int32).reshape(n, 3)
        if stones == 60:
            b[:] = np.asarray(boards)
        else:
            b[:] = 0
            for c, v in zip((112, 113, 97, 98)[:stones], (1, 2, 1, 2)):
                b[:, c >> 4] |= np.uint32(v << ((c & 15) * 2))
        ts = []
        for rep in range(60):
            t0 = time.perf_counter()
            gk.rollout_submit_host(0, b, 5, w, key=3, ctr_hi=rep, pos_base=0)
            gk.rollout_wait(0)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[10:])
        rows.append({"stones": stones, "positions": n, "rollouts": 5, "us_median": round(ts[len(ts) // 2] * 1e6, 1), "us_min": round(ts[0] * 1e6, 1)})
        L.gk_host_free(hb); L.gk_host_free(hw)
        print(rows[-1], file=sys.stderr, flush=True)
print(json.dumps(rows))

import sys, time, ctypes
sys.path.insert(0, ".")
import numpy as np
import gomokuai_b200 as gk
gk.init(0)
L = gk.lib()
b = np.zeros((1, 16), np.uint32); out = np.zeros((1, 3), np.int32)
bp, op = b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)
for R in (5, 64):
    for _ in range(200): L.gk_rollout_batch_host(bp, 1, R, ctypes.c_uint64(1), 0, 0, op)
    t0 = time.perf_counter()
    for i in range(2000): L.gk_rollout_batch_host(bp, 1, R, ctypes.c_uint64(1), i, 0, op)
    print("n=1 R=%d: %.1f us per call" % (R, (time.perf_counter() - t0) / 2000 * 1e6), out)
# mid-game position: shorter rollouts
bb, _, _ = gk.synth_positions(7, 1, want_moves=False)
bp2 = bb.ctypes.data_as(ctypes.c_void_p)
t0 = time.perf_counter()
for i in range(2000): L.gk_rollout_batch_host(bp2, 1, 5, ctypes.c_uint64(1), i, 0, op)
print("mid-game n=1 R=5: %.1f us per call" % ((time.perf_counter() - t0) / 2000 * 1e6))
# TraditionalPolicy::hybridSimulate for one leaf
t = gk.default_table()
L.gk_hybrid_simulate_batch_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
probs = np.zeros((1, 225), np.float32); value = np.zeros(1, np.float32)
pp, vp = probs.ctypes.data_as(ctypes.c_void_p), value.ctypes.data_as(ctypes.c_void_p)
for _ in range(200): L.gk_hybrid_simulate_batch_host(t.handle, bp2, 1, pp, vp, None)
t0 = time.perf_counter()
for i in range(2000): L.gk_hybrid_simulate_batch_host(t.handle, bp2, 1, pp, vp, None)
print("hybrid simulate n=1: %.1f us per call" % ((time.perf_counter() - t0) / 2000 * 1e6), float(value[0]), float(probs.sum()))

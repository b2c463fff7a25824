"""Quick GPU sanity run (development aid): parity of both kernels against the oracle + first timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gomokuai_b200 as gk
from oracle import pyoracle as po

gk.init(0)
print(gk.device_info(), gk.default_table().info(), flush=True)
P = po.port()
n = 4096
boards, moves, starts = gk.synth_positions(0, n)
t = time.time(); ref = P.eval_batch(moves, starts); print("oracle eval %.2fs" % (time.time() - t), "bad", ref["bad"], flush=True)
out = gk.eval_batch(boards)
torch.cuda.synchronize()
sc = out["scores"].cpu().numpy(); pt = out["pat_totals"].cpu().numpy().view(np.uint16); ct = out["cmp_totals"].cpu().numpy().view(np.uint16); w = out["winner"].cpu().numpy()
print("scores equal", np.array_equal(sc, ref["scores"]), "pat", np.array_equal(pt, ref["pat_totals"]), "cmp", np.array_equal(ct, ref["cmp_totals"]), "winner", np.array_equal(w, ref["winner"]), flush=True)
if not np.array_equal(sc, ref["scores"]):
    bad = np.nonzero((sc != ref["scores"]).any(axis=(1, 2)))[0]
    print("bad positions", len(bad), bad[:10])
    i = bad[0]; d = np.argwhere(sc[i] != ref["scores"][i]); print(d[:10], sc[i][tuple(d[0])], ref["scores"][i][tuple(d[0])])
# timing eval
big = np.tile(boards, (64, 1))  # 262144 boards
bt = torch.from_numpy(big.view(np.int32)).cuda()
o = gk.eval_batch(bt)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): gk.eval_batch(bt, out=o)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
print("eval %d boards: %.3f ms -> %.3e boards/s" % (len(big), ms, len(big) / ms * 1e3), flush=True)
# rollouts: parity
npos, R = 64, 32
t = time.time(); wn, ln, wdb = P.rollout_philox_batch(moves[:starts[npos]], starts[:npos + 1], R, gk.SYNTH_KEY); print("oracle rollouts %.2fs" % (time.time() - t))
r = gk.rollout_batch(boards[:npos], R, want_trace=True)
torch.cuda.synchronize()
print("rollout winners equal", np.array_equal(r["winners"].cpu().numpy(), wn), "lengths", np.array_equal(r["lengths"].cpu().numpy(), ln), "wdb", np.array_equal(r["wdb"].cpu().numpy(), wdb), flush=True)
# timing rollouts
bt4 = torch.from_numpy(boards.view(np.int32)).cuda()
r = gk.rollout_batch(bt4, 256)
torch.cuda.synchronize()
ev0.record(); r = gk.rollout_batch(bt4, 4096); ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
wd = r["wdb"].cpu().numpy()
print("rollouts 4096x4096: %.3f ms -> %.3e rollouts/s; wdb sum ok %s; mean W/D/B %s" % (ms, 4096 * 4096 / ms * 1e3, bool((wd.sum(1) == 4096).all()), wd.mean(0)), flush=True)

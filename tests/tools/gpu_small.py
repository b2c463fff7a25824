"""Tiny end-to-end run of every kernel (for compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gomokuai_b200 as gk
from oracle import pyoracle as po
gk.init(0)
P = po.port()
boards, moves, starts = gk.synth_positions(0, 600)
ref = P.eval_batch(moves, starts)
out = gk.eval_batch(boards); torch.cuda.synchronize()
assert np.array_equal(out["scores"].cpu().numpy(), ref["scores"])
wn, ln, wdb = P.rollout_philox_batch(moves[:starts[8]], starts[:9], 40, 5)
r = gk.rollout_batch(boards[:8], 40, key=5, want_trace=True); torch.cuda.synchronize()
assert np.array_equal(r["winners"].cpu().numpy(), wn) and np.array_equal(r["wdb"].cpu().numpy(), wdb)
rs = np.random.default_rng(0).integers(0, 225, size=(8, 3, 230)).astype(np.uint8)
gk.rollout_injected(boards[:8], rs); torch.cuda.synchronize()
codes = np.array([3, 4, 1, 1, 1, 4, 2, 3, 3], np.uint8)
gk.scan_batch(codes, np.array([0, 9], np.int64)); torch.cuda.synchronize()
print("small ok")
lists = [moves[starts[i]:starts[i + 1]].tolist() for i in range(40)]
o = gk.eval_policy_batch(boards[:40], want_scores=True); torch.cuda.synchronize()
for i in (0, 7, 39):
    rp, rv = po.policy_heads(P, lists[i])
    assert np.allclose(o["probs"][i].cpu().numpy(), rp, rtol=2e-5, atol=2e-6) and abs(float(o["value"][i]) - float(rv)) < 2e-5
g = gk.guided_rollout_batch(boards[:40], mode="sample"); torch.cuda.synchronize()
g = gk.guided_rollout_batch(np.zeros((3, 16), np.uint32), mode="max"); torch.cuda.synchronize()
assert int(g["length"].min()) > 8
last = np.full((40, 2), -1, np.int16)
pl, pr = gk.encode_states_batch(boards[:39], last[:39], augment=True, probs=np.ones((39, 225), np.float32)); torch.cuda.synchronize()
pl = gk.encode_states_batch(boards[:7], last[:7]); torch.cuda.synchronize()
print("small ok (heads, guided, encode)")

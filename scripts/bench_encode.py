"""Timing of encode_states_kernel (row f3) against the HBM roofline (development aid)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk

gk.init(0)
n = 1 << 18
boards, _, _ = gk.synth_positions(0, n, want_moves=False)
bt = torch.from_numpy(boards.view(np.int32)).cuda()
last = torch.randint(0, 225, (n, 2), dtype=torch.int16, device="cuda")
res = {}
for aug in (False, True):
    for _ in range(3):
        p = gk.encode_states_batch(bt, last, augment=aug)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        p = gk.encode_states_batch(bt, last, augment=aug)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = n * (64 + 4 + (10800 if aug else 1350)) / 1e9
    res["augment" if aug else "plain"] = {"ms": ms, "positions_per_s": n / ms * 1e3, "GBps": gb / ms * 1e3}
    del p
print(json.dumps(res))

"""Small fixed workload for ncu: a few launches of each hot kernel at the bench sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk

gk.init(0)
n = int(os.environ.get("GK_PROFILE_BOARDS", 1 << 18))
boards, _, _ = gk.synth_positions(0, n, want_moves=False)
bt = torch.from_numpy(boards.view(np.int32)).cuda()
out = gk.eval_batch(bt)
for _ in range(2):
    gk.eval_batch(bt, out=out)
r = None
for _ in range(2):
    r = gk.rollout_batch(bt[:4096].contiguous(), int(os.environ.get("GK_PROFILE_ROLLOUTS", 1024)))
torch.cuda.synchronize()
print("ok", int(out["pat_totals"].sum()), int(r["wdb"].sum()))

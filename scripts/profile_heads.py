"""ncu target: the policy-head and guided variants of ac_eval_kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk
gk.init(0)
n = 1 << 18
boards, _, _ = gk.synth_positions(0, n, want_moves=False)
bt = torch.from_numpy(boards.view(np.int32)).cuda()
for _ in range(2):
    o = gk.eval_policy_batch(bt)
g = gk.guided_rollout_batch(torch.zeros((8192, 16), dtype=torch.int32, device="cuda"), mode="sample")
torch.cuda.synchronize()
print("ok", float(o["value"].sum()), int(g["length"].sum()))

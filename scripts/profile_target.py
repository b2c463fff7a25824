"""Small fixed workload for ncu: a few launches of every kernel at the bench sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk

gk.init(0)
n = int(os.environ.get("GK_PROFILE_BOARDS", 1 << 18))
boards, _, _ = gk.synth_positions(0, n, want_moves=False)
bt = torch.from_numpy(boards.view(np.int32)).cuda()
once = bool(os.environ.get("GK_PROFILE_ONCE"))          # one launch per kernel: keeps an `ncu --set full` report small
out = gk.eval_batch(bt)
for _ in range(0 if once else 2):
    gk.eval_batch(bt, out=out)
r = None
for _ in range(1 if once else 2):
    r = gk.rollout_batch(bt[:4096].contiguous(), int(os.environ.get("GK_PROFILE_ROLLOUTS", 1024)))
if os.environ.get("GK_PROFILE_ALL"):
    m = min(n, 1 << 18)
    pol = gk.eval_policy_batch(bt[:m])
    hyb = gk.hybrid_simulate_batch(bt[:m])
    g = gk.guided_rollout_batch(torch.zeros((8192, 16), dtype=torch.int32, device="cuda"), mode="sample")     # guided_kernel
    g2 = gk.guided_rollout_batch(torch.zeros((1024, 16), dtype=torch.int32, device="cuda"), mode="sample")    # guided_pair_kernel
    last = torch.full((m, 2), -1, dtype=torch.int16, device="cuda")
    enc = gk.encode_states_batch(bt[:m], last, augment=True)
    if not once:
        enc1 = gk.encode_states_batch(bt[:m], last)
torch.cuda.synchronize()
print("ok", int(out["pat_totals"].sum()), int(r["wdb"].sum()))

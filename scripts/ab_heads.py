"""A/B timing of library variants (development aid): GK_LIB=<path> python scripts/ab_heads.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gomokuai_b200 as gk
if os.environ.get("GK_LIB"):
    gk.LIB_PATH = os.environ["GK_LIB"]
gk.init(0)
b, _, _ = gk.synth_positions(0, 1 << 19, want_moves=False)
bt = torch.from_numpy(b.view(np.int32)).cuda()
z = torch.zeros((65536, 16), dtype=torch.int32, device="cuda")
def t(f, n=4):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(os.environ.get("GK_LIB", "default"), "policy %.3f ms  hybrid %.3f ms  plain %.3f ms  guided65536 %.3f ms  guided8192 %.3f ms" % (
    t(lambda: gk.eval_policy_batch(bt)), t(lambda: gk.hybrid_simulate_batch(bt)), t(lambda: gk.eval_batch(bt)),
    t(lambda: gk.guided_rollout_batch(z, mode="sample")), t(lambda: gk.guided_rollout_batch(z[:8192], mode="sample"))))

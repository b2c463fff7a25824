"""ncu target: a few launches of the guided kernels at a small batch (latency-bound regime).  GK_GUIDED_N, GK_GUIDED_MODE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomokuai_b200 as gk
gk.init(0)
n = int(os.environ.get("GK_GUIDED_N", 1024))
mode = os.environ.get("GK_GUIDED_MODE", "max")
boards = torch.zeros((n, 16), dtype=torch.int32, device="cuda")
for full in (False, True):
    r = gk.guided_rollout_batch(boards, mode=mode, key=gk.SYNTH_KEY, full_rescan=full)
torch.cuda.synchronize()
print("ok", int(r["length"].sum()))

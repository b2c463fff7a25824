import os, sys, subprocess
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
import gomokuai_b200 as gk
gk.init(0)
boards, _, _ = gk.synth_positions(0, 4096, want_moves=False)
bt = torch.from_numpy(boards.view(np.int32)).cuda()
gk.rollout_batch(bt, 256); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(4):
    e0.record(); r = gk.rollout_batch(bt, 4096); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("ctas/sm", os.environ.get("GK_ROLLOUT_CTAS"), "ms %.3f -> %.3e rollouts/s" % (best, 4096 * 4096 / best * 1e3))
'''
for c in ("1", "2", "3"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GK_ROLLOUT_CTAS=c))

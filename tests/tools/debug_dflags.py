import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gomokuai_b200 as gk
from oracle import pyoracle
from test_heads import _positions, _dflag_bits
gk.init(0)
port = pyoracle.port()
lists = [m for m in _positions(41, 900) if port.eval_moves(m)["winner"] == 0]
mv, st = pyoracle.pack_moves(lists)
out = gk.hybrid_simulate_batch(gk.pack_moves(mv, st), want_flags=True)
dflags = out["dflags"].cpu().numpy().view(np.uint32)
for i, m in enumerate(lists):
    port.eval_moves(m)
    pf, cf, _ = port.eval_flags()
    want = np.array([_dflag_bits(pf[c], cf[c]) for c in range(225)], np.uint32)
    if not np.array_equal(dflags[i], want):
        bad = np.nonzero(dflags[i] != want)[0]
        print("pos", i, "moves", m)
        for c in bad[:6]:
            print("   cell", int(c), "gpu %08x oracle %08x" % (int(dflags[i][c]), int(want[c])), "pf", [hex(int(x)) for x in pf[c][4:8]], "cf", [hex(int(x)) for x in cf[c]])

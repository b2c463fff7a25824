"""One-off differential run of the three guided-playout kernels: many games from empty, random, clustered, dense and
decided start positions, arg-max and sampled, capped and uncapped -- guided_pair_kernel (two warps per game) and
guided_kernel (one) against ac_eval_kernel<true, true> (whole board re-evaluated after every move): winners, lengths,
moves and final boards must be identical, game for game.  A sample of the arg-max games is also replayed through the
COMPILED reference's Heuristic::MaxEvaluatedRollout when oracle/_ref is present.
    python tests/tools/fuzz_guided.py [n_lists] [seed] > profiles/rXX_fuzz_guided.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    import torch
    import gomokuai_b200 as gk
    from conftest import random_positions
    from oracle import pyoracle as po
    gk.init(0)
    ref = po.ref()
    t0 = time.time()
    games = moves = bad = ref_checked = ref_same = 0
    per = 2000
    for i in range(total // per):
        lists = random_positions(seed0 + i, per, lo=0 if i % 4 == 0 else 10, hi=40 if i % 4 == 0 else 200, clustered_every=2 if i % 2 else 3)
        mv, st = po.pack_moves(lists)
        boards = gk.pack_moves(mv, st)
        for mode in ("max", "sample"):
            for cap in (225, 1 + (i % 9)):
                full = gk.guided_rollout_batch(boards, mode=mode, key=gk.SYNTH_KEY + i, game_base=i * per, max_moves=cap, full_rescan=True)
                for kw in ({}, {"single_warp": True}, {"max_in_flight": 96}):
                    inc = gk.guided_rollout_batch(boards, mode=mode, key=gk.SYNTH_KEY + i, game_base=i * per, max_moves=cap, **kw)
                    same = all(torch.equal(inc[k], full[k]) for k in ("winner", "length", "moves", "final_boards"))
                    bad += 0 if same else 1
                games += 4 * per
                moves += 4 * int(full["length"].sum().item())
        if ref is not None and i % 5 == 0:                       # the reference itself, on a few games of this batch
            w, ln, mvs = full["winner"].cpu().numpy(), None, None
            inc = gk.guided_rollout_batch(boards[:40], mode="max", key=gk.SYNTH_KEY, max_moves=225)
            w, ln, mvs = inc["winner"].cpu().numpy(), inc["length"].cpu().numpy(), inc["moves"].cpu().numpy()
            for g in range(40):
                if ref.heads(lists[g]) is None and ref.eval_moves(lists[g])["winner"] == 0 and len(lists[g]) < 225:
                    continue
                rw, rp = ref.guided_rollout_max(lists[g])
                ref_checked += 1
                ref_same += int(rw == int(w[g]) and rp == mvs[g, :ln[g]].tolist())
    print(json.dumps({"start_positions": total, "games_per_kernel": games, "moves_per_kernel": moves, "differing_runs": bad,
                      "kernels": ["guided_pair_kernel", "guided_kernel", "guided_pair_kernel (queue of 96)"], "against": "ac_eval_kernel<true, true> (full rescan)",
                      "reference_max_rollouts_checked": ref_checked, "identical_to_reference": ref_same, "seconds": time.time() - t0}))
    sys.exit(1 if bad else 0)

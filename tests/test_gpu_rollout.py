"""K2 `rollout` through the C-ABI against the oracle: bit-exact winner and length under injected
start-index streams; statistical agreement with the reference's own RNG path.  Needs a B200."""
import numpy as np
import pytest

from conftest import random_positions
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def test_philox_stream_bit_exact(gpu, port):
    import torch
    n, R = 256, 64
    boards, moves, starts = gpu.synth_positions(0, n)
    wn, ln, wdb = port.rollout_philox_batch(moves, starts, R, gpu.SYNTH_KEY)
    r = gpu.rollout_batch(boards, R, want_trace=True)
    torch.cuda.synchronize()
    assert np.array_equal(_np(r["winners"]), wn)
    assert np.array_equal(_np(r["lengths"]), ln)
    assert np.array_equal(_np(r["wdb"]), wdb)
    assert 30 < ln.mean() < 80


def test_philox_stream_other_key_ctr_and_base(gpu, port):
    import torch
    boards, moves, starts = gpu.synth_positions(1000, 40)
    wn, ln, wdb = port.rollout_philox_batch(moves, starts, 33, 0x1234567890ABCDEF, ctr_hi=7, pos_base=500)
    r = gpu.rollout_batch(boards, 33, key=0x1234567890ABCDEF, ctr_hi=7, pos_base=500, want_trace=True)
    torch.cuda.synchronize()
    assert np.array_equal(_np(r["winners"]), wn) and np.array_equal(_np(r["lengths"]), ln)
    assert np.array_equal(_np(r["wdb"]), wdb)


def test_committed_reference_outcomes(gpu, ref_outputs):
    import torch
    g = ref_outputs["rollout_philox"]
    for case in g["cases"]:
        mv = ref_outputs["eval"][case["name"]]["moves"]
        m, st = po.pack_moves([mv])
        r = gpu.rollout_batch(gpu.pack_moves(m, st), 16, key=g["key"], ctr_hi=g["ctr_hi"], pos_base=case["position_index"], want_trace=True)
        torch.cuda.synchronize()
        assert _np(r["winners"])[0].tolist() == [o[0] for o in case["outcomes"]], case["name"]
        assert _np(r["lengths"])[0].tolist() == [o[1] for o in case["outcomes"]], case["name"]
    for ex in ref_outputs["rollout_explicit"]:
        m, st = po.pack_moves([ex["moves"]])
        r = gpu.rollout_injected(gpu.pack_moves(m, st), np.array(ex["r"], np.uint8).reshape(1, 1, -1))
        torch.cuda.synchronize()
        assert int(_np(r["winners"])[0, 0]) == ex["winner"] and int(_np(r["lengths"])[0, 0]) == ex["length"]


def test_explicit_injected_streams_bit_exact(gpu, port):
    import torch
    rng = np.random.default_rng(4)
    lists = random_positions(31, 96, lo=0, hi=100)
    lists = [l for l in lists if port.board_play(l)["cur_player"] != 0][:64]
    mv, st = po.pack_moves(lists)
    R, stride = 8, 230
    rs = rng.integers(0, 225, size=(len(lists), R, stride)).astype(np.uint8)
    rs[:, 0, :] = 224                                   # always wraps around the end of the board
    rs[:, 1, :] = 0
    r = gpu.rollout_injected(gpu.pack_moves(mv, st), rs)
    torch.cuda.synchronize()
    for i, l in enumerate(lists):
        for j in range(R):
            w, k = port.rollout_injected(l, rs[i, j])
            assert (int(_np(r["winners"])[i, j]), int(_np(r["lengths"])[i, j])) == (w, k), (i, j)


def test_edge_positions(gpu, port, kats):
    import torch
    tie = [y * 15 + x for y in kats["tie_row_order"] for x in range(15)]
    lists = [
        [],                                                   # empty board: ~100 moves
        tie[:-1],                                             # one empty cell: exactly one move
        tie[:-2], tie[:-7],
        tie,                                                  # full: 0 moves, draw
        [y * 15 + x for x, y in kats["black_win"]],           # decided: 0 moves
        [y * 15 + x for x, y in kats["white_win"]],
    ]
    mv, st = po.pack_moves(lists)
    wn, ln, wdb = port.rollout_philox_batch(mv, st, 40, 99)
    r = gpu.rollout_batch(gpu.pack_moves(mv, st), 40, key=99, want_trace=True)
    torch.cuda.synchronize()
    assert np.array_equal(_np(r["winners"]), wn) and np.array_equal(_np(r["lengths"]), ln)
    assert np.array_equal(_np(r["wdb"]), wdb)
    assert (ln[4] == 0).all() and (wn[4] == 0).all() and (wn[5] == 1).all() and (wn[6] == -1).all() and (ln[5] == 0).all()
    assert (ln[1] == 1).all()


def test_stream_exhaustion_is_reported(gpu):
    import torch
    rs = np.zeros((1, 2, 4), np.uint8)
    r = gpu.rollout_injected(np.zeros((1, 16), np.uint32), rs)
    torch.cuda.synchronize()
    assert _np(r["lengths"]).tolist() == [[255, 255]] and _np(r["winners"]).tolist() == [[0, 0]]


def test_full_size_properties_16m_rollouts(gpu, port):
    """BASELINE config 3 size: 4096 positions x 4096 rollouts."""
    import torch
    n, R = 4096, 4096
    boards, moves, starts = gpu.synth_positions(0, n)
    bt = torch.from_numpy(boards.view(np.int32)).cuda()
    a = gpu.rollout_batch(bt, R)["wdb"]
    torch.cuda.synchronize()
    assert torch.equal(a.sum(dim=1), torch.full((n,), R, dtype=torch.int64, device=a.device))   # every rollout counted once
    b = gpu.rollout_batch(bt, R)["wdb"]                                                          # deterministic
    # partition invariance: positions split across "ranks" with pos_base, as the multi-GPU path does
    c0 = gpu.rollout_batch(bt[:1000], R, pos_base=0)["wdb"]
    c1 = gpu.rollout_batch(bt[1000:], R, pos_base=1000)["wdb"]
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a[:1000], c0) and torch.equal(a[1000:], c1)
    # sampled bit-exact check of whole positions against the oracle
    sub = [5, 1234, 4095]
    mv, st = po.pack_moves([moves[starts[i]:starts[i + 1]] for i in sub])
    for k, i in enumerate(sub):
        _, _, wdb = port.rollout_philox_batch(mv[st[k]:st[k + 1]], np.array([0, st[k + 1] - st[k]]), R, gpu.SYNTH_KEY, pos_base=i)
        assert _np(a[i]).tolist() == wdb[0].tolist()


def test_win_rates_match_reference_rng_path(gpu, ref):
    """Free-running agreement with the reference's own mt19937 path (Board::getRandomMove):
    |p_gpu - p_cpu| <= 4.5 * sqrt(p(1-p)(1/n_g + 1/n_c)) per position."""
    import torch
    n, n_c, n_g = 48, 4096, 16384
    boards, moves, starts = gpu.synth_positions(9000, n)
    g = _np(gpu.rollout_batch(boards, n_g, key=2024)["wdb"]).astype(np.float64)
    torch.cuda.synchronize()
    worst = 0.0
    for i in range(n):
        wdb, total = ref.rollout_free(moves[starts[i]:starts[i + 1]], n_c)
        pg, pc = g[i, 2] / n_g, wdb[2] / n_c
        p = (g[i, 2] + wdb[2]) / (n_g + n_c)
        tol = 4.5 * np.sqrt(max(p * (1 - p), 1e-4) * (1 / n_g + 1 / n_c))
        worst = max(worst, abs(pg - pc) / tol)
        assert abs(pg - pc) <= tol, (i, pg, pc, tol)
    assert worst > 0.0


def test_async_slots_equal_the_synchronous_call(gpu):
    """gk_rollout_submit_host / gk_rollout_wait: four batches in flight on four slots give the counts of four
    synchronous calls (Philox streams depend on (position, rollout), not on the slot or the interleaving)."""
    boards, _, _ = gpu.synth_positions(2000, 4 * 96, want_moves=False)
    parts = [np.ascontiguousarray(boards[i * 96:(i + 1) * 96]) for i in range(4)]
    outs = [np.full((96, 3), -1, np.int32) for _ in range(4)]
    for rep in range(3):                                    # reuse of the slots' buffers
        for s in range(4):
            gpu.rollout_submit_host(s, parts[s], 24, outs[s], key=99, ctr_hi=rep, pos_base=s * 96)
        for s in (2, 0, 3, 1):
            gpu.rollout_wait(s)
        want = gpu.rollout_batch_host(boards, 24, key=99, ctr_hi=rep)
        assert np.array_equal(np.concatenate(outs), want)
        assert (want.sum(1) == 24).all()
    with pytest.raises(gpu.GomokuB200Error):
        gpu.rollout_submit_host(16, parts[0], 24, outs[0])
    gpu.rollout_wait(5)                                     # a slot never used: nothing to wait for


def test_concurrent_streams_do_not_share_scratch(gpu):
    """Rollout launches in flight on different streams each use their own slot-image scratch."""
    import torch
    sets = [gpu.synth_positions(3000 + 512 * i, 512, want_moves=False)[0] for i in range(3)]
    serial = [_np(gpu.rollout_batch(b, 64, pos_base=512 * i)["wdb"]) for i, b in enumerate(sets)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in sets]
    dev = [torch.from_numpy(b.view(np.int32)).cuda() for b in sets]
    torch.cuda.synchronize()
    for rep in range(4):
        res = []
        for i, (st, b) in enumerate(zip(streams, dev)):
            with torch.cuda.stream(st):
                res.append(gpu.rollout_batch(b, 64, pos_base=512 * i, stream=st)["wdb"])
        torch.cuda.synchronize()
        for got, want in zip(res, serial):
            assert np.array_equal(_np(got), want)


def test_c_abi_argument_errors(gpu):
    """The C-ABI never throws and never touches memory on bad arguments: negative status + gk_last_error() text."""
    import ctypes
    L = gpu.lib()
    boards = np.zeros((4, 16), np.uint32)
    wdb = np.zeros((4, 3), np.int32)
    bp, wp = boards.ctypes.data_as(ctypes.c_void_p), wdb.ctypes.data_as(ctypes.c_void_p)
    INVALID = -1
    assert L.gk_rollout_batch_host(None, 4, 8, ctypes.c_uint64(1), 0, 0, wp) == INVALID
    assert L.gk_rollout_batch_host(bp, -1, 8, ctypes.c_uint64(1), 0, 0, wp) == INVALID
    assert L.gk_rollout_batch_host(bp, 4, 0, ctypes.c_uint64(1), 0, 0, wp) == INVALID                 # no rollouts
    assert L.gk_rollout_batch_host(bp, 0, 8, ctypes.c_uint64(1), 0, 0, None) == 0                      # empty batch: nothing to do
    assert L.gk_rollout_submit_host(-1, bp, 4, 8, ctypes.c_uint64(1), 0, 0, wp) == INVALID
    assert L.gk_rollout_submit_host(0, bp, 4, 8, ctypes.c_uint64(1), 0, 0, None) == INVALID
    assert L.gk_rollout_wait(99) == INVALID
    assert b"slot" in L.gk_last_error()
    # n * rollouts_per_pos must stay below 2^31 per call (rollout tickets are 32-bit)
    import torch
    d = torch.zeros((2, 16), dtype=torch.int32, device="cuda")
    out = torch.zeros((2, 3), dtype=torch.int32, device="cuda")
    assert L.gk_rollout_batch(ctypes.c_void_p(d.data_ptr()), 2, 1 << 30, ctypes.c_uint64(1), 0, 0,
                              ctypes.c_void_p(out.data_ptr()), None, None, None) == INVALID
    assert L.gk_expand_games(None, None, None, None, 3, 225, None, None, None, None, None) == INVALID
    assert L.gk_expand_games(None, None, None, None, 0, 225, None, None, None, None, None) == 0
    v = ctypes.c_double()
    assert L.gk_measure_issue_peak(7, ctypes.byref(v)) == INVALID and L.gk_measure_issue_peak(0, None) == INVALID
    st = torch.zeros(10, dtype=torch.int64, device="cuda")
    assert L.gk_root_allreduce(None, ctypes.c_void_p(st.data_ptr()), None) in (-5, -6)                  # no communicator yet (or no NCCL)
    # after all that the library still works
    assert (gpu.rollout_batch_host(boards, 8).sum(1) == 8).all()


def test_fused_small_batch_launch_equals_the_big_kernel(gpu, port, kats):
    """gk_rollout_batch_host with a handful of positions runs ONE fused launch (image build + rollouts, one CTA per
    position): same Philox streams, same counts as the batched kernels and as the oracle -- including decided,
    full and nearly full boards, and every rollout count that changes the block shape."""
    tie = [y * 15 + x for y in kats["tie_row_order"] for x in range(15)]
    lists = [[], tie[:-1], tie[:-7], tie, [y * 15 + x for x, y in kats["black_win"]], [112, 113, 97, 98]]
    lists += [list(map(int, m)) for m in (gpu.synth_positions(900 + i, 1)[1] for i in range(6))]
    mv, st = po.pack_moves(lists)
    boards = gpu.pack_moves(mv, st)
    for R, base in ((1, 0), (5, 3), (32, 100), (33, 7), (256, 1 << 20)):
        small = gpu.rollout_batch_host(boards, R, key=77, ctr_hi=R, pos_base=base)          # 12 positions: fused path
        big = _np(gpu.rollout_batch(boards, R, key=77, ctr_hi=R, pos_base=base)["wdb"])     # device path: batched kernels
        _, _, want = port.rollout_philox_batch(mv, st, R, 77, ctr_hi=R, pos_base=base)
        assert np.array_equal(small, big) and np.array_equal(small, want), R
    one = gpu.rollout_batch_host(boards[5:6], 5, key=1, pos_base=9)
    assert np.array_equal(one, _np(gpu.rollout_batch(boards[5:6], 5, key=1, pos_base=9)["wdb"]))


def test_shutdown_and_reinit_in_a_fresh_process():
    """gk_shutdown gives everything back (pipes, async slots, staging, scratch) and the library can be bound again."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, gomokuai_b200 as gk\n"
        "gk.init(0)\n"
        "b = gk.synth_positions(0, 300, want_moves=False)[0]\n"
        "a = gk.rollout_batch_host(b, 8); s = gk.rollout_batch_host(b[:4], 5)\n"
        "out = np.zeros((300, 3), np.int32); gk.rollout_submit_host(2, b, 8, out); gk.rollout_wait(2)\n"
        "e = gk.eval_policy_batch_host(b[:3]); h = gk.eval_batch_host(b)\n"
        "gk.shutdown()\n"
        "assert gk.lib().gk_rollout_wait(0) == -5\n"                      # GK_ERR_NOT_INIT: nothing works until the next gk_init
        "gk.init(0)\n"
        "assert np.array_equal(gk.rollout_batch_host(b, 8), a) and np.array_equal(out, a)\n"
        "assert np.array_equal(gk.rollout_batch_host(b[:4], 5), s)\n"
        "assert np.array_equal(gk.eval_batch_host(b)['scores'], h['scores'])\n"
        "gk.shutdown(); print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]

"""The N > 1 path on CPU: two gloo ranks merge their root statistics with one allreduce, and
shard the synthetic position set the way bench.py does (disjoint ranges, pos_base)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gomokuai_b200 import root_parallel as rp
    import gomokuai_b200 as gk
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    local = rng.integers(0, 1000, size=(3, 225)).astype(np.int64)
    merged = rp.allreduce_root_stats(local)
    # each rank owns a disjoint range of the synthetic set (host generator, no GPU)
    boards, _, _ = gk.synth_positions(rank * 64, 64, want_moves=False)
    out.put((rank, local, merged, boards))
    dist.destroy_process_group()


def test_two_rank_allreduce_and_sharding():
    import multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    assert np.array_equal(res[0][2], total) and np.array_equal(res[1][2], total)
    sys.path.insert(0, ROOT)
    import gomokuai_b200 as gk
    whole, _, _ = gk.synth_positions(0, 128, want_moves=False)
    assert np.array_equal(np.concatenate([res[0][3], res[1][3]]), whole)

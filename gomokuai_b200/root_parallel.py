"""Root-parallel MCTS across the GPUs of one box (BASELINE config 4).

Every rank (one process per GPU) runs `trees_per_rank` independent search trees from the same
root with disjoint Philox streams; leaves are simulated in batches by the rollout kernel.  The
only exchange is ONE allreduce(sum) of the int64[3][225] root statistics per move; integer counts make
the result independent of the reduction order.  On GPUs the sum goes through the C-ABI
(`gk_root_allreduce`, the library's own NCCL communicator over NVLink; torch.distributed only carries
the 128-byte NCCL id to the other ranks once); the gloo path exists for the CPU tests.
"""
import math

import numpy as np


_gk_comm_ready = False


def _ensure_gk_comm(group=None):
    """Create the library's NCCL communicator once: rank 0 makes the id, the process group broadcasts it."""
    global _gk_comm_ready
    if _gk_comm_ready:
        return
    import torch.distributed as dist
    import gomokuai_b200 as gk
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [gk.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    gk.nccl_init(box[0], world, rank)
    _gk_comm_ready = True


def allreduce_root_stats(stats, group=None):
    """Sum int64[3,225] root statistics over all ranks. No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    stats = np.ascontiguousarray(stats, np.int64)
    if not (dist.is_available() and dist.is_initialized()):
        return stats
    t = torch.from_numpy(stats.copy())
    if dist.get_backend(group) == "nccl":
        import gomokuai_b200 as gk
        _ensure_gk_comm(group)
        t = gk.root_allreduce(t.cuda())
        torch.cuda.current_stream().synchronize()
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def best_move(stats):
    """Most visited root child, ties -> lowest cell (MCTS::stepForward, MCTS.cpp:129-134)."""
    visits = np.asarray(stats)[0]
    return int(np.argmax(visits)) if visits.max() > 0 else -1


_searchers = {}


def search(board, playouts_total, trees_per_rank=256, c_rollouts=5, c_puct=5.0, seed=1, noise=False, threads=0, group=None):
    """One move of root-parallel search. Returns (best cell, merged stats int64[3,225], local RootParallelSearch).
    The searcher (worker threads, tree arenas, page-locked buffers) is kept between moves; the number of moves already
    on the board is mixed into the seed, so consecutive moves of one game draw different playout streams."""
    import torch.distributed as dist
    from .core import RootParallelSearch
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if (dist.is_available() and dist.is_initialized()) else (0, 1)
    per_tree = max(1, math.ceil(playouts_total / (world * trees_per_rank)))
    key = (trees_per_rank, c_rollouts, c_puct, rank * trees_per_rank, threads, noise)
    s = _searchers.get(key)
    if s is None:
        _searchers.clear()
        s = _searchers[key] = RootParallelSearch(trees=trees_per_rank, c_rollouts=c_rollouts, c_puct=c_puct, seed=seed,
                                                 replica_base=rank * trees_per_rank, threads=threads, noise=noise)
    local = s.run(board, per_tree, (int(seed) * 1000003 + len(board.move_record)) & 0xffffffffffffffff)
    merged = allreduce_root_stats(local, group)
    return best_move(merged), merged, s

#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (board evals/s and rollouts/s on 15x15).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the hot path over one batch of synthetic random mid-game positions
(SURVEY.md 8d): the primary line is BASELINE.json configs[1], K1 `ac_eval` over 1,048,576
positions per GPU; the `rollouts` object of the same line is configs[2], K2 `rollout`,
4096 positions x 4096 playouts per GPU.  Ranks take disjoint position ranges and exchange nothing
on the data path (weak scaling).  Prints ONE JSON line on rank 0.

--impl reference times the reference's own CPU implementation of the same two paths
(oracle/_ref = the reference's sources compiled unmodified, else the C restatement) on all host
cores, one process per core, each step a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EVAL_POSITIONS = 1 << 20          # configs[1]
N_EVAL_PER_GPU = EVAL_POSITIONS
EVAL_WORKLOAD = ("configs[1]: batched AC pattern evaluation of 1,048,576 synthetic random mid-game 15x15 positions per GPU, "
                 "bit-exact scores")
EVAL_OUTPUTS = "int32 scores[4][225] + u16 totals[2][8] + u16 compounds[2][3] + winner per position"
ROLL_POSITIONS = 4096             # configs[2]
ROLL_PER_POS = 4096
EVAL_BYTES_IN, EVAL_BYTES_OUT = 64, 3648       # algorithmic bytes per board
EVAL_STEPS_ALGO = 1076                          # automaton symbol-steps per board
FALLBACK_HBM_GBS = 6650.0                       # B200_PROFILING.md, used only without MEASURED_PEAKS.json


# ------------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed regions
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None and self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation, one process per core
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, what, moves, starts, extra = args
    from oracle import pyoracle
    orc = pyoracle.ref() if kind == "reference" else pyoracle.port()
    t0 = time.perf_counter()
    if what == "eval":
        r = orc.eval_batch(moves, starts, want_scores=False)        # Evaluator::applyMove replay, the reference's algorithm
        assert r["bad"] == 0
        units = len(starts) - 1
    elif what == "linescan":
        orc.linescan_batch(moves, starts)                            # PatternSearch::execute over all 88 line strings of each position
        units = 88 * (len(starts) - 1)
    else:
        units = 0
        n_roll = extra
        for i in range(len(starts) - 1):
            mv = moves[starts[i]:starts[i + 1]]
            if kind == "reference":
                orc.rollout_free(mv, n_roll)                          # Board::getRandomMove with its own mt19937
            else:
                m, s = pyoracle.pack_moves([mv])
                orc.rollout_philox_batch(m, s, n_roll, 1, 0, i)
            units += n_roll
    return units, time.perf_counter() - t0


class CpuArm:
    def __init__(self):
        from oracle import pyoracle
        if pyoracle.ref() is not None:
            self.kind = "reference"
        else:
            pyoracle.port()
            self.kind = "port"
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.pool = mp.get_context("fork").Pool(self.cores)

    def _slices(self, moves, starts, n):
        per = (n + self.cores - 1) // self.cores
        out = []
        for c in range(self.cores):
            lo, hi = c * per, min(n, (c + 1) * per)
            if lo >= hi:
                break
            st = starts[lo:hi + 1] - starts[lo]
            out.append((moves[starts[lo]:starts[hi]], st))
        return out

    def run(self, what, moves, starts, n, extra=None):
        """units/s over all cores for one bounded sample (wall clock over the whole pool)."""
        jobs = [(self.kind, what, m, s, extra) for m, s in self._slices(moves, starts, n)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
        return sum(u for u, _ in res) / wall, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import gomokuai_b200 as gk                                          # host-only use: the synthetic position generator
    arm = CpuArm()
    n_eval = 1536 * arm.cores                                            # ~0.5 s of evaluator replay per core and step: long enough that the
    n_roll_pos, n_roll = 4 * arm.cores, 32768                            # pool's start-up does not weigh on the rate; ~0.5 s of rollouts
    _, moves, starts = gk.synth_positions(0, max(n_eval, n_roll_pos))
    ev, ro = [], []
    for i in range(args.warmup + args.steps):
        a, _ = arm.run("eval", moves, starts, n_eval)
        b, _ = arm.run("rollout", moves, starts, n_roll_pos, n_roll)
        if i >= args.warmup:
            ev.append(a)
            ro.append(b)
    arm.close()
    value, rvalue = statistics.mean(ev), statistics.mean(ro)
    sample = (f"each step replays the first {n_eval} positions of the set through Evaluator::applyMove on {arm.cores} host cores "
              f"(one process per core); rollouts: {n_roll_pos}x{n_roll} per step")
    config1 = reference_config1(max(1, min(args.steps, 3)))
    line = {
        "impl": "reference", "metric": "board evals/sec (15x15)", "value": value, "unit": "boards/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_eval / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": "boards/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "boards/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "rollouts": {"metric": "rollouts/sec (15x15)", "value": rvalue, "unit": "rollouts/s",
                     "cpu_baseline": {"value": rvalue, "unit": "rollouts/s", "cores": arm.cores, "kind": arm.kind,
                                      "sample": f"{n_roll_pos} positions x {n_roll} rollouts per step"},
                     "e2e": {"value": rvalue, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "config1": config1,
    }
    _emit(line)


def bench_config():
    """the `config` object: the SAME dict in both arms (what differs between the arms lives in other keys)"""
    return {"workload": EVAL_WORKLOAD, "positions_per_gpu": N_EVAL_PER_GPU, "outputs": EVAL_OUTPUTS}


CONFIG1_WORKLOAD = ("configs[0]: 15x15 pure-MCTS random-rollout policy, MCTS(c_iterations=10000, RandomPolicy(5.0, 5))"
                    ".get_action(Board()) from the empty board, single game")


def reference_config1(repeats):
    """configs[0] on the reference itself: MCTS.cpp + policies/Random.h compiled unmodified (oracle/_ref), one core --
    the reference's search is single-threaded.  None when oracle/_ref was not built."""
    from oracle import pyoracle
    ref = pyoracle.ref()
    if ref is None or not hasattr(ref.lib, "ref_mcts_get_action"):
        return None
    runs = [ref.mcts_get_action([], 10000, policy="random", c_puct=5.0, c_rollouts=5) for _ in range(repeats)]
    best = min(r["seconds"] for r in runs)
    return {"workload": CONFIG1_WORKLOAD, "value": 10000 / best, "unit": "playouts/s", "rollouts_per_sec": 50000 / best,
            "seconds": best, "cores": 1, "kind": "reference", "tree_nodes": runs[0]["size"],
            "sample": f"best of {repeats} runs of the whole call on one host core (Board, MCTS and RandomPolicy are the reference's own object code)"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


KERNEL_SOURCES = {"ac_eval_kernel": "gk_eval.cu", "rollout_kernel": "gk_rollout.cu"}


def git_blob_sha(path):
    """the git blob id of a file's current content (sha1 of 'blob <len>\\0' + bytes), without calling git"""
    import hashlib
    data = open(path, "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def ncu_summary(kernel):
    """Per-launch figures of the committed `ncu --set full` capture of the bench-size launch (profiles/ncu_summary.json:
    dram bytes, warp instructions, issue-slot utilisation).  The capture is stamped with the git blob ids of the kernel
    sources it was taken from (scripts/ncu_summary.py); when the source has changed since, the figures describe another
    kernel and {} is returned, so that `traffic` and `roofline_issue` read null instead of a stale number."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if not os.path.exists(path):
        return {}
    try:
        doc = json.load(open(path))
        src = KERNEL_SOURCES.get(kernel)
        if src and doc.get("_sources", {}).get(src) != git_blob_sha(os.path.join(ROOT, "gomokuai_b200", "csrc", src)):
            return {}
        return doc.get(kernel, {})
    except Exception:
        return {}


_ISSUE_PEAKS = {}


def measure_issue_peaks(gk):
    """Live microbenchmark (gk_measure_issue_peak, csrc/gk_peaks.cu): G warp-inst/s of register-only LOP3, IMAD and 1:1 streams."""
    if not _ISSUE_PEAKS:
        for name, mode in (("alu_only_lop3", 0), ("fma_only_imad", 1), ("mixed_lop3_imad", 2)):
            _ISSUE_PEAKS[name] = max(gk.measure_issue_peak(mode) for _ in range(3)) / 1e9
    return _ISSUE_PEAKS


def issue_roofline(kernel, launch_ms, sm_mhz, sm_count=148):
    """Integer-issue roofline (the one that binds): warp instructions of one launch (from the committed ncu
    capture) / the rate a balanced LOP3 + IMAD stream sustains on this GPU, measured live by the microbenchmark
    (nominally SMs x 4 schedulers x 1 inst/clk x SM clock; a stream that uses only the ALU pipe gets half of that)."""
    summ = ncu_summary(kernel)
    if not summ.get("warp_inst_per_launch") or not sm_mhz:
        return None
    nominal = sm_count * 4 * sm_mhz * 1e6 / 1e9
    peak = _ISSUE_PEAKS.get("mixed_lop3_imad") or nominal
    achieved = summ["warp_inst_per_launch"] / (launch_ms * 1e-3) / 1e9
    return {"bound": "int32 issue slots", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s",
            "frac": achieved / peak, "peak_source": "measured live: LOP3 + IMAD 1:1 register-only stream" if _ISSUE_PEAKS else "nominal",
            "peak_nominal": nominal, "peaks_measured": dict(_ISSUE_PEAKS),
            "alu_pipe_pct_ncu": summ.get("alu_pipe_pct"), "active_lanes_per_inst": summ.get("threads_per_inst"),
            "source": "warp instructions per launch from profiles/ncu_summary.json, time and peak measured live"}


def gpu_config1(gk):
    """configs[0] through the drop-in: one tree, one leaf per playout, every simulate a 5-rollout GPU call (one fused launch
    + one synchronisation).  Launch-latency bound by construction; a single game does not shard, so rank 0 alone runs it."""
    from gomokuai_b200 import core
    core.seed(1)
    best, move, size = None, None, None
    for _ in range(2):
        b = core.Board()
        m = core.MCTS(c_iterations=10000, policy=core.RandomPolicy(5.0, 5))
        t0 = time.perf_counter()
        move = int(m.get_action(b))
        dt = time.perf_counter() - t0
        best, size = (dt if best is None else min(best, dt)), int(m.size)
    return {"workload": CONFIG1_WORKLOAD + ", through gomokuai_b200.core (CorePyExt mirror, simulate slot on the GPU)",
            "value": 10000 / best, "unit": "playouts/s", "rollouts_per_sec": 50000 / best, "seconds": best, "gpu_launches": 10000,
            "tree_nodes": size, "move": move, "replicas": "a single game does not shard: rank 0 only"}


RP_TREES_TOTAL, RP_PLAYOUTS = 1024, 1_000_000


def gpu_root_parallel(gk, torch, dist, dev, rank, world, barrier, moves=4):
    """configs[3]: root-parallel MCTS, 1M playouts per move over ALL the GPUs of the job (strong scaling): RP_TREES_TOTAL
    independent trees, cut evenly over the ranks, every tree does 1M / RP_TREES_TOTAL playouts from the same root with its own
    Philox stream, leaves simulated by the rollout kernel in one batch per round, then ONE allreduce(sum) of the int64[3][225]
    root statistics (gk_root_allreduce on the library's NCCL communicator) -- inside the timed region."""
    import math
    from gomokuai_b200 import core, root_parallel as rp
    trees = RP_TREES_TOTAL // world
    per_tree = math.ceil(RP_PLAYOUTS / RP_TREES_TOTAL)
    cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    share = cores[rank * len(cores) // world:(rank + 1) * len(cores) // world] or cores
    if world > 1 and hasattr(os, "sched_setaffinity"):
        os.sched_setaffinity(0, share)                                   # the ranks' worker threads get disjoint cores
    threads = max(1, len(share))
    s = core.RootParallelSearch(trees=trees, c_rollouts=5, seed=1, replica_base=rank * trees, threads=threads)
    if world > 1:
        rp._ensure_gk_comm()
    h_stats = torch.zeros((3, 225), dtype=torch.int64).pin_memory()
    d_stats = torch.zeros((3, 225), dtype=torch.int64, device=dev)
    b = core.Board()
    for c in (112, 113, 97, 98):
        b.apply_move(c)
    rows, equal = [], True
    for mv in range(1 + moves):                                          # the first move warms up (arenas, streams, communicator)
        seed = 1000 + mv
        barrier()
        t0 = time.perf_counter()
        local = s.run(b, per_tree, seed)
        t1 = time.perf_counter()
        h_stats.copy_(torch.from_numpy(local))
        d_stats.copy_(h_stats, non_blocking=True)
        if world > 1:
            gk.root_allreduce(d_stats)
        h_stats.copy_(d_stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        merged = h_stats.numpy().copy()
        t = torch.tensor([t2 - t0, t2 - t1, s.seconds_gpu, s.seconds_total], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        move = rp.best_move(merged)
        if rank == 0:
            # the same search on ONE rank with all the trees (tree streams are numbered globally, counts are integers)
            alone = core.RootParallelSearch(trees=RP_TREES_TOTAL, c_rollouts=5, seed=1, replica_base=0, threads=2 if world == 1 else threads)
            equal = equal and bool(np.array_equal(alone.run(b, per_tree, seed), merged))
            del alone
        if mv > 0:
            rows.append({"seconds": float(t[0]), "exchange_us": float(t[1]) * 1e6, "worker_wait_for_gpu_s": float(t[2]),
                         "search_s": float(t[3]), "move": int(move), "playouts": int(merged[0].sum()) + RP_TREES_TOTAL})
        b.apply_move(move)
        barrier()
    if rank != 0:
        return None
    total_s = sum(r["seconds"] for r in rows)
    playouts = sum(r["playouts"] for r in rows)
    search_s = statistics.mean(r["search_s"] for r in rows)
    return {"workload": "configs[3]: root-parallel MCTS, 1M playouts/move across the GPUs of the job with visit-count allreduce",
            "metric": "root-parallel MCTS playouts/sec", "value": playouts / total_s, "unit": "playouts/s", "scaling": "strong",
            "rollouts_per_sec": 5 * playouts / total_s, "seconds_per_move": total_s / len(rows),
            "trees_total": RP_TREES_TOTAL, "trees_per_gpu": trees, "playouts_per_tree": per_tree, "c_rollouts": 5,
            "host_threads_per_rank": threads, "cores_per_rank": len(share), "moves_timed": len(rows),
            "exchange_us_median": statistics.median(r["exchange_us"] for r in rows),
            "exchange": "5.4 KB H2D + gk_root_allreduce (ncclAllReduce int64[675], sum) + 5.4 KB D2H, inside the timed region" if world > 1
                        else "5.4 KB H2D + D2H (one rank: no collective)",
            "merged_equals_single_rank_search": equal,
            "limiter": {"host_us_per_playout_per_thread": search_s * threads / (trees * per_tree) * 1e6,
                        "worker_wait_for_gpu_frac": statistics.mean(r["worker_wait_for_gpu_s"] for r in rows) / max(search_s, 1e-9),
                        "note": "the tree stays on the host (north_star): a move costs trees_per_gpu x playouts_per_tree host playouts / "
                                "threads (every thread does tree work; whoever finishes a group's last tree launches its leaf batch); the ranks "
                                "of one box share its cores, so more GPUs add no host throughput"},
            "gpu_launches_per_move": per_tree * max(1, min(8, trees // 16)), "moves": rows}


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import gomokuai_b200 as gk

    arm = CpuArm() if (rank == 0 and world == 1 and not args.no_cpu) else None   # fork the workers before CUDA exists
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gk.init(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream()
    table = gk.default_table()
    if rank == 0:
        measure_issue_peaks(gk)                                           # a few ms, before the timed regions

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: each rank owns a disjoint range of the synthetic position set -------------------------
    n_eval, n_roll = EVAL_POSITIONS, ROLL_POSITIONS
    h_boards, moves, starts = gk.synth_positions(rank * n_eval, n_eval, want_moves=(rank == 0))
    d_boards = torch.from_numpy(h_boards.view(np.int32)).to(dev)
    d_roll = d_boards[:n_roll].contiguous()
    out = gk.eval_batch(d_boards, table)                               # allocates outputs once
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    sampler = ClockSampler(local_rank)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sampler.start()
        for a, b in evs:
            flush_buf.fill_(1)                                          # L2 flush between timed iterations (untimed)
            a.record(stream)
            fn()
            b.record(stream)
        barrier()
        sampler.stop()
        ms = [a.elapsed_time(b) for a, b in evs]
        total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)                # max over ranks
        return float(total.item()) / steps, ms

    # ---- K1: resident inputs ---------------------------------------------------------------------------------
    eval_ms, eval_each = timed(lambda: gk.eval_batch(d_boards, table, out=out), args.steps, args.warmup)
    # ---- K2: resident inputs ---------------------------------------------------------------------------------
    roll_out = {}

    def roll_step():
        roll_out["r"] = gk.rollout_batch(d_roll, ROLL_PER_POS, pos_base=rank * n_roll)
    roll_ms, _ = timed(roll_step, args.steps, args.warmup)
    wdb = roll_out["r"]["wdb"]
    assert bool((wdb.sum(dim=1) == ROLL_PER_POS).all()), "rollout counts do not add up"
    mean_len = None
    if rank == 0:
        tr = gk.rollout_batch(d_roll[:256], 256, pos_base=0, want_trace=True)
        torch.cuda.synchronize()
        mean_len = float(tr["lengths"].float().mean().item())

    # ---- e2e: HOST buffers through the C-ABI, copies inside the timed region ---------------------------------------
    e2e_steps = max(2, min(args.steps, 5))
    h_pinned = torch.from_numpy(h_boards.view(np.int32)).pin_memory()
    h_out = {
        "scores": torch.empty((n_eval, 4, 225), dtype=torch.int32).pin_memory(),
        "pat_totals": torch.empty((n_eval, 2, 8), dtype=torch.int16).pin_memory(),
        "cmp_totals": torch.empty((n_eval, 2, 3), dtype=torch.int16).pin_memory(),
        "winner": torch.empty((n_eval,), dtype=torch.int8).pin_memory(),
    }
    h_wdb = torch.empty((n_roll, 3), dtype=torch.int32).pin_memory()

    def wall(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        barrier()
        return float(dt.item()) / steps

    e2e_eval_s = wall(lambda: gk.eval_batch_host(h_pinned, table, out=h_out), e2e_steps, 1)
    e2e_roll_s = wall(lambda: gk.rollout_batch_host(h_pinned[:n_roll], ROLL_PER_POS, pos_base=rank * n_roll, out=h_wdb),
                      e2e_steps, 1)
    if rank == 0:                                                       # the e2e result is the same data as the device path
        assert torch.equal(h_out["scores"][:4096], out["scores"][:4096].cpu())
        assert torch.equal(h_wdb, wdb.cpu())
    # The ceiling of that copy when several GPUs deliver into ONE host memory system: every rank streams memsets with its
    # share of the host cores at the same moment; the sum over ranks is what the box's DRAM takes.
    host_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    bw_threads = max(1, host_cores // world)
    barrier()
    host_bw = torch.tensor([gk.measure_host_write_bw(bw_threads, 128 << 20, 3)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(host_bw, op=dist.ReduceOp.SUM)
    host_bw = float(host_bw.item())
    # ... and the platform's device-to-host DMA ceiling: every rank copies 1 GiB out of HBM into pinned memory at the same
    # moment, nothing else running (one GPU: its PCIe link; several: whatever the box's I/O fabric and memory take together)
    d_big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    h_big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    h_big.copy_(d_big)                                                  # touch the pages
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_big.copy_(d_big, non_blocking=True)
    torch.cuda.synchronize()
    dma_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(dma_s, op=dist.ReduceOp.MAX)
    dma_gbs = world * 3 * (1 << 30) / float(dma_s.item()) / 1e9
    del d_big, h_big

    # ---- the other configurations of BASELINE.json, briefly (same timing rules; extra objects of the JSON line) -----
    extras = {}
    pol = gk.eval_policy_batch(d_boards, table)                        # row f1: evaluator + policy heads, 904 B out per board
    pol_ms, _ = timed(lambda: gk.eval_policy_batch(d_boards, table), max(3, args.steps // 4), 2)
    h_pol = (torch.empty((n_eval, 225), dtype=torch.float32).pin_memory(), torch.empty((n_eval,), dtype=torch.float32).pin_memory(),
             torch.empty((n_eval,), dtype=torch.int8).pin_memory())
    pol_e2e_s = wall(lambda: gk.eval_policy_batch_host(h_pinned, table, out=h_pol), e2e_steps, 1)
    if rank == 0:
        assert torch.equal(h_pol[0][:2048], pol["probs"][:2048].cpu())
    extras["policy_heads"] = {"metric": "board evals/sec with policy heads (probs + value, Heuristic.hpp:16-45)",
                              "value": world * n_eval / (pol_ms * 1e-3), "unit": "boards/s", "ms_per_step": pol_ms,
                              "e2e": {"value": world * n_eval / pol_e2e_s, "unit": "boards/s", "h2d_bytes_per_step": n_eval * 64,
                                      "d2h_bytes_per_step": n_eval * 905, "ms_per_step": pol_e2e_s * 1e3}}
    del pol, h_pol
    n_games = 8192 // world                                             # configs[4]: 8192 concurrent guided self-play games across the GPUs of the job
    d_empty = torch.zeros((n_games, 16), dtype=torch.int32, device=dev)
    sp = {}

    def selfplay_step():
        sp["r"] = gk.guided_rollout_batch(d_empty, mode="sample", key=gk.SYNTH_KEY, game_base=rank * n_games, want_moves=True)
    sp_ms, _ = timed(selfplay_step, max(3, args.steps // 4), 2)
    sp_moves = float(sp["r"]["length"].float().sum().item())
    sp_full_ms, _ = timed(lambda: gk.guided_rollout_batch(d_empty, mode="sample", key=gk.SYNTH_KEY, game_base=rank * n_games,
                                                          want_moves=True, full_rescan=True), 3, 1)
    # the same concurrency as a STREAM of games (self-play data generation runs continuously): 8 batches' worth of games queued,
    # at most n_games in flight per GPU -- a batch lasts as long as its longest game, a stream runs at the rate of the mean
    d_queue = torch.zeros((8 * n_games, 16), dtype=torch.int32, device=dev)
    sp_stream_ms, _ = timed(lambda: gk.guided_rollout_batch(d_queue, mode="sample", key=gk.SYNTH_KEY, game_base=rank * 8 * n_games,
                                                            want_moves=True, max_in_flight=n_games), 3, 1)
    del d_queue
    extras["selfplay"] = {"metric": "pattern-guided self-play games/sec (configs[4]: 8192 concurrent games in total, sampled moves)",
                          "games_per_gpu": n_games, "scaling": "strong",
                          "value": world * n_games / (sp_ms * 1e-3), "unit": "games/s", "ms_per_step": sp_ms,
                          "evaluated_moves_per_sec": world * sp_moves / (sp_ms * 1e-3), "mean_game_length": sp_moves / n_games,
                          "kernel": "guided_pair_kernel (two warps per game) up to 15 games per SM in flight, guided_kernel above: per move only the "
                                    "four 13-symbol windows around the new stone are re-scanned (before / after)",
                          "stream": {"value": world * 8 * n_games / (sp_stream_ms * 1e-3), "unit": "games/s", "games_per_gpu": 8 * n_games,
                                     "in_flight_per_gpu": n_games, "ms_per_step": sp_stream_ms,
                                     "what": "gk_guided_rollout_queue: the same number of games in flight, worked through as a queue"},
                          "full_rescan": {"value": world * n_games / (sp_full_ms * 1e-3), "unit": "games/s", "ms_per_step": sp_full_ms,
                                          "kernel": "ac_eval_kernel<true, true>: the whole board after every move (identical games)"}}
    n_enc = 1 << 18
    d_last = torch.full((n_enc, 2), -1, dtype=torch.int16, device=dev)
    enc_ms, _ = timed(lambda: gk.encode_states_batch(d_boards[:n_enc], d_last, augment=True), max(3, args.steps // 4), 2)
    enc_gbs = n_enc * (64 + 4 + 10800) / (enc_ms * 1e-3) / 1e9
    extras["encode"] = {"metric": "augmented feature planes (8 x 6 x 15 x 15 uint8 per position)", "value": world * n_enc / (enc_ms * 1e-3),
                        "unit": "positions/s", "ms_per_step": enc_ms,
                        "roofline": {"bound": "hbm", "achieved": enc_gbs, "unit": "GB/s", "kernel": "encode_states_kernel"}}

    clocks = sampler.summary()
    extras["config1"] = gpu_config1(gk) if rank == 0 else None
    extras["root_parallel"] = gpu_root_parallel(gk, torch, dist, dev, rank, world, barrier)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload --------------------------------------
    cpu_eval = cpu_roll = None
    if arm is not None:
        n_c = 6144 * arm.cores                                           # ~4 s per core of evaluator replay
        v, w = arm.run("eval", moves, starts, n_c)
        cpu_eval = {"value": v, "unit": "boards/s", "cores": arm.cores, "kind": arm.kind, "per_core": v / arm.cores,
                    "sample": f"first {n_c} of the {n_eval} positions, replayed through Evaluator::applyMove, {w:.1f} s wall"}
        lv, lw = arm.run("linescan", moves, starts, n_c)                  # the other CPU figure of SURVEY 8d: whole-line scans
        cpu_eval["full_line_scans_per_s"] = lv
        cpu_eval["full_line_scan_sample"] = (f"PatternSearch::execute over the 88 padded line strings of the same {n_c} positions "
                                             f"(no score bookkeeping), {lw:.1f} s wall")
        n_p = 16 * arm.cores                                             # ~3 s per core of rollouts
        v, w = arm.run("rollout", moves, starts, n_p, 32768)
        cpu_roll = {"value": v, "unit": "rollouts/s", "cores": arm.cores, "kind": arm.kind, "per_core": v / arm.cores,
                    "sample": f"first {n_p} positions x 32768 rollouts, {w:.1f} s wall"}
        arm.close()

    if arm is not None and extras["config1"] is not None:
        extras["config1"]["cpu_baseline"] = reference_config1(2)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    eval_value = world * n_eval / (eval_ms * 1e-3)
    roll_value = world * n_roll * ROLL_PER_POS / (roll_ms * 1e-3)
    eval_gbs = n_eval * (EVAL_BYTES_IN + EVAL_BYTES_OUT) / (eval_ms * 1e-3) / 1e9
    line = {
        "metric": "board evals/sec (15x15)", "value": eval_value, "unit": "boards/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": eval_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": bench_config(),
        "timing": {"l2": "256 MB write between timed iterations; each step also streams 3.9 GB, > L2",
                   "clock": "CUDA events on the launching stream per step, max over ranks of the sum"},
        "line_scans_per_sec": eval_value * 72,
        "clocks": clocks,
        "e2e": {"value": world * n_eval / e2e_eval_s, "unit": "boards/s", "h2d_bytes_per_step": n_eval * 64,
                "d2h_bytes_per_step": n_eval * (3600 + 32 + 12 + 1), "ms_per_step": e2e_eval_s * 1e3,
                "host": {"achieved_write_gbs": world * n_eval * 3645 / e2e_eval_s / 1e9, "per_gpu_gbs": n_eval * 3645 / e2e_eval_s / 1e9,
                         "d2h_copy_ceiling_gbs": dma_gbs, "frac_of_d2h_copy_ceiling": world * n_eval * 3645 / e2e_eval_s / 1e9 / dma_gbs,
                         "cpu_stream_write_gbs": host_bw, "cpu_stream_threads": bw_threads * world,
                         "note": "two ceilings measured in the same job: d2h_copy_ceiling = every rank copying 1 GiB from HBM to pinned host "
                                 "memory at the same moment (one GPU: its PCIe link; several: the box's I/O path into ONE host memory "
                                 "system), cpu_stream_write = every rank's share of the host cores streaming memsets (what the DRAM itself takes)"},
                "compact": {"what": "gk_eval_policy_batch_host: probs f32[225] + value + winner = 905 B per board instead of 3645",
                            "value": world * n_eval / pol_e2e_s, "unit": "boards/s"},
                "note": "gk_eval_batch_host: pinned host buffers, 3 chunks in flight"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": eval_gbs, "peak": peak, "unit": "GB/s", "frac": eval_gbs / peak,
                     "traffic": ncu_summary("ac_eval_kernel").get("dram_bytes_per_launch"), "kernel": "ac_eval_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_board": EVAL_BYTES_IN + EVAL_BYTES_OUT,
                     "symbol_steps_per_sec": eval_value / world * EVAL_STEPS_ALGO,
                     "note": "the kernel is integer-issue / shared-memory-latency bound, not HBM bound (SURVEY 8d)"},
        "roofline_issue": issue_roofline("ac_eval_kernel", eval_ms, clocks.get("sm_mhz")),
        "cpu_baseline": cpu_eval,
        "rollouts": {
            "metric": "rollouts/sec (15x15)", "value": roll_value, "unit": "rollouts/s", "ms_per_step": roll_ms,
            "config": {"workload": "configs[2]: 4096 positions x 4096 random playouts per GPU (16,777,216 rollouts)",
                       "mean_rollout_length_moves": mean_len},
            "moves_per_sec": roll_value * mean_len if mean_len else None,
            "e2e": {"value": world * n_roll * ROLL_PER_POS / e2e_roll_s, "unit": "rollouts/s",
                    "h2d_bytes_per_step": n_roll * 64, "d2h_bytes_per_step": n_roll * 12, "ms_per_step": e2e_roll_s * 1e3},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": n_roll * (64 + 12) / (roll_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": n_roll * (64 + 12) / (roll_ms * 1e-3) / 1e9 / peak,
                         "traffic": ncu_summary("rollout_kernel").get("dram_bytes_per_launch"), "kernel": "rollout_kernel",
                         "note": "compute-bound by construction: 76 algorithmic bytes per position, amortised over 4096 playouts"},
            "roofline_issue": issue_roofline("rollout_kernel", roll_ms, clocks.get("sm_mhz")),
            "cpu_baseline": cpu_roll,
        },
    }
    extras["encode"]["roofline"].update({"peak": peak, "frac": extras["encode"]["roofline"]["achieved"] / peak})
    line.update(extras)
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything a library writes to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line goes to the
    saved descriptor through _emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT, text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""The root-parallel scheduler without a GPU: csrc/host/root_parallel.cpp linked against a stub of the C-ABI
(scripts/host_bw/rp_host_bench.cpp: hash-derived rollout counts after a modelled latency).  Whatever the number of threads
and of leaf batches in flight, a search visits the same leaves and ends with the same statistics; and ThreadSanitizer sees
no race between the threads that take chunks, launch a group's batch and wait for it."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["scripts/host_bw/rp_host_bench.cpp", "gomokuai_b200/csrc/host/root_parallel.cpp", "gomokuai_b200/csrc/host/mcts.cpp"]


def _build(tmp, name, flags):
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++17", "-mavx2", *flags, "-I", os.path.join(ROOT, "include"), *[os.path.join(ROOT, s) for s in SOURCES],
           "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return exe if r.returncode == 0 else None, r.stderr


def _run(exe, trees, per_tree, threads, latency_us, groups, reps=1):
    r = subprocess.run([exe, str(trees), str(per_tree), str(threads), str(latency_us), str(groups), str(reps)],
                       capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr:            # the runtime could not set up its shadow memory on this kernel
        pytest.skip(r.stderr.strip().splitlines()[0])
    assert r.returncode == 0, r.stderr
    return r.stdout + r.stderr, re.findall(r"(\d+) playouts .* nodes (\d+), best (\d+), chk ([0-9a-f]+)", r.stdout)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe, err = _build(str(tmp_path_factory.mktemp("rp")), "rp_host", ["-O2"])
    assert exe, err
    return exe


def test_statistics_do_not_depend_on_threads_or_groups(harness):
    want = None
    for threads, groups in ((1, 1), (2, 0), (3, 4), (5, 6), (8, 8), (8, 3), (4, 2)):
        _, rows = _run(harness, 200, 60, threads, 3, groups, reps=2)
        assert len(rows) == 2 and rows[0] == rows[1], (threads, groups, rows)          # a searcher is reusable
        want = want or rows[0]
        assert rows[0] == want, (threads, groups, rows[0], want)
    assert int(want[0]) > 0


def test_more_threads_than_trees_and_one_tree(harness):
    _, a = _run(harness, 3, 50, 8, 0, 0)
    _, b = _run(harness, 3, 50, 1, 0, 1)
    assert a == b
    _, c = _run(harness, 1, 200, 4, 0, 0)
    _, d = _run(harness, 1, 200, 1, 0, 0)
    assert c == d


def test_thread_sanitizer_sees_no_race(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe, err = _build(str(tmp_path), "rp_tsan", ["-O1", "-g", "-fsanitize=thread"])
    if exe is None:
        pytest.skip("this toolchain has no ThreadSanitizer runtime: " + err[-200:])
    for threads, groups in ((4, 6), (6, 2), (3, 8)):
        out, rows = _run(exe, 96, 40, threads, 5, groups)
        assert "ThreadSanitizer" not in out, out[-3000:]
        assert rows

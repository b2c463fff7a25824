"""Builds gomokuai_b200/lib/libgomoku_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libgomoku_b200.so")
SOURCES = ["gk_table.cpp", "gk_eval.cu", "gk_rollout.cu", "gk_rollout_warp.cu", "gk_encode.cu", "gk_peaks.cu", "gk_capi.cu"]
HEADERS = ["gk_format.h", "gk_table.h", "gk_kernels.h", os.path.join("..", "..", "include", "gomoku_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall"]
# --use_fast_math only where every result is an integer.  gk_eval.cu carries the float policy heads, which mirror the
# reference's IEEE float (Eigen) arithmetic: correctly rounded division and square root, no flush-to-zero, so that only
# the summation ORDER differs from the reference (tests/test_heads.py states the tolerance); -fmad=false for the same
# reason: the reference's x86-64 build rounds the product and the sum of `0.6 a + 0.4 b` separately.
FAST_MATH = {"gk_rollout.cu", "gk_rollout_warp.cu", "gk_encode.cu", "gk_peaks.cu", "gk_capi.cu"}
IEEE_FLAGS = ["-fmad=false"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into one shared library. Returns its path."""
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(["--use_fast_math"] if src in FAST_MATH else IEEE_FLAGS), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], check=True)
    return LIB


HOST = os.path.join(CSRC, "host")
HOST_SOURCES = ["module.cpp", "mcts.cpp", "root_parallel.cpp"]
HOST_HEADERS = ["game.h", "mcts.h", "root_parallel.h"]


def ext_path():
    import sysconfig
    return os.path.join(HERE, "CorePyExt" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pyext(force=False):
    """Compile the pybind11 module CorePyExt (host C++17 mirror of Board / Policy / MCTS) against the library."""
    import sysconfig
    import pybind11
    out = ext_path()
    deps = [os.path.join(HOST, f) for f in HOST_SOURCES + HOST_HEADERS] + [LIB]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           *[os.path.join(HOST, f) for f in HOST_SOURCES], "-o", out,
           "-L" + LIB_DIR, "-lgomoku_b200", "-Wl,-rpath,$ORIGIN/lib", "-lpthread"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_pyext(force="--force" in sys.argv))

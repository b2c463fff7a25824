"""The C-ABI library loads and exports every symbol include/gomoku_b200.h declares; without a
GPU (or before gk_init) every compute call fails loudly instead of falling back.  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gomoku_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gk_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(gk):
    lib = gk.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_compute_fails_loudly_without_a_device(gk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    lib = gk.lib()
    lib.gk_last_error.restype = ctypes.c_char_p
    assert lib.gk_init(0) == -2                                   # GK_ERR_NO_DEVICE
    assert b"CUDA" in lib.gk_last_error() or b"device" in lib.gk_last_error()
    t = gk.default_table()
    boards = np.zeros((4, 16), np.uint32)
    out = np.zeros((4, 4, 225), np.int32)
    rc = lib.gk_eval_batch_host(t.handle, boards.ctypes.data_as(ctypes.c_void_p), 4, out.ctypes.data_as(ctypes.c_void_p), None, None, None)
    assert rc == -5                                               # GK_ERR_NOT_INIT: no silent CPU path
    rc = lib.gk_rollout_batch_host(boards.ctypes.data_as(ctypes.c_void_p), 4, 8, ctypes.c_uint64(1), 0, 0, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == -5
    with pytest.raises(gk.GomokuB200Error):
        gk.eval_batch_host(boards)


def test_product_never_touches_the_oracle():
    """No file of the product package mentions the oracle library or module."""
    pkg = os.path.join(ROOT, "gomokuai_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in text and "libgomoku_oracle" not in text and "libgomoku_ref" not in text, f
                assert "gomoku_oracle.h" not in text, f

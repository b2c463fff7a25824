"""gk_scan_batch (PatternSearch::matches on the GPU) against the oracle generator.  Needs a B200."""
import numpy as np
import pytest

from test_oracle_vs_ref import _strings

pytestmark = pytest.mark.gpu


def test_scan_batch_equals_oracle(gpu, port):
    import torch
    strings = _strings(23, 200000)
    for tail in ([1] * 5, [1] * 9, [2] * 6 + [4], [4, 1, 1, 1, 1, 1, 1, 2], [2] * 12, []):
        strings.append(np.array([3, 4] + tail, np.uint8))
    starts = np.zeros(len(strings) + 1, np.int64)
    starts[1:] = np.cumsum([len(s) for s in strings])
    codes = np.concatenate(strings)
    want_p, want_o, want_c = port.scan_many(codes, starts)
    pids, offs, counts = gpu.scan_batch(codes, starts, max_per_string=16)
    torch.cuda.synchronize()
    counts = counts.cpu().numpy()
    assert np.array_equal(counts, want_c) and counts.max() <= 16
    mask = np.arange(16)[None, :] < counts[:, None]
    assert np.array_equal(pids.cpu().numpy()[mask], want_p)
    assert np.array_equal(offs.cpu().numpy()[mask], want_o)


def test_board_lines_with_minimal_padding(gpu, port):
    """1 leading + 2 trailing '?' reproduce the emissions of the reference's 6 + 6 layout on every line
    of length >= 5 (the padding the eval kernel's tape uses); shorter lines never emit."""
    import torch
    rng = np.random.default_rng(8)
    long_a, long_b, short = [], [], []
    for _ in range(40000):
        n = int(rng.integers(5, 16))
        body = rng.choice([1, 2, 4], size=n, p=[.3, .3, .4]).astype(np.uint8)
        long_a.append(np.concatenate([[3], body, [3, 3]]).astype(np.uint8))
        long_b.append(np.concatenate([[3] * 6, body, [3] * 6]).astype(np.uint8))
    for n in range(1, 5):
        for code in range(3 ** n):
            body = np.array([[1, 2, 4][(code // 3 ** i) % 3] for i in range(n)], np.uint8)
            short.append(np.concatenate([[3] * 6, body, [3] * 6]).astype(np.uint8))

    def run(strs):
        starts = np.zeros(len(strs) + 1, np.int64)
        starts[1:] = np.cumsum([len(s) for s in strs])
        p, o, c = gpu.scan_batch(np.concatenate(strs), starts, max_per_string=16)
        torch.cuda.synchronize()
        return p.cpu().numpy(), o.cpu().numpy(), c.cpu().numpy()
    pa, oa, ca = run(long_a)
    pb, ob, cb = run(long_b)
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb)
    mask = np.arange(16)[None, :] < ca[:, None]
    assert np.array_equal(oa[mask] - 1, ob[mask] - 6)
    assert run(short)[2].sum() == 0


def test_issue_peak_microbenchmark(gpu):
    """gk_measure_issue_peak (the denominator of bench.py's integer-issue roofline): a LOP3-only stream is held to the
    ALU pipe's rate, a balanced LOP3 + IMAD stream gets close to one instruction per scheduler and clock."""
    info = gpu.device_info()
    alu, fma, mixed = (max(gpu.measure_issue_peak(m) for _ in range(2)) for m in (0, 1, 2))
    nominal = info["sm_count"] * 4 * 1.9e9
    assert 0.3 * nominal < alu < 0.75 * nominal and 0.3 * nominal < fma < 0.75 * nominal
    assert 0.7 * nominal < mixed < 1.15 * nominal and mixed > 1.4 * alu
    with pytest.raises(gpu.GomokuB200Error):
        gpu.measure_issue_peak(3)

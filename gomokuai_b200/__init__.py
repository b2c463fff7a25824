"""gomokuai_b200 -- B200-native hot path of Vigilans/GomokuAI (pattern evaluation + random rollouts).

This module is a thin ctypes binding over the C-ABI in ``include/gomoku_b200.h``
(``gomokuai_b200/lib/libgomoku_b200.so``, hand-written sm_100a CUDA).  PyTorch is used only to own
device memory and streams.  There is no CPU fallback: if the shared library is missing, or no
sm_100 GPU is visible, every compute entry point raises.

Build the library in-tree with ``python -m gomokuai_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgomoku_b200.so")

CELLS = 225
BOARD_WORDS = 16
PATTERN_TYPES = ("DeadOne", "LiveOne", "DeadTwo", "LiveTwo", "DeadThree", "LiveThree", "DeadFour", "LiveFour", "Five")
COMPOUND_TYPES = ("DoubleThree", "FourThree", "DoubleFour")
SYNTH_KEY = 0x474F4D4F4B5531


class GomokuB200Error(RuntimeError):
    pass


_lib = None
_device = None


def lib():
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GomokuB200Error(
                f"{LIB_PATH} is missing: build it with `python -m gomokuai_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.gk_last_error.restype = ctypes.c_char_p
        L.gk_version.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise GomokuB200Error(f"gk status {status}: {lib().gk_last_error().decode()}")


def init(device=None):
    """Bind this process to one GPU (one process per GPU). Returns the device index."""
    global _device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    _check(lib().gk_init(int(device)))
    _device = int(device)
    return _device


def shutdown():
    """Give back the library's streams, device buffers, page-locked staging and NCCL communicator (gk_shutdown)."""
    global _device
    _check(lib().gk_shutdown())
    _device = None


def _require_init():
    if _device is None:
        init()
    return _device


def device_info():
    _require_init()
    d, sm, maj, mnr = (ctypes.c_int() for _ in range(4))
    _check(lib().gk_device_info(ctypes.byref(d), ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)))
    return {"device": d.value, "sm_count": sm.value, "cc": (maj.value, mnr.value)}


# ---- tables ------------------------------------------------------------------------------------------
def measure_issue_peak(mode):
    """Sustained warp instructions / s of a register-only stream (0: LOP3 only, 1: IMAD only, 2: both 1:1)."""
    _require_init()
    v = ctypes.c_double()
    _check(lib().gk_measure_issue_peak(int(mode), ctypes.byref(v)))
    return v.value


class Table:
    """Compiled pattern automaton (``gk_table``). ``Table()`` is the reference's default pattern set."""

    def __init__(self, protos=None, types=None, scores=None):
        self._h = ctypes.c_void_p()
        self._owned = protos is not None
        if protos is None:
            _check(lib().gk_table_default(ctypes.byref(self._h)))
        else:
            n = len(protos)
            arr = (ctypes.c_char_p * n)(*[p.encode() for p in protos])
            _check(lib().gk_table_build(arr, (ctypes.c_int * n)(*types), (ctypes.c_int * n)(*scores), n,
                                        ctypes.byref(self._h)))

    def __del__(self):
        if getattr(self, "_owned", False) and self._h and _lib is not None:
            _lib.gk_table_free(self._h)

    @property
    def handle(self):
        return self._h

    def info(self):
        a, b, c, d = (ctypes.c_int() for _ in range(4))
        _check(lib().gk_table_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return {"n_states": a.value, "n_patterns": b.value, "trail_pad": c.value, "tape_steps": d.value}

    def pattern(self, i):
        s = ctypes.create_string_buffer(8)
        fav, typ, sc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().gk_table_pattern(self._h, int(i), s, ctypes.byref(fav), ctypes.byref(typ), ctypes.byref(sc)))
        return (s.raw.split(b"\0")[0].decode(), fav.value, typ.value, sc.value)

    def patterns(self):
        return [self.pattern(i) for i in range(self.info()["n_patterns"])]

    def entries(self):
        """Flat transition words, shape [n_states, 4] (see csrc/gk_format.h)."""
        n = self.info()["n_states"]
        out = np.zeros(n * 4, np.uint32)
        _check(lib().gk_table_entries(self._h, out.ctypes.data_as(ctypes.c_void_p), out.size))
        return out.reshape(n, 4)


    def device_format(self):
        """The automaton as the eval kernel reads it (csrc/gk_format.h): dict(next[n_rows,4] u16 row offsets
        indexed by raw cell value, erec[n_clones] u32, n_clones, root_off, start_off, list_cap, tape_steps)."""
        info = (ctypes.c_int * 6)()
        _check(lib().gk_table_device_format(self._h, info, None, 0, None, 0))
        nxt = np.zeros(info[0] * 4, np.uint16)
        erec = np.zeros(max(info[1], 1), np.uint32)
        _check(lib().gk_table_device_format(self._h, info, nxt.ctypes.data_as(ctypes.c_void_p), nxt.size,
                                            erec.ctypes.data_as(ctypes.c_void_p), erec.size))
        return {"next": nxt.reshape(-1, 4), "erec": erec[:info[1]], "n_clones": info[1], "root_off": info[2],
                "start_off": info[3], "list_cap": info[4], "tape_steps": info[5]}

    def flush(self):
        """Per state: pattern id owed when the input ends there (-1 = none)."""
        n = self.info()["n_states"]
        out = np.zeros(n, np.int16)
        _check(lib().gk_table_flush(self._h, out.ctypes.data_as(ctypes.c_void_p), out.size))
        return out


_default_table = None


def default_table():
    global _default_table
    if _default_table is None:
        _default_table = Table()
    return _default_table


# ---- host utilities ------------------------------------------------------------------------------------
def pack_moves(moves, starts):
    """Move lists (black first, alternating) -> uint32[n,16] packed boards (numpy, host)."""
    moves = np.ascontiguousarray(moves, np.int16)
    starts = np.ascontiguousarray(starts, np.int64)
    n = len(starts) - 1
    boards = np.zeros((n, BOARD_WORDS), np.uint32)
    _check(lib().gk_pack_moves(moves.ctypes.data_as(ctypes.c_void_p), starts.ctypes.data_as(ctypes.c_void_p), n,
                               boards.ctypes.data_as(ctypes.c_void_p)))
    return boards


def synth_positions(first, n, want_moves=True):
    """The synthetic random mid-game set of BASELINE.json. Returns (boards[n,16] u32, moves i16, starts i64)."""
    boards = np.zeros((n, BOARD_WORDS), np.uint32)
    moves = np.zeros(96 * max(n, 1), np.int16) if want_moves else None
    starts = np.zeros(n + 1, np.int64) if want_moves else None
    _check(lib().gk_synth_positions(ctypes.c_int64(first), int(n), boards.ctypes.data_as(ctypes.c_void_p),
                                    moves.ctypes.data_as(ctypes.c_void_p) if want_moves else None,
                                    starts.ctypes.data_as(ctypes.c_void_p) if want_moves else None))
    if want_moves:
        moves = moves[:int(starts[-1])].copy()
    return boards, moves, starts


def unpack_boards(boards):
    """uint32[n,16] -> uint8[n,225] cell values (0 empty, 1 black, 2 white)."""
    boards = np.asarray(boards, np.uint32).reshape(-1, BOARD_WORDS)
    c = np.arange(CELLS)
    return ((boards[:, c >> 4] >> ((c & 15) * 2).astype(np.uint32)) & 3).astype(np.uint8)


# ---- device entry points (torch tensors on the bound GPU) ---------------------------------------------
def nccl_unique_id():
    """128-byte NCCL id made by rank 0 (gk_nccl_unique_id); hand it to the other ranks by any side channel."""
    buf = (ctypes.c_uint8 * 128)()
    _check(lib().gk_nccl_unique_id(buf))
    return bytes(buf)


def nccl_init(unique_id, world_size, rank):
    """Join the library's own communicator (one process per GPU, after init())."""
    _require_init()
    buf = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
    _check(lib().gk_nccl_init(buf, int(world_size), int(rank)))


def nccl_shutdown():
    _check(lib().gk_nccl_shutdown())


def root_allreduce(stats, stream=None):
    """In-place sum of the int64[3,225] root statistics (a CUDA tensor) over the ranks of the library's
    communicator: the single collective of the path (gk_root_allreduce, NCCL over NVLink)."""
    torch = _torch()
    if not stats.is_cuda or stats.dtype != torch.int64 or stats.numel() != 3 * CELLS or not stats.is_contiguous():
        raise GomokuB200Error("stats must be a contiguous CUDA int64 tensor with 3*225 elements")
    _check(lib().gk_root_allreduce(None, _ptr(stats), _stream_ptr(stream)))
    return stats


def _torch():
    import torch
    return torch


def _stream_ptr(stream):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _as_board_tensor(boards):
    torch = _torch()
    dev = torch.device("cuda", _require_init())
    if isinstance(boards, np.ndarray):
        boards = torch.from_numpy(boards.view(np.int32)).to(dev)
    if boards.dtype != torch.int32 or boards.dim() != 2 or boards.shape[1] != BOARD_WORDS or not boards.is_contiguous():
        raise GomokuB200Error("boards must be a contiguous int32 tensor of shape [n, 16]")
    if not boards.is_cuda:
        boards = boards.to(dev)
    return boards


def eval_batch(boards, table=None, want_scores=True, out=None, stream=None):
    """Evaluate packed boards on the GPU. Returns dict(scores[n,4,225] i32, pat_totals[n,2,8] i16,
    cmp_totals[n,2,3] i16, winner[n] i8) of CUDA tensors (totals are uint16 bit patterns in int16)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    table = table or default_table()
    n = boards.shape[0]
    dev = boards.device
    if out is None:
        out = {
            "scores": torch.empty((n, 4, CELLS), dtype=torch.int32, device=dev) if want_scores else None,
            "pat_totals": torch.empty((n, 2, 8), dtype=torch.int16, device=dev),
            "cmp_totals": torch.empty((n, 2, 3), dtype=torch.int16, device=dev),
            "winner": torch.empty((n,), dtype=torch.int8, device=dev),
        }
    _check(lib().gk_eval_batch(table.handle, _ptr(boards), n, _ptr(out.get("scores")), _ptr(out["pat_totals"]),
                               _ptr(out["cmp_totals"]), _ptr(out["winner"]), _stream_ptr(stream)))
    return out


def eval_batch_host(boards, table=None, want_scores=True, out=None):
    """Same through HOST buffers (numpy arrays or pinned CPU tensors): copies are part of the call."""
    table = table or default_table()
    _require_init()
    b = boards.numpy() if hasattr(boards, "numpy") else boards
    b = np.ascontiguousarray(b).view(np.uint32).reshape(-1, BOARD_WORDS)
    n = b.shape[0]
    if out is None:
        out = {
            "scores": np.empty((n, 4, CELLS), np.int32) if want_scores else None,
            "pat_totals": np.empty((n, 2, 8), np.uint16),
            "cmp_totals": np.empty((n, 2, 3), np.uint16),
            "winner": np.empty((n,), np.int8),
        }

    def hp(x):
        if x is None:
            return None
        return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else x.ctypes.data_as(ctypes.c_void_p)

    _check(lib().gk_eval_batch_host(table.handle, b.ctypes.data_as(ctypes.c_void_p), n, hp(out.get("scores")),
                                    hp(out["pat_totals"]), hp(out["cmp_totals"]), hp(out["winner"])))
    return out


def eval_policy_batch(boards, table=None, want_scores=False, stream=None):
    """gk_eval_policy_batch: the evaluator's policy heads for the side to move (Heuristic::EvaluationProbs /
    EvaluationValue, include/algorithms/Heuristic.hpp:16-45).  Returns dict(probs[n,225] f32, value[n] f32,
    winner[n] i8, pat_totals, cmp_totals and, if want_scores, scores[n,4,225] i32) of CUDA tensors."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    table = table or default_table()
    n, dev = boards.shape[0], boards.device
    out = {
        "probs": torch.empty((n, CELLS), dtype=torch.float32, device=dev),
        "value": torch.empty((n,), dtype=torch.float32, device=dev),
        "scores": torch.empty((n, 4, CELLS), dtype=torch.int32, device=dev) if want_scores else None,
        "pat_totals": torch.empty((n, 2, 8), dtype=torch.int16, device=dev),
        "cmp_totals": torch.empty((n, 2, 3), dtype=torch.int16, device=dev),
        "winner": torch.empty((n,), dtype=torch.int8, device=dev),
    }
    _check(lib().gk_eval_policy_batch(table.handle, _ptr(boards), n, _ptr(out["probs"]), _ptr(out["value"]),
                                      _ptr(out["scores"]), _ptr(out["pat_totals"]), _ptr(out["cmp_totals"]),
                                      _ptr(out["winner"]), _stream_ptr(stream)))
    return out


def hybrid_simulate_batch(boards, table=None, want_flags=False, stream=None):
    """TraditionalPolicy::hybridSimulate for a batch (include/policies/Traditional.h:49-69): probs after
    Heuristic::DecisiveFilter, value, winner (+ the per-cell decisive flag words if want_flags)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    table = table or default_table()
    n, dev = boards.shape[0], boards.device
    out = {"probs": torch.empty((n, CELLS), dtype=torch.float32, device=dev),
           "value": torch.empty((n,), dtype=torch.float32, device=dev),
           "winner": torch.empty((n,), dtype=torch.int8, device=dev),
           "dflags": torch.empty((n, CELLS), dtype=torch.int32, device=dev) if want_flags else None}
    _check(lib().gk_hybrid_simulate_batch(table.handle, _ptr(boards), n, _ptr(out["probs"]), _ptr(out["value"]),
                                          _ptr(out["winner"]), _ptr(out["dflags"]), _stream_ptr(stream)))
    return out


def eval_policy_batch_host(boards, table=None, decisive=False, out=None):
    """Same through HOST buffers: returns (probs[n,225] f32, value[n] f32, winner[n] i8) -- numpy arrays, or the
    pinned CPU tensors passed as out=(probs, value, winner).  decisive=True applies Heuristic::DecisiveFilter
    (= gk_hybrid_simulate_batch_host)."""
    table = table or default_table()
    _require_init()
    b = boards.numpy() if hasattr(boards, "numpy") else boards
    b = np.ascontiguousarray(b).view(np.uint32).reshape(-1, BOARD_WORDS)
    n = b.shape[0]
    if out is None:
        out = (np.empty((n, CELLS), np.float32), np.empty((n,), np.float32), np.empty((n,), np.int8))

    def hp(x):
        return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else x.ctypes.data_as(ctypes.c_void_p)

    fn = lib().gk_hybrid_simulate_batch_host if decisive else lib().gk_eval_policy_batch_host
    _check(fn(table.handle, b.ctypes.data_as(ctypes.c_void_p), n, hp(out[0]), hp(out[1]), hp(out[2])))
    return out


GUIDED_FULL_RESCAN = 0x100
GUIDED_SINGLE_WARP = 0x200


def guided_rollout_batch(boards, mode="max", key=SYNTH_KEY, ctr_hi=0, game_base=0, max_moves=CELLS, want_moves=True,
                         table=None, stream=None, full_rescan=False, max_in_flight=0, single_warp=False):
    """Pattern-guided playouts (Heuristic::EvaluatedRollout, include/algorithms/Heuristic.hpp:61-91), one warp per game,
    whole games inside one kernel.  mode "max" = MaxEvaluatedRollout, "sample" = RandomEvaluatedRollout.
    full_rescan: re-evaluate the whole board after every move instead of the four lines through the new stone (slower;
    both play identical games).  max_in_flight > 0: at most that many games are played at the same time, the others queue
    (gk_guided_rollout_queue: a stream of games with fixed concurrency); the results do not depend on it.
    single_warp: keep every game on one warp even when few are in flight (else two warps per game there; same games).
    Returns dict(winner[n] i8, length[n] i16, moves[n,max_moves] i16 (-1 padded), final_boards[n,16] i32)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    table = table or default_table()
    n, dev = boards.shape[0], boards.device
    out = {
        "winner": torch.empty((n,), dtype=torch.int8, device=dev),
        "length": torch.empty((n,), dtype=torch.int16, device=dev),
        "moves": torch.full((n, max_moves), -1, dtype=torch.int16, device=dev) if want_moves else None,
        "final_boards": torch.empty((n, BOARD_WORDS), dtype=torch.int32, device=dev),
    }
    _check(lib().gk_guided_rollout_queue(table.handle, _ptr(boards), n, int(max_in_flight),
                                         {"max": 1, "sample": 2}[mode] | (GUIDED_FULL_RESCAN if full_rescan else 0) | (GUIDED_SINGLE_WARP if single_warp else 0),
                                         ctypes.c_uint64(key),
                                         ctypes.c_uint32(ctr_hi), int(game_base), int(max_moves), _ptr(out["winner"]),
                                         _ptr(out["length"]), _ptr(out["moves"]), _ptr(out["final_boards"]),
                                         _stream_ptr(stream)))
    return out


def expand_games(boards, games, stream=None):
    """Every position the games of a guided_rollout_batch result went through (the position BEFORE each ply).
    Returns dict(boards[p,16] i32, last_moves[p,2] i16, z[p] i8 outcome for the side to move, game[p] i64, starts[n] i64)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    n = boards.shape[0]
    lengths = games["length"]
    starts = torch.cumsum(lengths.to(torch.int64), 0) - lengths.to(torch.int64)
    total = int(lengths.to(torch.int64).sum().item())
    dev = boards.device
    out = {"boards": torch.empty((total, BOARD_WORDS), dtype=torch.int32, device=dev),
           "last_moves": torch.empty((total, 2), dtype=torch.int16, device=dev),
           "z": torch.empty((total,), dtype=torch.int8, device=dev), "starts": starts,
           "game": torch.repeat_interleave(torch.arange(n, device=dev), lengths.to(torch.int64))}
    moves = games["moves"]
    _check(lib().gk_expand_games(_ptr(boards), _ptr(moves), _ptr(lengths), _ptr(games["winner"]), n, int(moves.shape[1]), _ptr(starts),
                                 _ptr(out["boards"]), _ptr(out["last_moves"]), _ptr(out["z"]), _stream_ptr(stream)))
    return out


def rollout_batch(boards, rollouts_per_pos, key=SYNTH_KEY, ctr_hi=0, pos_base=0, want_trace=False, stream=None):
    """Random playouts. Returns dict(wdb[n,3] i32 = {white, draw, black}, winners[n,R] i8, lengths[n,R] u8)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    n = boards.shape[0]
    dev = boards.device
    wdb = torch.empty((n, 3), dtype=torch.int32, device=dev)
    winners = torch.empty((n, rollouts_per_pos), dtype=torch.int8, device=dev) if want_trace else None
    lengths = torch.empty((n, rollouts_per_pos), dtype=torch.uint8, device=dev) if want_trace else None
    _check(lib().gk_rollout_batch(_ptr(boards), n, int(rollouts_per_pos), ctypes.c_uint64(key), ctypes.c_uint32(ctr_hi),
                                  int(pos_base), _ptr(wdb), _ptr(winners), _ptr(lengths), _stream_ptr(stream)))
    return {"wdb": wdb, "winners": winners, "lengths": lengths}


def rollout_batch_host(boards, rollouts_per_pos, key=SYNTH_KEY, ctr_hi=0, pos_base=0, out=None):
    _require_init()
    b = boards.numpy() if hasattr(boards, "numpy") else boards
    b = np.ascontiguousarray(b).view(np.uint32).reshape(-1, BOARD_WORDS)
    n = b.shape[0]
    wdb = out if out is not None else np.empty((n, 3), np.int32)
    p = ctypes.c_void_p(wdb.data_ptr()) if hasattr(wdb, "data_ptr") else wdb.ctypes.data_as(ctypes.c_void_p)
    _check(lib().gk_rollout_batch_host(b.ctypes.data_as(ctypes.c_void_p), n, int(rollouts_per_pos), ctypes.c_uint64(key),
                                       ctypes.c_uint32(ctr_hi), int(pos_base), p))
    return wdb


def measure_host_write_bw(threads, bytes_per_thread=256 << 20, repeats=3):
    """GB/s of `threads` CPU threads streaming memsets into host memory (gk_measure_host_write_bw); needs no GPU."""
    out = ctypes.c_double()
    _check(lib().gk_measure_host_write_bw(int(threads), ctypes.c_size_t(bytes_per_thread), int(repeats), ctypes.byref(out)))
    return out.value


def rollout_trace_host(board, rollouts, key=SYNTH_KEY, ctr_hi=0, pos=0):
    """gk_rollout_trace_host: `rollouts` (<= 256) playouts of ONE position with their move lists.
    Returns dict(winners i8[R], lengths u8[R], moves u8[R,225])."""
    _require_init()
    b = np.ascontiguousarray(board).view(np.uint32).reshape(BOARD_WORDS)
    winners = np.zeros(rollouts, np.int8)
    lengths = np.zeros(rollouts, np.uint8)
    moves = np.zeros((rollouts, CELLS), np.uint8)
    _check(lib().gk_rollout_trace_host(b.ctypes.data_as(ctypes.c_void_p), int(rollouts), ctypes.c_uint64(key), ctypes.c_uint32(ctr_hi),
                                       int(pos), winners.ctypes.data_as(ctypes.c_void_p), lengths.ctypes.data_as(ctypes.c_void_p),
                                       moves.ctypes.data_as(ctypes.c_void_p)))
    return {"winners": winners, "lengths": lengths, "moves": moves}


def rollout_submit_host(slot, boards, rollouts_per_pos, wdb, key=SYNTH_KEY, ctr_hi=0, pos_base=0):
    """Asynchronous gk_rollout_batch_host on stream `slot` (0..7): returns at once; `boards` (uint32[n,16]) and `wdb`
    (int32[n,3]) are numpy arrays that must stay alive and untouched until rollout_wait(slot)."""
    _require_init()
    b = boards.view(np.uint32).reshape(-1, BOARD_WORDS)
    if not (b.flags.c_contiguous and wdb.flags.c_contiguous and wdb.dtype == np.int32 and wdb.size == 3 * b.shape[0]):
        raise GomokuB200Error("boards must be contiguous uint32[n,16] and wdb contiguous int32[n,3]")
    _check(lib().gk_rollout_submit_host(int(slot), b.ctypes.data_as(ctypes.c_void_p), b.shape[0], int(rollouts_per_pos),
                                        ctypes.c_uint64(key), ctypes.c_uint32(ctr_hi), int(pos_base),
                                        wdb.ctypes.data_as(ctypes.c_void_p)))


def rollout_wait(slot):
    _check(lib().gk_rollout_wait(int(slot)))


def rollout_injected(boards, r_stream, stream=None):
    """Playouts driven by an explicit start-index stream r_stream[n, R, stride] (uint8, values 0..224)."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    n = boards.shape[0]
    if isinstance(r_stream, np.ndarray):
        r_stream = torch.from_numpy(np.ascontiguousarray(r_stream, np.uint8)).to(boards.device)
    if r_stream.dim() != 3 or r_stream.shape[0] != n or r_stream.dtype != torch.uint8 or not r_stream.is_contiguous():
        raise GomokuB200Error("r_stream must be a contiguous uint8 tensor of shape [n, R, stride]")
    R, stride = r_stream.shape[1], r_stream.shape[2]
    winners = torch.empty((n, R), dtype=torch.int8, device=boards.device)
    lengths = torch.empty((n, R), dtype=torch.uint8, device=boards.device)
    _check(lib().gk_rollout_injected(_ptr(boards), n, R, _ptr(r_stream), stride, _ptr(winners), _ptr(lengths),
                                     _stream_ptr(stream)))
    return {"winners": winners, "lengths": lengths}


def encode_states_batch(boards, last_moves=None, augment=False, probs=None, stream=None):
    """Board.encoded_states() for a batch (core/py_ext/src/game_ext.hpp:87-104) and, with augment=True,
    the 8 rotations / reflections of augment_game_data (network/data_helper.py:36-55).
    last_moves: int16[n,2] = (last, second-to-last) cell ids, -1 = none.  probs: float32[n,225] (optional).
    Returns planes uint8[n,V,6,15,15] (V = 8 or 1) and, if probs is given, float32[n,V,225]."""
    torch = _torch()
    boards = _as_board_tensor(boards)
    n, dev, V = boards.shape[0], boards.device, (8 if augment else 1)
    if last_moves is not None:
        if isinstance(last_moves, np.ndarray):
            last_moves = torch.from_numpy(np.ascontiguousarray(last_moves, np.int16))
        last_moves = last_moves.to(dev).contiguous()
        if last_moves.dtype != torch.int16 or tuple(last_moves.shape) != (n, 2):
            raise GomokuB200Error("last_moves must be int16[n, 2]")
    probs_out = None
    if probs is not None:
        if isinstance(probs, np.ndarray):
            probs = torch.from_numpy(np.ascontiguousarray(probs, np.float32))
        probs = probs.to(dev).contiguous()
        if probs.dtype != torch.float32 or tuple(probs.shape) != (n, CELLS):
            raise GomokuB200Error("probs must be float32[n, 225]")
        probs_out = torch.empty((n, V, CELLS), dtype=torch.float32, device=dev)
    planes = torch.empty((n, V, 6, 15, 15), dtype=torch.uint8, device=dev)
    _check(lib().gk_encode_states_batch(_ptr(boards), _ptr(last_moves), n, int(bool(augment)), _ptr(planes),
                                        _ptr(probs), _ptr(probs_out), _stream_ptr(stream)))
    return (planes, probs_out) if probs is not None else planes


def scan_batch(codes, starts, table=None, max_per_string=16, stream=None):
    """PatternSearch::matches for a batch of symbol strings (codes 1..4). Returns (pids, offsets, counts)."""
    torch = _torch()
    dev = torch.device("cuda", _require_init())
    table = table or default_table()
    codes_t = torch.from_numpy(np.ascontiguousarray(codes, np.uint8)).to(dev)
    starts_t = torch.from_numpy(np.ascontiguousarray(starts, np.int64)).to(dev)
    ns = len(starts) - 1
    pids = torch.full((ns, max_per_string), -1, dtype=torch.int32, device=dev)
    offs = torch.full((ns, max_per_string), -1, dtype=torch.int32, device=dev)
    counts = torch.zeros((ns,), dtype=torch.int32, device=dev)
    _check(lib().gk_scan_batch(table.handle, _ptr(codes_t), _ptr(starts_t), ns, int(max_per_string), _ptr(pids),
                               _ptr(offs), _ptr(counts), _stream_ptr(stream)))
    return pids, offs, counts

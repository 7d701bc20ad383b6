"""ctypes binding of libnoc_b200.so — the C-ABI in include/lqr_control/noc_b200.h.

This is the Python face of the drop-in boundary: every call goes through the shared library's
`extern "C"` entry points (plain pointers and sizes).  Arguments may be numpy arrays (host buffers:
the library stages them over PCIe) or torch CUDA tensors (device buffers: used in place).

There is no CPU fallback.  If the library is missing, or no B200-class GPU is present,
`Solver()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libnoc_b200.so")

# enums (mirror the header)
OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
COMM_ID_BYTES = 128
STATUS_OK, STATUS_NOT_PD, STATUS_MAX_ITER, STATUS_LS_FAIL, STATUS_REG_FAIL = 0, 1, 2, 3, 4
MODEL_QUAD12, MODEL_DUAL24, MODEL_LTV_USER = 0, 1, 2
LAYOUT_A_THEN_B, LAYOUT_ROWS_AB = 0, 1
FLAG_NONE, FLAG_ASYNC, FLAG_TIME_PARALLEL = 0, 1, 2
OPT_SCRATCH_LIMIT_BYTES, OPT_CHUNK_WAVES, OPT_SINGLE_STREAM, OPT_TIME_PARALLEL_DFMA, OPT_TIME_PARALLEL_CHUNKS, \
    OPT_NO_SMALL_BATCH_KERNEL, OPT_NO_TENSOR_SWEEP, OPT_TENSOR_SWEEP_VARIANT, OPT_NO_WAVE_PLAN, \
    OPT_DEFERRED_BACKTRACK, OPT_NO_SPECULATIVE_BACKTRACK = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
REG_X10, REG_ADAPTIVE = 0, 1

EXPORTS = [
    "noc_create", "noc_destroy", "noc_last_error", "noc_version", "noc_set_stream", "noc_synchronize", "noc_set_option",
    "noc_launch_count", "noc_device_alloc", "noc_device_free", "noc_memcpy", "noc_problem_defaults", "noc_lqr_solve_batch", "noc_linearize_batch",
    "noc_rollout_open_loop", "noc_lqr_solve_from_x0", "noc_lqr_solve_along", "noc_dare_solve_batch", "noc_dare_solve_at", "noc_slq_solve_batch", "noc_slq_solve_warm",
    "noc_argmin_cost", "noc_argmin_cost_async", "noc_comm_unique_id", "noc_comm_create_rank", "noc_comm_create_all",
    "noc_comm_destroy", "noc_comm_rank", "noc_comm_size", "noc_comm_last_error", "noc_comm_group_begin",
    "noc_comm_group_end", "noc_comm_allgather", "noc_comm_best_rollout", "noc_comm_lqr_solve_from_x0",
    "noc_profile_enable", "noc_profile_read", "noc_profile_reset", "noc_selftest_rcp", "noc_plan_time_parallel",
]
KERNEL_SETUP, KERNEL_ROLLOUT, KERNEL_LINEARIZE, KERNEL_RICCATI, KERNEL_FORWARD, KERNEL_ARGMIN = range(6)


class Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32), ("model", C.c_int32),
                ("layout", C.c_int32), ("max_iter", C.c_int32), ("n_alpha", C.c_int32), ("flags", C.c_int32),
                ("dt", C.c_double), ("tol", C.c_double), ("reg0", C.c_double), ("armijo", C.c_double),
                ("u_min", C.c_double), ("u_max", C.c_double), ("reg_schedule", C.c_int32), ("reserved_", C.c_int32)]


class CostIndex(C.Structure):
    """noc_cost_index_t: the winner of a "pick the best rollout" (16 bytes, also the device-side layout)."""
    _fields_ = [("cost", C.c_double), ("index", C.c_int64)]


class NocError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnoc_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """dlopen the product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C noc_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    pp = C.POINTER(Problem)
    lib.noc_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.noc_create.restype = C.c_int
    lib.noc_destroy.argtypes = [vp]
    lib.noc_destroy.restype = None
    lib.noc_last_error.argtypes = [vp]
    lib.noc_last_error.restype = C.c_char_p
    lib.noc_version.restype = C.c_char_p
    lib.noc_set_stream.argtypes = [vp, vp]
    lib.noc_synchronize.argtypes = [vp]
    lib.noc_set_option.argtypes = [vp, C.c_int, i64]
    lib.noc_launch_count.argtypes = [vp]
    lib.noc_launch_count.restype = i64
    lib.noc_problem_defaults.argtypes = [pp, i32, i32]
    lib.noc_problem_defaults.restype = None
    lib.noc_lqr_solve_batch.argtypes = [vp, pp, i64, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp]
    lib.noc_linearize_batch.argtypes = [vp, pp, i64, dp, dp, dp]
    lib.noc_rollout_open_loop.argtypes = [vp, pp, i64, dp, dp, dp]
    lib.noc_lqr_solve_from_x0.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, dp]
    lib.noc_lqr_solve_along.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.noc_dare_solve_batch.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp]
    lib.noc_dare_solve_at.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, dp]
    lib.noc_slq_solve_batch.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.noc_slq_solve_warm.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.noc_argmin_cost.argtypes = [vp, dp, i64, C.POINTER(i64), C.POINTER(C.c_double)]
    lib.noc_argmin_cost_async.argtypes = [vp, dp, i64, i64, dp]
    lib.noc_comm_unique_id.argtypes = [C.c_char_p]
    lib.noc_comm_create_rank.argtypes = [C.POINTER(vp), vp, C.c_char_p, C.c_int, C.c_int]
    lib.noc_comm_create_all.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int]
    lib.noc_comm_destroy.argtypes = [vp]
    lib.noc_comm_destroy.restype = None
    lib.noc_comm_rank.argtypes = [vp]
    lib.noc_comm_size.argtypes = [vp]
    lib.noc_comm_last_error.argtypes = [vp]
    lib.noc_comm_last_error.restype = C.c_char_p
    lib.noc_comm_allgather.argtypes = [vp, dp, dp, i64, i64]
    lib.noc_comm_best_rollout.argtypes = [vp, dp, i64, i64, dp, C.POINTER(CostIndex)]
    lib.noc_comm_lqr_solve_from_x0.argtypes = [vp, pp, i64, dp, dp, dp, dp, dp, dp, C.POINTER(CostIndex)]
    lib.noc_profile_enable.argtypes = [vp, C.c_int]
    lib.noc_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.noc_profile_reset.argtypes = [vp]
    lib.noc_selftest_rcp.argtypes = [vp, dp, dp, dp, i64]
    _lib = lib
    return lib


def problem(model=MODEL_QUAD12, N=100, **kw) -> Problem:
    p = Problem()
    load_library().noc_problem_defaults(C.byref(p), model, N)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _ptr(a, dtype=np.float64, count=None, name="argument"):
    """Raw pointer of a numpy array (host) or torch tensor (host or CUDA); None -> NULL.  The library reads / writes
    `count` elements of `dtype` behind the pointer and cannot see the array, so type and size are checked here."""
    if a is None:
        return None
    want = np.dtype(dtype)
    if isinstance(a, np.ndarray):
        if a.dtype != want:
            raise TypeError(f"{name}: expected {want}, got {a.dtype}")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name}: array must be C-contiguous")
        if count is not None and a.size < count:
            raise ValueError(f"{name}: needs at least {count} elements, got {a.size}")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        tname = str(a.dtype).replace("torch.", "")
        if tname != want.name:
            raise TypeError(f"{name}: expected {want}, got {a.dtype}")
        if not a.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        if count is not None and a.numel() < count:
            raise ValueError(f"{name}: needs at least {count} elements, got {a.numel()}")
        return a.data_ptr()
    raise TypeError(f"{name}: {type(a)}")


def _f(a, count, name):
    return _ptr(a, np.float64, count, name)


def _i(a, count, name):
    return _ptr(a, np.int32, count, name)


class Solver:
    """One noc_ctx bound to one GPU (single caller)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.noc_create(C.byref(h), device)
        if rc != OK:
            raise NocError(rc, "noc_create failed (no usable sm_100 CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.noc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise NocError(rc, self._lib.noc_last_error(self._h).decode())

    # -- plumbing
    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.noc_set_stream(self._h, cuda_stream_handle))

    def synchronize(self):
        self._check(self._lib.noc_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.noc_launch_count(self._h))

    def profile_enable(self, on=True):
        self._check(self._lib.noc_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(self._lib.noc_profile_reset(self._h))

    def profile_read(self, kernel_id):
        """(total device ms, launches) of one kernel id since the last reset."""
        ms, cnt = C.c_double(), C.c_int64()
        self._check(self._lib.noc_profile_read(self._h, kernel_id, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    @staticmethod
    def version():
        return load_library().noc_version().decode()

    def set_option(self, option, value):
        self._check(self._lib.noc_set_option(self._h, option, int(value)))

    # -- hot path (every array is checked for dtype and minimum size before its pointer crosses the C-ABI)
    def lqr_solve_batch(self, prob, batch, AB, QRQf, x0dev=None, K=None, K0=None, P0=None, cost=None,
                        status=None, qr_per_instance=False):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        qr = 2 * n * n + m * m
        self._check(self._lib.noc_lqr_solve_batch(
            self._h, C.byref(prob), batch, _f(AB, B * N * (n * n + n * m), "AB"),
            _f(QRQf, qr * (B if qr_per_instance else 1), "QRQf"), int(qr_per_instance), _f(x0dev, B * n, "x0dev"),
            _f(K, B * N * m * n, "K"), _f(K0, B * m * n, "K0"), _f(P0, B * n * n, "P0"), _f(cost, B, "cost"),
            _i(status, B, "status")))

    def linearize_batch(self, prob, batch, xbar, ubar, AB):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        self._check(self._lib.noc_linearize_batch(self._h, C.byref(prob), batch, _f(xbar, B * (N + 1) * n, "xbar"),
                                                  _f(ubar, B * N * m, "ubar"), _f(AB, B * N * (n * n + n * m), "AB")))

    def rollout_open_loop(self, prob, batch, x0, ubar, xbar):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        self._check(self._lib.noc_rollout_open_loop(self._h, C.byref(prob), batch, _f(x0, B * n, "x0"),
                                                    _f(ubar, B * N * m, "ubar"), _f(xbar, B * (N + 1) * n, "xbar")))

    def lqr_solve_from_x0(self, prob, batch, x0, QRQf, K=None, K0=None, P0=None, cost=None, status=None):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        self._check(self._lib.noc_lqr_solve_from_x0(
            self._h, C.byref(prob), batch, _f(x0, B * n, "x0"), _f(QRQf, 2 * n * n + m * m, "QRQf"),
            _f(K, B * N * m * n, "K"), _f(K0, B * m * n, "K0"), _f(P0, B * n * n, "P0"), _f(cost, B, "cost"),
            _i(status, B, "status")))

    def lqr_solve_along(self, prob, batch, x0, ubar, QRQf, xbar=None, K=None, K0=None, P0=None, cost=None, status=None):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        self._check(self._lib.noc_lqr_solve_along(
            self._h, C.byref(prob), batch, _f(x0, B * n, "x0"), _f(ubar, B * N * m, "ubar"),
            _f(QRQf, 2 * n * n + m * m, "QRQf"), _f(xbar, B * (N + 1) * n, "xbar"), _f(K, B * N * m * n, "K"),
            _f(K0, B * m * n, "K0"), _f(P0, B * n * n, "P0"), _f(cost, B, "cost"), _i(status, B, "status")))

    def dare_solve_batch(self, prob, batch, AB, QR, P, K, iters, status):
        n, m, B = prob.n, prob.m, batch
        self._check(self._lib.noc_dare_solve_batch(
            self._h, C.byref(prob), batch, _f(AB, B * (n * n + n * m), "AB"), _f(QR, n * n + m * m, "QR"),
            _f(P, B * n * n, "P"), _f(K, B * m * n, "K"), _i(iters, B, "iters"), _i(status, B, "status")))

    def dare_solve_at(self, prob, batch, xbar, ubar, QR, P, K, iters, status):
        n, m, B = prob.n, prob.m, batch
        self._check(self._lib.noc_dare_solve_at(
            self._h, C.byref(prob), batch, _f(xbar, B * n, "xbar"), _f(ubar, B * m, "ubar"),
            _f(QR, n * n + m * m, "QR"), _f(P, B * n * n, "P"), _f(K, B * m * n, "K"), _i(iters, B, "iters"),
            _i(status, B, "status")))

    def slq_solve_batch(self, prob, batch, x0, xgoal, QRQf, x=None, u=None, K=None, cost=None, iters=None,
                        alpha_idx=None, status=None, u_init=None):
        n, m, N, B = prob.n, prob.m, prob.N, batch
        self._check(self._lib.noc_slq_solve_warm(
            self._h, C.byref(prob), batch, _f(x0, B * n, "x0"), _f(xgoal, B * n, "xgoal"),
            _f(QRQf, 2 * n * n + m * m, "QRQf"), _f(u_init, B * N * m, "u_init"), _f(x, B * (N + 1) * n, "x"),
            _f(u, B * N * m, "u"), _f(K, B * N * m * n, "K"), _f(cost, B, "cost"), _i(iters, B, "iters"),
            _i(alpha_idx, B * prob.max_iter, "alpha_idx"), _i(status, B, "status")))

    def selftest_rcp(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        fast, ieee = np.empty_like(x), np.empty_like(x)
        self._check(self._lib.noc_selftest_rcp(self._h, _ptr(x), _ptr(fast), _ptr(ieee), x.size))
        return fast, ieee

    def argmin_cost_async(self, cost_dev, batch, best_dev, index_offset=0):
        """Device-resident pick: best_dev is a 16-byte device buffer (e.g. a 2-element float64 CUDA tensor viewed as
        {double cost; int64 index}); no host synchronisation."""
        self._check(self._lib.noc_argmin_cost_async(self._h, _f(cost_dev, batch, "cost"), batch, index_offset,
                                                    best_dev.data_ptr()))

    def argmin_cost(self, cost, batch):
        idx, val = C.c_int64(), C.c_double()
        self._check(self._lib.noc_argmin_cost(self._h, _f(cost, batch, "cost"), batch, C.byref(idx), C.byref(val)))
        return idx.value, val.value


def plan_time_parallel(N: int):
    """(chunks, len) of the time-parallel sweep's chunk plan for horizon N (host-side, no GPU needed)."""
    c, l = C.c_int32(), C.c_int32()
    rc = load_library().noc_plan_time_parallel(N, C.byref(c), C.byref(l))
    if rc != OK:
        raise NocError(rc, "noc_plan_time_parallel: N >= 1 required")
    return c.value, l.value


def comm_unique_id() -> bytes:
    """128-byte NCCL id, created on rank 0 and distributed by the launcher (torch.distributed, MPI, a file)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = load_library().noc_comm_unique_id(buf)
    if rc != OK:
        raise NocError(rc, "noc_comm_unique_id failed (NCCL not loadable?)")
    return buf.raw


class Comm:
    """One NCCL rank bound to a Solver (noc_comm): the in-library multi-GPU plane — device-resident gathers of costs
    and gains over NVLink and the cross-GPU "pick the best rollout" (SURVEY.md §8e)."""

    def __init__(self, solver: Solver, unique_id: bytes, rank: int, nranks: int):
        self._lib = load_library()
        self.solver = solver
        h = C.c_void_p()
        rc = self._lib.noc_comm_create_rank(C.byref(h), solver._h, unique_id, rank, nranks)
        solver._check(rc)
        self._h = h
        self.rank, self.nranks = rank, nranks

    @classmethod
    def from_torch_distributed(cls, solver: Solver, dist, device):
        """Rank 0 creates the id, torch.distributed broadcasts its 128 bytes (that is all torch is used for here)."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(solver, bytes(t.cpu().numpy().tobytes()), rank, world)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.noc_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allgather(self, src_dev, dst_dev, count, elem_bytes=8):
        self.solver._check(self._lib.noc_comm_allgather(self._h, src_dev.data_ptr(), dst_dev.data_ptr(), count, elem_bytes))

    def best_rollout(self, cost_dev, count, index_offset, best_dev, want_host=False):
        host = CostIndex() if want_host else None
        self.solver._check(self._lib.noc_comm_best_rollout(self._h, _f(cost_dev, count, "cost"), count, index_offset,
                                                           best_dev.data_ptr(), C.byref(host) if want_host else None))
        return (host.index, host.cost) if want_host else None

    def lqr_solve_from_x0(self, prob, count, x0_dev, qr_dev, best_dev, cost_all=None, K0_all=None, status=None,
                          want_host=False):
        n, m = prob.n, prob.m
        host = CostIndex() if want_host else None
        self.solver._check(self._lib.noc_comm_lqr_solve_from_x0(
            self._h, C.byref(prob), count, _f(x0_dev, count * n, "x0"), _f(qr_dev, 2 * n * n + m * m, "QRQf"),
            _f(cost_all, self.nranks * count, "cost_all"), _f(K0_all, self.nranks * count * m * n, "K0_all"),
            _i(status, count, "status"), best_dev.data_ptr(), C.byref(host) if want_host else None))
        return (host.index, host.cost) if want_host else None

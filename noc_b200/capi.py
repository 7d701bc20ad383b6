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
OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM = 0, -1, -2, -3, -4, -5
STATUS_OK, STATUS_NOT_PD, STATUS_MAX_ITER, STATUS_LS_FAIL, STATUS_REG_FAIL = 0, 1, 2, 3, 4
MODEL_QUAD12, MODEL_DUAL24, MODEL_LTV_USER = 0, 1, 2
LAYOUT_A_THEN_B, LAYOUT_ROWS_AB = 0, 1
FLAG_NONE, FLAG_ASYNC = 0, 1

EXPORTS = [
    "noc_create", "noc_destroy", "noc_last_error", "noc_version", "noc_set_stream", "noc_synchronize",
    "noc_launch_count", "noc_problem_defaults", "noc_lqr_solve_batch", "noc_linearize_batch",
    "noc_rollout_open_loop", "noc_lqr_solve_from_x0", "noc_lqr_solve_along", "noc_dare_solve_batch", "noc_dare_solve_at", "noc_slq_solve_batch", "noc_slq_solve_warm",
    "noc_argmin_cost", "noc_profile_enable", "noc_profile_read", "noc_profile_reset", "noc_selftest_rcp",
]
KERNEL_SETUP, KERNEL_ROLLOUT, KERNEL_LINEARIZE, KERNEL_RICCATI, KERNEL_FORWARD, KERNEL_ARGMIN = range(6)


class Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32), ("model", C.c_int32),
                ("layout", C.c_int32), ("max_iter", C.c_int32), ("n_alpha", C.c_int32), ("flags", C.c_int32),
                ("dt", C.c_double), ("tol", C.c_double), ("reg0", C.c_double), ("armijo", C.c_double),
                ("u_min", C.c_double), ("u_max", C.c_double)]


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


def _ptr(a, dtype=None):
    """Raw pointer of a numpy array (host) or torch tensor (host or CUDA); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {a.dtype}")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(type(a))


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

    # -- hot path
    def lqr_solve_batch(self, prob, batch, AB, QRQf, x0dev=None, K=None, K0=None, P0=None, cost=None,
                        status=None, qr_per_instance=False):
        self._check(self._lib.noc_lqr_solve_batch(self._h, C.byref(prob), batch, _ptr(AB), _ptr(QRQf),
                                                  int(qr_per_instance), _ptr(x0dev), _ptr(K), _ptr(K0), _ptr(P0),
                                                  _ptr(cost), _ptr(status)))

    def linearize_batch(self, prob, batch, xbar, ubar, AB):
        self._check(self._lib.noc_linearize_batch(self._h, C.byref(prob), batch, _ptr(xbar), _ptr(ubar), _ptr(AB)))

    def rollout_open_loop(self, prob, batch, x0, ubar, xbar):
        self._check(self._lib.noc_rollout_open_loop(self._h, C.byref(prob), batch, _ptr(x0), _ptr(ubar), _ptr(xbar)))

    def lqr_solve_from_x0(self, prob, batch, x0, QRQf, K=None, K0=None, P0=None, cost=None, status=None):
        self._check(self._lib.noc_lqr_solve_from_x0(self._h, C.byref(prob), batch, _ptr(x0), _ptr(QRQf), _ptr(K),
                                                    _ptr(K0), _ptr(P0), _ptr(cost), _ptr(status)))

    def lqr_solve_along(self, prob, batch, x0, ubar, QRQf, xbar=None, K=None, K0=None, P0=None, cost=None, status=None):
        self._check(self._lib.noc_lqr_solve_along(self._h, C.byref(prob), batch, _ptr(x0), _ptr(ubar), _ptr(QRQf),
                                                  _ptr(xbar), _ptr(K), _ptr(K0), _ptr(P0), _ptr(cost), _ptr(status)))

    def dare_solve_batch(self, prob, batch, AB, QR, P, K, iters, status):
        self._check(self._lib.noc_dare_solve_batch(self._h, C.byref(prob), batch, _ptr(AB), _ptr(QR), _ptr(P),
                                                   _ptr(K), _ptr(iters), _ptr(status)))

    def dare_solve_at(self, prob, batch, xbar, ubar, QR, P, K, iters, status):
        self._check(self._lib.noc_dare_solve_at(self._h, C.byref(prob), batch, _ptr(xbar), _ptr(ubar), _ptr(QR),
                                                _ptr(P), _ptr(K), _ptr(iters), _ptr(status)))

    def slq_solve_batch(self, prob, batch, x0, xgoal, QRQf, x=None, u=None, K=None, cost=None, iters=None,
                        alpha_idx=None, status=None, u_init=None):
        self._check(self._lib.noc_slq_solve_warm(self._h, C.byref(prob), batch, _ptr(x0), _ptr(xgoal), _ptr(QRQf),
                                                 _ptr(u_init), _ptr(x), _ptr(u), _ptr(K), _ptr(cost), _ptr(iters),
                                                 _ptr(alpha_idx), _ptr(status)))

    def selftest_rcp(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        fast, ieee = np.empty_like(x), np.empty_like(x)
        self._check(self._lib.noc_selftest_rcp(self._h, _ptr(x), _ptr(fast), _ptr(ieee), x.size))
        return fast, ieee

    def argmin_cost(self, cost, batch):
        idx, val = C.c_int64(), C.c_double()
        self._check(self._lib.noc_argmin_cost(self._h, _ptr(cost), batch, C.byref(idx), C.byref(val)))
        return idx.value, val.value

"""ctypes binding of the CPU oracle (oracle/libnoc_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package (noc_b200/) never does.  PARITY UNPINNED against KahnShi/NOC (the reference
snapshot has no solver source); the oracle restates DESIGN.md §2 and is pinned against scipy / numpy
longdouble / closed forms in tests/test_oracle_*.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnoc_oracle.so")

MODEL_QUAD12, MODEL_DUAL24 = 0, 1
LAYOUT_A_THEN_B, LAYOUT_ROWS_AB = 0, 1
OK, NOT_PD, MAX_ITER, LS_FAIL, REG_FAIL = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only). Returns the .so path."""
    src = os.path.join(_HERE, "noc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "noc_oracle.h"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class SlqOpts(C.Structure):
    _fields_ = [("model", C.c_int), ("N", C.c_int), ("max_iter", C.c_int), ("n_alpha", C.c_int),
                ("dt", C.c_double), ("tol", C.c_double), ("reg0", C.c_double), ("armijo", C.c_double),
                ("u_min", C.c_double), ("u_max", C.c_double), ("reg_schedule", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.noc_oracle_forward.restype = C.c_double
        _lib.noc_oracle_traj_cost.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def dims(model):
    n, m = C.c_int(), C.c_int()
    assert lib().noc_oracle_model_dims(model, C.byref(n), C.byref(m)) == 0
    return n.value, m.value


def max_threads():
    return lib().noc_oracle_max_threads()


def sincos(x):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    s = np.empty_like(x)
    c = np.empty_like(x)
    sv, cv = C.c_double(), C.c_double()
    f = lib().noc_oracle_sincos
    for i, xi in enumerate(x):
        f(C.c_double(float(xi)), C.byref(sv), C.byref(cv))
        s[i], c[i] = sv.value, cv.value
    return s, c


def hover_input(model):
    n, m = dims(model)
    u = np.empty(m)
    lib().noc_oracle_hover_input(model, _p(u))
    return u


def step(model, x, u, dt):
    x, u = _c(x), _c(u)
    out = np.empty_like(x)
    lib().noc_oracle_step(model, _p(x), _p(u), C.c_double(dt), _p(out))
    return out


def linearize(model, x, u, dt):
    n, m = dims(model)
    x, u = _c(x), _c(u)
    A, B = np.empty((n, n)), np.empty((n, m))
    lib().noc_oracle_linearize(model, _p(x), _p(u), C.c_double(dt), _p(A), _p(B))
    return A, B


def rollout_open_loop(model, N, dt, x0, ubar):
    n, m = dims(model)
    x0, ubar = _c(x0), _c(ubar)
    xbar = np.empty((N + 1, n))
    lib().noc_oracle_rollout_open_loop(model, N, C.c_double(dt), _p(x0), _p(ubar), _p(xbar))
    return xbar


def linearize_traj(model, N, dt, xbar, ubar, layout=LAYOUT_A_THEN_B):
    n, m = dims(model)
    xbar, ubar = _c(xbar), _c(ubar)
    AB = np.empty((N, n * n + n * m))
    lib().noc_oracle_linearize_traj(model, N, C.c_double(dt), _p(xbar), _p(ubar), layout, _p(AB))
    return AB


def pack_qrqf(Q, R, Qf):
    return np.concatenate([np.asarray(Q, float).ravel(), np.asarray(R, float).ravel(), np.asarray(Qf, float).ravel()])


def lqr_sweep(n, m, N, AB, Q, R, Qf, layout=LAYOUT_A_THEN_B, x0dev=None):
    AB, Q, R, Qf, x0dev = _c(AB), _c(Q), _c(R), _c(Qf), _c(x0dev)
    K = np.empty((N, m, n))
    P0 = np.empty((n, n))
    cost = C.c_double(float("nan"))
    st = lib().noc_oracle_lqr_sweep(n, m, N, _p(AB), layout, _p(Q), _p(R), _p(Qf), _p(K), _p(P0),
                                    _p(x0dev), C.byref(cost))
    return K, P0, cost.value, st


def lqr_batch(n, m, N, AB, QRQf, layout=LAYOUT_A_THEN_B, x0dev=None, want_K=True, threads=0,
              qr_per_instance=False):
    AB, QRQf, x0dev = _c(AB), _c(QRQf), _c(x0dev)
    B = AB.shape[0]
    K = np.empty((B, N, m, n)) if want_K else None
    K0 = np.empty((B, m, n))
    P0 = np.empty((B, n, n))
    cost = np.full(B, np.nan)
    status = np.zeros(B, dtype=np.int32)
    lib().noc_oracle_lqr_batch(n, m, N, C.c_int64(B), _p(AB), layout, _p(QRQf), int(qr_per_instance),
                               _p(x0dev), _p(K), _p(K0), _p(P0), _p(cost), _pi(status), threads)
    return dict(K=K, K0=K0, P0=P0, cost=cost, status=status)


def lqr_from_x0_batch(model, N, dt, x0, QRQf, threads=0, want_P0=True):
    n, m = dims(model)
    x0, QRQf = _c(x0), _c(QRQf)
    B = x0.shape[0]
    K0 = np.empty((B, m, n))
    P0 = np.empty((B, n, n)) if want_P0 else None
    cost = np.full(B, np.nan)
    status = np.zeros(B, dtype=np.int32)
    lib().noc_oracle_lqr_from_x0_batch(model, N, C.c_double(dt), C.c_int64(B), _p(x0), _p(QRQf), _p(K0),
                                       _p(P0), _p(cost), _pi(status), threads)
    return dict(K0=K0, P0=P0, cost=cost, status=status)


def make_ab_batch(model, N, dt, x0, layout=LAYOUT_A_THEN_B, threads=0):
    n, m = dims(model)
    x0 = _c(x0)
    B = x0.shape[0]
    AB = np.empty((B, N, n * n + n * m))
    lib().noc_oracle_make_ab_batch(model, N, C.c_double(dt), C.c_int64(B), _p(x0), layout, _p(AB), threads)
    return AB


def dare(A, B, Q, R, tol=1e-12, max_iter=10000):
    A, B, Q, R = _c(A), _c(B), _c(Q), _c(R)
    n, m = B.shape
    P = np.empty((n, n))
    K = np.empty((m, n))
    it = lib().noc_oracle_dare(n, m, _p(A), _p(B), _p(Q), _p(R), C.c_double(tol), max_iter, _p(P), _p(K))
    return P, K, it


def slq_opts(model, N, dt=0.02, tol=1e-8, max_iter=50, n_alpha=11, reg0=0.0, armijo=1e-4, u_min=-np.inf, u_max=np.inf,
             reg_schedule=0):
    return SlqOpts(model, N, max_iter, n_alpha, dt, tol, reg0, armijo, u_min, u_max, reg_schedule)


def slq(opts, x0, xgoal, QRQf, want_K=True):
    n, m = dims(opts.model)
    N = opts.N
    x0, xgoal, QRQf = _c(x0), _c(xgoal), _c(QRQf)
    x = np.empty((N + 1, n))
    u = np.empty((N, m))
    K = np.empty((N, m, n)) if want_K else None
    kff = np.empty((N, m))
    cost, reg = C.c_double(), C.c_double()
    iters = C.c_int32()
    aidx = np.full(opts.max_iter, -1, dtype=np.int32)
    st = lib().noc_oracle_slq(C.byref(opts), _p(x0), _p(xgoal), _p(QRQf), _p(x), _p(u), _p(K), _p(kff),
                              C.byref(cost), C.byref(iters), _pi(aidx), C.byref(reg))
    return dict(x=x, u=u, K=K, kff=kff, cost=cost.value, iters=iters.value, alpha_idx=aidx, status=st,
                reg=reg.value)


def slq_batch(opts, x0, xgoal, QRQf, want_traj=True, want_K=False, threads=0, u_init=None):
    n, m = dims(opts.model)
    N = opts.N
    x0, xgoal, QRQf = _c(x0), _c(xgoal), _c(QRQf)
    B = x0.shape[0]
    x = np.empty((B, N + 1, n)) if want_traj else None
    u = np.empty((B, N, m)) if want_traj else None
    K = np.empty((B, N, m, n)) if want_K else None
    cost = np.empty(B)
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    aidx = np.full((B, opts.max_iter), -1, dtype=np.int32)
    u_init = _c(u_init)
    lib().noc_oracle_slq_batch_warm(C.byref(opts), C.c_int64(B), _p(x0), _p(xgoal), _p(QRQf), _p(u_init), _p(x), _p(u),
                                    _p(K), _p(cost), _pi(iters), _pi(aidx), _pi(status), threads)
    return dict(x=x, u=u, K=K, cost=cost, iters=iters, alpha_idx=aidx, status=status)


def backward_affine(model, N, dt, xbar, ubar, xgoal, QRQf, mu=0.0):
    n, m = dims(model)
    xbar, ubar, xgoal, QRQf = _c(xbar), _c(ubar), _c(xgoal), _c(QRQf)
    K = np.empty((N, m, n))
    kff = np.empty((N, m))
    dV = np.empty(2)
    st = lib().noc_oracle_backward_affine(model, N, C.c_double(dt), _p(xbar), _p(ubar), _p(xgoal), _p(QRQf),
                                          C.c_double(mu), _p(K), _p(kff), _p(dV))
    return K, kff, dV, st


def forward(model, N, dt, xbar, ubar, xgoal, QRQf, K, kff, alpha):
    n, m = dims(model)
    xbar, ubar, xgoal, QRQf, K, kff = _c(xbar), _c(ubar), _c(xgoal), _c(QRQf), _c(K), _c(kff)
    xn = np.empty((N + 1, n))
    un = np.empty((N, m))
    J = lib().noc_oracle_forward(model, N, C.c_double(dt), _p(xbar), _p(ubar), _p(xgoal), _p(QRQf), _p(K),
                                 _p(kff), C.c_double(alpha), _p(xn), _p(un))
    return J, xn, un


def traj_cost(model, N, x, u, xgoal, QRQf):
    x, u, xgoal, QRQf = _c(x), _c(u), _c(xgoal), _c(QRQf)
    return lib().noc_oracle_traj_cost(model, N, _p(x), _p(u), _p(xgoal), _p(QRQf))


def box_qp(H, g, lo, hi, x0=None):
    """min 0.5 x'Hx + g'x, lo <= x <= hi by the oracle's projected Newton; returns x, clamped"""
    H, g, lo, hi = _c(H), _c(g), _c(lo), _c(hi)
    m = g.shape[0]
    x = np.zeros(m) if x0 is None else np.array(x0, dtype=np.float64)
    cl = np.zeros(m, dtype=np.int32)
    lib().noc_oracle_box_qp(m, _p(H), _p(g), _p(lo), _p(hi), _p(x), _pi(cl))
    return x, cl.astype(bool)

"""The time-parallel Riccati sweep (NOC_FLAG_TIME_PARALLEL; noc_b200/csrc/riccati_scan.cuh) against the
CPU oracle's sequential sweep.  The scan associates the arithmetic differently (interval elements, a pivoted 12x12 solve
per chunk boundary), so the bar here is north_star's TOLERANCE, written below: 1e-9 relative on gains, cost-to-go and
cost — not bitwise.  (Parity statement as everywhere: kernel vs in-repo oracle, oracle vs scipy; parity unpinned.)"""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def solver():
    from noc_b200 import capi
    s = capi.Solver(0)
    yield s
    s.close()


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _traj_AB(oracle, x0, N):
    ubar = np.tile(pb.hover_input(), (N, 1))
    return np.stack([oracle.linearize_traj(0, N, 0.02, oracle.rollout_open_loop(0, N, 0.02, x, ubar), ubar, 0) for x in x0])


@pytest.mark.parametrize("dfma,chunks", [(0, 0), (1, 0), (0, 1), (0, 5)])
@pytest.mark.parametrize("batch,N", [(1, 1), (2, 2), (3, 7), (5, 23), (64, 100), (9, 200), (2, 517)])
def test_time_parallel_sweep_matches_the_sequential_oracle(oracle, solver, batch, N, dfma, chunks):
    """dfma = 1: the same kernel with DFMA chains instead of FP64 tensor-core tiles; chunks: forced plan (1 = the plain
    warp-per-instance recursion, no scan)."""
    from noc_b200 import capi
    solver.set_option(capi.OPT_TIME_PARALLEL_DFMA, dfma)
    solver.set_option(capi.OPT_TIME_PARALLEL_CHUNKS, chunks)
    qr = pb.pack_qrqf(*pb.default_weights())
    x0 = pb.sample_x0(batch, seed=4000 + N)
    AB = _traj_AB(oracle, x0, N)
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12))
    cost = np.empty(batch); status = np.full(batch, -1, dtype=np.int32)
    prob = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_TIME_PARALLEL)
    solver.lqr_solve_batch(prob, batch, AB, qr, x0dev=x0, K=K, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0) and np.array_equal(status, ref["status"])
    # per time step, so that a small gain late in the horizon is held to the same relative bar as a large early one
    for k in range(N):
        assert _rel(K[:, k], ref["K"][:, k]) < RTOL, k
    assert _rel(K0, ref["K0"]) < RTOL and _rel(P0, ref["P0"]) < RTOL
    assert np.max(np.abs(cost - ref["cost"]) / np.abs(ref["cost"])) < RTOL
    assert np.array_equal(P0, P0.transpose(0, 2, 1))
    assert np.array_equal(K[:, 0], K0)
    solver.set_option(capi.OPT_TIME_PARALLEL_DFMA, 0)
    solver.set_option(capi.OPT_TIME_PARALLEL_CHUNKS, 0)


def test_time_parallel_dense_weights_and_random_ltv(oracle, solver):
    """Dense SPD weights and unstructured A_k, B_k (not a quadrotor), unstable open loop: the interval elements grow."""
    from noc_b200 import capi
    rng = np.random.default_rng(5)
    batch, N = 6, 60
    def spd(k, s):
        M = rng.normal(size=(k, k)); return s * (M @ M.T + k * np.eye(k))
    Q, R, Qf = spd(12, 0.1), spd(4, 0.05), spd(12, 1.0)
    qr = pb.pack_qrqf(Q, R, Qf)
    A = np.eye(12) + 0.08 * rng.normal(size=(batch, N, 12, 12))
    B = 0.1 * rng.normal(size=(batch, N, 12, 4))
    AB = np.concatenate([A.reshape(batch, N, 144), B.reshape(batch, N, 48)], axis=2)
    x0 = rng.normal(size=(batch, 12))
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_TIME_PARALLEL), batch, AB, qr, x0dev=x0, K=K, P0=P0,
                           cost=cost, status=status)
    assert np.all(status == 0)
    for k in range(N):
        assert _rel(K[:, k], ref["K"][:, k]) < RTOL, k
    assert _rel(P0, ref["P0"]) < RTOL and np.max(np.abs(cost - ref["cost"]) / np.abs(ref["cost"])) < RTOL


def test_time_parallel_singular_state_weight(oracle, solver):
    """Q and Qf only positive SEMI-definite (velocities and rates unweighted): the boundary solve is a pivoted
    Gauss-Jordan on I + S C, which needs no definiteness of S."""
    from noc_b200 import capi
    batch, N = 4, 80
    Q, R, Qf = pb.default_weights()
    Q = Q.copy(); Qf = Qf.copy()
    for i in (3, 4, 5, 9, 10, 11):
        Q[i, i] = 0.0; Qf[i, i] = 0.0
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=77)
    AB = _traj_AB(oracle, x0, N)
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_TIME_PARALLEL), batch, AB, qr, K=K, P0=P0, status=status)
    assert np.all(status == 0)
    for k in range(N):
        assert _rel(K[:, k], ref["K"][:, k]) < RTOL, k
    assert _rel(P0, ref["P0"]) < RTOL


@pytest.mark.parametrize("batch,N", [(1, 100), (33, 50)])
def test_time_parallel_from_x0_and_along(oracle, solver, batch, N):
    """x0 -> nominal rollout -> time-parallel linearisation -> time-parallel sweep (the small-batch MPC call)."""
    from noc_b200 import capi
    qr = pb.pack_qrqf(*pb.default_weights())
    x0 = pb.sample_x0(batch, seed=11)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr)
    K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch); status = np.full(batch, -1, dtype=np.int32)
    prob = capi.problem(capi.MODEL_QUAD12, N, flags=capi.FLAG_TIME_PARALLEL)
    solver.lqr_solve_from_x0(prob, batch, x0, qr, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0)
    assert _rel(K0, ref["K0"]) < RTOL and _rel(P0, ref["P0"]) < RTOL
    assert np.max(np.abs(cost - ref["cost"]) / np.abs(ref["cost"])) < RTOL
    # the same call without the flag is the bitwise path
    K0s = np.empty_like(K0)
    solver.lqr_solve_from_x0(capi.problem(capi.MODEL_QUAD12, N), batch, x0, qr, K0=K0s)
    assert np.array_equal(K0s, ref["K0"])
    # along a given input sequence, nominal returned
    rng = np.random.default_rng(3)
    ubar = pb.hover_input() + 0.3 * rng.normal(size=(batch, N, 4))
    xbar = np.empty((batch, N + 1, 12)); K = np.empty((batch, N, 4, 12))
    solver.lqr_solve_along(prob, batch, x0, ubar, qr, xbar=xbar, K=K, status=status)
    Ks = np.empty_like(K); xs = np.empty_like(xbar)
    solver.lqr_solve_along(capi.problem(capi.MODEL_QUAD12, N), batch, x0, ubar, qr, xbar=xs, K=Ks)
    assert np.array_equal(xbar, xs) and np.all(status == 0)
    for k in range(N):
        assert _rel(K[:, k], Ks[:, k]) < RTOL, k


def test_time_parallel_not_pd_and_unsupported(oracle, solver):
    from noc_b200 import capi
    batch, N = 3, 40
    Q, R, Qf = pb.default_weights()
    Rb = R.copy(); Rb[2, 2] = -50.0
    x0 = pb.sample_x0(batch, seed=8)
    AB = _traj_AB(oracle, x0, N)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch); status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_TIME_PARALLEL), batch, AB, pb.pack_qrqf(Q, Rb, Qf),
                           x0dev=x0, K=K, P0=P0, cost=cost, status=status)
    assert np.all(status == 1) and np.all(np.isnan(K)) and np.all(np.isnan(P0)) and np.all(np.isnan(cost))
    with pytest.raises(capi.NocError):      # ROWS_AB records
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, layout=1, flags=capi.FLAG_TIME_PARALLEL), batch, AB,
                               pb.pack_qrqf(Q, R, Qf), K=K)
    with pytest.raises(capi.NocError):      # the dual model
        solver.lqr_solve_from_x0(capi.problem(capi.MODEL_DUAL24, N, flags=capi.FLAG_TIME_PARALLEL), batch,
                                 pb.sample_x0(batch, model=1), pb.pack_qrqf(*pb.default_weights(1)), K0=np.empty((batch, 8, 24)))

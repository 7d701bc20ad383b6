"""n=24 / m=8 (spring-coupled dual UAV, BASELINE.json configs[4]): CUDA vs the CPU oracle through the C-ABI,
bit-exact, for the LTV sweep, the from-x0 pipeline and SLQ (iteration counts and step indices identical)."""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu
M = pb.MODEL_DUAL24


@pytest.fixture(scope="module", params=[0, 1], ids=["tensor-core-sweep", "dfma-sweep"])
def solver(request):
    """Every test of this module runs twice: on k_ric24_mma (one instance per warp, products on the FP64 tensor cores, step
    records by k_prep24 in the SLQ backward pass — the default for A_THEN_B records and the built-in model) and, with
    NOC_OPT_NO_TENSOR_SWEEP, on k_ric24 (DFMA chains, two instances per warp).  Both must reproduce the oracle bit for bit."""
    from noc_b200 import capi
    s = capi.Solver(0)
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, request.param)
    yield s
    s.close()


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("batch,N", [(1, 1), (3, 7), (34, 40)])
def test_lqr24_sweep_bitexact(oracle, solver, layout, batch, N):
    from noc_b200 import capi
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=300 + batch, model=M)
    AB = oracle.make_ab_batch(M, N, 0.02, x0, layout=layout)
    ref = oracle.lqr_batch(24, 8, N, AB, qr, layout=layout, x0dev=x0)
    K = np.empty((batch, N, 8, 24)); K0 = np.empty((batch, 8, 24)); P0 = np.empty((batch, 24, 24))
    cost = np.empty(batch); status = np.full(batch, -1, dtype=np.int32)
    prob = capi.problem(capi.MODEL_LTV_USER, N, layout=layout, n=24, m=8)
    solver.lqr_solve_batch(prob, batch, AB, qr, x0dev=x0, K=K, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0) and np.array_equal(status, ref["status"])
    assert np.array_equal(K, ref["K"]) and np.array_equal(K0, ref["K0"])
    assert np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])
    assert np.array_equal(P0, P0.transpose(0, 2, 1))


def test_lqr24_dense_weights_random_ltv_and_not_pd(oracle, solver):
    from noc_b200 import capi
    rng = np.random.default_rng(15)
    batch, N = 7, 12

    def spd(k, s):
        Mx = rng.normal(size=(k, k)); return s * (Mx @ Mx.T + k * np.eye(k))
    Q, R, Qf = spd(24, 0.1), spd(8, 0.05), spd(24, 1.0)
    A = np.eye(24) + 0.03 * rng.normal(size=(batch, N, 24, 24))
    B = 0.1 * rng.normal(size=(batch, N, 24, 8))
    AB = np.concatenate([A.reshape(batch, N, 576), B.reshape(batch, N, 192)], axis=2)
    x0 = rng.normal(size=(batch, 24))
    for Rm in (R, R - 3.0 * np.eye(8)):       # second: indefinite R -> NOT_PD for (some) instances
        qr = pb.pack_qrqf(Q, Rm, Qf)
        ref = oracle.lqr_batch(24, 8, N, AB, qr, x0dev=x0)
        K = np.empty((batch, N, 8, 24)); P0 = np.empty((batch, 24, 24)); cost = np.empty(batch)
        status = np.full(batch, -1, dtype=np.int32)
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, n=24, m=8), batch, AB, qr, x0dev=x0, K=K, P0=P0,
                               cost=cost, status=status)
        assert np.array_equal(status, ref["status"])
        assert np.array_equal(np.isnan(K), np.isnan(ref["K"]))
        ok = ~np.isnan(ref["K"])
        assert np.array_equal(K[ok], ref["K"][ok])
        okp = ~np.isnan(ref["P0"])
        assert np.array_equal(np.isnan(P0), ~okp) and np.array_equal(P0[okp], ref["P0"][okp])


def test_from_x0_dual24_wild_states_match_dense_oracle(oracle, solver):
    """Both n=24 kernels skip structural zeros of the dual model's Jacobian — k_ric24 entry by entry, k_ric24_mma the four
    all-zero 4x8 tiles of [B | A] — while the oracle multiplies them.  Same bits as long as the intermediates are finite,
    also far from hover (every quadrant of the sincos reduction, pitch near +-pi/2, body rates of tens of rad/s); states
    that make the sweep fail must fail identically (status NOT_PD, NaN outputs)."""
    from noc_b200 import capi
    batch, N = 48, 40
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    rng = np.random.default_rng(78)
    x0 = pb.sample_x0(batch, seed=44, model=M)
    for v in (0, 12):
        x0[:, v + 6:v + 9] = rng.uniform(-3.1, 3.1, size=(batch, 3))
        x0[:, v + 9:v + 12] = rng.uniform(-40.0, 40.0, size=(batch, 3))
    x0[:16, 7] = np.pi / 2 + rng.uniform(-2e-3, 2e-3, size=16)
    x0[16:20, 21:24] *= 1e3
    ref = oracle.lqr_from_x0_batch(M, N, 0.02, x0, qr)
    K = np.empty((batch, N, 8, 24)); K0 = np.empty((batch, 8, 24)); P0 = np.empty((batch, 24, 24)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(M, N), batch, x0, qr, K=K, K0=K0, P0=P0, cost=cost, status=status)
    assert np.array_equal(status, ref["status"])
    assert (status == 0).sum() >= batch // 2
    assert np.array_equal(K0, ref["K0"], equal_nan=True) and np.array_equal(P0, ref["P0"], equal_nan=True)
    assert np.array_equal(cost, ref["cost"], equal_nan=True)
    refK = oracle.lqr_batch(24, 8, N, oracle.make_ab_batch(M, N, 0.02, x0), qr, x0dev=x0)["K"]
    assert np.array_equal(K, refK, equal_nan=True)


def test_from_x0_dual24_bitexact(oracle, solver):
    from noc_b200 import capi
    batch, N = 19, 50
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=41, model=M)
    ref = oracle.lqr_from_x0_batch(M, N, 0.02, x0, qr)
    K0 = np.empty((batch, 8, 24)); P0 = np.empty((batch, 24, 24)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(M, N), batch, x0, qr, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])


@pytest.mark.parametrize("kw", [dict(), dict(max_iter=4), dict(u_min=0.5, u_max=4.0), dict(u_min=1.5, u_max=3.2, max_iter=20),
                                dict(reg_schedule=1, reg0=0.5, max_iter=25), dict(reg_schedule=1, u_min=1.5, u_max=3.2, max_iter=20)])
def test_slq_dual24_bitexact(oracle, solver, kw):
    from noc_b200 import capi
    batch, N = 9, 30
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch, model=M)
    xg = pb.sample_goals(batch, seed=77, model=M)
    prob = capi.problem(M, N, **kw)
    x = np.empty((batch, N + 1, 24)); u = np.empty((batch, N, 8)); K = np.empty((batch, N, 8, 24))
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    aidx = np.full((batch, prob.max_iter), -7, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, x=x, u=u, K=K, cost=cost, iters=iters, alpha_idx=aidx, status=status)
    opts = oracle.slq_opts(M, N, max_iter=prob.max_iter, u_min=prob.u_min, u_max=prob.u_max, reg0=prob.reg0,
                           reg_schedule=prob.reg_schedule)
    ref = oracle.slq_batch(opts, x0, xg, qr, want_traj=True, want_K=True)
    assert np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(aidx, ref["alpha_idx"]) and np.array_equal(cost, ref["cost"])
    assert np.array_equal(x, ref["x"]) and np.array_equal(u, ref["u"]) and np.array_equal(K, ref["K"])
    if "u_min" in kw:     # control-limited (DESIGN.md 2.4b, 8x8 box QP): inside the box, at a bound somewhere, gain rows off
        assert u.min() >= kw["u_min"] and u.max() <= kw["u_max"]
        assert np.any(u == kw["u_max"]) or np.any(u == kw["u_min"])
        assert np.any(np.all(K == 0.0, axis=-1))


@pytest.mark.parametrize("batch,N,box", [(1, 1, False), (2, 2, True), (3, 3, False), (5, 9, True)])
def test_slq_dual24_tiny_horizons_and_batches(oracle, solver, batch, N, box):
    """Edge shapes through k_prep24 + k_ric24_mma<RIC_AFFINE> (step records, double-buffered bulk copies, the first / last
    step of a sweep) and their DFMA partner: horizons of one to three steps, single instances, with and without the box."""
    from noc_b200 import capi
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=80 + N, model=M)
    xg = pb.sample_goals(batch, seed=81 + N, model=M)
    kw = dict(u_min=1.0, u_max=3.5) if box else {}
    prob = capi.problem(M, N, max_iter=6, **kw)
    x = np.empty((batch, N + 1, 24)); u = np.empty((batch, N, 8)); K = np.empty((batch, N, 8, 24))
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    aidx = np.full((batch, prob.max_iter), -7, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, x=x, u=u, K=K, cost=cost, iters=iters, alpha_idx=aidx, status=status)
    opts = oracle.slq_opts(M, N, max_iter=prob.max_iter, u_min=prob.u_min, u_max=prob.u_max, reg0=prob.reg0,
                           reg_schedule=prob.reg_schedule)
    ref = oracle.slq_batch(opts, x0, xg, qr, want_traj=True, want_K=True)
    assert np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(aidx, ref["alpha_idx"]) and np.array_equal(cost, ref["cost"])
    assert np.array_equal(x, ref["x"]) and np.array_equal(u, ref["u"]) and np.array_equal(K, ref["K"])


def test_lqr24_multiwarp_config_bitexact(oracle, solver):
    """Batches above 2 x 2 x SM-count use the 8-warp CTA configuration of k_ric24 (smaller ones one warp per CTA)."""
    from noc_b200 import capi
    batch, N = 1500, 12
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=91, model=M)
    ref = oracle.lqr_from_x0_batch(M, N, 0.02, x0, qr)
    K0 = np.empty((batch, 8, 24)); P0 = np.empty((batch, 24, 24)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(M, N), batch, x0, qr, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])
    AB = oracle.make_ab_batch(M, N, 0.02, x0)
    ref2 = oracle.lqr_batch(24, 8, N, AB, qr, x0dev=x0, want_K=True)
    K = np.empty((batch, N, 8, 24))
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, n=24, m=8), batch, AB, qr, K=K)
    assert np.array_equal(K, ref2["K"])


@pytest.mark.parametrize("box", [dict(), dict(u_min=1.0, u_max=3.5)])
def test_slq_dual24_multiwarp_config(oracle, solver, box):
    from noc_b200 import capi
    batch, N = 1300, 10
    Q, R, Qf = pb.default_weights(M)
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch, model=M)
    xg = pb.sample_goals(batch, seed=78, model=M)
    prob = capi.problem(M, N, max_iter=6, **box)
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, cost=cost, iters=iters, status=status)
    ref = oracle.slq_batch(oracle.slq_opts(M, N, max_iter=6, **box), x0, xg, qr, want_traj=False)
    assert np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"]) and np.array_equal(cost, ref["cost"])

"""Ring 2: CUDA kernels vs the CPU oracle through the C-ABI, bit-exact.

Parity statement: "kernel vs in-repo oracle, oracle vs scipy" — not "vs KahnShi/NOC" (parity unpinned,
the reference has no solver source).  Tolerance written here: 0 (np.array_equal) for everything the spec
orders; north_star's bar is 1e-9 relative.
"""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(0, 0, 0), (1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 1, 0)],
                ids=["small-batch-kernel", "tensor-core-sweep", "tensor-core-sweep-interlock", "tensor-core-sweep-direct", "dfma-sweep"])
def solver(request):
    """Every test of this module runs three times: batches of at most two instances per SM go to the warp-per-instance
    recursion (k_ric12_tp<DFMA>, one chunk) by default; with NOC_OPT_NO_SMALL_BATCH_KERNEL they go to the
    8-instances-per-warp sweeps — k_ric12_mma (products on the FP64 tensor cores, A_THEN_B records; its three staging
    variants) and, with NOC_OPT_NO_TENSOR_SWEEP, k_ric12 (DFMA chains).  All must reproduce the oracle bit for bit."""
    from noc_b200 import capi
    s = capi.Solver(0)
    s.set_option(capi.OPT_NO_SMALL_BATCH_KERNEL, request.param[0])
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, request.param[1])
    s.set_option(capi.OPT_TENSOR_SWEEP_VARIANT, request.param[2])
    yield s
    s.close()


def _traj_AB(oracle, x0, N, layout):
    ubar = np.tile(pb.hover_input(), (N, 1))
    return np.stack([oracle.linearize_traj(0, N, 0.02, oracle.rollout_open_loop(0, N, 0.02, x, ubar), ubar, layout)
                     for x in x0])


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("batch,N", [(1, 1), (3, 5), (64, 100), (37, 50)])
def test_lqr_sweep_bitexact(oracle, solver, layout, batch, N):
    from noc_b200 import capi
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=1000 + batch)
    AB = _traj_AB(oracle, x0, N, layout)
    ref = oracle.lqr_batch(12, 4, N, AB, qr, layout=layout, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12))
    cost = np.empty(batch); status = np.full(batch, -1, dtype=np.int32)
    prob = capi.problem(capi.MODEL_LTV_USER, N, layout=layout)
    solver.lqr_solve_batch(prob, batch, AB, qr, x0dev=x0, K=K, K0=K0, P0=P0, cost=cost, status=status)
    assert np.array_equal(status, ref["status"]) and np.all(status == 0)
    assert np.array_equal(K, ref["K"])
    assert np.array_equal(K0, ref["K0"])
    assert np.array_equal(P0, ref["P0"])
    assert np.array_equal(cost, ref["cost"])
    assert np.array_equal(P0, P0.transpose(0, 2, 1))


def test_lqr_dense_weights_and_random_ltv(oracle, solver):
    """General dense SPD Q, R, Qf and unstructured random A_k, B_k (not a quadrotor)."""
    from noc_b200 import capi
    rng = np.random.default_rng(5)
    batch, N = 10, 30
    def spd(k, s):
        M = rng.normal(size=(k, k)); return s * (M @ M.T + k * np.eye(k))
    Q, R, Qf = spd(12, 0.1), spd(4, 0.05), spd(12, 1.0)
    qr = pb.pack_qrqf(Q, R, Qf)
    A = np.eye(12) + 0.05 * rng.normal(size=(batch, N, 12, 12))
    B = 0.1 * rng.normal(size=(batch, N, 12, 4))
    AB = np.concatenate([A.reshape(batch, N, 144), B.reshape(batch, N, 48)], axis=2)
    x0 = rng.normal(size=(batch, 12))
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N), batch, AB, qr, x0dev=x0, K=K, P0=P0, cost=cost,
                           status=status)
    assert np.all(status == 0)
    assert np.array_equal(K, ref["K"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])


def test_lqr_not_pd_rule(oracle, solver):
    """An indefinite R makes R + B'PB fail at some step for some instances: NaN outputs, status 1."""
    from noc_b200 import capi
    batch, N = 6, 12
    Q, R, Qf = pb.default_weights()
    R = R.copy(); R[2, 2] = -40.0
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=21)
    AB = _traj_AB(oracle, x0, N, 0)
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    assert np.any(ref["status"] == 1)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N), batch, AB, qr, x0dev=x0, K=K, P0=P0, cost=cost,
                           status=status)
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(np.isnan(K), np.isnan(ref["K"]))
    ok = ~np.isnan(ref["K"])
    assert np.array_equal(K[ok], ref["K"][ok])
    assert np.array_equal(np.isnan(P0), np.isnan(ref["P0"])) and np.array_equal(np.isnan(cost), np.isnan(ref["cost"]))


@pytest.mark.parametrize("model", [0, 1])
def test_rollout_and_linearize_bitexact(oracle, solver, model):
    from noc_b200 import capi
    n, m = pb.dims(model)
    batch, N = 9, 23
    x0 = pb.sample_x0(batch, seed=31, model=model)
    rng = np.random.default_rng(7)
    ubar = pb.hover_input(model) + rng.uniform(-0.5, 0.5, (batch, N, m))
    prob = capi.problem(model, N)
    xbar = np.empty((batch, N + 1, n))
    solver.rollout_open_loop(prob, batch, x0, ubar, xbar)
    ref = np.stack([oracle.rollout_open_loop(model, N, 0.02, x0[b], ubar[b]) for b in range(batch)])
    assert np.array_equal(xbar, ref)
    xh = np.empty_like(xbar)
    solver.rollout_open_loop(prob, batch, x0, None, xh)
    uh = np.tile(pb.hover_input(model), (N, 1))
    assert np.array_equal(xh, np.stack([oracle.rollout_open_loop(model, N, 0.02, x0[b], uh) for b in range(batch)]))
    for layout in (0, 1):
        prob = capi.problem(model, N, layout=layout)
        AB = np.empty((batch, N, n * n + n * m))
        solver.linearize_batch(prob, batch, xbar, ubar, AB)
        refAB = np.stack([oracle.linearize_traj(model, N, 0.02, xbar[b], ubar[b], layout) for b in range(batch)])
        assert np.array_equal(AB, refAB)


def test_from_x0_pipeline_bitexact(oracle, solver):
    from noc_b200 import capi
    batch, N = 50, 100
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=41)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr)
    K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])


def test_from_x0_wild_states_match_dense_oracle(oracle, solver):
    """The fused sweep skips the structural zeros of the Jacobian (model.cuh quad_jac_nz_*); the oracle multiplies
    them.  Both give the same bits as long as the intermediates are finite — also far from hover: angles in every
    quadrant of the sincos reduction, pitch near +-pi/2 (sec ~ 1e3), body rates of tens of rad/s, and states that make
    the sweep fail (status NOT_PD, NaN outputs) must fail identically."""
    from noc_b200 import capi
    batch, N = 96, 60
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    rng = np.random.default_rng(77)
    x0 = pb.sample_x0(batch, seed=43)
    x0[:, 6:9] = rng.uniform(-3.1, 3.1, size=(batch, 3))
    x0[:32, 7] = np.pi / 2 + rng.uniform(-2e-3, 2e-3, size=32)
    x0[:, 9:12] = rng.uniform(-40.0, 40.0, size=(batch, 3))
    x0[32:40, 9:12] *= 1e3
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr)
    K = np.empty((batch, N, 4, 12)); K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K=K, K0=K0, P0=P0, cost=cost, status=status)
    assert np.array_equal(status, ref["status"])
    ok = status == 0
    assert ok.sum() >= batch // 2
    assert np.array_equal(K0, ref["K0"], equal_nan=True) and np.array_equal(P0, ref["P0"], equal_nan=True)
    assert np.array_equal(cost, ref["cost"], equal_nan=True)
    refK = oracle.lqr_batch(12, 4, N, oracle.make_ab_batch(0, N, 0.02, x0), qr, x0dev=x0)["K"]
    assert np.array_equal(K, refK, equal_nan=True)   # all gains, incl. the NaN fill of failed instances


def test_device_pointers_and_argmin(oracle, solver):
    """torch CUDA tensors are used in place (no staging); argmin matches numpy and skips NaN."""
    import torch
    from noc_b200 import capi
    batch, N = 33, 20
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=51)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr)
    dx0 = torch.from_numpy(x0).cuda(); dqr = torch.from_numpy(qr).cuda()
    dcost = torch.empty(batch, dtype=torch.float64, device="cuda")
    dK0 = torch.empty(batch, 4, 12, dtype=torch.float64, device="cuda")
    solver.lqr_solve_from_x0(capi.problem(0, N), batch, dx0, dqr, K0=dK0, cost=dcost)
    assert np.array_equal(dcost.cpu().numpy(), ref["cost"]) and np.array_equal(dK0.cpu().numpy(), ref["K0"])
    idx, val = solver.argmin_cost(dcost, batch)
    assert idx == int(np.argmin(ref["cost"])) and val == ref["cost"].min()
    c = ref["cost"].copy(); c[int(np.argmin(c))] = np.nan
    idx2, val2 = solver.argmin_cost(c, batch)
    assert idx2 == int(np.nanargmin(c)) and val2 == np.nanmin(c)


def test_large_batch_properties(solver):
    """Full config-2 shape is too slow for the scalar oracle; check size-independent properties:
    determinism across two runs, symmetry / positive-definiteness of P0, status all OK, and agreement of a
    strided sample with the oracle."""
    from noc_b200 import capi
    from oracle import pyoracle as oracle
    batch, N = 16384, 100
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=61)
    out = []
    for _ in range(2):
        K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
        status = np.full(batch, -1, dtype=np.int32)
        solver.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K0=K0, P0=P0, cost=cost, status=status)
        out.append((K0, P0, cost, status))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    K0, P0, cost, status = out[0]
    assert np.all(status == 0) and np.all(cost > 0)
    assert np.array_equal(P0, P0.transpose(0, 2, 1))
    assert np.all(np.linalg.eigvalsh(P0[::97]) > 0)
    sel = np.arange(0, batch, 331)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0[sel], qr)
    assert np.array_equal(K0[sel], ref["K0"]) and np.array_equal(P0[sel], ref["P0"]) and np.array_equal(cost[sel], ref["cost"])


def test_lqr_tail_and_big_mixed(oracle, solver):
    """Batch sizes that are not a multiple of the 8 instances a warp owns / the 32 a CTA owns."""
    from noc_b200 import capi
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    for batch, N in [(7, 9), (9, 3), (33, 17), (257, 11)]:
        x0 = pb.sample_x0(batch, seed=77 + batch)
        AB = oracle.make_ab_batch(0, N, 0.02, x0)
        ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
        K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
        status = np.full(batch, -1, dtype=np.int32)
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N), batch, AB, qr, x0dev=x0, K=K, P0=P0, cost=cost,
                               status=status)
        assert np.all(status == 0)
        assert np.array_equal(K, ref["K"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])


@pytest.mark.parametrize("n,m", [(1, 1), (2, 1), (6, 2), (7, 3), (12, 8), (20, 4), (24, 3)])
@pytest.mark.parametrize("layout", [0, 1])
def test_generic_dimensions_bitexact(oracle, solver, n, m, layout):
    """NOC_MODEL_LTV_USER with dimensions other than the tuned (12,4) / (24,8): the thread-per-instance kernel."""
    from noc_b200 import capi
    rng = np.random.default_rng(100 * n + m)
    batch, N = 70, 9

    def spd(k, s):
        Mx = rng.normal(size=(k, k)); return s * (Mx @ Mx.T + k * np.eye(k))
    Q, R, Qf = spd(n, 0.1), spd(m, 0.05), spd(n, 1.0)
    qr = pb.pack_qrqf(Q, R, Qf)
    A = np.eye(n) + 0.05 * rng.normal(size=(batch, N, n, n))
    B = 0.2 * rng.normal(size=(batch, N, n, m))
    if layout == 0:
        AB = np.concatenate([A.reshape(batch, N, n * n), B.reshape(batch, N, n * m)], axis=2)
    else:
        AB = np.concatenate([A, B], axis=3).reshape(batch, N, n * (n + m))
    x0 = rng.normal(size=(batch, n))
    ref = oracle.lqr_batch(n, m, N, AB, qr, layout=layout, x0dev=x0)
    K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, n=n, m=m, layout=layout), batch, AB, qr, x0dev=x0, K=K,
                           K0=K0, P0=P0, cost=cost, status=status)
    assert np.all(status == 0) and np.array_equal(status, ref["status"])
    assert np.array_equal(K, ref["K"]) and np.array_equal(K0, ref["K0"])
    assert np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])


def test_generic_dimensions_not_pd(oracle, solver):
    from noc_b200 import capi
    rng = np.random.default_rng(8)
    n, m, batch, N = 5, 2, 12, 7
    Q = np.eye(n); R = np.diag([0.1, -5.0]); Qf = np.eye(n)
    qr = pb.pack_qrqf(Q, R, Qf)
    AB = np.concatenate([(np.eye(n) + 0.1 * rng.normal(size=(batch, N, n, n))).reshape(batch, N, n * n),
                         (0.3 * rng.normal(size=(batch, N, n, m))).reshape(batch, N, n * m)], axis=2)
    ref = oracle.lqr_batch(n, m, N, AB, qr)
    assert np.any(ref["status"] == 1)
    K = np.empty((batch, N, m, n)); P0 = np.empty((batch, n, n)); status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, n=n, m=m), batch, AB, qr, K=K, P0=P0, status=status)
    assert np.array_equal(status, ref["status"]) and np.array_equal(np.isnan(K), np.isnan(ref["K"]))
    ok = ~np.isnan(ref["K"])
    assert np.array_equal(K[ok], ref["K"][ok]) and np.array_equal(np.isnan(P0), np.isnan(ref["P0"]))


@pytest.mark.parametrize("n,m", [(12, 4), (24, 8), (5, 2)])
def test_per_instance_weights_bitexact(oracle, solver, n, m):
    """qr_per_instance=1: every instance brings its own Q | R | Qf block; each must equal the oracle run with that
    instance's weights ((12,4): k_ric12_mma<PERW>; the other shapes take the thread-per-instance kernel)."""
    from noc_b200 import capi
    rng = np.random.default_rng(500 + n)
    batch, N = 7, 12
    AB = np.empty((batch, N, n * n + n * m))
    for b in range(batch):
        for k in range(N):
            A = np.eye(n) + 0.05 * rng.standard_normal((n, n)); Bm = 0.2 * rng.standard_normal((n, m))
            AB[b, k] = np.concatenate([A.ravel(), Bm.ravel()])
    qr = np.empty((batch, 2 * n * n + m * m))
    for b in range(batch):
        Q = np.diag(rng.uniform(0.5, 5.0, n)); R = np.diag(rng.uniform(0.05, 1.0, m)); Qf = 10.0 * Q
        M = 0.1 * rng.standard_normal((n, n)); Q = Q + M @ M.T       # dense symmetric PSD part
        qr[b] = np.concatenate([Q.ravel(), R.ravel(), Qf.ravel()])
    x0 = rng.standard_normal((batch, n))
    K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    prob = capi.problem(capi.MODEL_LTV_USER, N, n=n, m=m)
    solver.lqr_solve_batch(prob, batch, AB, qr, x0dev=x0, K=K, K0=K0, P0=P0, cost=cost, status=status, qr_per_instance=True)
    for b in range(batch):
        ref = oracle.lqr_batch(n, m, N, AB[b:b + 1], qr[b], x0dev=x0[b:b + 1])
        assert status[b] == ref["status"][0] == 0
        assert np.array_equal(K[b], ref["K"][0]) and np.array_equal(K0[b], ref["K0"][0])
        assert np.array_equal(P0[b], ref["P0"][0]) and cost[b] == ref["cost"][0]


@pytest.mark.parametrize("n,m,batch,N", [(12, 4, 1, 1), (12, 4, 33, 3), (12, 4, 1003, 40), (12, 4, 2 * 148 * 32 + 13, 9),
                                         (24, 8, 1, 1), (24, 8, 33, 3), (24, 8, 8 * 148 + 5, 6), (24, 8, 12 * 148 + 7, 4)])
def test_per_instance_weights_tensor_sweep_bitexact(oracle, n, m, batch, N):
    """(12,4) / (24,8) with A_THEN_B records and qr_per_instance=1 run on k_ric12_mma<PERW> (W' fragments per instance-step
    from the caller's QRQf array) / k_ric24_mma<PERW> (the warp stages its instance's W' tiles): ragged last warp / CTA,
    one-warp and six-warp CTAs, more than one wave, device-resident weights; bit-identical to the oracle
    run per instance, to the thread-per-instance kernel (NOC_OPT_NO_TENSOR_SWEEP), and — with the UNUSED triangles of
    Q, R, Qf overwritten by NaN — still the same (only Q[i][j], i <= j, R[a][b], b <= a, Qf[i][j], i <= j are read)."""
    from noc_b200 import capi
    import torch
    solver = capi.Solver(0)          # (its own context: the test toggles NOC_OPT_NO_TENSOR_SWEEP)
    rng = np.random.default_rng(900 + batch + n)
    A = np.eye(n)[None, None] + 0.05 * rng.standard_normal((batch, N, n, n))
    Bm = 0.2 * rng.standard_normal((batch, N, n, m))
    AB = np.concatenate([A.reshape(batch, N, -1), Bm.reshape(batch, N, -1)], axis=2)
    d = rng.uniform(0.5, 5.0, (batch, n)); M = 0.1 * rng.standard_normal((batch, n, n))
    Q = np.einsum("bij,bkj->bik", M, M) + d[:, :, None] * np.eye(n)
    L = 0.05 * rng.standard_normal((batch, m, m))
    R = np.einsum("bij,bkj->bik", L, L) + rng.uniform(0.05, 1.0, (batch, m))[:, :, None] * np.eye(m)
    Qf = 10.0 * Q + np.eye(n)
    qr = np.concatenate([Q.reshape(batch, -1), R.reshape(batch, -1), Qf.reshape(batch, -1)], axis=1)
    x0 = rng.standard_normal((batch, n))
    prob = capi.problem(capi.MODEL_LTV_USER, N, n=n, m=m)

    def run(qr_arg, no_tensor=0):
        K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
        status = np.full(batch, -1, dtype=np.int32)
        solver.set_option(capi.OPT_NO_TENSOR_SWEEP, no_tensor)
        try:
            solver.lqr_solve_batch(prob, batch, AB, qr_arg, x0dev=x0, K=K, K0=K0, P0=P0, cost=cost, status=status,
                                   qr_per_instance=True)
        finally:
            solver.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        return K, K0, P0, cost, status

    got = run(qr)
    assert np.all(got[4] == 0)
    ref = oracle.lqr_batch(n, m, N, AB, qr, x0dev=x0, qr_per_instance=True)
    for a, key in zip(got, ("K", "K0", "P0", "cost", "status")):
        assert np.array_equal(a, ref[key]), key
    if batch * N * n <= 1003 * 40 * 12:
        for a, b in zip(got, run(qr, no_tensor=1)):
            assert np.array_equal(a, b)
    # device-resident weights with poisoned unused triangles
    Qp = Q.copy(); Rp = R.copy(); Qfp = Qf.copy()
    il = np.tril_indices(n, -1); iu = np.triu_indices(m, 1)
    Qp[:, il[0], il[1]] = np.nan; Qfp[:, il[0], il[1]] = np.nan; Rp[:, iu[0], iu[1]] = np.nan
    qrp = torch.from_numpy(np.concatenate([Qp.reshape(batch, -1), Rp.reshape(batch, -1), Qfp.reshape(batch, -1)], axis=1)).cuda()
    for a, b in zip(got, run(qrp)):
        assert np.array_equal(a, b)
    solver.close()


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("want_xbar", [False, True])
def test_lqr_along_nominal_inputs_bitexact(oracle, solver, model, want_xbar):
    """noc_lqr_solve_along: time-varying LQR around the open-loop rollout of a given input sequence (fused
    linearisation) == oracle rollout + linearisation + sweep, bit for bit; with and without the nominal returned."""
    from noc_b200 import capi
    n, m = pb.dims(model)
    batch, N = 21, 30
    rng = np.random.default_rng(60 + model)
    qr = pb.pack_qrqf(*pb.default_weights(model))
    x0 = pb.sample_x0(batch, seed=61, model=model)
    ubar = pb.hover_input(model)[None, None, :] + 0.4 * rng.standard_normal((batch, N, m))
    xbar = np.empty((batch, N + 1, n)) if want_xbar else None
    K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
    status = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_along(capi.problem(model, N), batch, x0, ubar, qr, xbar=xbar, K=K, K0=K0, P0=P0, cost=cost, status=status)
    for b in range(batch):
        xr = oracle.rollout_open_loop(model, N, 0.02, x0[b], ubar[b])
        AB = oracle.linearize_traj(model, N, 0.02, xr, ubar[b])
        ref = oracle.lqr_batch(n, m, N, AB[None], qr, x0dev=x0[b:b + 1])
        assert status[b] == 0 == ref["status"][0]
        assert np.array_equal(K[b], ref["K"][0]) and np.array_equal(K0[b], ref["K0"][0])
        assert np.array_equal(P0[b], ref["P0"][0]) and cost[b] == ref["cost"][0]
        if want_xbar:
            assert np.array_equal(xbar[b], xr)


def test_lqr_along_large_host_batch_chunks(oracle, solver):
    """The overlap schedule (host outputs, batch >= 8192, two streams) with a nominal input sequence and the nominal
    returned: sample of instances against the oracle."""
    from noc_b200 import capi
    batch, N = 20000, 12
    rng = np.random.default_rng(70)
    qr = pb.pack_qrqf(*pb.default_weights())
    x0 = pb.sample_x0(batch, seed=71)
    ubar = pb.hover_input()[None, None, :] + 0.3 * rng.standard_normal((batch, N, 4))
    xbar = np.empty((batch, N + 1, 12)); K0 = np.empty((batch, 4, 12)); st = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_along(capi.problem(0, N), batch, x0, ubar, qr, xbar=xbar, K0=K0, status=st)
    assert np.all(st == 0)
    for b in range(0, batch, 997):
        xr = oracle.rollout_open_loop(0, N, 0.02, x0[b], ubar[b])
        ref = oracle.lqr_batch(12, 4, N, oracle.linearize_traj(0, N, 0.02, xr, ubar[b])[None], qr)
        assert np.array_equal(xbar[b], xr) and np.array_equal(K0[b], ref["K0"][0])

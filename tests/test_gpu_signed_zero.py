"""Bit patterns, not values: `np.array_equal` treats +0.0 and -0.0 as equal, so the parity tests could not see a gain that
is an exact zero with the other sign.  Such zeros are common in structurally sparse problems — hover nominals, diagonal
weights, the last step of a sweep — and the kernels' negated back substitution (a chain on -z instead of the oracle's
z, negated at the end) produced +0 where the oracle writes -0 until the end of round 2 (found by the memcmp of
tests/cpp/facade_test.cpp).  These tests compare the int64 views of every output on exactly those problems, through
every kernel family: the LTV sweeps (k_ric12_tp one-chunk, k_ric12_mma and its PERW build, k_ric12, k_ric24_mma (+PERW),
k_ric24, k_ric_generic), the fused from-x0 sweeps, SLQ (k_ric12<AFFINE>, k_ric12_wpi, k_ric24_mma<AFFINE>, with and
without the input box) and the DARE iterations."""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.int64), b.view(np.int64))


OPTS = {"default": (0, 0), "no-small-batch-kernel": (1, 0), "dfma": (1, 1)}


@pytest.fixture(params=list(OPTS), ids=list(OPTS))
def solver(request):
    from noc_b200 import capi
    s = capi.Solver(0)
    s.set_option(capi.OPT_NO_SMALL_BATCH_KERNEL, OPTS[request.param][0])
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, OPTS[request.param][1])
    yield s
    s.close()


def test_there_are_negative_zeros_to_find(oracle):
    """The problems below do contain exact zeros of both signs in the oracle's gains (otherwise this file tests nothing)."""
    qr = pb.pack_qrqf(*pb.default_weights(0))
    x0 = pb.hover_state(3)
    AB = oracle.make_ab_batch(0, 10, 0.02, x0)
    K = oracle.lqr_batch(12, 4, 10, AB, qr, x0dev=x0)["K"]
    z = K == 0.0
    assert z.any() and np.signbit(K[z]).any()


@pytest.mark.parametrize("model,batch,N", [(0, 5, 10), (0, 700, 7), (1, 3, 6), (1, 300, 4)])
@pytest.mark.parametrize("per_instance", [False, True])
def test_ltv_sweep_zero_signs(oracle, solver, model, batch, N, per_instance):
    from noc_b200 import capi
    n, m = pb.dims(model)
    qr = pb.pack_qrqf(*pb.default_weights(model))
    x0 = pb.hover_state(batch, model=model)
    x0[1::2] = pb.sample_x0(batch, seed=77, model=model)[1::2]       # every other instance away from hover
    AB = oracle.make_ab_batch(model, N, 0.02, x0)
    if per_instance:
        qr = np.tile(qr, (batch, 1)) * np.linspace(0.5, 2.0, batch)[:, None]
    ref = oracle.lqr_batch(n, m, N, AB, qr, x0dev=x0, qr_per_instance=per_instance)
    K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
    st = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N, n=n, m=m), batch, AB, qr, x0dev=x0, K=K, K0=K0, P0=P0,
                           cost=cost, status=st, qr_per_instance=per_instance)
    assert np.all(st == 0) and np.all(ref["status"] == 0)
    assert same_bits(K, ref["K"]) and same_bits(K0, ref["K0"]) and same_bits(P0, ref["P0"]) and same_bits(cost, ref["cost"])


@pytest.mark.parametrize("model,batch,N", [(0, 9, 12), (0, 1500, 5), (1, 4, 8), (1, 400, 3)])
def test_fused_from_x0_zero_signs(oracle, solver, model, batch, N):
    from noc_b200 import capi
    n, m = pb.dims(model)
    qr = pb.pack_qrqf(*pb.default_weights(model))
    x0 = pb.hover_state(batch, model=model)
    x0[1::2] = pb.sample_x0(batch, seed=78, model=model)[1::2]
    AB = oracle.make_ab_batch(model, N, 0.02, x0)
    ref = oracle.lqr_batch(n, m, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, m, n)); K0 = np.empty((batch, m, n)); P0 = np.empty((batch, n, n)); cost = np.empty(batch)
    st = np.full(batch, -1, dtype=np.int32)
    solver.lqr_solve_from_x0(capi.problem(model, N), batch, x0, qr, K=K, K0=K0, P0=P0, cost=cost, status=st)
    assert np.all(st == 0)
    assert same_bits(K, ref["K"]) and same_bits(K0, ref["K0"]) and same_bits(P0, ref["P0"]) and same_bits(cost, ref["cost"])


@pytest.mark.parametrize("model,batch,N", [(0, 6, 20), (0, 12 * 148 + 30, 6), (1, 3, 10), (1, 1300, 4)])
@pytest.mark.parametrize("box", [None, (0.0, 3.5)])
def test_slq_zero_signs(oracle, solver, model, batch, N, box):
    """Hover starts, waypoint goals: x, u, K and the cost of the last accepted iteration, bit patterns."""
    from noc_b200 import capi
    n, m = pb.dims(model)
    qr = pb.pack_qrqf(*pb.default_weights(model))
    x0 = pb.hover_state(batch, model=model); xg = pb.sample_goals(batch, seed=21, model=model)
    kw = dict(max_iter=6)
    if box:
        kw.update(u_min=box[0], u_max=box[1])
    prob = capi.problem(model, N, **kw)
    x = np.empty((batch, N + 1, n)); u = np.empty((batch, N, m)); K = np.empty((batch, N, m, n)); cost = np.empty(batch)
    iters = np.zeros(batch, dtype=np.int32); st = np.full(batch, -1, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, x=x, u=u, K=K, cost=cost, iters=iters, status=st)
    opts = oracle.slq_opts(model, N, dt=prob.dt, tol=prob.tol, max_iter=prob.max_iter, n_alpha=prob.n_alpha, reg0=prob.reg0,
                           armijo=prob.armijo, u_min=prob.u_min, u_max=prob.u_max, reg_schedule=prob.reg_schedule)
    ref = oracle.slq_batch(opts, x0, xg, qr, want_traj=True, want_K=True)
    assert np.array_equal(st, ref["status"]) and np.array_equal(iters, ref["iters"])
    assert same_bits(x, ref["x"]) and same_bits(u, ref["u"]) and same_bits(cost, ref["cost"])
    assert same_bits(K, ref["K"])


@pytest.mark.parametrize("model", [0, 1])
def test_dare_zero_signs(oracle, solver, model):
    from noc_b200 import capi
    n, m = pb.dims(model)
    Q, R, _ = pb.default_weights(model)
    xh = pb.hover_state(1, model=model)[0]; uh = pb.hover_input(model)
    A, B = oracle.linearize(model, xh, uh, 0.02)
    Pr, Kr, it = oracle.dare(A, B, Q, R)
    batch = 5
    AB = np.tile(np.concatenate([A.ravel(), B.ravel()]), (batch, 1))
    P = np.empty((batch, n, n)); K = np.empty((batch, m, n)); iters = np.zeros(batch, dtype=np.int32)
    st = np.full(batch, -1, dtype=np.int32)
    solver.dare_solve_batch(capi.problem(capi.MODEL_LTV_USER, 1, n=n, m=m, tol=1e-12, max_iter=10000), batch, AB, np.concatenate([Q.ravel(), R.ravel()]),
                            P, K, iters, st)
    assert np.all(st == 0) and np.all(iters == it)
    for b in range(batch):
        assert same_bits(P[b], Pr) and same_bits(K[b], Kr)

"""Ring 2 for the infinite-horizon LQR (row a5) and SLQ / iLQR (rows a1-a4) paths: CUDA vs the CPU
oracle through the C-ABI.  Bit-exact values; iteration counts, accepted step indices and statuses identical.
Parity statement: "kernel vs in-repo oracle, oracle vs scipy" (parity unpinned — the reference has no source)."""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(0, 0, 0, 0), (1, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)],
                ids=["default-schedule", "fixed-shapes+deferred-backtracking", "dfma-sweeps", "no-speculative-line-search"])
def solver(request):
    """Every test of this module runs under both SLQ schedules — the default (balanced wave plan, the line search settled
    before the next sweep) and round 1's fixed launch shapes with the line search's backtracking round deferred beside
    the next sweep (NOC_OPT_NO_WAVE_PLAN, NOC_OPT_DEFERRED_BACKTRACK) — and on both families of backward-pass kernels: the
    tensor-core ones (default: short n=12 work lists on k_ric12_wpi, n=24 on k_ric24_mma, both fed by step-record kernels)
    and, with NOC_OPT_NO_TENSOR_SWEEP, the DFMA kernels k_ric12 / k_ric24; and with the line search of the instances that
    backtracked last time run speculatively beside the others' first trial (default) or not
    (NOC_OPT_NO_SPECULATIVE_BACKTRACK).  None of this ever changes a result."""
    from noc_b200 import capi
    s = capi.Solver(0)
    s.set_option(capi.OPT_NO_WAVE_PLAN, request.param[0])
    s.set_option(capi.OPT_DEFERRED_BACKTRACK, request.param[1])
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, request.param[2])
    s.set_option(capi.OPT_NO_SPECULATIVE_BACKTRACK, request.param[3])
    yield s
    s.close()


def _hover_AB(oracle, batch, rng, layout=0):
    """LTI records: linearisation at hover with a random yaw and small random attitude per instance."""
    recs = []
    for _ in range(batch):
        x = np.zeros(12)
        x[6:8] = rng.uniform(-0.2, 0.2, 2)
        x[8] = rng.uniform(-np.pi, np.pi)
        A, B = oracle.linearize(0, x, pb.hover_input(), 0.02)
        if layout == 0:
            recs.append(np.concatenate([A.ravel(), B.ravel()]))
        else:
            recs.append(np.concatenate([A, B], axis=1).ravel())
    return np.stack(recs)


@pytest.mark.parametrize("layout", [0, 1])
def test_dare_bitexact_and_iteration_counts(oracle, solver, layout):
    from noc_b200 import capi
    rng = np.random.default_rng(11)
    batch = 21
    Q, R, _ = pb.default_weights()
    AB = _hover_AB(oracle, batch, rng, layout)
    qr = np.concatenate([Q.ravel(), R.ravel()])
    prob = capi.problem(capi.MODEL_LTV_USER, 1, layout=layout, tol=1e-12, max_iter=10000)
    P = np.empty((batch, 12, 12)); K = np.empty((batch, 4, 12))
    iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    solver.dare_solve_batch(prob, batch, AB, qr, P, K, iters, status)
    for b in range(batch):
        if layout == 0:
            A, B = AB[b, :144].reshape(12, 12), AB[b, 144:].reshape(12, 4)
        else:
            F = AB[b].reshape(12, 16); A, B = F[:, :12].copy(), F[:, 12:].copy()
        Pr, Kr, it = oracle.dare(A, B, Q, R, tol=1e-12, max_iter=10000)
        assert status[b] == 0 and iters[b] == it
        assert np.array_equal(P[b], Pr) and np.array_equal(K[b], Kr)
    assert 150 < iters.min() and iters.max() < 400   # 218 at exact hover


@pytest.mark.parametrize("with_u", [False, True])
def test_dare_at_operating_points_bitexact(oracle, solver, with_u):
    """noc_dare_solve_at: the built-in model linearised in the kernel (pruned Jacobian) == linearise + DARE of the
    oracle, bit for bit, iteration counts included; several warps, ragged tail."""
    from noc_b200 import capi
    rng = np.random.default_rng(21)
    batch = 77
    Q, R, _ = pb.default_weights()
    xb = np.zeros((batch, 12))
    xb[:, 6:8] = rng.uniform(-0.25, 0.25, (batch, 2)); xb[:, 8] = rng.uniform(-np.pi, np.pi, batch)
    xb[:, 9:12] = rng.uniform(-0.3, 0.3, (batch, 3))
    ub = (pb.hover_input()[None, :] + rng.uniform(-0.2, 0.2, (batch, 4))) if with_u else None
    qr = np.concatenate([Q.ravel(), R.ravel()])
    prob = capi.problem(capi.MODEL_QUAD12, 1, tol=1e-12, max_iter=10000)
    P = np.empty((batch, 12, 12)); K = np.empty((batch, 4, 12))
    iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    solver.dare_solve_at(prob, batch, xb, ub, qr, P, K, iters, status)
    for b in range(batch):
        A, B = oracle.linearize(0, xb[b], ub[b] if with_u else pb.hover_input(), 0.02)
        Pr, Kr, it = oracle.dare(A, B, Q, R, tol=1e-12, max_iter=10000)
        assert status[b] == 0 and iters[b] == it
        assert np.array_equal(P[b], Pr) and np.array_equal(K[b], Kr)


def test_dare_max_iter_status(oracle, solver):
    from noc_b200 import capi
    rng = np.random.default_rng(12)
    batch = 5
    Q, R, _ = pb.default_weights()
    AB = _hover_AB(oracle, batch, rng)
    qr = np.concatenate([Q.ravel(), R.ravel()])
    prob = capi.problem(capi.MODEL_LTV_USER, 1, tol=1e-12, max_iter=17)
    P = np.empty((batch, 12, 12)); K = np.empty((batch, 4, 12))
    iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    solver.dare_solve_batch(prob, batch, AB, qr, P, K, iters, status)
    assert np.all(status == 2) and np.all(iters == 17)
    for b in range(batch):
        Pr, Kr, it = oracle.dare(AB[b, :144].reshape(12, 12), AB[b, 144:].reshape(12, 4), Q, R, tol=1e-12, max_iter=17)
        assert it == 17 and np.array_equal(P[b], Pr) and np.array_equal(K[b], Kr)


@pytest.mark.parametrize("n,m,layout", [(24, 8, 0), (24, 8, 1), (6, 2, 0), (3, 1, 1)])
def test_dare_other_dimensions_bitexact(oracle, solver, n, m, layout):
    """noc_dare_solve_batch beyond the tuned (12,4) shape: the dual-UAV model linearised near hover (n=24, m=8) and
    small random stable systems, through the thread-per-instance kernel; P, K, iteration count, and the
    NOT_PD / MAX_ITER outcomes as the oracle's."""
    from noc_b200 import capi
    rng = np.random.default_rng(100 + n)
    batch = 9
    recs, ABs = [], []
    for b in range(batch):
        if n == 24:
            x = pb.hover_state(1, model=1)[0].copy()
            x[[6, 7, 18, 19]] = rng.uniform(-0.2, 0.2, 4)
            x[[8, 20]] = rng.uniform(-np.pi, np.pi, 2)
            A, B = oracle.linearize(1, x, pb.hover_input(model=1), 0.02)
        else:
            A = np.eye(n) + 0.05 * rng.standard_normal((n, n))
            B = 0.3 * rng.standard_normal((n, m))
        ABs.append((np.ascontiguousarray(A), np.ascontiguousarray(B)))
        recs.append(np.concatenate([A.ravel(), B.ravel()]) if layout == 0 else np.concatenate([A, B], axis=1).ravel())
    AB = np.stack(recs)
    if n == 24:
        Q, R, _ = pb.default_weights(1)
    else:
        Q = np.diag(rng.uniform(0.5, 2.0, n)); R = np.diag(rng.uniform(0.1, 1.0, m))
    qr = np.concatenate([Q.ravel(), R.ravel()])
    # (24, 8) with A_THEN_B records runs on k_ric24_mma<RIC_DARE> (tensor cores); with the option set, on k_ric_generic
    for max_iter, no_tensor in ((10000, 0), (7, 0), (10000, 1)):
        solver.set_option(capi.OPT_NO_TENSOR_SWEEP, no_tensor)
        prob = capi.problem(capi.MODEL_LTV_USER, 1, n=n, m=m, layout=layout, tol=1e-12, max_iter=max_iter)
        P = np.empty((batch, n, n)); K = np.empty((batch, m, n))
        iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
        solver.dare_solve_batch(prob, batch, AB, qr, P, K, iters, status)
        solver.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        for b in range(batch):
            Pr, Kr, it = oracle.dare(ABs[b][0], ABs[b][1], Q, R, tol=1e-12, max_iter=max_iter)
            assert it > 0 and iters[b] == it
            assert status[b] == (2 if (it == max_iter and max_iter == 7) else 0)
            assert np.array_equal(P[b], Pr) and np.array_equal(K[b], Kr)
    # an indefinite R makes the first pivot fail: NOT_PD, NaN outputs, the failing iteration is counted
    Rbad = R.copy(); Rbad[0, 0] = -1e6
    qrb = np.concatenate([Q.ravel(), Rbad.ravel()])
    prob = capi.problem(capi.MODEL_LTV_USER, 1, n=n, m=m, layout=layout, tol=1e-12, max_iter=50)
    solver.dare_solve_batch(prob, batch, AB, qrb, P, K, iters, status)
    Pr, Kr, it = oracle.dare(ABs[0][0], ABs[0][1], Q, Rbad, tol=1e-12, max_iter=50)
    assert it < 0 and np.all(status == 1) and np.all(iters == -it)
    assert np.all(np.isnan(P)) and np.all(np.isnan(K))


def _slq_compare(oracle, solver, batch, N, x0, xg, qr, **kw):
    from noc_b200 import capi
    prob = capi.problem(capi.MODEL_QUAD12, N, **kw)
    x = np.empty((batch, N + 1, 12)); u = np.empty((batch, N, 4)); K = np.empty((batch, N, 4, 12))
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    aidx = np.full((batch, prob.max_iter), -7, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, x=x, u=u, K=K, cost=cost, iters=iters, alpha_idx=aidx, status=status)
    opts = oracle.slq_opts(0, N, dt=prob.dt, tol=prob.tol, max_iter=prob.max_iter, n_alpha=prob.n_alpha,
                           reg0=prob.reg0, armijo=prob.armijo, u_min=prob.u_min, u_max=prob.u_max,
                           reg_schedule=prob.reg_schedule)
    ref = oracle.slq_batch(opts, x0, xg, qr, want_traj=True, want_K=True)
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(iters, ref["iters"])          # iteration counts identical (north_star)
    assert np.array_equal(aidx, ref["alpha_idx"])       # accepted line-search indices identical
    assert np.array_equal(cost, ref["cost"])
    assert np.array_equal(x, ref["x"]) and np.array_equal(u, ref["u"])
    # K of the last backward pass of every instance
    assert np.array_equal(K, ref["K"])
    return dict(status=status, iters=iters, aidx=aidx, cost=cost, u=u, K=K)


def test_slq_waypoints_bitexact(oracle, solver):
    """Config-3 style: hover start, waypoint goals; every instance converges."""
    batch, N = 40, 60
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=5)
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr)
    assert np.all(out["status"] == 0) and out["iters"].min() >= 2
    assert len(set(out["iters"].tolist())) > 1           # instances leave the work list at different iterations


@pytest.mark.parametrize("batch,N", [(1, 1), (1, 2), (3, 3), (5, 17)])
def test_slq_tiny_horizons_and_batches(oracle, solver, batch, N):
    """Edge shapes through the warp-per-instance backward pass (k_prep12 + k_ric12_wpi: step records, double-buffered bulk
    copies, the first / last step of a sweep): horizons of one to three steps, single instances."""
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=70 + N); xg = pb.sample_goals(batch, seed=71 + N)
    _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=8)


def test_slq_list_crosses_the_short_list_threshold(oracle, solver):
    """A batch above twelve instances per SM whose work list shrinks below it during the solve: the backward pass moves
    from k_ric12<RIC_AFFINE> (eight instances per warp) to k_ric12_wpi (one per warp) between iterations, and the
    line search of the instances that backtracked runs speculatively beside the others — same bits throughout."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    batch, N = 12 * sms + 40, 12
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    rng = np.random.default_rng(12)
    x0 = pb.sample_x0(batch, seed=124)
    x0[: batch // 8, 6:8] = rng.uniform(-0.9, 0.9, (batch // 8, 2))     # some aggressive starts: backtracking
    xg = pb.sample_goals(batch, seed=16)
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=12)
    assert len(set(out["iters"].tolist())) > 2 and np.any(out["aidx"] > 0)


def test_slq_backtracking_and_max_iter(oracle, solver):
    """Aggressive starts force alpha < 1; a small max_iter leaves some instances at MAX_ITER."""
    batch, N = 24, 40
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    rng = np.random.default_rng(9)
    x0 = pb.sample_x0(batch, seed=123)
    x0[:, 6:8] = rng.uniform(-0.9, 0.9, (batch, 2))
    xg = pb.sample_goals(batch, seed=6) * 2.0
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=6)
    assert np.any(out["aidx"] > 0)
    assert np.any(out["status"] == 2)


@pytest.mark.parametrize("u_min,u_max", [(0.0, 4.0), (1.5, 3.2), (-np.inf, 3.0), (-1e6, 1e6)])
def test_slq_input_box_bitexact(oracle, solver, u_min, u_max):
    """Control-limited SLQ: box QP for the feed-forward, zero feedback rows on clamped inputs, clamped
    rollouts.  Iteration counts, step indices, trajectories, gains and costs as the oracle's, bit for bit; the inputs
    stay inside the box; several warps and a ragged tail (batch 44)."""
    batch, N = 44, 60
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=15)
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr, u_min=u_min, u_max=u_max)
    assert out["u"].min() >= u_min and out["u"].max() <= u_max
    if u_max < 100.0:
        assert np.any(out["u"] == u_max) or np.any(out["u"] == u_min)     # the box is active somewhere
        assert np.any(np.all(out["K"] == 0.0, axis=-1))                   # ... and some feedback rows are switched off
    else:
        ref = _slq_compare(oracle, solver, batch, N, x0, xg, qr)          # a box nobody touches changes nothing
        assert np.array_equal(out["cost"], ref["cost"]) and np.array_equal(out["iters"], ref["iters"])


@pytest.mark.parametrize("u_min,u_max", [(3.0, 5.0), (0.0, 2.0), (2.4, 2.5)])
def test_slq_input_box_excluding_or_hugging_hover(oracle, solver, u_min, u_max):
    """Boxes that exclude the hover thrust the nominal starts from (the first QP starts outside its box, several
    instances end with LS_FAIL in iteration 1 and keep the infeasible nominal) or barely contain it: same outcomes."""
    batch, N = 20, 60
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=34)
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr, u_min=u_min, u_max=u_max)
    moved = out["iters"] > 1
    assert moved.any() and out["u"][moved].min() >= u_min and out["u"][moved].max() <= u_max


def test_slq_input_box_with_backtracking_and_failures(oracle, solver):
    """Aggressive starts under a tight box: alpha < 1, LS_FAIL and MAX_ITER outcomes all occur and match."""
    batch, N = 24, 40
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    rng = np.random.default_rng(19)
    x0 = pb.sample_x0(batch, seed=124)
    x0[:, 6:8] = rng.uniform(-0.9, 0.9, (batch, 2))
    xg = pb.sample_goals(batch, seed=16) * 2.0
    out = _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=30, u_min=1.0, u_max=3.5)
    assert np.any(out["aidx"] > 0)
    assert {0, 2, 3} <= set(out["status"].tolist())      # converged, MAX_ITER and LS_FAIL all occur


def test_slq_regularisation_retry(oracle, solver):
    """A negative R makes Quu indefinite at mu = 0: the backward pass must raise mu exactly like the oracle."""
    batch, N = 10, 20
    Q, R, Qf = pb.default_weights()
    R = R.copy(); R[1, 1] = -0.5
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=31); xg = pb.sample_goals(batch, seed=7)
    _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=5)
    _slq_compare(oracle, solver, batch, N, x0, xg, qr, max_iter=5, u_min=1.0, u_max=3.5)   # mu retries under the input box


def test_slq_adaptive_regularisation_bitexact(oracle, solver):
    """NOC_REG_ADAPTIVE: mu lowered after accepted steps, raised on failed factorisations AND failed
    line searches (iteration repeated from the same nominal, alpha_idx = -2).  Bit-identical to the oracle."""
    from noc_b200 import capi
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    # (a) started from mu0 = 1: the schedule walks mu down, fewer iterations than the fixed rule
    batch, N = 24, 60
    x0 = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=5150)
    ada = _slq_compare(oracle, solver, batch, N, x0, xg, qr, reg_schedule=capi.REG_ADAPTIVE, reg0=1.0)
    fix = _slq_compare(oracle, solver, batch, N, x0, xg, qr, reg0=1.0)
    assert ada["iters"].sum() < fix["iters"].sum()
    # (b) tight input box: line searches fail and are repeated with a larger mu; several warps, ragged tail
    pick = [284, 217, 252, 60, 89, 101, 214, 226, 293, 490] + list(range(30))
    xg = pb.sample_goals(512, seed=33)[pick]; x0 = pb.hover_state(len(pick))
    out = _slq_compare(oracle, solver, len(pick), N, x0, xg, qr, reg_schedule=capi.REG_ADAPTIVE, u_min=0.0, u_max=4.0)
    assert (out["aidx"] == -2).any(axis=1).sum() >= 8
    # (c) an indefinite R: factorisation retries inside the sweep follow the adaptive rule too
    Rb = R.copy(); Rb[1, 1] = -0.5
    x0 = pb.sample_x0(10, seed=31); xg = pb.sample_goals(10, seed=7)
    _slq_compare(oracle, solver, 10, 20, x0, xg, pb.pack_qrqf(Q, Rb, Qf), max_iter=6, reg_schedule=capi.REG_ADAPTIVE)
    _slq_compare(oracle, solver, 10, 20, x0, xg, pb.pack_qrqf(Q, Rb, Qf), max_iter=6, reg_schedule=capi.REG_ADAPTIVE,
                 u_min=1.0, u_max=3.5)


def test_slq_optional_outputs(oracle, solver):
    from noc_b200 import capi
    batch, N = 12, 30
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=8)
    prob = capi.problem(capi.MODEL_QUAD12, N)
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x0, xg, qr, cost=cost, iters=iters)
    ref = oracle.slq_batch(oracle.slq_opts(0, N), x0, xg, qr, want_traj=False)
    assert np.array_equal(cost, ref["cost"]) and np.array_equal(iters, ref["iters"])


def test_slq_multiwarp_config_and_worklist_compaction(oracle, solver):
    """A batch large enough for the 7-warp CTA configuration of the backward pass (> 2,368 instances); instances
    converge at different iterations, so the sorted work-list compaction is exercised at scale."""
    from noc_b200 import capi
    batch, N = 3000, 24
    qr = pb.pack_qrqf(*pb.default_weights())
    xs = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=17)
    prob = capi.problem(0, N)
    cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    aidx = np.zeros((batch, prob.max_iter), dtype=np.int32)
    solver.slq_solve_batch(prob, batch, xs, xg, qr, cost=cost, iters=iters, alpha_idx=aidx, status=status)
    ref = oracle.slq_batch(oracle.slq_opts(0, N), xs, xg, qr, want_traj=False)
    assert np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(aidx, ref["alpha_idx"]) and np.array_equal(cost, ref["cost"])
    assert len(set(iters.tolist())) > 3


@pytest.mark.parametrize("model", [0, 1])
def test_slq_warm_start_bitexact_and_pays_off(oracle, solver, model):
    """noc_slq_solve_warm: start from a given input sequence (receding-horizon MPC: the previous solution shifted by
    one step, the state advanced by one step).  Same bits as the oracle started from the same sequence, and the
    warm-started solve needs fewer iterations than the cold one; u_init may alias the u output."""
    from noc_b200 import capi
    n, m = pb.dims(model)
    batch, N = 12, 40
    qr = pb.pack_qrqf(*pb.default_weights(model))
    x0 = pb.hover_state(batch, model=model); xg = pb.sample_goals(batch, seed=91, model=model)
    prob = capi.problem(model, N)
    cold = oracle.slq_batch(oracle.slq_opts(model, N), x0, xg, qr, want_traj=True)
    # one MPC step later: state x_1 of the solution, inputs shifted, last input repeated
    x1 = np.ascontiguousarray(cold["x"][:, 1]); u_init = np.concatenate([cold["u"][:, 1:], cold["u"][:, -1:]], axis=1)
    ref = oracle.slq_batch(oracle.slq_opts(model, N), x1, xg, qr, want_traj=True, u_init=u_init)
    x = np.empty((batch, N + 1, n)); u = u_init.copy(); cost = np.empty(batch)
    iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
    aidx = np.full((batch, prob.max_iter), -7, dtype=np.int32)
    solver.slq_solve_batch(prob, batch, x1, xg, qr, x=x, u=u, cost=cost, iters=iters, alpha_idx=aidx, status=status, u_init=u)
    assert np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(aidx, ref["alpha_idx"]) and np.array_equal(cost, ref["cost"])
    assert np.array_equal(x, ref["x"]) and np.array_equal(u, ref["u"])
    cold1 = oracle.slq_batch(oracle.slq_opts(model, N), x1, xg, qr, want_traj=False)
    assert iters.sum() < cold1["iters"].sum()

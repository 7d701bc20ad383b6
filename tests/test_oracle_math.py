"""Ring 1: the CPU oracle against independent math.

The reference holds no golden vectors (parity unpinned), so the oracle is pinned
against scipy / numpy-longdouble / closed forms / finite differences here.  CPU only.
"""
import numpy as np
import pytest
import scipy.linalg

from noc_b200 import problems as pb


# ---------------------------------------------------------------- elementary functions
def test_sincos_vs_libm(oracle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-4, 4, 4000), rng.uniform(-100, 100, 1000),
                        np.array([0.0, -0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 1e-300, 1e-8])])
    s, c = oracle.sincos(x)
    # <= 2 ulp of the libm value, absolute scale 1
    assert np.max(np.abs(s - np.sin(x))) < 3e-16
    assert np.max(np.abs(c - np.cos(x))) < 3e-16
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5e-16
    s0, c0 = oracle.sincos(0.0)
    assert s0[0] == 0.0 and c0[0] == 1.0


# ---------------------------------------------------------------- models
def _f_numpy(x, u):
    """Independent quadrotor dynamics: rotation matrices built with numpy (not the oracle's formulas)."""
    mass, g, L, c = 1.0, 9.81, 0.2, 0.02
    J = np.diag([0.01, 0.01, 0.02])
    ph, th, ps = x[6:9]
    Rz = np.array([[np.cos(ps), -np.sin(ps), 0], [np.sin(ps), np.cos(ps), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ph), -np.sin(ph)], [0, np.sin(ph), np.cos(ph)]])
    Rm = Rz @ Ry @ Rx
    M = np.array([[1, 1, 1, 1], [0, L, 0, -L], [-L, 0, L, 0], [c, -c, c, -c]], float)
    Tt = M @ u
    W = np.array([[1, np.sin(ph) * np.tan(th), np.cos(ph) * np.tan(th)],
                  [0, np.cos(ph), -np.sin(ph)],
                  [0, np.sin(ph) / np.cos(th), np.cos(ph) / np.cos(th)]])
    w = x[9:12]
    xd = np.zeros(12)
    xd[0:3] = x[3:6]
    xd[3:6] = Tt[0] / mass * Rm[:, 2] - np.array([0, 0, g])
    xd[6:9] = W @ w
    xd[9:12] = np.linalg.solve(J, Tt[1:] - np.cross(w, J @ w))
    return xd


def _f_dual_numpy(x, u):
    ks, cs, d = 2.0, 0.5, np.array([1.0, 0, 0])
    xd = np.concatenate([_f_numpy(x[:12], u[:4]), _f_numpy(x[12:], u[4:])])
    c = ks * (x[12:15] - x[0:3] - d) + cs * (x[15:18] - x[3:6])
    xd[3:6] += c
    xd[15:18] -= c
    return xd


@pytest.mark.parametrize("model", [pb.MODEL_QUAD12, pb.MODEL_DUAL24])
def test_dynamics_vs_numpy(oracle, model):
    rng = np.random.default_rng(2)
    n, m = pb.dims(model)
    f = _f_numpy if model == pb.MODEL_QUAD12 else _f_dual_numpy
    for _ in range(50):
        x = pb.sample_x0(1, seed=int(rng.integers(1 << 30)), model=model)[0]
        u = pb.hover_input(model) + rng.uniform(-1, 1, m)
        xn = oracle.step(model, x, u, 0.02)
        ref = x + 0.02 * f(x, u)
        assert np.allclose(xn, ref, rtol=0, atol=1e-13)


@pytest.mark.parametrize("model", [pb.MODEL_QUAD12, pb.MODEL_DUAL24])
def test_jacobians_vs_finite_differences(oracle, model):
    rng = np.random.default_rng(3)
    n, m = pb.dims(model)
    f = _f_numpy if model == pb.MODEL_QUAD12 else _f_dual_numpy
    dt = 0.02
    for _ in range(10):
        x = pb.sample_x0(1, seed=int(rng.integers(1 << 30)), model=model)[0]
        u = pb.hover_input(model) + rng.uniform(-1, 1, m)
        A, B = oracle.linearize(model, x, u, dt)
        h = 1e-6
        Afd = np.empty((n, n))
        Bfd = np.empty((n, m))
        for j in range(n):
            e = np.zeros(n); e[j] = h
            Afd[:, j] = ((x + e + dt * f(x + e, u)) - (x - e + dt * f(x - e, u))) / (2 * h)
        for j in range(m):
            e = np.zeros(m); e[j] = h
            Bfd[:, j] = dt * (f(x, u + e) - f(x, u - e)) / (2 * h)
        assert np.allclose(A, Afd, rtol=0, atol=2e-9)
        assert np.allclose(B, Bfd, rtol=0, atol=2e-9)


def test_hover_is_equilibrium(oracle):
    x = np.zeros(12)
    x[0:3] = [1.0, -2.0, 0.5]
    xn = oracle.step(pb.MODEL_QUAD12, x, oracle.hover_input(pb.MODEL_QUAD12), 0.02)
    assert np.array_equal(xn, x)


def test_layouts_agree(oracle):
    x0 = pb.sample_x0(1, seed=5)[0]
    N = 7
    ubar = np.tile(pb.hover_input(), (N, 1))
    xbar = oracle.rollout_open_loop(pb.MODEL_QUAD12, N, 0.02, x0, ubar)
    ab0 = oracle.linearize_traj(pb.MODEL_QUAD12, N, 0.02, xbar, ubar, oracle.LAYOUT_A_THEN_B)
    ab1 = oracle.linearize_traj(pb.MODEL_QUAD12, N, 0.02, xbar, ubar, oracle.LAYOUT_ROWS_AB)
    A = ab0[:, :144].reshape(N, 12, 12)
    B = ab0[:, 144:].reshape(N, 12, 4)
    F = ab1.reshape(N, 12, 16)
    assert np.array_equal(F[:, :, :12], A) and np.array_equal(F[:, :, 12:], B)
    Q, R, Qf = pb.default_weights()
    K0, P0, _, st0 = oracle.lqr_sweep(12, 4, N, ab0, Q, R, Qf, oracle.LAYOUT_A_THEN_B)
    K1, P1, _, st1 = oracle.lqr_sweep(12, 4, N, ab1, Q, R, Qf, oracle.LAYOUT_ROWS_AB)
    assert st0 == st1 == 0 and np.array_equal(K0, K1) and np.array_equal(P0, P1)


# ---------------------------------------------------------------- Riccati sweep
def _sweep_longdouble(A, B, Q, R, Qf):
    """Textbook recursion in numpy longdouble, different association from the oracle."""
    ld = np.longdouble
    N = A.shape[0]
    P = Qf.astype(ld)
    Ks = []
    for k in range(N - 1, -1, -1):
        Ak, Bk = A[k].astype(ld), B[k].astype(ld)
        S = R.astype(ld) + Bk.T @ P @ Bk
        Kk = -np.linalg.solve(S.astype(np.float64), (Bk.T @ P @ Ak).astype(np.float64)).astype(ld)
        # one step of iterative refinement in long double
        res = -(Bk.T @ P @ Ak) - S @ Kk
        Kk = Kk + np.linalg.solve(S.astype(np.float64), res.astype(np.float64)).astype(ld)
        Acl = Ak + Bk @ Kk
        P = Q.astype(ld) + Kk.T @ R.astype(ld) @ Kk + Acl.T @ P @ Acl   # Joseph form
        P = (P + P.T) / 2
        Ks.append(Kk)
    return np.array(Ks[::-1], dtype=np.float64), P.astype(np.float64)


def test_sweep_vs_longdouble_joseph(oracle):
    N = 100
    Q, R, Qf = pb.default_weights()
    worst = 0.0
    for seed in range(4):
        x0 = pb.sample_x0(1, seed=100 + seed)[0]
        ubar = np.tile(pb.hover_input(), (N, 1))
        xbar = oracle.rollout_open_loop(pb.MODEL_QUAD12, N, 0.02, x0, ubar)
        ab = oracle.linearize_traj(pb.MODEL_QUAD12, N, 0.02, xbar, ubar)
        K, P0, cost, st = oracle.lqr_sweep(12, 4, N, ab, Q, R, Qf, x0dev=x0)
        assert st == 0
        A = ab[:, :144].reshape(N, 12, 12)
        B = ab[:, 144:].reshape(N, 12, 4)
        Kr, Pr = _sweep_longdouble(A, B, Q, R, Qf)
        eK = np.max(np.abs(K - Kr)) / np.max(np.abs(Kr))
        eP = np.max(np.abs(P0 - Pr)) / np.max(np.abs(Pr))
        worst = max(worst, eK, eP)
        assert np.array_equal(P0, P0.T)
        assert np.all(np.linalg.eigvalsh(P0) > 0)
        assert abs(cost - 0.5 * x0 @ Pr @ x0) <= 1e-10 * abs(cost)
    assert worst < 1e-10, worst


def test_scalar_closed_form(oracle):
    # n=m=1: p' = q + a^2 p - (a b p)^2/(r + b^2 p), k = -a b p/(r + b^2 p)
    a, b, q, r, qf, N = 1.1, 0.5, 2.0, 0.3, 4.0, 25
    AB = np.tile(np.array([[a, b]]), (N, 1))
    K, P0, cost, st = oracle.lqr_sweep(1, 1, N, AB, [[q]], [[r]], [[qf]], x0dev=np.array([2.0]))
    p = qf
    ks = []
    for _ in range(N):
        s = r + b * b * p
        ks.append(-a * b * p / s)
        p = q + a * a * p - (a * b * p) ** 2 / s
    assert st == 0
    assert abs(P0[0, 0] - p) < 1e-12 * p
    assert np.allclose(K[:, 0, 0], ks[::-1], rtol=1e-12, atol=0)
    assert abs(cost - 0.5 * 4.0 * p) < 1e-12 * cost


def test_cost_identity_closed_loop(oracle):
    """0.5 x0' P0 x0 == sum of stage costs of the closed-loop LTV rollout (ties backward to forward)."""
    N = 60
    Q, R, Qf = pb.default_weights()
    x0 = pb.sample_x0(1, seed=77)[0]
    ubar = np.tile(pb.hover_input(), (N, 1))
    xbar = oracle.rollout_open_loop(pb.MODEL_QUAD12, N, 0.02, x0, ubar)
    ab = oracle.linearize_traj(pb.MODEL_QUAD12, N, 0.02, xbar, ubar)
    dx0 = np.random.default_rng(4).normal(size=12)
    K, P0, cost, st = oracle.lqr_sweep(12, 4, N, ab, Q, R, Qf, x0dev=dx0)
    A = ab[:, :144].reshape(N, 12, 12)
    B = ab[:, 144:].reshape(N, 12, 4)
    x = dx0.copy()
    J = 0.0
    for k in range(N):
        u = K[k] @ x
        J += 0.5 * (x @ Q @ x + u @ R @ u)
        x = A[k] @ x + B[k] @ u
    J += 0.5 * x @ Qf @ x
    assert abs(J - cost) < 1e-10 * cost


def test_not_pd_rule(oracle):
    # R negative enough that R + B'PB is indefinite at the first processed step
    N = 5
    AB = np.tile(np.array([[1.0, 0.1]]), (N, 1))
    K, P0, cost, st = oracle.lqr_sweep(1, 1, N, AB, [[1.0]], [[-5.0]], [[1.0]], x0dev=np.array([1.0]))
    assert st == oracle.NOT_PD
    assert np.all(np.isnan(K)) and np.all(np.isnan(P0)) and np.isnan(cost)


def test_batch_matches_single(oracle):
    N, Bn = 20, 6
    Q, R, Qf = pb.default_weights()
    x0 = pb.sample_x0(Bn, seed=9)
    ubar = np.tile(pb.hover_input(), (N, 1))
    AB = np.stack([oracle.linearize_traj(0, N, 0.02, oracle.rollout_open_loop(0, N, 0.02, x0[b], ubar), ubar)
                   for b in range(Bn)])
    out = oracle.lqr_batch(12, 4, N, AB, pb.pack_qrqf(Q, R, Qf), x0dev=x0, threads=3)
    out2 = oracle.lqr_from_x0_batch(0, N, 0.02, x0, pb.pack_qrqf(Q, R, Qf), threads=2)
    for b in range(Bn):
        K, P0, c, st = oracle.lqr_sweep(12, 4, N, AB[b], Q, R, Qf, x0dev=x0[b])
        assert np.array_equal(out["K"][b], K) and np.array_equal(out["P0"][b], P0)
        assert out["cost"][b] == c and out["status"][b] == 0
        assert np.array_equal(out["K0"][b], K[0])
        assert np.array_equal(out2["K0"][b], K[0]) and np.array_equal(out2["P0"][b], P0)
        assert out2["cost"][b] == c


# ---------------------------------------------------------------- infinite horizon (config 1)
@pytest.mark.parametrize("dt,iters_expected", [(0.02, None), (0.05, None)])
def test_dare_vs_scipy(oracle, dt, iters_expected):
    x = np.zeros(12)
    A, B = oracle.linearize(pb.MODEL_QUAD12, x, oracle.hover_input(0), dt)
    Q, R, _ = pb.default_weights()
    P, K, it = oracle.dare(A, B, Q, R, tol=1e-12, max_iter=5000)
    assert 0 < it < 5000
    Ps = scipy.linalg.solve_discrete_are(A, B, Q, R)
    assert np.max(np.abs(P - Ps)) / np.max(np.abs(Ps)) < 1e-9
    Ks = -np.linalg.solve(R + B.T @ Ps @ B, B.T @ Ps @ A)
    assert np.max(np.abs(K - Ks)) / np.max(np.abs(Ks)) < 1e-8
    rho = np.max(np.abs(np.linalg.eigvals(A + B @ K)))
    assert rho < 1.0
    # P is monotone in the horizon: more iterations never decrease x'Px
    P2, _, _ = oracle.dare(A, B, Q, R, tol=1e-6, max_iter=5000)
    assert np.all(np.linalg.eigvalsh(P - P2) > -1e-6)


# ---------------------------------------------------------------- SLQ / iLQR
def test_slq_converges_and_is_consistent(oracle):
    N = 60
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    opts = oracle.slq_opts(pb.MODEL_QUAD12, N)
    goals = pb.sample_goals(4, seed=11)
    for g in goals:
        x0 = np.zeros(12)
        r = oracle.slq(opts, x0, g, qr)
        assert r["status"] == oracle.OK, r
        assert 1 <= r["iters"] <= 50
        used = r["alpha_idx"][: r["iters"]]
        assert np.all(used >= 0) and np.all(r["alpha_idx"][r["iters"]:] == -1)
        # returned trajectory is dynamically feasible and its cost matches
        x, u = r["x"], r["u"]
        for k in range(N):
            assert np.array_equal(oracle.step(0, x[k], u[k], 0.02), x[k + 1])
        assert oracle.traj_cost(0, N, x, u, g, qr) == r["cost"]
        # cost is far below the initial (hover) cost and the feed-forward has vanished
        u0 = np.tile(pb.hover_input(), (N, 1))
        J0 = oracle.traj_cost(0, N, oracle.rollout_open_loop(0, N, 0.02, x0, u0), u0, g, qr)
        assert r["cost"] < 0.5 * J0
        K, kff, dV, st = oracle.backward_affine(0, N, 0.02, x, u, g, qr)
        assert st == 0 and dV[0] <= 0 and dV[1] >= 0
        assert abs(dV[0]) < 1e-5 * r["cost"]


def test_slq_cost_monotone(oracle):
    """Run with increasing max_iter: accepted costs never increase (Armijo)."""
    N = 40
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    g = pb.sample_goals(1, seed=3)[0]
    prev = np.inf
    for mi in range(1, 8):
        r = oracle.slq(oracle.slq_opts(0, N, max_iter=mi), np.zeros(12), g, qr)
        assert r["cost"] <= prev
        prev = r["cost"]


def test_slq_on_lq_problem_one_newton_step(oracle):
    """Near hover with a tiny goal offset the problem is almost LQ: alpha=1 accepted, few iterations."""
    N = 30
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    g = np.zeros(12); g[0] = 1e-3
    r = oracle.slq(oracle.slq_opts(0, N), np.zeros(12), g, qr)
    assert r["status"] == 0 and r["iters"] <= 3 and r["alpha_idx"][0] == 0


def test_slq_dual24_runs(oracle):
    N = 30
    Q, R, Qf = pb.default_weights(pb.MODEL_DUAL24)
    qr = pb.pack_qrqf(Q, R, Qf)
    g = pb.sample_goals(1, seed=5, model=pb.MODEL_DUAL24)[0]
    x0 = pb.hover_state(1, pb.MODEL_DUAL24)[0]
    r = oracle.slq(oracle.slq_opts(pb.MODEL_DUAL24, N), x0, g, qr)
    assert r["status"] == 0
    b = oracle.slq_batch(oracle.slq_opts(pb.MODEL_DUAL24, N), x0[None], g[None], qr, threads=1)
    assert b["cost"][0] == r["cost"] and b["iters"][0] == r["iters"]
    assert np.array_equal(b["alpha_idx"][0], r["alpha_idx"])


# ---------------------------------------------------------------- SLQ against an independent numpy iLQR
def _numpy_ilqr(oracle, x0, xg, Q, R, Qf, N, dt=0.02, tol=1e-8, max_iter=50, n_alpha=11, armijo=1e-4):
    """Textbook iLQR written with numpy matrix algebra (np.linalg.solve, no LDL^T, no fma chains): the same
    specification in a different arithmetic, so agreement is to rounding, not to the bit."""
    uh = pb.hover_input()
    f = lambda x, u: x + dt * _f_numpy(x, u)

    def cost(xs, us):
        J = 0.0
        for k in range(N):
            ex, eu = xs[k] - xg, us[k] - uh
            J += 0.5 * (ex @ Q @ ex + eu @ R @ eu)
        ex = xs[N] - xg
        return J + 0.5 * ex @ Qf @ ex

    us = np.tile(uh, (N, 1))
    xs = np.empty((N + 1, 12)); xs[0] = x0
    for k in range(N):
        xs[k + 1] = f(xs[k], us[k])
    J = cost(xs, us)
    alphas_taken = []
    for it in range(max_iter):
        Vxx, Vx = Qf.copy(), Qf @ (xs[N] - xg)
        Ks, ks = np.empty((N, 4, 12)), np.empty((N, 4))
        dV1 = dV2 = 0.0
        for k in range(N - 1, -1, -1):
            A, B = oracle.linearize(0, xs[k], us[k], dt)
            Qx = Q @ (xs[k] - xg) + A.T @ Vx
            Qu = R @ (us[k] - uh) + B.T @ Vx
            Qxx = Q + A.T @ Vxx @ A
            Qux = B.T @ Vxx @ A
            Quu = R + B.T @ Vxx @ B
            Ks[k] = -np.linalg.solve(Quu, Qux)
            ks[k] = -np.linalg.solve(Quu, Qu)
            Vxx = Qxx + Qux.T @ Ks[k]
            Vxx = 0.5 * (Vxx + Vxx.T)
            Vx = Qx + Qux.T @ ks[k]
            dV1 += ks[k] @ Qu
            dV2 += 0.5 * ks[k] @ Quu @ ks[k]
        accepted = None
        for j in range(n_alpha):
            a = 0.5 ** j
            xn = np.empty_like(xs); un = np.empty_like(us); xn[0] = x0
            for k in range(N):
                un[k] = us[k] + a * ks[k] + Ks[k] @ (xn[k] - xs[k])
                xn[k + 1] = f(xn[k], un[k])
            Jn = cost(xn, un)
            if J - Jn >= armijo * (-(a * dV1 + a * a * dV2)):
                accepted = j
                break
        if accepted is None:
            return J, it + 1, alphas_taken, 3
        alphas_taken.append(accepted)
        rel = (J - Jn) / max(J, 1e-12)
        xs, us, J = xn, un, Jn
        if rel < tol:
            return J, it + 1, alphas_taken, 0
    return J, max_iter, alphas_taken, 2


def test_slq_vs_independent_numpy_ilqr(oracle):
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    N = 30
    goals = pb.sample_goals(4, seed=404)
    for b in range(4):
        x0 = pb.hover_state(1)[0] if b < 2 else pb.sample_x0(1, seed=50 + b)[0]
        ref = oracle.slq(oracle.slq_opts(0, N), x0, goals[b], qr)
        J, iters, alphas, status = _numpy_ilqr(oracle, x0, goals[b], Q, R, Qf, N)
        assert status == ref["status"]                                   # converged, or MAX_ITER for the hard start
        assert iters == ref["iters"]                                   # identical iteration counts
        assert alphas == [int(a) for a in ref["alpha_idx"][:iters]]     # identical accepted step lengths
        assert abs(J - ref["cost"]) <= 1e-9 * abs(ref["cost"])          # north_star's tolerance


def test_box_qp_vs_scipy(oracle):
    """The projected-Newton box QP of the control-limited backward pass against scipy's bounded
    least squares on the Cholesky factor: same minimiser, KKT conditions hold, clamped set consistent."""
    from scipy.optimize import lsq_linear
    rng = np.random.default_rng(0)
    for _ in range(200):
        m = int(rng.choice([4, 8]))
        M = rng.standard_normal((m, m)); H = M @ M.T + 0.1 * np.eye(m); g = 3 * rng.standard_normal(m)
        lo = -rng.uniform(0.05, 1.0, m); hi = rng.uniform(0.05, 1.0, m)
        x, cl = oracle.box_qp(H, g, lo, hi, x0=-np.linalg.solve(H, g))
        L = np.linalg.cholesky(H)
        r = lsq_linear(L.T, -np.linalg.solve(L, g), bounds=(lo, hi), tol=1e-14, max_iter=500)
        assert np.abs(x - r.x).max() < 1e-6
        grad = H @ x + g
        for a in range(m):
            at_lo, at_hi = x[a] <= lo[a], x[a] >= hi[a]
            assert abs(grad[a]) < 1e-7 or (at_lo and grad[a] > 0) or (at_hi and grad[a] < 0)
            assert cl[a] == ((at_lo and grad[a] > 0) or (at_hi and grad[a] < 0))
        assert np.all(x >= lo) and np.all(x <= hi)


def test_slq_input_box_behaviour(oracle):
    """Control-limited SLQ on waypoint problems: inputs stay in the box, every instance still converges, the optimum
    costs more than the unconstrained one and less the wider the box; an untouched box is bitwise a no-op."""
    batch, N = 8, 100
    qr = pb.pack_qrqf(*pb.default_weights())
    xs = pb.hover_state(batch); xg = pb.sample_goals(batch, seed=33)
    free = oracle.slq_batch(oracle.slq_opts(0, N), xs, xg, qr, want_traj=True)
    wide = oracle.slq_batch(oracle.slq_opts(0, N, u_min=-1e6, u_max=1e6), xs, xg, qr, want_traj=True)
    assert np.array_equal(free["cost"], wide["cost"]) and np.array_equal(free["u"], wide["u"])
    assert free["u"].min() < 0.0 and free["u"].max() > 6.0          # the unconstrained optimum needs negative thrust
    b6 = oracle.slq_batch(oracle.slq_opts(0, N, u_min=0.0, u_max=6.0), xs, xg, qr, want_traj=True)
    b4 = oracle.slq_batch(oracle.slq_opts(0, N, u_min=0.0, u_max=4.0), xs, xg, qr, want_traj=True)
    for r, hi in ((b6, 6.0), (b4, 4.0)):
        assert r["u"].min() >= 0.0 and r["u"].max() <= hi
        assert np.all(np.isin(r["status"], (0, 2)))
    conv = (free["status"] == 0) & (b6["status"] == 0) & (b4["status"] == 0)
    assert conv.sum() >= batch // 2
    assert np.all(b6["cost"][conv] >= free["cost"][conv] * (1 - 1e-9))
    assert np.all(b4["cost"][conv] >= b6["cost"][conv] * (1 - 1e-6))

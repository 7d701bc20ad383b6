"""The oracle's SLQ against an INDEPENDENT numpy iLQR on the parts VERDICT r1 called thinly pinned: the dual-UAV model
(n=24, m=8) and the input box (control-limited DDP).  Nothing below calls the oracle for arithmetic:

  * dynamics: rotation matrices multiplied out with numpy (dtype-generic, so they also run on complex numbers);
  * Jacobians: complex-step differentiation of those dynamics (exact to rounding, no analytic formulas, no
    `oracle.linearize`);
  * backward pass: textbook matrix algebra with `np.linalg.solve` / Cholesky (no LDL^T, no fma chains);
  * box QP: SciPy's BVLS active-set solver on the Cholesky factor of Quu (not the oracle's projected Newton).

Same specification in a different arithmetic: iteration counts, accepted step-length indices and
status must be IDENTICAL, costs within north_star's 1e-9 relative.  The oracle is only the thing being checked.
"""
import numpy as np
import pytest

from noc_b200 import problems as pb

DT = 0.02


def _quad_f(x, u):
    """Quadrotor of DESIGN.md §2.1 for real or complex x, u."""
    mass, g, L, c = 1.0, 9.81, 0.2, 0.02
    Jd = np.array([0.01, 0.01, 0.02])
    ph, th, ps = x[6], x[7], x[8]
    one, zero = np.ones_like(ph), np.zeros_like(ph)
    Rz = np.array([[np.cos(ps), -np.sin(ps), zero], [np.sin(ps), np.cos(ps), zero], [zero, zero, one]])
    Ry = np.array([[np.cos(th), zero, np.sin(th)], [zero, one, zero], [-np.sin(th), zero, np.cos(th)]])
    Rx = np.array([[one, zero, zero], [zero, np.cos(ph), -np.sin(ph)], [zero, np.sin(ph), np.cos(ph)]])
    Rm = Rz @ Ry @ Rx
    T = u[0] + u[1] + u[2] + u[3]
    tau = np.array([L * (u[1] - u[3]), L * (u[2] - u[0]), c * (u[0] - u[1] + u[2] - u[3])])
    W = np.array([[one, np.sin(ph) * np.tan(th), np.cos(ph) * np.tan(th)],
                  [zero, np.cos(ph), -np.sin(ph)],
                  [zero, np.sin(ph) / np.cos(th), np.cos(ph) / np.cos(th)]])
    w = x[9:12]
    xd = np.empty(12, dtype=np.result_type(x.dtype, u.dtype))
    xd[0:3] = x[3:6]
    xd[3:6] = T / mass * Rm[:, 2]
    xd[5] = xd[5] - g
    xd[6:9] = W @ w
    xd[9:12] = (tau - np.cross(w, Jd * w)) / Jd
    return xd


def _dual_f(x, u):
    ks, cs, d = 2.0, 0.5, np.array([1.0, 0.0, 0.0])
    xd = np.concatenate([_quad_f(x[:12], u[:4]), _quad_f(x[12:], u[4:])])
    cpl = ks * (x[12:15] - x[0:3] - d) + cs * (x[15:18] - x[3:6])
    xd[3:6] = xd[3:6] + cpl
    xd[15:18] = xd[15:18] - cpl
    return xd


def _jac_complex_step(f, x, u):
    """A = I + dt df/dx, B = dt df/du by complex-step differentiation (h = 1e-30: no subtractive cancellation)."""
    n, m, h = x.size, u.size, 1e-30
    A, B = np.empty((n, n)), np.empty((n, m))
    xc, uc = x.astype(complex), u.astype(complex)
    for j in range(n):
        xp = xc.copy(); xp[j] += 1j * h
        A[:, j] = f(xp, uc).imag / h
    for j in range(m):
        up = uc.copy(); up[j] += 1j * h
        B[:, j] = f(xc, up).imag / h
    return np.eye(n) + DT * A, DT * B


def _box_qp_bvls(H, g, lo, hi):
    """min 0.5 k'Hk + g'k, lo <= k <= hi, by SciPy's BVLS on H = L L'."""
    from scipy.optimize import lsq_linear
    L = np.linalg.cholesky(H)
    r = lsq_linear(L.T, -np.linalg.solve(L, g), bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=200)
    # BVLS reports its active set; it may leave an active component one ulp inside the bound, so snap those
    k = np.where(r.active_mask < 0, lo, np.where(r.active_mask > 0, hi, r.x))
    return k, r.active_mask != 0


def _numpy_ilqr(f, x0, xg, uh, Q, R, Qf, N, box=None, tol=1e-8, max_iter=50, n_alpha=11, armijo=1e-4, adaptive=False,
                mu0=0.0):
    """adaptive = the Tassa-style schedule of DESIGN.md §2.4c (mu raised on failed factorisations / line searches,
    lowered after accepted steps); rejected iterations are recorded as -2."""
    n, m = x0.size, uh.size
    mu, delta = mu0, 1.0

    def increase():
        nonlocal mu, delta
        delta = max(2.0, delta * 2.0)
        mu = max(1e-6, mu * delta)

    def decrease():
        nonlocal mu, delta
        delta = min(0.5, delta / 2.0)
        mu = mu * delta if mu * delta >= 1e-6 else 0.0

    step = lambda x, u: x + DT * f(x, u)
    clamp = (lambda u: np.minimum(np.maximum(u, box[0]), box[1])) if box else (lambda u: u)

    def cost(xs, us):
        J = 0.0
        for k in range(N):
            ex, eu = xs[k] - xg, us[k] - uh
            J += 0.5 * (ex @ Q @ ex + eu @ R @ eu)
        ex = xs[N] - xg
        return J + 0.5 * ex @ Qf @ ex

    us = np.tile(uh, (N, 1))
    xs = np.empty((N + 1, n)); xs[0] = x0
    for k in range(N):
        xs[k + 1] = step(xs[k], us[k])
    J = cost(xs, us)
    taken = []
    for it in range(max_iter):
        Vxx, Vx = Qf.copy(), Qf @ (xs[N] - xg)
        Ks, ks = np.zeros((N, m, n)), np.zeros((N, m))
        dV1 = dV2 = 0.0
        for k in range(N - 1, -1, -1):
            A, B = _jac_complex_step(f, xs[k], us[k])
            Qx = Q @ (xs[k] - xg) + A.T @ Vx
            Qu = R @ (us[k] - uh) + B.T @ Vx
            Qxx = Q + A.T @ Vxx @ A
            Qux = B.T @ Vxx @ A
            Quu = R + B.T @ Vxx @ B + mu * np.eye(m)
            np.linalg.cholesky(Quu)                       # positive definite: no factorisation retries on these problems
            kk = -np.linalg.solve(Quu, Qu)
            KK = -np.linalg.solve(Quu, Qux)
            if box is not None:
                lo, hi = box[0] - us[k], box[1] - us[k]
                if not np.all((kk > lo) & (kk < hi)):
                    kk, cl = _box_qp_bvls(Quu, Qu, lo, hi)
                    fr = ~cl
                    KK = np.zeros((m, n))
                    if fr.any():
                        KK[fr] = -np.linalg.solve(Quu[np.ix_(fr, fr)], Qux[fr])
            Ks[k], ks[k] = KK, kk
            Vxx = Qxx + Qux.T @ KK
            Vxx = 0.5 * (Vxx + Vxx.T)
            Vx = Qx + Qux.T @ kk
            dV1 += kk @ Qu
            dV2 += 0.5 * kk @ Quu @ kk
        accepted = None
        for j in range(n_alpha):
            a = 0.5 ** j
            xn = np.empty_like(xs); un = np.empty_like(us); xn[0] = x0
            for k in range(N):
                un[k] = clamp(us[k] + a * ks[k] + Ks[k] @ (xn[k] - xs[k]))
                xn[k + 1] = step(xn[k], un[k])
            Jn = cost(xn, un)
            if J - Jn >= armijo * (-(a * dV1 + a * a * dV2)):
                accepted = j
                break
        if accepted is None:
            if not adaptive:
                return J, it + 1, taken, 3
            increase()
            if mu > 1e6:
                return J, it + 1, taken, 3
            taken.append(-2)
            continue
        taken.append(accepted)
        if adaptive:
            decrease()
        rel = (J - Jn) / max(J, 1e-12)
        xs, us, J = xn, un, Jn
        if rel < tol:
            return J, it + 1, taken, 0
    return J, max_iter, taken, 2


def _compare(oracle, model, N, count, seed, box=None, max_iter=50, adaptive=False, mu0=0.0, pick=None):
    n, m = pb.dims(model)
    Q, R, Qf = pb.default_weights(model)
    qr = pb.pack_qrqf(Q, R, Qf)
    f = _quad_f if model == pb.MODEL_QUAD12 else _dual_f
    xs = pb.hover_state(count, model=model)
    xg = pb.sample_goals(count, seed=seed, model=model)
    if pick is not None:      # a few named instances of a larger sample
        xs, xg, count = xs[pick], xg[pick], len(pick)
    kw = dict(u_min=box[0], u_max=box[1]) if box else {}
    ref = oracle.slq_batch(oracle.slq_opts(model, N, max_iter=max_iter, reg_schedule=int(adaptive), reg0=mu0, **kw), xs, xg, qr,
                           want_traj=False)
    worst = 0.0
    for b in range(count):
        J, iters, alphas, status = _numpy_ilqr(f, xs[b], xg[b], pb.hover_input(model), Q, R, Qf, N, box=box, max_iter=max_iter,
                                               adaptive=adaptive, mu0=mu0)
        assert status == ref["status"][b], (b, status, ref["status"][b])
        assert iters == ref["iters"][b], (b, iters, ref["iters"][b])
        assert alphas == [int(a) for a in ref["alpha_idx"][b][:iters]], b
        rel = abs(J - ref["cost"][b]) / abs(ref["cost"][b])
        worst = max(worst, rel)
        assert rel <= 1e-9, (b, rel)
    return worst, ref


def test_complex_step_jacobians_match_the_oracles_analytic_ones(oracle):
    """The independent Jacobians agree with the oracle's analytic A_k, B_k to rounding, for both models."""
    rng = np.random.default_rng(5)
    for model, f in ((pb.MODEL_QUAD12, _quad_f), (pb.MODEL_DUAL24, _dual_f)):
        n, m = pb.dims(model)
        for _ in range(10):
            x = pb.sample_x0(1, seed=int(rng.integers(1 << 30)), model=model)[0]
            u = pb.hover_input(model) + rng.uniform(-1, 1, m)
            A, B = oracle.linearize(model, x, u, DT)
            Ac, Bc = _jac_complex_step(f, x, u)
            assert np.abs(A - Ac).max() < 1e-14 and np.abs(B - Bc).max() < 1e-14


def test_dual24_slq_vs_independent_numpy_ilqr(oracle):
    """n=24, m=8, N=50, eight waypoint instances: identical iteration counts / step lengths / status, cost to 1e-9."""
    worst, ref = _compare(oracle, pb.MODEL_DUAL24, N=50, count=8, seed=2024, max_iter=30)
    assert ref["iters"].min() >= 3      # real multi-iteration solves, not trivial ones


def test_box_slq_vs_independent_numpy_ilqr(oracle):
    """Control-limited SLQ (0 <= u <= 4 N per rotor), n=12, N=60, eight instances, BVLS as the independent QP."""
    worst, ref = _compare(oracle, pb.MODEL_QUAD12, N=60, count=8, seed=77, box=(0.0, 4.0), max_iter=40)
    assert ref["iters"].min() >= 3


def test_box_slq_dual24_vs_independent_numpy_ilqr(oracle):
    """The 8x8 box QP of the dual model inside SLQ: four instances, N=50."""
    _compare(oracle, pb.MODEL_DUAL24, N=50, count=4, seed=99, box=(0.0, 4.0), max_iter=25)


def test_survey_sampler_states_slq_vs_independent_numpy_ilqr(oracle):
    """SURVEY.md §8d's literal sampler (std::mt19937_64 per instance, body rates U(-0.5,0.5)) as start states, goal =
    origin hover: the harder second test set."""
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    N, count = 50, 8
    x0 = pb.sample_x0_survey(count, first_instance=1000)
    xg = np.zeros((count, 12))
    ref = oracle.slq_batch(oracle.slq_opts(0, N, max_iter=30), x0, xg, qr, want_traj=False)
    for b in range(count):
        J, iters, alphas, status = _numpy_ilqr(_quad_f, x0[b], xg[b], pb.hover_input(), Q, R, Qf, N, max_iter=30)
        assert (status, iters) == (ref["status"][b], ref["iters"][b]), b
        assert alphas == [int(a) for a in ref["alpha_idx"][b][:iters]], b
        assert abs(J - ref["cost"][b]) <= 1e-9 * abs(ref["cost"][b]), b


def test_adaptive_regularisation_vs_independent_numpy_ilqr(oracle):
    """DESIGN.md §2.4c (NOC_REG_ADAPTIVE) restated in numpy: started from mu0 = 1 the schedule has to walk mu down
    (1 -> 0.5 -> 0.125 -> ... -> 0) over the accepted steps; under the tight input box some line searches fail and are
    repeated with a larger mu (-2 entries).  Same counts / step indices / status, cost to 1e-9."""
    worst, ref = _compare(oracle, pb.MODEL_QUAD12, N=60, count=6, seed=5150, adaptive=True, mu0=1.0, max_iter=40)
    fixed = oracle.slq_batch(oracle.slq_opts(0, 60, max_iter=40, reg0=1.0), pb.hover_state(6), pb.sample_goals(6, seed=5150),
                             pb.pack_qrqf(*pb.default_weights()), want_traj=False)
    assert ref["iters"].sum() < fixed["iters"].sum()      # a mu that never comes down needs more iterations
    # instances of a boxed waypoint sample in which a line search fails and is repeated with a larger mu
    _, ref = _compare(oracle, pb.MODEL_QUAD12, N=60, count=512, seed=33, box=(0.0, 4.0), adaptive=True, max_iter=50,
                      pick=[284, 217, 252])
    assert (ref["alpha_idx"] == -2).any(axis=1).all() and np.all(ref["status"] == 0)

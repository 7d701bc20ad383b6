"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the scalar oracle is too
slow for 65,536 x 100 or 1,048,576 x 50): determinism, symmetry / definiteness of P0, the cost identity
0.5 x0' P0 x0 == closed-loop cost of the LTV system under the returned gains, and a strided sample against the
oracle, bit-exact."""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    from noc_b200 import capi
    s = capi.Solver(0)
    yield s
    s.close()


def test_config2_full_size_ltv_sweep(oracle, solver):
    import torch
    from noc_b200 import capi
    B, N = 65536, 100
    dev = torch.device("cuda", 0)
    Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(B, seed=2026)
    f64 = dict(dtype=torch.float64, device=dev)
    d_x0 = torch.from_numpy(x0).to(dev); d_qr = torch.from_numpy(qr).to(dev)
    d_xbar = torch.empty(B, N + 1, 12, **f64); d_AB = torch.empty(B, N, 192, **f64)
    pm = capi.problem(0, N)
    solver.rollout_open_loop(pm, B, d_x0, None, d_xbar)
    solver.linearize_batch(pm, B, d_xbar, None, d_AB)
    pl = capi.problem(capi.MODEL_LTV_USER, N)
    outs = []
    for _ in range(2):
        K = torch.empty(B, N, 4, 12, **f64); P0 = torch.empty(B, 12, 12, **f64); cost = torch.empty(B, **f64)
        st = torch.full((B,), -1, dtype=torch.int32, device=dev)
        solver.lqr_solve_batch(pl, B, d_AB, d_qr, x0dev=d_x0, K=K, P0=P0, cost=cost, status=st)
        outs.append((K, P0, cost, st))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))          # deterministic
    K, P0, cost, st = outs[0]
    assert int((st != 0).sum()) == 0 and bool(torch.isfinite(K).all())
    assert torch.equal(P0, P0.transpose(1, 2))                               # bitwise symmetric
    assert bool((torch.linalg.eigvalsh(P0[::257]) > 0).all())
    # cost identity on a strided sample: roll the LTV system forward under u = K x and add up the stage costs
    sel = torch.arange(0, B, 4099, device=dev)
    A = d_AB[sel, :, :144].reshape(-1, N, 12, 12); Bm = d_AB[sel, :, 144:].reshape(-1, N, 12, 4)
    Qt, Rt, Qft = (torch.from_numpy(M).to(dev) for M in (Q, R, Qf))
    x = d_x0[sel].clone(); J = torch.zeros(len(sel), **f64)
    for k in range(N):
        u = torch.einsum("bij,bj->bi", K[sel, k], x)
        J += 0.5 * (torch.einsum("bi,ij,bj->b", x, Qt, x) + torch.einsum("bi,ij,bj->b", u, Rt, u))
        x = torch.einsum("bij,bj->bi", A[:, k], x) + torch.einsum("bij,bj->bi", Bm[:, k], u)
    J += 0.5 * torch.einsum("bi,ij,bj->b", x, Qft, x)
    assert float(((J - cost[sel]).abs() / cost[sel]).max()) < 1e-10
    # strided sample against the oracle, bit-exact
    s2 = np.arange(0, B, 8191)
    ref = oracle.lqr_batch(12, 4, N, d_AB[torch.from_numpy(s2).to(dev)].cpu().numpy(), qr, x0dev=x0[s2])
    assert np.array_equal(K[torch.from_numpy(s2).to(dev)].cpu().numpy(), ref["K"])
    assert np.array_equal(cost[torch.from_numpy(s2).to(dev)].cpu().numpy(), ref["cost"])


def test_config4_full_size_mpc_batch(oracle, solver):
    import torch
    from noc_b200 import capi
    B, N = 1 << 20, 50
    dev = torch.device("cuda", 0)
    qr = pb.pack_qrqf(*pb.default_weights())
    x0 = pb.sample_x0(B, seed=4)
    d_x0 = torch.from_numpy(x0).to(dev); d_qr = torch.from_numpy(qr).to(dev)
    K0 = torch.empty(B, 4, 12, dtype=torch.float64, device=dev); cost = torch.empty(B, dtype=torch.float64, device=dev)
    st = torch.full((B,), -1, dtype=torch.int32, device=dev)
    solver.lqr_solve_from_x0(capi.problem(0, N), B, d_x0, d_qr, K0=K0, cost=cost, status=st)
    assert int((st != 0).sum()) == 0 and bool(torch.isfinite(K0).all()) and bool((cost > 0).all())
    idx, val = solver.argmin_cost(cost, B)
    c = cost.cpu().numpy()
    assert idx == int(np.argmin(c)) and val == c.min()
    sel = np.arange(0, B, 65521)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0[sel], qr, want_P0=False)
    assert np.array_equal(K0.cpu().numpy()[sel], ref["K0"]) and np.array_equal(c[sel], ref["cost"])


def test_config3_size_slq_counts_on_a_sample(oracle, solver):
    """Config 3 shape at a quarter of its batch (4,096 x N=200): statuses sane, costs decrease, and the first 48
    instances reproduce the oracle's iteration counts, step indices and costs exactly."""
    from noc_b200 import capi
    B, N = 4096, 200
    qr = pb.pack_qrqf(*pb.default_weights())
    xs = pb.hover_state(B); xg = pb.sample_goals(B, seed=33)
    prob = capi.problem(0, N)
    cost = np.empty(B); iters = np.zeros(B, dtype=np.int32); st = np.full(B, -1, dtype=np.int32)
    aidx = np.zeros((B, prob.max_iter), dtype=np.int32)
    solver.slq_solve_batch(prob, B, xs, xg, qr, cost=cost, iters=iters, alpha_idx=aidx, status=st)
    assert set(np.unique(st)).issubset({0, 2, 3}) and (st == 0).mean() > 0.95
    assert np.all(iters >= 1) and np.all(iters <= prob.max_iter) and np.all(np.isfinite(cost))
    m = 48
    ref = oracle.slq_batch(oracle.slq_opts(0, N), xs[:m], xg[:m], qr, want_traj=False)
    assert np.array_equal(iters[:m], ref["iters"]) and np.array_equal(aidx[:m], ref["alpha_idx"])
    assert np.array_equal(cost[:m], ref["cost"]) and np.array_equal(st[:m], ref["status"])


@pytest.mark.parametrize("box", [dict(), dict(u_min=0.0, u_max=4.0)])
def test_config5_size_slq_counts_on_a_sample(oracle, solver, box):
    """Config 5 shape (dual UAV, n=24, m=8, N=100) at a quarter of its batch, without and with the input box: statuses
    sane, and the first 16 instances reproduce the oracle's iteration counts, step indices and costs exactly."""
    from noc_b200 import capi
    B, N, M = 2048, 100, 1
    qr = pb.pack_qrqf(*pb.default_weights(M))
    xs = pb.hover_state(B, model=M); xg = pb.sample_goals(B, seed=35, model=M)
    prob = capi.problem(M, N, **box)
    cost = np.empty(B); iters = np.zeros(B, dtype=np.int32); st = np.full(B, -1, dtype=np.int32)
    aidx = np.zeros((B, prob.max_iter), dtype=np.int32)
    u = np.empty((B, N, 8))
    solver.slq_solve_batch(prob, B, xs, xg, qr, u=u, cost=cost, iters=iters, alpha_idx=aidx, status=st)
    assert set(np.unique(st)).issubset({0, 2, 3}) and (st == 0).mean() > 0.8
    assert np.all(iters >= 1) and np.all(iters <= prob.max_iter) and np.all(np.isfinite(cost))
    if box:
        assert u.min() >= box["u_min"] and u.max() <= box["u_max"]
    m = 16
    ref = oracle.slq_batch(oracle.slq_opts(M, N, **box), xs[:m], xg[:m], qr, want_traj=False)
    assert np.array_equal(iters[:m], ref["iters"]) and np.array_equal(aidx[:m], ref["alpha_idx"])
    assert np.array_equal(cost[:m], ref["cost"]) and np.array_equal(st[:m], ref["status"])

"""C-ABI contract on a GPU: error codes instead of crashes, empty and degenerate shapes, optional outputs,
host/device pointer mixing, determinism.  (errors are return codes, per-instance outcomes go to
status[].)"""
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


def _qr():
    return pb.pack_qrqf(*pb.default_weights())


def test_empty_batch_is_ok(solver):
    from noc_b200 import capi
    e = np.empty(0)
    solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, 5), 0, np.empty((0, 5, 192)), _qr(), K=np.empty((0, 5, 4, 12)))
    solver.lqr_solve_from_x0(capi.problem(0, 5), 0, np.empty((0, 12)), _qr(), cost=e)
    solver.slq_solve_batch(capi.problem(0, 5), 0, np.empty((0, 12)), np.empty((0, 12)), _qr(), cost=e)
    solver.dare_solve_batch(capi.problem(capi.MODEL_LTV_USER, 1), 0, np.empty((0, 192)), _qr()[:160], np.empty((0, 144)),
                            np.empty((0, 48)), np.empty(0, dtype=np.int32), np.empty(0, dtype=np.int32))


def test_bad_arguments_return_codes(solver):
    from noc_b200 import capi
    x0 = pb.sample_x0(4, seed=1)
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_from_x0(capi.problem(0, 5), 4, None, _qr())
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_from_x0(capi.problem(0, 0), 4, x0, _qr())            # horizon 0
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_from_x0(capi.problem(0, 5, n=10), 4, x0, _qr())       # model / dims mismatch
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, 5, n=30, m=2), 4, np.zeros((4, 5, 960)), np.zeros(1804))
    assert e.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.NocError) as e:   # the DARE path has the same bounds (thread-local arrays of 24 x 24)
        solver.dare_solve_batch(capi.problem(capi.MODEL_LTV_USER, 1, n=25, m=9), 2, np.zeros((2, 850)), np.zeros(706),
                                np.zeros((2, 625)), np.zeros((2, 225)), np.zeros(2, dtype=np.int32), np.zeros(2, dtype=np.int32))
    assert e.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.NocError) as e:
        solver.dare_solve_batch(capi.problem(capi.MODEL_LTV_USER, 1, n=0, m=1), 2, np.zeros(8), np.zeros(8),
                                np.zeros(8), np.zeros(8), np.zeros(2, dtype=np.int32), np.zeros(2, dtype=np.int32))
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.slq_solve_batch(capi.problem(0, 5, n_alpha=40), 4, x0, x0, _qr())
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.slq_solve_batch(capi.problem(0, 5, u_min=3.0, u_max=2.0), 4, x0, x0, _qr())          # empty input box
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.slq_solve_batch(capi.problem(capi.MODEL_LTV_USER, 5), 4, x0, x0, _qr())
    assert e.value.code == capi.ERR_BAD_ARG
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, 5, layout=7), 4, np.zeros((4, 5, 192)), _qr())
    assert e.value.code == capi.ERR_BAD_ARG
    # the context is still usable after errors
    cost = np.empty(4)
    solver.lqr_solve_from_x0(capi.problem(0, 5), 4, x0, _qr(), cost=cost)
    assert np.all(np.isfinite(cost))


def test_misaligned_device_records_are_rejected(solver):
    import torch
    from noc_b200 import capi
    buf = torch.zeros(4 * 5 * 192 + 1, dtype=torch.float64, device="cuda")
    with pytest.raises(capi.NocError) as e:
        solver.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, 5), 4, buf[1:], torch.from_numpy(_qr()).cuda(),
                               K0=torch.empty(4, 48, dtype=torch.float64, device="cuda"))
    assert e.value.code == capi.ERR_BAD_ARG


def test_horizon_one_and_single_instance(oracle, solver):
    from noc_b200 import capi
    qr = _qr()
    x0 = pb.sample_x0(1, seed=2)
    ref = oracle.lqr_from_x0_batch(0, 1, 0.02, x0, qr)
    K0 = np.empty((1, 4, 12)); P0 = np.empty((1, 12, 12)); cost = np.empty(1)
    solver.lqr_solve_from_x0(capi.problem(0, 1), 1, x0, qr, K0=K0, P0=P0, cost=cost)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(cost, ref["cost"])
    xg = pb.sample_goals(3, seed=3)
    xs = pb.hover_state(3)
    for N in (1, 2):
        c = np.empty(3); it = np.zeros(3, dtype=np.int32); st = np.zeros(3, dtype=np.int32)
        solver.slq_solve_batch(capi.problem(0, N), 3, xs, xg, qr, cost=c, iters=it, status=st)
        r = oracle.slq_batch(oracle.slq_opts(0, N), xs, xg, qr, want_traj=False)
        assert np.array_equal(c, r["cost"]) and np.array_equal(it, r["iters"]) and np.array_equal(st, r["status"])


def test_mixed_host_and_device_pointers_and_async(oracle, solver):
    import torch
    from noc_b200 import capi
    batch, N = 40, 16
    qr = _qr()
    x0 = pb.sample_x0(batch, seed=4)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr)
    d_x0 = torch.from_numpy(x0).cuda()
    K0 = np.empty((batch, 4, 12))                      # host output
    d_cost = torch.empty(batch, dtype=torch.float64, device="cuda")   # device output
    solver.lqr_solve_from_x0(capi.problem(0, N), batch, d_x0, qr, K0=K0, cost=d_cost)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(d_cost.cpu().numpy(), ref["cost"])
    # async flag with device pointers only: returns before completion, noc_synchronize() finishes it
    d_K0 = torch.empty(batch, 4, 12, dtype=torch.float64, device="cuda")
    d_qr = torch.from_numpy(qr).cuda()
    solver.lqr_solve_from_x0(capi.problem(0, N, flags=capi.FLAG_ASYNC), batch, d_x0, d_qr, K0=d_K0, cost=d_cost)
    solver.synchronize()
    assert np.array_equal(d_K0.cpu().numpy(), ref["K0"])


def test_per_kernel_profile_counters(solver):
    from noc_b200 import capi
    x0 = pb.sample_x0(640, seed=5)
    solver.profile_reset(); solver.profile_enable(True)
    n0 = solver.launch_count
    solver.lqr_solve_from_x0(capi.problem(0, 10), 640, x0, _qr(), cost=np.empty(640))
    solver.profile_enable(False)
    ms, cnt = solver.profile_read(capi.KERNEL_RICCATI)
    assert cnt == 1 and ms > 0 and solver.launch_count - n0 == 3      # rollout, weight packing, fused sweep
    assert solver.profile_read(capi.KERNEL_LINEARIZE)[1] == 0          # linearisation is fused into the sweep
    # at most two instances per SM: the small-batch path (time-parallel linearisation + one instance per warp)
    solver.profile_reset(); solver.profile_enable(True)
    n0 = solver.launch_count
    solver.lqr_solve_from_x0(capi.problem(0, 10), 64, x0[:64], _qr(), cost=np.empty(64))
    solver.profile_enable(False)
    assert solver.launch_count - n0 == 4 and solver.profile_read(capi.KERNEL_LINEARIZE)[1] == 1


def test_outer_chunking_of_the_nominal_scratch(oracle):
    """NOC_OPT_SCRATCH_LIMIT_BYTES caps the scratch budget so that a modest batch is cut into several outer chunks (each
    with its own rollout launch) and, with host outputs, inner copy chunks: results must not depend on the cut."""
    from noc_b200 import capi
    batch, N = 9000, 40
    qr = _qr()
    x0 = pb.sample_x0(batch, seed=6)
    s = capi.Solver(0)
    s.set_option(capi.OPT_SCRATCH_LIMIT_BYTES, 8 << 20)   # 8 MB / (41 x 96 B) = ~2,100 instances per outer chunk
    K0 = np.empty((batch, 4, 12)); cost = np.empty(batch); st = np.full(batch, -1, dtype=np.int32)
    s.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K0=K0, cost=cost, status=st)
    s.close()
    sel = np.arange(0, batch, 97)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0[sel], qr, want_P0=False)
    assert np.all(st == 0)
    assert np.array_equal(K0[sel], ref["K0"]) and np.array_equal(cost[sel], ref["cost"])
    s2 = capi.Solver(0)
    K0b = np.empty((batch, 4, 12)); costb = np.empty(batch)
    s2.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K0=K0b, cost=costb)
    s2.close()
    assert np.array_equal(K0, K0b) and np.array_equal(cost, costb)


def test_inner_chunks_on_two_streams_do_not_change_results(oracle):
    """Host outputs and batch >= 8192: the batch is cut into one-wave inner chunks that alternate between two streams
    while a third copies finished chunks back.  Same bits as the single-stream schedule, as two-wave chunks, and as
    the oracle (noc_set_option: NOC_OPT_SINGLE_STREAM / NOC_OPT_CHUNK_WAVES)."""
    from noc_b200 import capi
    batch, N = 24000, 20
    qr = _qr()
    x0 = pb.sample_x0(batch, seed=8)
    outs = []
    for opts in ({}, {capi.OPT_SINGLE_STREAM: 1}, {capi.OPT_CHUNK_WAVES: 2}):
        s = capi.Solver(0)
        for k, v in opts.items():
            s.set_option(k, v)
        K0 = np.empty((batch, 4, 12)); P0 = np.empty((batch, 12, 12)); cost = np.empty(batch)
        st = np.full(batch, -1, dtype=np.int32)
        for _ in range(2):   # the second call reuses scratch, events and both streams
            s.lqr_solve_from_x0(capi.problem(0, N), batch, x0, qr, K0=K0, P0=P0, cost=cost, status=st)
        s.close()
        assert np.all(st == 0)
        outs.append((K0, P0, cost))
    for K0, P0, cost in outs[1:]:
        assert np.array_equal(K0, outs[0][0]) and np.array_equal(P0, outs[0][1]) and np.array_equal(cost, outs[0][2])
    sel = np.arange(0, batch, 211)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0[sel], qr)
    assert np.array_equal(outs[0][0][sel], ref["K0"]) and np.array_equal(outs[0][1][sel], ref["P0"])
    assert np.array_equal(outs[0][2][sel], ref["cost"])


def test_failed_call_leaves_no_pending_copies(oracle):
    """a call that fails AFTER staging a host output must not leave its copy-back queued — the next call
    would write the old byte count from reused device buffers into the previous call's (possibly freed) host array."""
    import torch
    from noc_b200 import capi
    s = capi.Solver(0)
    qr = _qr()
    N = 12
    x0 = pb.sample_x0(300, seed=11)
    stale = np.full((300, N + 1, 12), 7.0)                  # host xbar output of the failing call
    bad_K0 = torch.zeros(300 * 48 + 1, dtype=torch.float64, device="cuda")[1:]   # misaligned device output -> BAD_ARG
    with pytest.raises(capi.NocError) as e:
        s.lqr_solve_along(capi.problem(0, N), 300, x0, None, qr, xbar=stale, K0=bad_K0)
    assert e.value.code == capi.ERR_BAD_ARG
    # a smaller, valid call on the same context: only its own outputs may be written
    x1 = pb.sample_x0(40, seed=12)
    K0 = np.empty((40, 4, 12)); cost = np.empty(40)
    s.lqr_solve_from_x0(capi.problem(0, N), 40, x1, qr, K0=K0, cost=cost)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, x1, qr, want_P0=False)
    assert np.array_equal(K0, ref["K0"]) and np.array_equal(cost, ref["cost"])
    assert np.all(stale == 7.0)                             # the failed call's array was never touched
    s.close()


def test_host_AB_stream_goes_through_the_chunked_pipeline(oracle):
    """noc_lqr_solve_batch with a HOST A_k,B_k stream above 32 MB: two-buffer copy-in / sweep / copy-out pipeline,
    chunked to the scratch budget (here capped so that there are many chunks and a ragged tail).  Same bits as the
    oracle, with host, device and absent outputs mixed, shared and per-instance weights."""
    import torch
    from noc_b200 import capi
    s = capi.Solver(0)
    s.set_option(capi.OPT_SCRATCH_LIMIT_BYTES, 24 << 20)
    qr = _qr()
    batch, N = 3001, 9                                       # 3001 x 9 x 1536 B = 41 MB
    x0 = pb.sample_x0(batch, seed=13)
    AB = oracle.make_ab_batch(0, N, 0.02, x0)
    ref = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    K = np.empty((batch, N, 4, 12)); P0 = np.empty((batch, 12, 12)); st = np.full(batch, -1, dtype=np.int32)
    d_cost = torch.empty(batch, dtype=torch.float64, device="cuda")
    s.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N), batch, AB, qr, x0dev=x0, K=K, P0=P0, cost=d_cost, status=st)
    assert np.all(st == 0)
    assert np.array_equal(K, ref["K"]) and np.array_equal(P0, ref["P0"]) and np.array_equal(d_cost.cpu().numpy(), ref["cost"])
    # per-instance weights (thread-per-instance kernel) through the same pipeline
    qrs = np.tile(qr, (batch, 1)); qrs[:, :144] *= np.linspace(0.5, 2.0, batch)[:, None]
    K0 = np.empty((batch, 4, 12)); cost = np.empty(batch)
    s.lqr_solve_batch(capi.problem(capi.MODEL_LTV_USER, N), batch, AB, qrs, x0dev=x0, K0=K0, cost=cost, qr_per_instance=True)
    sel = np.arange(0, batch, 250)
    for b in sel:
        r = oracle.lqr_batch(12, 4, N, AB[b:b + 1], qrs[b], x0dev=x0[b:b + 1])
        assert np.array_equal(K0[b], r["K0"][0]) and cost[b] == r["cost"][0]
    s.close()


def test_slq_outer_chunks_do_not_change_results(oracle):
    """SLQ batches larger than the scratch budget are solved slice by slice (noc_slq_solve_batch used to allocate the
    whole working set or fail with NOC_ERR_OOM)."""
    from noc_b200 import capi
    batch, N = 150, 30
    qr = _qr()
    xg = pb.sample_goals(batch, seed=14); xs = pb.hover_state(batch)
    outs = []
    for cap in (0, 2 << 20):                                 # 2 MB / ~21 KB per instance = ~96 instances per chunk
        s = capi.Solver(0)
        s.set_option(capi.OPT_SCRATCH_LIMIT_BYTES, cap)
        x = np.empty((batch, N + 1, 12)); u = np.empty((batch, N, 4)); c = np.empty(batch)
        it = np.zeros(batch, dtype=np.int32); st = np.zeros(batch, dtype=np.int32); ai = np.zeros((batch, 50), dtype=np.int32)
        s.slq_solve_batch(capi.problem(0, N), batch, xs, xg, qr, x=x, u=u, cost=c, iters=it, alpha_idx=ai, status=st)
        s.close()
        outs.append((x, u, c, it, st, ai))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    r = oracle.slq_batch(oracle.slq_opts(0, N), xs, xg, qr)
    assert np.array_equal(outs[1][2], r["cost"]) and np.array_equal(outs[1][3], r["iters"])

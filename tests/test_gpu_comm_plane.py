"""The in-library multi-GPU plane (noc_comm_* in include/lqr_control/noc_b200.h) on ONE GPU: an NCCL
communicator of size 1 exercises the whole path — dlopen'd NCCL, chunked sweep with per-chunk gathers on the side
stream, multi-CTA local argmin, gather of (cost, index) pairs, final pick — and must reproduce the oracle bit for
bit.  The 2-rank version of the same checks is tests/cpp/facade_test.cpp (two devices, when the box has them) and
tools/comm_check.py under torchrun."""
import numpy as np
import pytest

from noc_b200 import problems as pb

pytestmark = pytest.mark.gpu


def _qr():
    return pb.pack_qrqf(*pb.default_weights())


def test_argmin_async_multi_cta_matches_numpy():
    import torch
    from noc_b200 import capi
    s = capi.Solver(0)
    rng = np.random.default_rng(3)
    for n in (1, 31, 4097, 1 << 20):
        c = rng.uniform(1.0, 2.0, n)
        c[rng.integers(0, n, max(1, n // 7))] = np.nan               # failed instances are skipped
        if n > 100:
            c[[n // 3, n // 2]] = 0.5                                # a tie: the smaller index wins
        d = torch.from_numpy(c).cuda()
        best = torch.zeros(2, dtype=torch.float64, device="cuda")
        s.argmin_cost_async(d, n, best, index_offset=1000)
        s.synchronize()
        cost = float(best[0].item()); idx = int(best.view(torch.int64)[1].item())
        fin = np.where(np.isfinite(c), c, np.inf)
        if np.isfinite(fin.min()):
            assert idx == 1000 + int(np.argmin(fin)) and cost == fin.min()
        else:
            assert idx == -1 and np.isnan(cost)
        i2, v2 = s.argmin_cost(d, n)                                  # the synchronous entry point, same kernel
        assert (i2 + 1000 == idx and v2 == cost) or (i2 == -1 and idx == -1)
    allnan = torch.full((777,), float("nan"), dtype=torch.float64, device="cuda")
    assert s.argmin_cost(allnan, 777)[0] == -1
    s.close()


def test_comm_plane_single_rank_bitexact(oracle):
    import torch
    from noc_b200 import capi
    s = capi.Solver(0)
    comm = capi.Comm(s, capi.comm_unique_id(), 0, 1)
    assert (comm.rank, comm.nranks) == (0, 1)
    qr = _qr()
    for batch, N in ((200, 12), (20000, 6)):                          # one launch; several wave-sized chunks
        x0 = pb.sample_x0(batch, seed=21)
        ref = oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr, want_P0=False)
        d_x0 = torch.from_numpy(x0).cuda(); d_qr = torch.from_numpy(qr).cuda()
        cost_all = torch.empty(batch, dtype=torch.float64, device="cuda")
        K0_all = torch.empty(batch, 4, 12, dtype=torch.float64, device="cuda")
        st = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
        best = torch.zeros(2, dtype=torch.float64, device="cuda")
        idx, val = comm.lqr_solve_from_x0(capi.problem(0, N), batch, d_x0, d_qr, best, cost_all=cost_all, K0_all=K0_all,
                                          status=st, want_host=True)
        assert np.array_equal(cost_all.cpu().numpy(), ref["cost"]) and np.array_equal(K0_all.cpu().numpy(), ref["K0"])
        assert int((st != 0).sum()) == 0
        assert idx == int(np.argmin(ref["cost"])) and val == ref["cost"].min()
        # device-resident result, no host pointer: read it back separately
        best.zero_()
        comm.lqr_solve_from_x0(capi.problem(0, N, flags=capi.FLAG_ASYNC), batch, d_x0, d_qr, best)
        s.synchronize()
        assert float(best[0].item()) == val and int(best.view(torch.int64)[1].item()) == idx
        # the building blocks on their own
        g = torch.empty_like(cost_all)
        comm.allgather(cost_all, g, batch, 8)
        i2, v2 = comm.best_rollout(cost_all, batch, 0, best, want_host=True)
        assert torch.equal(g, cost_all) and (i2, v2) == (idx, val)
    comm.close()
    s.close()

"""Randomised parity sweep of the SLQ paths against the CPU oracle (bit-exact: status, iteration counts, accepted step
indices, costs, trajectories, gains): random models, horizons, batch sizes, input boxes, regularisation schedules, indefinite
R (REG_FAIL / NOT_PD paths), aggressive starts (backtracking, failed line searches), kernel families and schedules.
    python tools/fuzz_slq.py [cases] [seed]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noc_b200 import capi, problems as pb
from oracle import pyoracle as oracle

def same_bits(a, b):
    """float64 arrays: equal bit patterns (so the sign of an exact zero counts); NaNs in the same places count as equal"""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and bool(np.all(na == nb)) and bool(np.all((a.view(np.int64) == b.view(np.int64)) | na))


def run(cases=60, seed=1, verbose=True):
    rng = np.random.default_rng(seed)
    s = capi.Solver(0)
    bad = 0
    seen = {"status": {}, "backtracked": 0, "ls_repeat": 0, "boxed": 0, "adaptive": 0, "instances": 0}
    t0 = time.time()
    for c in range(cases):
        model = int(rng.integers(0, 2))
        n, m = pb.dims(model)
        N = int(rng.choice([1, 2, 5, 17, 40, 64]))
        batch = int(rng.choice([1, 3, 9, 33, 70]))
        Q, R, Qf = pb.default_weights(model)
        kw = {}
        if rng.random() < 0.4:
            lo = float(rng.choice([0.0, 1.0, 1.8])); kw.update(u_min=lo, u_max=lo + float(rng.choice([0.9, 2.0, 4.0])))
        if rng.random() < 0.4:
            kw.update(reg_schedule=capi.REG_ADAPTIVE if hasattr(capi, "REG_ADAPTIVE") else 1, reg0=float(rng.choice([0.0, 1.0])))
        if rng.random() < 0.2:
            R = R.copy(); R[0, 0] = -float(rng.choice([0.05, 5.0]))          # indefinite R: regularisation retries / REG_FAIL
        kw.update(max_iter=int(rng.choice([3, 8, 15])), n_alpha=int(rng.choice([1, 4, 11])))
        qr = pb.pack_qrqf(Q, R, Qf)
        x0 = pb.sample_x0(batch, seed=int(rng.integers(1 << 30)), model=model)
        if rng.random() < 0.5:
            for v in range(n // 12):
                x0[:, 12 * v + 6:12 * v + 8] = rng.uniform(-1.0, 1.0, (batch, 2))
        xg = pb.sample_goals(batch, seed=int(rng.integers(1 << 30)), model=model) * float(rng.choice([1.0, 2.5]))
        opts = tuple(int(v) for v in rng.integers(0, 2, 4))
        for o, val in zip((capi.OPT_NO_WAVE_PLAN, capi.OPT_DEFERRED_BACKTRACK, capi.OPT_NO_TENSOR_SWEEP, capi.OPT_NO_SPECULATIVE_BACKTRACK), opts):
            s.set_option(o, val)
        prob = capi.problem(model, N, **kw)
        x = np.empty((batch, N + 1, n)); u = np.empty((batch, N, m)); K = np.empty((batch, N, m, n))
        cost = np.empty(batch); iters = np.zeros(batch, dtype=np.int32); status = np.full(batch, -1, dtype=np.int32)
        aidx = np.full((batch, prob.max_iter), -7, dtype=np.int32)
        s.slq_solve_batch(prob, batch, x0, xg, qr, x=x, u=u, K=K, cost=cost, iters=iters, alpha_idx=aidx, status=status)
        o = oracle.slq_opts(model, N, dt=prob.dt, tol=prob.tol, max_iter=prob.max_iter, n_alpha=prob.n_alpha, reg0=prob.reg0,
                            armijo=prob.armijo, u_min=prob.u_min, u_max=prob.u_max, reg_schedule=prob.reg_schedule)
        ref = oracle.slq_batch(o, x0, xg, qr, want_traj=True, want_K=True)
        ok_inst = ref["status"] != 4      # (REG_FAIL leaves its gains unspecified)
        same = (np.array_equal(status, ref["status"]) and np.array_equal(iters, ref["iters"]) and np.array_equal(aidx, ref["alpha_idx"])
                and same_bits(cost, ref["cost"]) and same_bits(x, ref["x"]) and same_bits(u, ref["u"])
                and same_bits(K[ok_inst], ref["K"][ok_inst]))
        for k_, v_ in zip(*np.unique(ref["status"], return_counts=True)): seen["status"][int(k_)] = seen["status"].get(int(k_), 0) + int(v_)
        seen["backtracked"] += int((ref["alpha_idx"] > 0).any(axis=1).sum()); seen["ls_repeat"] += int((ref["alpha_idx"] == -2).any(axis=1).sum())
        seen["boxed"] += batch * int("u_min" in kw); seen["adaptive"] += batch * int("reg_schedule" in kw); seen["instances"] += batch
        if not same:
            bad += 1
            print(json.dumps({"case": c, "model": model, "N": N, "batch": batch, "kw": {k: float(v) for k, v in kw.items()}, "opts": opts,
                              "status": np.unique(status).tolist(), "ref_status": np.unique(ref["status"]).tolist(),
                              "iters_eq": bool(np.array_equal(iters, ref["iters"])), "aidx_eq": bool(np.array_equal(aidx, ref["alpha_idx"])),
                              "cost_bits": same_bits(cost, ref["cost"]), "x_bits": same_bits(x, ref["x"]), "u_bits": same_bits(u, ref["u"]),
                              "K_bits": same_bits(K[ok_inst], ref["K"][ok_inst]),
                              "K_values": bool(np.array_equal(K[ok_inst], ref["K"][ok_inst], equal_nan=True))}), flush=True)
    s.close()
    out = {"cases": cases, "mismatches": bad, "seconds": round(time.time() - t0, 1), "covered": seen}
    if verbose: print(json.dumps(out))
    return out


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 1)

#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle (DESIGN.md §2 frozen spec).

The reference (KahnShi/NOC) holds no golden vectors (parity unpinned), so these fixtures freeze
THIS repo's specification: any later change of the arithmetic order shows up as a golden mismatch in both the
CPU suite (oracle vs golden) and the GPU suite (CUDA vs golden).  Re-run only when DESIGN.md §2 changes:
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noc_b200 import problems as pb  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    oracle.build()
    os.makedirs(OUT, exist_ok=True)
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)

    # finite-horizon LTV LQR (configs 2 / 4 shape, small)
    batch, N = 5, 25
    x0 = pb.sample_x0(batch, seed=20261019)
    AB = oracle.make_ab_batch(0, N, pb.DT, x0)
    r = oracle.lqr_batch(12, 4, N, AB, qr, x0dev=x0)
    np.savez_compressed(os.path.join(OUT, "lqr_n12_N25.npz"), x0=x0, qrqf=qr, N=N, K=r["K"], K0=r["K0"], P0=r["P0"],
                        cost=r["cost"], status=r["status"], AB_first=AB[:, 0], AB_last=AB[:, -1])

    # infinite-horizon LQR at exact hover (config 1)
    A, B = oracle.linearize(0, np.zeros(12), pb.hover_input(), pb.DT)
    P, K, it = oracle.dare(A, B, Q, R, tol=1e-12, max_iter=10000)
    np.savez_compressed(os.path.join(OUT, "dare_hover.npz"), A=A, B=B, Q=Q, R=R, P=P, K=K, iters=it)

    # SLQ to waypoints (config 3 shape, small)
    batch, N = 4, 40
    xs = pb.hover_state(batch)
    xg = pb.sample_goals(batch, seed=20261020)
    o = oracle.slq_opts(0, N)
    s = oracle.slq_batch(o, xs, xg, qr, want_traj=True, want_K=False)
    np.savez_compressed(os.path.join(OUT, "slq_n12_N40.npz"), x0=xs, xgoal=xg, qrqf=qr, N=N, cost=s["cost"],
                        iters=s["iters"], alpha_idx=s["alpha_idx"], status=s["status"], x_final=s["x"][:, -1],
                        u_first=s["u"][:, 0], x=s["x"], u=s["u"])

    # the same problems with the input box 0 <= u <= 4 N (control-limited SLQ)
    ob = oracle.slq_opts(0, N, u_min=0.0, u_max=4.0)
    sb = oracle.slq_batch(ob, xs, xg, qr, want_traj=True, want_K=False)
    np.savez_compressed(os.path.join(OUT, "slq_box_n12_N40.npz"), x0=xs, xgoal=xg, qrqf=qr, N=N, u_min=0.0, u_max=4.0,
                        cost=sb["cost"], iters=sb["iters"], alpha_idx=sb["alpha_idx"], status=sb["status"],
                        x=sb["x"], u=sb["u"])

    # hover nominals with one Q | R | Qf block per instance (n=12 and n=24): gains with exact zeros — the oracle writes
    # them as -0.0, and the fixtures keep the sign (tests compare bit patterns)
    for model, name, N in ((0, "n12", 8), (1, "n24", 5)):
        n, m = pb.dims(model)
        batch = 4
        xh = pb.hover_state(batch, model=model)
        xh[2:] = pb.sample_x0(batch, seed=20261021, model=model)[2:]
        ABh = oracle.make_ab_batch(model, N, pb.DT, xh)
        qrb = np.tile(pb.pack_qrqf(*pb.default_weights(model)), (batch, 1)) * np.array([0.5, 1.0, 1.5, 2.0])[:, None]
        rh = oracle.lqr_batch(n, m, N, ABh, qrb, x0dev=xh, qr_per_instance=True)
        assert np.signbit(rh["K"][rh["K"] == 0.0]).any()
        np.savez_compressed(os.path.join(OUT, f"lqr_hover_perw_{name}.npz"), x0=xh, qrqf=qrb, N=N, AB=ABh, K=rh["K"], K0=rh["K0"],
                            P0=rh["P0"], cost=rh["cost"], status=rh["status"])

    # model: one step, one linearisation at a generic state (both models)
    for model, name in ((0, "quad12"), (1, "dual24")):
        x = pb.sample_x0(1, seed=7, model=model)[0]
        u = pb.hover_input(model) + 0.3 * np.sin(np.arange(pb.dims(model)[1]) + 1.0)
        xn = oracle.step(model, x, u, pb.DT)
        Aj, Bj = oracle.linearize(model, x, u, pb.DT)
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), x=x, u=u, x_next=xn, A=Aj, B=Bj)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()

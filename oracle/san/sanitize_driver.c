/* sanitize_driver.c — runs every entry point of the CPU oracle on small problems under AddressSanitizer and
 * UndefinedBehaviorSanitizer (host -fsanitize=address,undefined on the oracle; compute-sanitizer is
 * closed on the GPU pool, so this is the memory/UB tool that is left).  Built and run by oracle/san/run.sh and by
 * tests/test_oracle_sanitizers.py; exit code 0 and no sanitizer report = clean.  TEST INFRASTRUCTURE ONLY. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../noc_oracle.h"

static unsigned long long s_rng = 0x4E4F43ULL;
static double urand(double lo, double hi) {
  s_rng = s_rng * 6364136223846793005ULL + 1442695040888963407ULL;
  return lo + (hi - lo) * ((double)(s_rng >> 11) * (1.0 / 9007199254740992.0));
}
/* exact-size heap blocks, so that any out-of-bounds access lands in an ASan red zone */
static double* dalloc(size_t n) { double* p = (double*)malloc(sizeof(double) * (n ? n : 1)); for (size_t i = 0; i < n; ++i) p[i] = NAN; return p; }
static int32_t* ialloc(size_t n) { int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1)); for (size_t i = 0; i < n; ++i) p[i] = -7; return p; }

static void weights(int n, int m, double* qrqf) {
  const double q[12] = {10, 10, 10, 1, 1, 1, 5, 5, 5, 0.1, 0.1, 0.1};
  memset(qrqf, 0, sizeof(double) * (size_t)(2 * n * n + m * m));
  for (int i = 0; i < n; ++i) { qrqf[i * n + i] = q[i % 12]; qrqf[n * n + m * m + i * n + i] = 10.0 * q[i % 12]; }
  for (int i = 0; i < m; ++i) qrqf[n * n + i * m + i] = 0.1;
}

static int run_model(int model) {
  int n, m;
  if (noc_oracle_model_dims(model, &n, &m)) return 1;
  const int N = 17, B = 5, rec = n * n + n * m;
  double* qr = dalloc((size_t)(2 * n * n + m * m)); weights(n, m, qr);
  double* uh = dalloc(m); noc_oracle_hover_input(model, uh);
  double* x0 = dalloc((size_t)B * n); double* xg = dalloc((size_t)B * n);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < n; ++i) {
      const int j = i % 12;
      x0[b * n + i] = j < 3 ? urand(-2, 2) : j < 6 ? urand(-1, 1) : j < 8 ? urand(-0.3, 0.3) : j == 8 ? urand(-3, 3) : urand(-0.5, 0.5);
      xg[b * n + i] = j < 3 ? urand(-3, 3) : j == 8 ? urand(-1.5, 1.5) : 0.0;
    }
  /* models, rollout, linearisation (both layouts) */
  double* xn = dalloc(n); noc_oracle_step(model, x0, uh, 0.02, xn);
  double* A = dalloc((size_t)n * n); double* Bm = dalloc((size_t)n * m); noc_oracle_linearize(model, x0, uh, 0.02, A, Bm);
  double* ub = dalloc((size_t)N * m); for (int k = 0; k < N; ++k) memcpy(ub + k * m, uh, sizeof(double) * m);
  double* xb = dalloc((size_t)(N + 1) * n); noc_oracle_rollout_open_loop(model, N, 0.02, x0, ub, xb);
  double* AB = dalloc((size_t)N * rec);
  for (int layout = 0; layout < 2; ++layout) {
    noc_oracle_linearize_traj(model, N, 0.02, xb, ub, layout, AB);
    double* K = dalloc((size_t)N * m * n); double* P0 = dalloc((size_t)n * n); double cost = 0;
    if (noc_oracle_lqr_sweep(n, m, N, AB, layout, qr, qr + n * n, qr + n * n + m * m, K, P0, x0, &cost) != 0 || !(cost > 0)) return 2;
    if (noc_oracle_lqr_sweep(n, m, N, AB, layout, qr, qr + n * n, qr + n * n + m * m, NULL, NULL, NULL, NULL) != 0) return 3;
    free(K); free(P0);
  }
  /* batch drivers, shared and per-instance weights, every optional output absent once */
  double* ABb = dalloc((size_t)B * N * rec); noc_oracle_make_ab_batch(model, N, 0.02, B, x0, 0, ABb, 2);
  double* Kb = dalloc((size_t)B * N * m * n); double* K0 = dalloc((size_t)B * m * n); double* P0b = dalloc((size_t)B * n * n);
  double* cb = dalloc(B); int32_t* st = ialloc(B);
  noc_oracle_lqr_batch(n, m, N, B, ABb, 0, qr, 0, x0, Kb, K0, P0b, cb, st, 2);
  noc_oracle_lqr_batch(n, m, N, B, ABb, 0, qr, 0, NULL, NULL, K0, NULL, NULL, st, 1);
  double* qrs = dalloc((size_t)B * (2 * n * n + m * m));
  for (int b = 0; b < B; ++b) memcpy(qrs + (size_t)b * (2 * n * n + m * m), qr, sizeof(double) * (size_t)(2 * n * n + m * m));
  noc_oracle_lqr_batch(n, m, N, B, ABb, 0, qrs, 1, x0, Kb, K0, P0b, cb, st, 2);
  noc_oracle_lqr_from_x0_batch(model, N, 0.02, B, x0, qr, K0, P0b, cb, st, 2);
  noc_oracle_lqr_from_x0_batch(model, N, 0.02, B, x0, qr, K0, NULL, cb, st, 1);
  for (int b = 0; b < B; ++b) if (st[b] != 0) return 4;
  /* a sweep that must fail: negative R -> NOT_PD, outputs NaN */
  double* qbad = dalloc((size_t)(2 * n * n + m * m)); weights(n, m, qbad);
  for (int i = 0; i < m; ++i) qbad[n * n + i * m + i] = -1e6;
  noc_oracle_lqr_batch(n, m, N, 1, ABb, 0, qbad, 0, x0, Kb, K0, P0b, cb, st, 1);
  if (st[0] != NOC_O_NOT_PD || cb[0] == cb[0]) return 5;
  /* DARE at the first operating point */
  double* P = dalloc((size_t)n * n); double* Kd = dalloc((size_t)m * n);
  double* xh = dalloc(n); memset(xh, 0, sizeof(double) * n); if (model == NOC_O_MODEL_DUAL24) xh[12] = 1.0;
  noc_oracle_linearize(model, xh, uh, 0.02, A, Bm);
  const int it = noc_oracle_dare(n, m, A, Bm, qr, qr + n * n, 1e-12, 10000, P, Kd);
  if (it <= 0) return 6;
  /* SLQ: cold, warm, boxed; single and batch; optional outputs absent */
  noc_oracle_slq_opts o = {model, N, 12, 11, 0.02, 1e-8, 0.0, 1e-4, -INFINITY, INFINITY};
  double* x = dalloc((size_t)B * (N + 1) * n); double* u = dalloc((size_t)B * N * m); double* Ks = dalloc((size_t)B * N * m * n);
  double* kff = dalloc((size_t)N * m); int32_t* iters = ialloc(B); int32_t* ai = ialloc((size_t)B * o.max_iter);
  double c1, reg;
  int32_t it1;
  noc_oracle_slq(&o, xh, xg, qr, x, u, Ks, kff, &c1, &it1, ai, &reg);
  noc_oracle_slq(&o, xh, xg, qr, x, u, NULL, NULL, &c1, &it1, ai, &reg);
  noc_oracle_slq_warm(&o, xh, xg, qr, u, x, u, Ks, kff, &c1, &it1, ai, &reg);
  double* xs = dalloc((size_t)B * n); for (int b = 0; b < B; ++b) memcpy(xs + b * n, xh, sizeof(double) * n);
  noc_oracle_slq_batch(&o, B, xs, xg, qr, x, u, Ks, cb, iters, ai, st, 2);
  noc_oracle_slq_batch_warm(&o, B, xs, xg, qr, u, x, u, NULL, cb, iters, ai, st, 2);
  noc_oracle_slq_batch(&o, B, xs, xg, qr, NULL, NULL, NULL, cb, iters, ai, st, 1);
  noc_oracle_slq_opts ob = o; ob.u_min = 0.0; ob.u_max = 3.5;
  noc_oracle_slq_batch(&ob, B, xs, xg, qr, x, u, Ks, cb, iters, ai, st, 2);
  /* single passes */
  double dV[2];
  double* Kp = dalloc((size_t)N * m * n); double* kp = dalloc((size_t)N * m);
  noc_oracle_backward_affine(model, N, 0.02, xb, ub, xg, qr, 0.0, Kp, kp, dV);
  noc_oracle_backward_affine_box(model, N, 0.02, xb, ub, xg, qr, 1e-3, 0.0, 3.0, Kp, kp, dV);
  double* xw = dalloc((size_t)(N + 1) * n); double* uw = dalloc((size_t)N * m);
  const double J1 = noc_oracle_forward(model, N, 0.02, xb, ub, xg, qr, Kp, kp, 0.5, xw, uw);
  const double J2 = noc_oracle_forward_box(model, N, 0.02, xb, ub, xg, qr, Kp, kp, 0.25, 0.0, 3.0, xw, uw);
  const double J3 = noc_oracle_traj_cost(model, N, xw, uw, xg, qr);
  if (!(J1 == J1) || !(J2 == J2) || J2 != J3) return 7;
  /* box QP, sizes 1..8 */
  for (int mm = 1; mm <= 8; ++mm) {
    double H[64], g[8], lo[8], hi[8], xq[8];
    int32_t cl[8];
    for (int i = 0; i < mm; ++i) {
      for (int j = 0; j < mm; ++j) H[i * mm + j] = (i == j) ? 2.0 + i : 0.1;
      g[i] = urand(-3, 3); lo[i] = -0.5; hi[i] = 0.5; xq[i] = 0.0;
    }
    noc_oracle_box_qp(mm, H, g, lo, hi, xq, cl);
    for (int i = 0; i < mm; ++i) if (!(xq[i] >= -0.5 && xq[i] <= 0.5)) return 8;
  }
  free(qr); free(uh); free(x0); free(xg); free(xn); free(A); free(Bm); free(ub); free(xb); free(AB); free(ABb); free(Kb);
  free(K0); free(P0b); free(cb); free(st); free(qrs); free(qbad); free(P); free(Kd); free(xh); free(x); free(u); free(Ks);
  free(kff); free(iters); free(ai); free(xs); free(Kp); free(kp); free(xw); free(uw);
  return 0;
}

int main(void) {
  double s, c;
  for (double a = -40.0; a < 40.0; a += 0.37) noc_oracle_sincos(a, &s, &c);
  /* generic shapes of the dense sweep: n=1..3, m=1..2 LTI double-integrator-like records */
  for (int n = 1; n <= 3; ++n)
    for (int m = 1; m <= 2; ++m) {
      const int N = 9, rec = n * n + n * m;
      double* AB = dalloc((size_t)N * rec); double* qr = dalloc((size_t)(2 * n * n + m * m));
      memset(qr, 0, sizeof(double) * (size_t)(2 * n * n + m * m));
      for (int i = 0; i < n; ++i) { qr[i * n + i] = 1.0; qr[n * n + m * m + i * n + i] = 2.0; }
      for (int i = 0; i < m; ++i) qr[n * n + i * m + i] = 0.5;
      for (int k = 0; k < N; ++k)
        for (int e = 0; e < rec; ++e) AB[k * rec + e] = (e < n * n && e % (n + 1) == 0) ? 1.0 : 0.05 * (1 + (e + k) % 3);
      double* K = dalloc((size_t)N * m * n); double* P0 = dalloc((size_t)n * n); double* x0 = dalloc(n); double cost;
      for (int i = 0; i < n; ++i) x0[i] = 1.0;
      if (noc_oracle_lqr_sweep(n, m, N, AB, 0, qr, qr + n * n, qr + n * n + m * m, K, P0, x0, &cost) != 0) return 20;
      free(AB); free(qr); free(K); free(P0); free(x0);
    }
  int rc = run_model(NOC_O_MODEL_QUAD12);
  if (rc) { printf("quad12 failed: %d\n", rc); return rc; }
  rc = run_model(NOC_O_MODEL_DUAL24);
  if (rc) { printf("dual24 failed: %d\n", 100 + rc); return 100 + rc; }
  printf("ORACLE SANITIZE OK (threads %d)\n", noc_oracle_max_threads());
  return 0;
}

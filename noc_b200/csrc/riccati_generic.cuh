// riccati_generic.cuh — finite-horizon LTV Riccati sweep, and the DARE fixed-point iteration, for ANY 1 <= n <= 24,
// 1 <= m <= 8 (NOC_MODEL_LTV_USER with dimensions other than the tuned shapes; DARE: everything but n=12, m=4).  One thread per instance, matrices in thread-local memory: this is
// the completeness path of the general dense solver the C-ABI promises, not a tuned kernel — (12,4) and (24,8) go to
// riccati12.cuh / riccati24.cuh.  Same arithmetic contract, same order of operations as
// oracle/noc_oracle.c (form_H, chol_lower, chol_solve_neg, gains_and_value): bit-identical results.
#pragma once
#include "common.cuh"
#include "riccati12.cuh"   // rcp_rn_normal, pivot range

namespace noc {

constexpr int RG_NMAX = 24, RG_MMAX = 8, RG_NMMAX = RG_NMAX + RG_MMAX;

struct RicGenArgs {
  const double* AB;     // [batch][N][n*n+n*m]
  const double* QRQf;   // Q[n*n] | R[m*m] | Qf[n*n]; one block shared by the batch, or [batch][...] when qr_stride != 0
  long long qr_stride;  // doubles between the weight blocks of consecutive instances (0 = shared)
  const double* x0;     // [batch][n] or nullptr
  double* K;            // [batch][N][m*n] or nullptr
  double* K0;           // [batch][m*n] or nullptr
  double* P0;           // [batch][n*n] or nullptr
  double* cost;         // [batch] or nullptr
  int32_t* status;      // [batch] or nullptr
  long long batch;
  int n, m, N, layout;
  // dare != 0 (noc_dare_solve_batch): AB is one LTI record per instance, QRQf holds Q | R only, the step is iterated
  // from P = Q until max|P'-P| < tol max|P'| (oracle: noc_oracle_dare); P0 receives P, K0 the stationary gain
  int dare, max_iter;
  double tol;
  int32_t* iters;       // [batch] or nullptr
};

__global__ void __launch_bounds__(64) k_ric_generic(const RicGenArgs a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  const int n = a.n, m = a.m, nm = n + m, N = a.N;
  const double* Q = a.QRQf + (size_t)b * (size_t)a.qr_stride;
  const double* R = Q + n * n;
  const double* Qf = a.dare ? Q : Q + n * n + m * m;
  double P[RG_NMAX * RG_NMAX], G[RG_NMAX * RG_NMMAX], Hxx[RG_NMAX * RG_NMAX], Hux[RG_MMAX * RG_NMAX];
  double Huu[RG_MMAX * RG_MMAX], L[RG_MMAX * RG_MMAX], D[RG_MMAX], rd[RG_MMAX], Kk[RG_MMAX * RG_NMAX];
  for (int e = 0; e < n * n; ++e) P[e] = Qf[e];
  const size_t rec = (size_t)n * n + (size_t)n * m;
  bool failed = false;
  int kfail = -1;
  // F[l][c] of the record at r: layout 0 = A (n x n) then B (n x m); layout 1 = rows of [A row | B row]
  auto Fat = [&](const double* r, int l, int c) -> double {
    if (a.layout == 1) return r[l * nm + c];
    return c < n ? r[l * n + c] : r[n * n + l * m + (c - n)];
  };
  const int steps = a.dare ? a.max_iter : N;
  int dare_iters = 0;
  bool converged = false;
  for (int it = 0; it < steps && !failed; ++it) {
    const int k = a.dare ? 0 : N - 1 - it;
    const double* r = a.AB + (a.dare ? (size_t)b : (size_t)b * N + k) * rec;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < nm; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc = fmad(P[i * n + l], Fat(r, l, j), acc);
        G[i * nm + j] = acc;
      }
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) {
        double acc = Q[i * n + j];
        for (int l = 0; l < n; ++l) acc = fmad(Fat(r, l, i), G[l * nm + j], acc);
        Hxx[i * n + j] = acc;
      }
    for (int c = 0; c < m; ++c)
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc = fmad(Fat(r, l, n + c), G[l * nm + j], acc);
        Hux[c * n + j] = acc;
      }
    for (int c = 0; c < m; ++c)
      for (int d = 0; d <= c; ++d) {
        double acc = R[c * m + d];
        for (int l = 0; l < n; ++l) acc = fmad(Fat(r, l, n + c), G[l * nm + n + d], acc);
        Huu[c * m + d] = acc;   // mu = 0 in the LQR sweep
      }
    // L D L^T
    for (int j = 0; j < m && !failed; ++j) {
      double v[RG_MMAX];
      for (int l = 0; l < j; ++l) v[l] = L[j * m + l] * D[l];
      double s = Huu[j * m + j];
      for (int l = 0; l < j; ++l) s = fmad(-L[j * m + l], v[l], s);
      if (!(s > R12_PIVOT_MIN) || !(s < R12_PIVOT_MAX)) { failed = true; kfail = k; break; }
      D[j] = s;
      rd[j] = rcp_rn_normal(s);
      for (int i = j + 1; i < m; ++i) {
        double t = Huu[i * m + j];
        for (int l = 0; l < j; ++l) t = fmad(-L[i * m + l], v[l], t);
        L[i * m + j] = t * rd[j];
      }
    }
    if (failed) break;
    for (int j = 0; j < n; ++j) {
      double y[RG_MMAX], z[RG_MMAX];
      for (int i = 0; i < m; ++i) {
        double s = Hux[i * n + j];
        for (int l = 0; l < i; ++l) s = fmad(-L[i * m + l], y[l], s);
        y[i] = s;
      }
      for (int i = m - 1; i >= 0; --i) {
        double s = y[i] * rd[i];
        for (int l = i + 1; l < m; ++l) s = fmad(-L[l * m + i], z[l], s);
        z[i] = s;
      }
      for (int i = 0; i < m; ++i) Kk[i * n + j] = -z[i];
    }
    double dmax = 0.0, pmax = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) {
        double acc = Hxx[i * n + j];
        for (int c = 0; c < m; ++c) acc = fmad(Hux[c * n + i], Kk[c * n + j], acc);
        dmax = fmax(dmax, fabs(acc - P[i * n + j]));   // P and P' are symmetric: the upper triangle sees every value
        pmax = fmax(pmax, fabs(acc));
        P[i * n + j] = acc;
        P[j * n + i] = acc;
      }
    if (a.dare) {
      dare_iters = it + 1;
      if (dmax < a.tol * pmax) { converged = true; break; }
      continue;
    }
    if (a.K) {
      double* kp = a.K + ((size_t)b * N + k) * m * n;
      for (int e = 0; e < m * n; ++e) kp[e] = Kk[e];
    }
    if (k == 0 && a.K0) {
      double* kp = a.K0 + (size_t)b * m * n;
      for (int e = 0; e < m * n; ++e) kp[e] = Kk[e];
    }
  }
  const double nanv = qnan();
  if (a.dare) {
    if (failed) dare_iters += 1;   // the failing iteration counts (oracle returns -it)
    if (a.K0) {
      double* kp = a.K0 + (size_t)b * m * n;
      for (int e = 0; e < m * n; ++e) kp[e] = failed ? nanv : Kk[e];
    }
    if (a.P0) {
      double* po = a.P0 + (size_t)b * n * n;
      for (int e = 0; e < n * n; ++e) po[e] = failed ? nanv : P[e];
    }
    if (a.iters) a.iters[b] = dare_iters;
    if (a.status) a.status[b] = failed ? 1 : (converged ? 0 : 2);
    return;
  }
  if (failed) {
    if (a.K)
      for (int k2 = 0; k2 <= kfail; ++k2) {
        double* kp = a.K + ((size_t)b * N + k2) * m * n;
        for (int e = 0; e < m * n; ++e) kp[e] = nanv;
      }
    if (a.K0) {
      double* kp = a.K0 + (size_t)b * m * n;
      for (int e = 0; e < m * n; ++e) kp[e] = nanv;
    }
  }
  if (a.P0) {
    double* po = a.P0 + (size_t)b * n * n;
    for (int e = 0; e < n * n; ++e) po[e] = failed ? nanv : P[e];
  }
  if (a.cost && a.x0) {
    const double* x = a.x0 + (size_t)b * n;
    double c = 0.0;
    for (int i = 0; i < n; ++i) {
      double rr = 0.0;
      for (int j = 0; j < n; ++j) rr = fmad(P[i * n + j], x[j], rr);
      c = fmad(x[i], rr, c);
    }
    a.cost[b] = failed ? nanv : 0.5 * c;
  }
  if (a.status) a.status[b] = failed ? 1 : 0;
}

}  // namespace noc

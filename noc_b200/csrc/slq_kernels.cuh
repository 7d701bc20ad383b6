// slq_kernels.cuh — the forward half of SLQ / iLQR: initial rollout + cost, and the
// closed-loop rollouts with backtracking line search, acceptance test, convergence test and work-list compaction.
// Arithmetic: DESIGN.md §2.4, restated on the CPU by oracle/noc_oracle.c (stage_cost, terminal_cost,
// noc_oracle_forward, noc_oracle_slq); bit-identical, so iteration counts and step indices match exactly.
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace noc {

// weights in shared memory: Q[n*n] | R[m*m] | Qf[n*n]; every lane reads the same address (broadcast)
template <int MODEL>
struct CostW {
  static constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  const double* Q;
  const double* R;
  const double* Qf;
  __device__ __forceinline__ double stage(const double* x, const double* u, const double* xg) const {
    double ex[n], eu[m];
#pragma unroll
    for (int i = 0; i < n; ++i) ex[i] = x[i] - xg[i];
#pragma unroll
    for (int c = 0; c < m; ++c) eu[c] = u[c] - model::kHover;
    double qx = 0.0, qu = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) r = fmad(Q[i * n + j], ex[j], r);
      qx = fmad(ex[i], r, qx);
    }
#pragma unroll
    for (int c = 0; c < m; ++c) {
      double r = 0.0;
#pragma unroll
      for (int d = 0; d < m; ++d) r = fmad(R[c * m + d], eu[d], r);
      qu = fmad(eu[c], r, qu);
    }
    return 0.5 * (qx + qu);
  }
  __device__ __forceinline__ double terminal(const double* x, const double* xg) const {
    double ex[n];
#pragma unroll
    for (int i = 0; i < n; ++i) ex[i] = x[i] - xg[i];
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) r = fmad(Qf[i * n + j], ex[j], r);
      c = fmad(ex[i], r, c);
    }
    return 0.5 * c;
  }
};

template <int MODEL>
__device__ __forceinline__ void load_weights(double* sw, const double* __restrict__ qrqf, CostW<MODEL>& w) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  for (int i = threadIdx.x; i < 2 * n * n + m * m; i += blockDim.x) sw[i] = qrqf[i];
  __syncthreads();
  w.Q = sw;
  w.R = sw + n * n;
  w.Qf = sw + n * n + m * m;
}

// SLQ start (oracle: head of noc_oracle_slq): u = hover, x = open-loop rollout from x0, J = trajectory cost;
// also resets the per-instance solver state.  One thread per instance.
template <int MODEL>
__global__ void __launch_bounds__(128) k_slq_init(const double* __restrict__ x0, const double* __restrict__ xgoal,
                                                  const double* __restrict__ qrqf, double* __restrict__ x,
                                                  double* __restrict__ u, double* __restrict__ J,
                                                  double* __restrict__ mu, int32_t* __restrict__ status,
                                                  int32_t* __restrict__ iters, int32_t* __restrict__ alpha_idx,
                                                  int32_t* __restrict__ idx, long long batch, int N, double dt,
                                                  double reg0, int max_iter) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  extern __shared__ __align__(16) double sw[];
  CostW<MODEL> w;
  load_weights<MODEL>(sw, qrqf, w);
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double xk[n], xn[n], uk[m], xg[n];
#pragma unroll
  for (int i = 0; i < n; ++i) { xk[i] = x0[b * n + i]; xg[i] = xgoal[b * n + i]; }
#pragma unroll
  for (int c = 0; c < m; ++c) uk[c] = model::kHover;
  double* xo = x + (size_t)b * (N + 1) * n;
  double* uo = u + (size_t)b * N * m;
  double cost = 0.0;
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int i = 0; i < n; ++i) xo[(size_t)k * n + i] = xk[i];
#pragma unroll
    for (int c = 0; c < m; ++c) uo[(size_t)k * m + c] = uk[c];
    cost = cost + w.stage(xk, uk, xg);
    model::model_step<MODEL>(xk, uk, dt, xn);
#pragma unroll
    for (int i = 0; i < n; ++i) xk[i] = xn[i];
  }
#pragma unroll
  for (int i = 0; i < n; ++i) xo[(size_t)N * n + i] = xk[i];
  cost = cost + w.terminal(xk, xg);
  J[b] = cost;
  mu[b] = reg0;
  status[b] = 2;  // NOC_STATUS_MAX_ITER until something else happens
  iters[b] = 0;
  idx[b] = (int32_t)b;
  if (alpha_idx)
    for (int i = 0; i < max_iter; ++i) alpha_idx[(size_t)b * max_iter + i] = -1;
}

struct SlqFwdArgs {
  const int32_t* idx;       // [count] active instances
  int32_t* next_idx;        // compacted list of the instances that continue
  int32_t* next_count;      // atomic counter (zeroed by the host before the launch)
  const double* xgoal;      // [batch][n]
  const double* qrqf;
  double* x;                // [batch][N+1][n]  in: nominal, out: accepted trajectory (in place)
  double* u;                // [batch][N][m]
  const double* K;          // [batch][N][m*n]
  const double* kff;        // [batch][N][m]
  const double* dV;         // [batch][2]
  double* J;                // [batch]
  const int32_t* bstatus;   // [batch] backward-pass outcome (0 or NOC_STATUS_REG_FAIL)
  int32_t* status;          // [batch]
  int32_t* iters;           // [batch]
  int32_t* alpha_idx;       // [batch][max_iter] or nullptr
  int count;
  int N;
  int iter;                 // 0-based index of this SLQ iteration
  int max_iter;
  int n_alpha;              // <= 32
  double dt, tol, armijo;
};

// One warp per active instance; lane j rolls out the candidate alpha_j = 2^-j (all candidates at once instead
// of the oracle's sequential backtracking: the first accepted j is the same).  Pass 2 re-runs the accepted
// candidate (bit-identical recomputation) and lane 0 writes the trajectory in place.
template <int MODEL>
__global__ void __launch_bounds__(128) k_slq_forward(const SlqFwdArgs a) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  extern __shared__ __align__(16) double sw[];
  CostW<MODEL> w;
  load_weights<MODEL>(sw, a.qrqf, w);
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= a.count) return;
  const long long b = a.idx[item];
  const int N = a.N;
  double* xb = a.x + (size_t)b * (N + 1) * n;
  double* ub = a.u + (size_t)b * N * m;
  const double* Kb = a.K + (size_t)b * N * m * n;
  const double* kb = a.kff + (size_t)b * N * m;
  double xg[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xg[i] = a.xgoal[b * n + i];

  if (a.bstatus[b] != 0) {  // regularisation exhausted in the backward pass
    if (lane == 0) { a.status[b] = 4; a.iters[b] = a.iter + 1; }
    return;
  }
  const double Jold = a.J[b];
  const double dV1 = a.dV[b * 2], dV2 = a.dV[b * 2 + 1];

  double alpha = 1.0;
  for (int j = 0; j < lane; ++j) alpha = alpha * 0.5;
  int accepted = -1;
  double Jacc = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    double xk[n], xn[n], uk[m];
#pragma unroll
    for (int i = 0; i < n; ++i) xk[i] = xb[i];
    double Jn = 0.0;
    for (int k = 0; k < N; ++k) {
      double dx[n];
#pragma unroll
      for (int i = 0; i < n; ++i) dx[i] = xk[i] - xb[(size_t)k * n + i];
#pragma unroll
      for (int c = 0; c < m; ++c) {
        double tk = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) tk = fmad(Kb[(size_t)k * m * n + c * n + i], dx[i], tk);
        uk[c] = (ub[(size_t)k * m + c] + alpha * kb[(size_t)k * m + c]) + tk;
      }
      Jn = Jn + w.stage(xk, uk, xg);
      model::model_step<MODEL>(xk, uk, a.dt, xn);
      if (pass == 1) {
        __syncwarp();  // every lane has read the nominal x_k, u_k before lane 0 overwrites them
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < n; ++i) xb[(size_t)k * n + i] = xk[i];
#pragma unroll
          for (int c = 0; c < m; ++c) ub[(size_t)k * m + c] = uk[c];
        }
      }
#pragma unroll
      for (int i = 0; i < n; ++i) xk[i] = xn[i];
    }
    Jn = Jn + w.terminal(xk, xg);
    if (pass == 1) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < n; ++i) xb[(size_t)N * n + i] = xk[i];
      }
      break;
    }
    const double expected = -(alpha * dV1 + (alpha * alpha) * dV2);
    const bool ok = (lane < a.n_alpha) && (Jold - Jn >= a.armijo * expected);
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask == 0u) break;
    accepted = __ffs(mask) - 1;
    Jacc = __shfl_sync(0xffffffffu, Jn, accepted);
    alpha = __shfl_sync(0xffffffffu, alpha, accepted);
  }
  if (lane != 0) return;
  if (accepted < 0) {
    a.status[b] = 3;  // NOC_STATUS_LS_FAIL
    a.iters[b] = a.iter + 1;
    return;
  }
  if (a.alpha_idx) a.alpha_idx[(size_t)b * a.max_iter + a.iter] = accepted;
  const double denom = (Jold > 1e-12) ? Jold : 1e-12;
  const double rel = (Jold - Jacc) / denom;
  a.J[b] = Jacc;
  a.iters[b] = a.iter + 1;
  if (rel < a.tol) {
    a.status[b] = 0;
    return;
  }
  if (a.iter + 1 < a.max_iter) {
    const int pos = atomicAdd(a.next_count, 1);
    a.next_idx[pos] = (int32_t)b;
  }
}

}  // namespace noc

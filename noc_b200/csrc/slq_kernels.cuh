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
                                                  const double* __restrict__ qrqf,
                                                  const double* __restrict__ u_init,   // [batch][N][m] or nullptr (hover)
                                                  double* __restrict__ x,
                                                  double* __restrict__ u, double* __restrict__ J,
                                                  double* __restrict__ mu, double* __restrict__ delta,
                                                  int32_t* __restrict__ status,
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
    if (u_init) {
#pragma unroll
      for (int c = 0; c < m; ++c) uk[c] = u_init[((size_t)b * N + k) * m + c];
    }
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
  delta[b] = 1.0;
  status[b] = 2;  // NOC_STATUS_MAX_ITER until something else happens
  iters[b] = 0;
  idx[b] = (int32_t)b;
  if (alpha_idx)
    for (int i = 0; i < max_iter; ++i) alpha_idx[(size_t)b * max_iter + i] = -1;
}

constexpr int SLQ_FLAG_SPEC_NEXT = 16;   // backtracked in this iteration: speculate in the next one
constexpr int SLQ_FLAG_SPEC_NOW = 32;    // handled by the speculative round of this iteration: the ordinary trial 0 skips it

struct SlqFwdArgs {
  const int32_t* idx;       // [count] active instances
  int cont_bit;             // the flag bit an instance that continues gets (1; 4 or 8 from a deferred backtracking round)
  int32_t* flags;           // [batch] bit 0: continues with the next SLQ iteration; bit 1: trial rejected, try the next
                            // step length.  k_compact_flags turns the bits into SORTED work lists: a list in
                            // instance order keeps the per-instance streams of a warp at a regular stride (an
                            // atomicAdd-ordered list made the full-size rollouts 1.6x slower).
  int trial;                // j of this launch: alpha = 2^-j for every instance of the work list
  // speculative round (cand_J > 0): work item = (list position i, candidate c), step length 2^-(c+1); the
  // trajectories go to the candidate scratch instead of x / u and k_slq_accept picks the first accepted c
  int cand_J;               // candidates per instance in this round, 0 = ordinary round
  int cand_first;           // trial index of candidate 0 (step length 2^-cand_first); candidates are consecutive trials
  int cand_last;            // 1: no step lengths remain after this round (an instance without a winner fails);
                            // 0: it is flagged for the next round instead
  int spec;                 // 1: SPECULATIVE round — the instances that needed a shorter step in their previous iteration get ALL
                            // their step lengths (cand_first = 0) as candidates beside everybody else's trial 0, on the auxiliary
                            // stream: the nominal is read from x / u themselves (no trial 0 has overwritten it) and stays there
                            // if no candidate passes.  The first passing candidate in order is the oracle's sequential search.
  int spec_mark;            // 1: instances that end an iteration on a step index > 0 (or a failed line search they repeat) are
                            // flagged SLQ_FLAG_SPEC_NEXT; their next forward pass is the speculative round
  double* xc;               // [count/cand_J][cand_J][N+1][n]
  double* uc;               // [count/cand_J][cand_J][N][m]
  double* lcc;              // [count/cand_J][cand_J][N+1]
  const double* xgoal;      // [batch][n]
  const double* qrqf;
  double* x;                // [batch][N+1][n]  in: nominal, out: accepted trajectory (in place)
  double* u;                // [batch][N][m]
  double* xn;               // [batch][N+1][n]  scratch: the nominal, saved by trial 0 while x / u are overwritten
  double* un;               // [batch][N][m]
  const double* K;          // [batch][N][m*n]
  const double* kff;        // [batch][N][m]
  const double* dV;         // [batch][2]
  double* lc;               // [batch][N+1] stage costs of the trial (k < N) and its terminal cost (k = N)
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
  double u_min, u_max;      // input box of the closed-loop rollouts (+-inf = none)
  double* mu;               // [batch] Levenberg-Marquardt mu (adaptive schedule: updated after every line search)
  double* delta;            // [batch] its scaling factor
  int reg_schedule;         // noc_reg_schedule
};

// The forward pass is three kernels per step length: (1) k_slq_rollout, the sequential closed-loop rollout, only
// the state/input recursion (latency-critical); (2) k_slq_stage_cost, the stage and terminal costs of the new
// trajectory, TIME-PARALLEL over (instance, k); (3) k_slq_accept, the ordered sum of the stage costs, Armijo test
// and bookkeeping.  Splitting the cost out of the rollout removes 60 % of the instructions from the latency-bound
// chain (profiles/launches_cfg3_slq_r01.csv: 1.7 ms per 200-step rollout before).
// Rollout mapping: FOUR lanes per instance, eight instances per warp, one warp per CTA.  The first
// thread-per-instance versions were bound by the LSU, not by HBM or the dependent chain: every 16-byte access of a
// warp went to 32 different instances, i.e. 32 sectors in 32 cache lines per instruction, ~1600 line requests per
// warp and step (profiles/ncu_r01_k_slq_rollout_v4.txt: lg_throttle 5.1 + short/long scoreboard 8.7 stall cycles
// per issue, 4.4 us per step).  With four lanes per instance a warp instruction covers eight 64-byte segments; lane
// t copies every fourth 16-byte chunk of the step's inputs (K_k, kff_k, xbar_k, ubar_k) into a STAGES-deep
// cp.async ring in shared memory, computes the feedback rows c = t (, t+4) of u = ubar + alpha k + K (x - xbar)
// as the oracle's ascending fma chain, the four lanes exchange their u by shuffle, and every lane steps the model
// redundantly (bitwise the same result), so no further exchange is needed.
// Backtracking is sequential like the oracle's (alpha_j = 2^-j until the Armijo test passes) but runs as one
// LAUNCH per step length: trial 0 goes over the whole work list, writes its trajectory IN PLACE over the nominal
// (x_k, u_k are consumed before they are overwritten; the ring prefetches only later steps) and saves the nominal
// it replaces to xn / un; the 1 % of instances whose trial is rejected go to a retry list that the speculative
// round settles (nominal read from xn / un).  alpha = 1 is accepted in 99 % of the iterations, so no warp ever
// repeats a rollout because one of its instances backtracks.
template <int MODEL>
struct FwdCfg {
  static constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  static constexpr int lanes = 4;                    // per instance
  static constexpr int inst = 8;                     // instances per warp = per CTA
  static constexpr int threads = 32;
  static constexpr int rows = m / lanes;             // feedback rows per lane
  static constexpr int stages = (MODEL == 0) ? 3 : 2;
  static constexpr int cK = m * n / 2, cX = n / 2, cU = m / 2;    // 16-byte chunks per step: K | xbar | ubar | kff
  static constexpr int chunks = cK + cX + 2 * cU;
  static_assert((cX + cU) % 4 == 0, "the store list x_k | u_k must fill whole lane groups");
  static constexpr size_t ring_bytes = (size_t)stages * inst * chunks * 16;
};

// (1) the sequential part: closed-loop rollout of trial a.trial, no cost evaluation.
template <int MODEL>
__global__ void __launch_bounds__(FwdCfg<MODEL>::threads) k_slq_rollout(const SlqFwdArgs a) {
  using C = FwdCfg<MODEL>;
  constexpr int n = C::n, m = C::m, S = C::stages;
  extern __shared__ __align__(16) double sw[];
  double2* ring = reinterpret_cast<double2*>(sw);
  const int lane = threadIdx.x, q = lane >> 2, t = lane & 3;
  const long long item_raw = (long long)blockIdx.x * C::inst + q;
  bool valid = item_raw < a.count;          // invalid slots shadow the last item: all 32 lanes stay in the shuffles
  const int item = (int)(valid ? item_raw : a.count - 1);
  const bool cand = a.cand_J > 0;
  const long long b = a.idx[cand ? item / a.cand_J : item];
  const int trial = cand ? a.cand_first + item % a.cand_J : a.trial;
  const int N = a.N;
  if (trial == 0 && a.bstatus[b] != 0) valid = false;   // regularisation exhausted: k_slq_accept records it
  if (!cand && (a.flags[b] & SLQ_FLAG_SPEC_NOW)) valid = false;   // this instance's step lengths are candidates of the speculative round
  double* xb = cand ? a.xc + (size_t)item * (N + 1) * n : a.x + (size_t)b * (N + 1) * n;   // receives the trajectory
  double* ub = cand ? a.uc + (size_t)item * N * m : a.u + (size_t)b * N * m;
  double* xs = a.xn + (size_t)b * (N + 1) * n;   // holds the nominal from trial 0 on
  double* us = a.un + (size_t)b * N * m;
  const bool first = trial == 0 && !a.spec;      // the in-place trial: reads the nominal it overwrites, and saves it
  const double* xnom = a.spec ? a.x + (size_t)b * (N + 1) * n : (first ? xb : xs);   // where this trial reads the nominal
  const double* unom = a.spec ? a.u + (size_t)b * N * m : (first ? ub : us);
  const double* Kb = a.K + (size_t)b * N * m * n;
  const double* kb = a.kff + (size_t)b * N * m;
  double alpha = 1.0;
  for (int j = 0; j < trial; ++j) alpha = alpha * 0.5;

  auto prefetch = [&](int k) {
    if (k < N) {
      double2* st = ring + ((size_t)(k % S) * C::inst + q) * C::chunks;
      const double* srcK = Kb + (size_t)k * m * n;
      const double* srcX = xnom + (size_t)k * n;
      const double* srcU = unom + (size_t)k * m;
      const double* srcF = kb + (size_t)k * m;
#pragma unroll
      for (int c0 = 0; c0 < C::cK; c0 += 4) cp_async16(smem_u32(st + c0 + t), srcK + 2 * (c0 + t));
      // xbar | ubar | kff: one virtual list of cX + 2 cU chunks (contiguous in the ring, three arrays in global
      // memory); lane t takes entry 4g + t of group g, so the four lanes of an instance copy neighbouring chunks
      // in ONE instruction (a per-lane `if` here compiles to four divergent paths of single-lane copies)
#pragma unroll
      for (int g = 0; g < (C::cX + 2 * C::cU + 3) / 4; ++g) {
        const int e = 4 * g + t;
        const double* src = e < C::cX ? srcX + 2 * e : e < C::cX + C::cU ? srcU + 2 * (e - C::cX) : srcF + 2 * (e - C::cX - C::cU);
        if (e < C::cX + 2 * C::cU) cp_async16(smem_u32(st + C::cK + e), src);
      }
    }
    cp_async_commit();   // one group per step, also when empty: keeps the wait_group arithmetic uniform
  };

  double xk[n], xn[n], uk[m];
#pragma unroll
  for (int i = 0; i < n; ++i) xk[i] = xnom[i];
#pragma unroll
  for (int p = 0; p < S - 1; ++p) prefetch(p);
  for (int k = 0; k < N; ++k) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(S - 2) : "memory");   // this lane's copies of step k have landed
    __syncwarp();   // ... and everybody else's; every lane is done reading the stage the next prefetch overwrites
    prefetch(k + S - 1);
    const double2* st = ring + ((size_t)(k % S) * C::inst + q) * C::chunks;
    double dx[n];
#pragma unroll
    for (int i = 0; i < n; i += 2) {
      const double2 v = st[C::cK + i / 2];
      dx[i] = xk[i] - v.x; dx[i + 1] = xk[i + 1] - v.y;
    }
    double ur[C::rows];
#pragma unroll
    for (int r = 0; r < C::rows; ++r) {
      const int c = t + 4 * r;   // own feedback row
      const double2* krow = st + (c * n) / 2;
      double tk = 0.0;
#pragma unroll
      for (int i = 0; i < n; i += 2) {
        const double2 kv = krow[i / 2];
        tk = fmad(kv.x, dx[i], tk);
        tk = fmad(kv.y, dx[i + 1], tk);
      }
      const double* up = reinterpret_cast<const double*>(st + C::cK + C::cX);
      const double v = (up[c] + alpha * up[m + c]) + tk;
      ur[r] = v < a.u_min ? a.u_min : (v > a.u_max ? a.u_max : v);   // the comparisons are false for NaN: a diverged rollout stays NaN
    }
#pragma unroll
    for (int c = 0; c < m; ++c) uk[c] = __shfl_sync(0xffffffffu, ur[c / 4], (lane & ~3) | (c & 3));
    model::model_step<MODEL>(xk, uk, a.dt, xn);
    if (valid) {
      // x_k | u_k as one list of cX + cU chunks; lane t stores entry 4g + t of group g (value picked from the
      // registers by t), and on trial 0 first saves the nominal chunk it replaces (still in the ring)
#pragma unroll
      for (int g = 0; g < (C::cX + C::cU) / 4; ++g) {
        const int e = 4 * g + t;
        const bool isx = e < C::cX;
        const size_t off = isx ? (size_t)k * n + 2 * e : (size_t)k * m + 2 * (e - C::cX);
        if (first) *reinterpret_cast<double2*>((isx ? xs : us) + off) = st[C::cK + e];
        auto chunk = [&](int ee) { return ee < C::cX ? make_double2(xk[2 * ee], xk[2 * ee + 1]) : make_double2(uk[2 * (ee - C::cX)], uk[2 * (ee - C::cX) + 1]); };
        const double2 v0 = chunk(4 * g), v1 = chunk(4 * g + 1), v2 = chunk(4 * g + 2), v3 = chunk(4 * g + 3);
        const double2 lo = (t & 1) ? v1 : v0, hi = (t & 1) ? v3 : v2;
        *reinterpret_cast<double2*>((isx ? xb : ub) + off) = (t & 2) ? hi : lo;
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) xk[i] = xn[i];
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  if (valid) {
#pragma unroll
    for (int g = 0; g < (C::cX + 3) / 4; ++g) {
      const int e = 4 * g + t;
      auto chunk = [&](int ee) { return ee < C::cX ? make_double2(xk[2 * ee], xk[2 * ee + 1]) : make_double2(0.0, 0.0); };
      const double2 v0 = chunk(4 * g), v1 = chunk(4 * g + 1), v2 = chunk(4 * g + 2), v3 = chunk(4 * g + 3);
      const double2 lo = (t & 1) ? v1 : v0, hi = (t & 1) ? v3 : v2;
      if (e < C::cX) {
        if (first) *reinterpret_cast<double2*>(xs + (size_t)N * n + 2 * e) = *reinterpret_cast<const double2*>(xb + (size_t)N * n + 2 * e);
        *reinterpret_cast<double2*>(xb + (size_t)N * n + 2 * e) = (t & 2) ? hi : lo;
      }
    }
  }
}

// (2) the time-parallel part: stage costs l_k (k < N) and the terminal cost (k = N) of the trial trajectories,
// one thread per (work item, k) -> lc[b][k]
template <int MODEL>
__global__ void __launch_bounds__(128) k_slq_stage_cost(const SlqFwdArgs a) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  extern __shared__ __align__(16) double sw[];
  CostW<MODEL> w;
  load_weights<MODEL>(sw, a.qrqf, w);   // once per CTA: the grid is capped and every thread strides over the points
  const int N = a.N;
  const bool cand = a.cand_J > 0;
  const long long total = (long long)a.count * (N + 1);
  for (long long tw = (long long)blockIdx.x * blockDim.x + threadIdx.x; tw < total; tw += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(tw / (N + 1));
    const int k = (int)(tw - (long long)item * (N + 1));
    const long long b = a.idx[cand ? item / a.cand_J : item];
    if (!cand && a.trial == 0 && a.bstatus[b] != 0) continue;
    if (!cand && (a.flags[b] & SLQ_FLAG_SPEC_NOW)) continue;
    double x[n], xg[n], u[m];
    const double* xp = cand ? a.xc + ((size_t)item * (N + 1) + k) * n : a.x + ((size_t)b * (N + 1) + k) * n;
    const double* gp = a.xgoal + (size_t)b * n;
#pragma unroll
    for (int i = 0; i < n; i += 2) {   // (16-byte loads: rows of n doubles, n even, buffers 16-byte aligned)
      const double2 v = *reinterpret_cast<const double2*>(xp + i), gq = __ldg(reinterpret_cast<const double2*>(gp + i));
      x[i] = v.x; x[i + 1] = v.y; xg[i] = gq.x; xg[i + 1] = gq.y;
    }
    double c;
    if (k < N) {
      const double* up = cand ? a.uc + ((size_t)item * N + k) * m : a.u + ((size_t)b * N + k) * m;
#pragma unroll
      for (int j = 0; j < m; j += 2) { const double2 v = *reinterpret_cast<const double2*>(up + j); u[j] = v.x; u[j + 1] = v.y; }
      c = w.stage(x, u, xg);
    } else {
      c = w.terminal(x, xg);
    }
    if (cand) a.lcc[(size_t)item * (N + 1) + k] = c;
    else a.lc[(size_t)b * (N + 1) + k] = c;
  }
}

// (3) per instance: J = sum of the stage costs in ascending k (the oracle's order), Armijo test, bookkeeping,
// convergence test, next work list / retry list
// bookkeeping of an accepted step (trial index `accepted`, cost Jacc): records, convergence test, next work list
// (The iteration index is the INSTANCE's own count, iters[b]: with deferred backtracking, noc_b200.cu, an instance that
// needed a shorter step rejoins the work list one sweep later than the instances that did not.)
__device__ __forceinline__ void slq_record_step(const SlqFwdArgs& a, long long b, double Jold, double Jacc, int accepted) {
  const int itb = a.iters[b];
  if (a.alpha_idx) a.alpha_idx[(size_t)b * a.max_iter + itb] = accepted;
  if (a.reg_schedule == 1) {   // accepted step: relax the regularisation
    double mu = a.mu[b], dl = a.delta[b];
    reg_decrease(mu, dl);
    a.mu[b] = mu; a.delta[b] = dl;
  }
  const double denom = (Jold > 1e-12) ? Jold : 1e-12;
  const double rel = (Jold - Jacc) / denom;
  a.J[b] = Jacc;
  a.iters[b] = itb + 1;
  if (rel < a.tol) {
    a.status[b] = 0;
    return;
  }
  if (itb + 1 < a.max_iter) a.flags[b] |= a.cont_bit;
}

// no step length passed (the nominal has been put back).  Default schedule: the instance stops with LS_FAIL.  Adaptive
// schedule: mu is raised and the iteration is repeated from the same nominal (recorded as alpha_idx = -2) unless mu
// leaves its range.  Called by one lane.
__device__ __forceinline__ void slq_line_search_failed(const SlqFwdArgs& a, long long b) {
  const int itb = a.iters[b];
  a.iters[b] = itb + 1;
  if (a.reg_schedule == 1) {
    double mu = a.mu[b], dl = a.delta[b];
    reg_increase(mu, dl);
    a.mu[b] = mu; a.delta[b] = dl;
    if (!(mu > 1e6)) {
      if (a.alpha_idx) a.alpha_idx[(size_t)b * a.max_iter + itb] = -2;
      if (itb + 1 < a.max_iter) a.flags[b] |= a.cont_bit;
      return;
    }
  }
  a.status[b] = 3;  // NOC_STATUS_LS_FAIL
}

// ordinary round (one step length per launch): one WARP per instance.  The stage costs are read with coalesced
// loads, 32 at a time, and summed in ascending k by every lane (shuffle broadcast of lane i's value, then the add —
// the oracle's chain); lane 0 does the bookkeeping.  (A thread per instance read its 201 costs with a stride of
// 1.6 KB between lanes: 35 us per launch, profiles/launches_cfg3_slq_r01d.csv.)
template <int MODEL>
__global__ void __launch_bounds__(128) k_slq_accept(const SlqFwdArgs a) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  const int lane = threadIdx.x & 31;
  const int item = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (item >= a.count) return;   // warp-uniform
  const long long b = a.idx[item];
  const int N = a.N;
  if (a.flags[b] & SLQ_FLAG_SPEC_NOW) return;   // settled by the speculative round (warp-uniform)
  if (a.trial == 0 && a.bstatus[b] != 0) {  // regularisation exhausted in the backward pass
    if (lane == 0) {
      a.status[b] = 4;
      a.iters[b] = a.iters[b] + 1;
    }
    return;
  }
  const double* lc = a.lc + (size_t)b * (N + 1);
  double Jacc = 0.0;
  for (int k0 = 0; k0 <= N; k0 += 32) {
    const double v = (k0 + lane <= N) ? lc[k0 + lane] : 0.0;
    const int cnt = (N + 1 - k0 < 32) ? N + 1 - k0 : 32;
    for (int i = 0; i < cnt; ++i) Jacc = Jacc + __shfl_sync(0xffffffffu, v, i);
  }
  const double Jold = a.J[b];
  const double dV1 = a.dV[b * 2], dV2 = a.dV[b * 2 + 1];
  double alpha = 1.0;
  for (int j = 0; j < a.trial; ++j) alpha = alpha * 0.5;
  const double expected = -(alpha * dV1 + (alpha * alpha) * dV2);
  if (!(Jold - Jacc >= a.armijo * expected)) {
    if (a.trial + 1 < a.n_alpha) {   // next step length(s), next launch
      if (lane == 0) a.flags[b] |= 2;
      return;
    }
    // no step length passed: put the nominal back (oracle: x, u unchanged) and stop this instance
    const double2* sx = reinterpret_cast<const double2*>(a.xn + (size_t)b * (N + 1) * n);
    const double2* su = reinterpret_cast<const double2*>(a.un + (size_t)b * N * m);
    double2* dxp = reinterpret_cast<double2*>(a.x + (size_t)b * (N + 1) * n);
    double2* dup = reinterpret_cast<double2*>(a.u + (size_t)b * N * m);
    for (int i = lane; i < (N + 1) * n / 2; i += 32) dxp[i] = sx[i];
    for (int i = lane; i < N * m / 2; i += 32) dup[i] = su[i];
    if (lane == 0) {
      slq_line_search_failed(a, b);
      if (a.spec_mark && (a.flags[b] & a.cont_bit)) a.flags[b] |= SLQ_FLAG_SPEC_NEXT;
    }
    return;
  }
  if (lane == 0) {
    slq_record_step(a, b, Jold, Jacc, a.trial);
    if (a.spec_mark && a.trial > 0 && (a.flags[b] & a.cont_bit)) a.flags[b] |= SLQ_FLAG_SPEC_NEXT;
  }
}

// speculative round (cand_J > 0): one WARP per rejected instance.  Lane c sums the stage costs of candidate c
// (alpha = 2^-(cand_first+c)) in ascending k and runs its Armijo test; the first passing candidate in order wins (the
// oracle's sequential backtracking stops at the same one); then the warp copies the winning trajectory (or, if none
// passed, the saved nominal) over x, u with coalesced 16-byte accesses.  (One thread per instance did both jobs
// serially at first: 0.5-0.6 ms per launch for a handful of instances, profiles/ncu_r01_k_slq_rollout_v4.txt.)
template <int MODEL>
__global__ void __launch_bounds__(128) k_slq_accept_cand(const SlqFwdArgs a) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  const int lane = threadIdx.x & 31;
  const int item = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (item >= a.count) return;   // warp-uniform
  const long long b = a.idx[item];
  const int N = a.N;
  if (a.spec && a.bstatus[b] != 0) {   // regularisation exhausted in the backward pass (k_slq_accept's rule for trial 0)
    if (lane == 0) {
      a.status[b] = 4;
      a.iters[b] = a.iters[b] + 1;
    }
    return;
  }
  const double Jold = a.J[b];
  const double dV1 = a.dV[b * 2], dV2 = a.dV[b * 2 + 1];
  double Jc = 0.0;
  bool pass = false;
  if (lane < a.cand_J) {
    // ordered sum (the oracle's order); the loads of a lane go to its own cache lines, so they are issued 16 at a time —
    // one at a time the N+1 dependent load latencies, not the additions, were the cost of this kernel
    const double* lc = a.lcc + ((size_t)item * a.cand_J + lane) * (N + 1);
    int k = 0;
    for (; k + 16 <= N + 1; k += 16) {
      double v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = __ldcs(lc + k + q);
#pragma unroll
      for (int q = 0; q < 16; ++q) Jc = Jc + v[q];
    }
    for (; k <= N; ++k) Jc = Jc + lc[k];
    double alpha = 1.0;
    for (int j = 0; j < a.cand_first + lane; ++j) alpha = alpha * 0.5;
    const double expected = -(alpha * dV1 + (alpha * alpha) * dV2);
    pass = Jold - Jc >= a.armijo * expected;
  }
  const unsigned won = __ballot_sync(0xffffffffu, pass);
  const int c = won ? __ffs((int)won) - 1 : -1;
  if (c < 0 && !a.cand_last) {   // no winner among these step lengths, shorter ones remain: next round
    if (lane == 0) a.flags[b] |= 2;
    return;
  }
  const double Jacc = __shfl_sync(0xffffffffu, Jc, c < 0 ? 0 : c);
  const size_t slot = (size_t)item * a.cand_J + (c < 0 ? 0 : c);
  const double2* sx = reinterpret_cast<const double2*>(c >= 0 ? a.xc + slot * (N + 1) * n : a.xn + (size_t)b * (N + 1) * n);
  const double2* su = reinterpret_cast<const double2*>(c >= 0 ? a.uc + slot * N * m : a.un + (size_t)b * N * m);
  double2* dxp = reinterpret_cast<double2*>(a.x + (size_t)b * (N + 1) * n);
  double2* dup = reinterpret_cast<double2*>(a.u + (size_t)b * N * m);
  if (c >= 0 || !a.spec) {   // the winning trajectory over the nominal (speculative round without a winner: x / u still hold it)
    // eight 16-byte loads in flight per lane
    const int nx = (N + 1) * n / 2, nu = N * m / 2;
    int i = lane;
    for (; i + 7 * 32 < nx; i += 8 * 32) {
      double2 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = sx[i + 32 * q];
#pragma unroll
      for (int q = 0; q < 8; ++q) dxp[i + 32 * q] = v[q];
    }
    for (; i < nx; i += 32) dxp[i] = sx[i];
    i = lane;
    for (; i + 7 * 32 < nu; i += 8 * 32) {
      double2 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = su[i + 32 * q];
#pragma unroll
      for (int q = 0; q < 8; ++q) dup[i + 32 * q] = v[q];
    }
    for (; i < nu; i += 32) dup[i] = su[i];
  }
  if (lane != 0) return;
  if (c < 0) {   // the nominal has been put back
    slq_line_search_failed(a, b);
    if (a.spec_mark && (a.flags[b] & a.cont_bit)) a.flags[b] |= SLQ_FLAG_SPEC_NEXT;
    return;
  }
  slq_record_step(a, b, Jold, Jacc, a.cand_first + c);
  if (a.spec_mark && a.cand_first + c > 0 && (a.flags[b] & a.cont_bit)) a.flags[b] |= SLQ_FLAG_SPEC_NEXT;
}

// SLQ_FLAG_SPEC_NOW on (set = 1) or off the instances of a (short) list.  The bit is set before the forward pass forks and
// cleared after it has joined: the ordinary kernels on the main stream must see it whenever they run.
__global__ void k_slq_mark_spec(const int32_t* __restrict__ list, int count, int32_t* __restrict__ flags, int set) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (set) flags[list[i]] |= SLQ_FLAG_SPEC_NOW;
  else flags[list[i]] &= ~SLQ_FLAG_SPEC_NOW;
}

// Ordered stream compaction of one flag bit: out = ascending list of the i < n with flags[i] & mask, the bit is
// cleared, *out_count = length.  Single CTA (work lists are at most a few hundred thousand entries).
__global__ void __launch_bounds__(1024) k_compact_flags(int32_t* __restrict__ flags, int mask, long long n,
                                                        int32_t* __restrict__ out, int32_t* __restrict__ out_count) {
  constexpr int ITEMS = 8;
  __shared__ int warp_sums[32];
  __shared__ int round_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int base = 0;
  for (long long start = 0; start < n; start += 1024 * ITEMS) {
    const long long i0 = start + (long long)threadIdx.x * ITEMS;
    bool f[ITEMS];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const long long i = i0 + j;
      f[j] = false;
      if (i < n) {
        const int v = flags[i];
        f[j] = (v & mask) != 0;
        if (f[j]) flags[i] = v & ~mask;
      }
      cnt += f[j] ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += v;
      }
      warp_sums[lane] = wi - w;   // exclusive prefix of the warp totals
      if (lane == 31) round_total = wi;
    }
    __syncthreads();
    int pos = base + warp_sums[wid] + incl - cnt;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (f[j]) out[pos++] = (int32_t)(i0 + j);
    base += round_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_count = base;
}

}  // namespace noc

// riccati24.cuh — backward discrete Riccati recursion for n=24, m=8 (BASELINE.json configs[4], the spring-coupled
// dual-UAV model).  Same arithmetic contract as riccati12.cuh (oracle/noc_oracle.c), same
// column-ownership idea, different tiling because F=[A|B] is 24x32 and P is 24x24:
//   * a warp owns TWO instances; instance q = lane>>4 is a group of 16 lanes, t = lane&15.
//   * lanes t<12 ("x-lanes") own the x-columns jA = t and jB = 23-t; lanes t>=12 ("u-lanes") own the u-columns
//     uA = t-12 and uB = 19-t (F columns 24+uA, 24+uB).  All lanes run the same code: per column the upper
//     rows of H = W + F^T G (12 rows for slot A, 24 for slot B) and the eight u-rows (Hux for x-lanes, Huu for
//     u-lanes): 52 accumulators.  For u-lanes the x-row accumulators are dead weight (the price of uniform code).
//   phase 1  G[:,c] = P F[:,c] for the two own columns: 48 accumulators, G never leaves registers.
//   phase 2  H rows; Hux / Huu exchanged through shared memory.
//   phase 3  8x8 L D L^T redundantly in every lane (registers; G is dead by then), gains of the own x-columns
//            (+ the feed-forward column in RIC_AFFINE).
//   phase 4  P' for own x-columns, stored at [i][j] and mirrored (predicated stores).
//   A_k,B_k records (6144 B) arrive by TMA bulk copy, one per instance and step, on a per-warp mbarrier.
#pragma once
#include "common.cuh"
#include "model.cuh"
#include "riccati12.cuh"  // mbarrier / bulk-copy helpers, rcp_rn_normal, RIC_* modes

namespace noc {

constexpr int R24_NX = 24, R24_NU = 8, R24_NXU = 32;
constexpr int R24_REC = R24_NX * R24_NX + R24_NX * R24_NU;  // 768 doubles = 6144 B per step
// per-instance shared memory, in doubles
constexpr int R24_FS = 0;       // raw record: layout 0 = A[24][24] | B[24][8]; layout 1 = F[24][32]
constexpr int R24_PS = 768;     // P[24][24]
constexpr int R24_HS = 1344;    // Hux[8][24]
constexpr int R24_US = 1536;    // Huu[8][8]
constexpr int R24_XS = 1600;    // affine: xbar_k[24] | ubar_k[8]
constexpr int R24_ES = 1632;    // affine: [xbar_k - xgoal ; ubar_k - u_h]
constexpr int R24_VS = 1664;    // affine: Vx[24]
constexpr int R24_HU = 1688;    // affine: h_u[8]
constexpr int R24_XG = 1696;    // affine: xgoal[24]
constexpr int R24_QP = 1720;     // BOX: workspace of the 8x8 box QP (QP_WS_DOUBLES(8) = 112)
constexpr int R24_STRIDE_LQR = 1602, R24_STRIDE_LQR_FUSED = 1634, R24_STRIDE_AFF = 1722, R24_STRIDE_AFF_BOX = 1834;   // 16 B modulo 32 B: the two instances of a warp differ in bank quad
constexpr int R24_WLD = 34;                                   // padded row stride of W
constexpr int R24_WS_DOUBLES = 32 * R24_WLD;
constexpr int R24_QF_DOUBLES = 576;
constexpr int R24_BAR_DOUBLES = 8;

struct Ric24Args {
  const double* AB;      // [batch][N][768]
  const double* W;       // [32][32] symmetric blockdiag(Q,R)
  const double* Qf;      // [24][24]
  long long w_stride;    // k_ric24_mma<PERW>: W = the per-instance QRQf array, this many doubles (1216) between instances
  const double* x0;      // [batch][24] or nullptr
  double* K;             // [batch][N][192] or nullptr
  double* K0;            // [batch][192] or nullptr
  double* P0;            // [batch][576] or nullptr
  double* cost;          // [batch] or nullptr
  int32_t* status;       // [batch] or nullptr
  const int32_t* idx;    // RIC_AFFINE work list or nullptr
  const double* xbar;    // nominal states: x_k of instance b at xbar + b*xbar_sb + k*xbar_sk ([batch][N+1][24] when 0/0)
  long long xbar_sb, xbar_sk;   // strides in doubles; 0 = the dense instance-major layout
  const double* ubar;    // [batch][N][8]
  const double* xgoal;   // [batch][24]
  double dt;             // FUSED: Euler step of the built-in model
  double u_min, u_max;   // BOX: input box of the SLQ backward pass
  double* kff;           // [batch][N][8]
  double* dV;            // [batch][2]
  double* mu;            // [batch]
  double* delta;         // [batch] in/out: adaptive schedule's scaling factor (reg_schedule == 1)
  int reg_schedule;      // noc_reg_schedule
  long long batch;
  int N;
  // RIC_DARE (k_ric24_mma): AB is one LTI record per instance, Qf holds Q; P0 / K0 receive P and the stationary gain
  int32_t* iters;        // [batch]
  double tol;
  int max_iter;
};

template <int MODE, bool FUSED, bool BOX = false>
__host__ __device__ constexpr int r24_stride() {
  return MODE == RIC_AFFINE ? (BOX ? R24_STRIDE_AFF_BOX : R24_STRIDE_AFF) : (FUSED ? R24_STRIDE_LQR_FUSED : R24_STRIDE_LQR);
}
template <int MODE, bool FUSED, bool BOX = false>
inline size_t ric24_smem_bytes(int warps) {
  return sizeof(double) *
         (size_t)(R24_WS_DOUBLES + R24_QF_DOUBLES + R24_BAR_DOUBLES + warps * 2 * r24_stride<MODE, FUSED, BOX>());
}

// M x M square-root-free Cholesky in registers (oracle: chol_lower / chol_solve_neg), fully unrolled
template <int M>
struct LdlReg {
  double L[M][M];  // strictly lower part used
  double D[M], r[M];
  // U: the matrix in shared memory, row-major M x M; only the lower triangle U[i*M+j], j <= i, is read
  __device__ __forceinline__ bool factor(const double* U) {
    bool bad = false;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double v[M];
#pragma unroll
      for (int l = 0; l < j; ++l) v[l] = L[j][l] * D[l];
      double s = U[j * M + j];
#pragma unroll
      for (int l = 0; l < j; ++l) s = fmad(-L[j][l], v[l], s);
      bad |= !(s > R12_PIVOT_MIN) || !(s < R12_PIVOT_MAX);
      D[j] = s;
      r[j] = rcp_rn_normal(s);
#pragma unroll
      for (int i = j + 1; i < M; ++i) {
        double tt = U[i * M + j];
#pragma unroll
        for (int l = 0; l < j; ++l) tt = fmad(-L[i][l], v[l], tt);
        L[i][j] = tt * r[j];
      }
    }
    return bad;
  }
  __device__ __forceinline__ void solve_neg(const double (&rhs)[M], double (&out)[M]) const {
    double y[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double s = rhs[i];
#pragma unroll
      for (int l = 0; l < i; ++l) s = fmad(-L[i][l], y[l], s);
      y[i] = s;
    }
    // back substitution on z, negated at the end like the oracle (a chain on -z_i flips the sign of exact zeros)
    double z[M];
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
      double s = y[i] * r[i];
#pragma unroll
      for (int l = i + 1; l < M; ++l) s = fmad(-L[l][i], z[l], s);
      z[i] = s;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = neg_bits(z[i]);
  }
};

// ---- input box: the oracle's box_qp for any M.  A rare, divergent path (only the instances whose
// feed-forward leaves the box run it), with dynamic loops, so its operands cannot live in registers.  They live in
// SHARED memory — `ws`, a per-instance scratch of QP_WS_DOUBLES(M) doubles (R24_QP, only in the BOX instantiation's
// slice) — not in local memory: all lanes of an instance run the QP redundantly on identical values, so every access is
// a broadcast (one wavefront), every lane reads back what it (and, identically, its neighbours) wrote, and no
// synchronisation is needed.  (Round 1 kept them in thread-local arrays: 16 lanes x 1 KB of L1-backed local memory
// per instance made the boxed n=24 sweep 3.9x slower than the unconstrained one, profiles/ncu_r01_k_ric12_affine_box.txt.)
// U: Quu, row-major M x M, lower triangle valid (shared memory).
template <int M>
__host__ __device__ constexpr int QP_WS_DOUBLES() { return M * M + 6 * M; }
template <int M>
__device__ __forceinline__ void ldl_mask(const double* U, const bool (&cl)[M], double* Um) {
  for (int a = 0; a < M; ++a)
    for (int b = 0; b <= a; ++b) Um[a * M + b] = (cl[a] || cl[b]) ? (a == b ? 1.0 : 0.0) : U[a * M + b];
}
template <int M>
__device__ __forceinline__ double qp_eval_gen(const double* U, const double (&g)[M], const double* x, double* r) {
  for (int a = 0; a < M; ++a) {
    double s = 0.0;
    for (int b = 0; b < M; ++b) s = fmad(b <= a ? U[a * M + b] : U[b * M + a], x[b], s);
    r[a] = s;
  }
  double f = 0.0;
  for (int a = 0; a < M; ++a) f = fmad(x[a], fmad(0.5, r[a], g[a]), f);
  return f;
}
template <int M>
__device__ __forceinline__ void box_qp_gen(const double* U, const double (&g)[M], const double (&lo)[M], const double (&hi)[M],
                                        double (&xio)[M], bool (&cl)[M], double* ws) {
  double* Um = ws;                 // M x M masked matrix
  double* r = ws + M * M;          // U x
  double* grad = r + M;
  double* dx = grad + M;
  double* xc = dx + M;
  double* rc = xc + M;
  double* x = rc + M;              // the iterate
  double gmax = 0.0;
  for (int a = 0; a < M; ++a) {
    x[a] = r12_clamp(xio[a], lo[a], hi[a]);
    if (fabs(g[a]) > gmax) gmax = fabs(g[a]);
  }
  double f = qp_eval_gen<M>(U, g, x, r);
  auto finish = [&]() {
#pragma unroll
    for (int a = 0; a < M; ++a) xio[a] = x[a];
  };
  for (int it = 0; it < R12_QP_MAXIT; ++it) {
    int nfree = 0;
    double gn = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      const double ga = r[a] + g[a];
      grad[a] = ga;
      cl[a] = (x[a] <= lo[a] && ga > 0.0) || (x[a] >= hi[a] && ga < 0.0);
      if (!cl[a]) { ++nfree; if (fabs(ga) > gn) gn = fabs(ga); }
    }
    if (nfree == 0 || gn < R12_QP_GTOL * (1.0 + gmax)) { finish(); return; }
    ldl_mask<M>(U, cl, Um);
    LdlReg<M> chm;
    if (chm.factor(Um)) { finish(); return; }
    double rhs[M], dxr[M];
#pragma unroll
    for (int a = 0; a < M; ++a) rhs[a] = cl[a] ? 0.0 : grad[a];
    chm.solve_neg(rhs, dxr);
    double sdotg = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) { dx[a] = dxr[a]; sdotg = fmad(grad[a], dxr[a], sdotg); }
    if (!(sdotg < 0.0)) { finish(); return; }
    double step = 1.0;
    bool ok = false;
    for (int ls = 0; ls < R12_QP_MAXLS; ++ls) {
      for (int a = 0; a < M; ++a) xc[a] = r12_clamp(fmad(step, dx[a], x[a]), lo[a], hi[a]);
      const double fc = qp_eval_gen<M>(U, g, xc, rc);
      if (fc - f <= R12_QP_ARMIJO * (step * sdotg)) {
        for (int a = 0; a < M; ++a) { x[a] = xc[a]; r[a] = rc[a]; }
        f = fc;
        ok = true;
        break;
      }
      step = step * 0.5;
    }
    if (!ok) { finish(); return; }
  }
#pragma unroll
  for (int a = 0; a < M; ++a) {   // iteration cap: the clamped set of the last accepted point
    const double ga = r[a] + g[a];
    cl[a] = (x[a] <= lo[a] && ga > 0.0) || (x[a] >= hi[a] && ga < 0.0);
  }
  finish();
}

template <int LAYOUT>
__device__ __forceinline__ int r24_foff(int l, int c) {
  if (LAYOUT == 1) return l * 32 + c;
  return c < 24 ? l * 24 + c : 576 + l * 8 + (c - 24);
}

template <int ROWS>
__device__ __forceinline__ void r24_store_mirror(double* Ps, const double* h, int j, bool on) {
#pragma unroll
  for (int i = 0; i < ROWS; i += 2)
    if (on && i <= j) sts128(Ps + j * 24 + i, h[i], h[i + 1]);
}
template <int ROWS>
__device__ __forceinline__ void r24_store_direct(double* Ps, const double* h, int j, bool on) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i)
    if (on && i < j) Ps[i * 24 + j] = h[i];
}

// FUSED (as in riccati12.cuh): the record is computed in the kernel from xbar_k, ubar_k instead of being read from
// HBM.  Lanes 0..7 of an instance evaluate vehicle 0's Jacobian, lanes 8..15 vehicle 1's (same code, different
// data); each stores one eighth of the state-dependent entries.  Zeros, the state-independent entries and the
// (constant) spring-damper coupling entries are written once per sweep.
template <int LAYOUT, int MODE, bool FUSED, int WARPS, bool BOX = false>
__global__ void __launch_bounds__(WARPS * 32, 1) k_ric24(const Ric24Args a) {
  extern __shared__ __align__(128) double smem[];
  static_assert(!FUSED || LAYOUT == 0, "fused linearisation: A_THEN_B records");
  constexpr int STRIDE = r24_stride<MODE, FUSED, BOX>();
  double* Ws = smem;
  double* Qfs = smem + R24_WS_DOUBLES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 4, t = lane & 15;
  double* base = smem + R24_WS_DOUBLES + R24_QF_DOUBLES + R24_BAR_DOUBLES + (warp * 2 + q) * STRIDE;
  double* Fs = base + R24_FS;
  double* Ps = base + R24_PS;
  double* Hs = base + R24_HS;
  double* Us = base + R24_US;
  const unsigned bar = smem_u32(smem + R24_WS_DOUBLES + R24_QF_DOUBLES) + 8u * warp;

  for (int i = threadIdx.x; i < 1024; i += WARPS * 32) Ws[(i >> 5) * R24_WLD + (i & 31)] = a.W[i];
  for (int i = threadIdx.x; i < 576; i += WARPS * 32) Qfs[i] = a.Qf[i];
  if (lane == 0) mbar_init(bar, 1);
  mbar_fence_init();
  __syncthreads();

  const long long w_raw = ((long long)blockIdx.x * WARPS + warp) * 2 + q;
  const bool valid = w_raw < a.batch;
  const long long w = valid ? w_raw : a.batch - 1;
  const long long b = (MODE == RIC_AFFINE && a.idx) ? (long long)a.idx[w] : w;
  const int N = a.N;

  const bool xl = t < 12;                         // x-lane?
  const int cA = xl ? t : 24 + (t - 12);          // own F / G / H columns
  const int cB = xl ? 23 - t : 24 + (19 - t);
  const int jA = xl ? t : 0, jB = xl ? 23 - t : 0;   // x-column numbers (x-lanes)
  const int uA = t - 12, uB = 19 - t;                // u-column numbers (u-lanes)
  // F[l][c] = Fs[off + l*str] for the two own columns
  const int offA = (LAYOUT == 1) ? cA : (xl ? cA : 576 + (cA - 24));
  const int offB = (LAYOUT == 1) ? cB : (xl ? cB : 576 + (cB - 24));
  const int strAB = (LAYOUT == 1) ? 32 : (xl ? 24 : 8);

  const double* rec = FUSED ? nullptr : a.AB + (size_t)b * N * R24_REC;
  const unsigned fs_u32 = smem_u32(Fs);
  unsigned phase = 0;

  const size_t xsb = a.xbar_sb ? (size_t)a.xbar_sb : (size_t)(N + 1) * 24, xsk = a.xbar_sk ? (size_t)a.xbar_sk : 24;
  auto fetch = [&](int k) {
    if (!FUSED) {
      if (lane == 0) mbar_expect_tx(bar, 2u * 6144u);
      __syncwarp();
      if (t == 0) bulk_g2s(fs_u32, rec + (size_t)k * R24_REC, 6144u, bar);
    }
    if (FUSED || MODE == RIC_AFFINE) {
      // xbar_k (192 B) | ubar_k (64 B, optional): sixteen 16-byte chunks, one per lane
      const double* xs = a.xbar + (size_t)b * xsb + (size_t)k * xsk;
      if (t < 12) cp_async16(smem_u32(base + R24_XS + 2 * t), xs + 2 * t);
      else if (a.ubar) cp_async16(smem_u32(base + R24_XS + 24 + 2 * (t - 12)), a.ubar + ((size_t)b * N + k) * 8 + 2 * (t - 12));
      cp_async_commit();
    }
  };

  double mu = 0.0;
  if (MODE == RIC_AFFINE) mu = a.mu[b];
  bool failed = false, reg_fail = false;
  double dV1 = 0.0, dV2 = 0.0;

  for (;;) {  // (re)start of a sweep
    failed = false;
    dV1 = 0.0; dV2 = 0.0;
    int kfail = -1;
    double* Kp = (a.K && valid) ? a.K + ((size_t)b * N + (N - 1)) * 192 : nullptr;
    for (int e = t; e < 576; e += 16) Ps[e] = Qfs[e];
    if (FUSED) {
      for (int e = t; e < R24_REC; e += 16) Fs[e] = 0.0;
      __syncwarp();
      if (t == 0) {   // state-independent entries: per-vehicle constants first, then the coupling (it overwrites (3+i,3+i))
        for (int v = 0; v < 2; ++v)
          model::quad_jac_constants(a.dt, [&](int i, int j, double val) { Fs[(12 * v + i) * 24 + 12 * v + j] = val; },
                                    [&](int i, int j, double val) { Fs[576 + (12 * v + i) * 8 + 4 * v + j] = val; });
        model::dual_coupling_jac(a.dt, [&](int i, int j, double val) { Fs[i * 24 + j] = val; });
      }
    }
    if (MODE == RIC_AFFINE) {
      // xgoal -> shared; Vx <- Qf (xbar_N - xgoal): lane t computes rows t and t+16 (<24)
      for (int i = t; i < 24; i += 16) base[R24_XG + i] = a.xgoal[(size_t)b * 24 + i];
      __syncwarp();
      const double* xN = a.xbar + (size_t)b * xsb + (size_t)N * xsk;
      for (int r = t; r < 24; r += 16) {
        double s = 0.0;
        for (int j = 0; j < 24; ++j) s = fmad(Qfs[r * 24 + j], xN[j] - base[R24_XG + j], s);
        base[R24_VS + r] = s;
      }
    }
    fetch(N - 1);

    for (int it = 0; it < N; ++it) {
      const int k = N - 1 - it;
      if (!FUSED) {
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
      if (FUSED || MODE == RIC_AFFINE) cp_async_wait_all();
      __syncwarp();
      if (FUSED) {
        // a1 in the sweep: vehicle v = t>>3 of this instance, lane t&7 stores every eighth state-dependent entry
        const int v = t >> 3, t8 = t & 7;
        const double* Xs = base + R24_XS;
        double xk[12], uk[4];
#pragma unroll
        for (int i = 0; i < 12; ++i) xk[i] = Xs[12 * v + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) uk[c] = a.ubar ? Xs[24 + 4 * v + c] : model::kHover;
        double* FA = Fs + (12 * v) * 24 + 12 * v;
        double* FB = Fs + 576 + (12 * v) * 8 + 4 * v;
        int e = 0;
        model::quad_jac_discrete(
            xk, uk, a.dt,
            [&](int i, int j, double val) {
              if (!model::quad_jac_is_const_A(i, j) && !model::dual_coupled_diag(i, j) && ((e++) & 7) == t8) FA[i * 24 + j] = val;
            },
            [&](int i, int j, double val) {
              if (!model::quad_jac_is_const_B(i, j) && ((e++) & 7) == t8) FB[i * 8 + j] = val;
            });
      }

      double hacc[2] = {0.0, 0.0};
      if (MODE == RIC_AFFINE) {
        // e = [xbar_k - xgoal ; ubar_k - u_h] -> shared, then h_c = W[c][:] e for the own columns (W is block
        // diagonal, so the foreign block contributes exact zeros to the chain)
        const double* Xs = base + R24_XS;
        base[R24_ES + t] = Xs[t] - base[R24_XG + t];
        if (t < 8) base[R24_ES + 16 + t] = Xs[16 + t] - base[R24_XG + 16 + t];
        else base[R24_ES + 16 + t] = Xs[16 + t] - model::kHover;
        __syncwarp();
        const double* wa = Ws + cA * R24_WLD;
        const double* wb = Ws + cB * R24_WLD;
        const int lo = xl ? 0 : 24, hi = xl ? 24 : 32;   // own block of the row (the rest is exact zeros)
        double sa = 0.0, sb = 0.0;
        for (int i = lo; i < hi; ++i) {
          const double e = base[R24_ES + i];
          sa = fmad(wa[i], e, sa);
          sb = fmad(wb[i], e, sb);
        }
        hacc[0] = sa; hacc[1] = sb;
      }

      if (FUSED) {
        __syncwarp();  // the record is complete and every lane has consumed xbar_k, ubar_k
        if (it + 1 < N) fetch(k - 1);
      }
      // ---------------- phase 1: G[:,c] = P F[:,c], own two columns
      double g[24][2];
#pragma unroll
      for (int r = 0; r < 24; ++r) { g[r][0] = 0.0; g[r][1] = 0.0; }
      if (!FUSED) {
#pragma unroll 4
        for (int l = 0; l < 24; ++l) {
          const double fa = Fs[offA + l * strAB];
          const double fb = Fs[offB + l * strAB];
#pragma unroll
          for (int r = 0; r < 24; r += 2) {
            const double2 v = lds128(Ps + l * 24 + r);
            g[r][0] = fmad(v.x, fa, g[r][0]);
            g[r][1] = fmad(v.x, fb, g[r][1]);
            g[r + 1][0] = fmad(v.y, fa, g[r + 1][0]);
            g[r + 1][1] = fmad(v.y, fb, g[r + 1][1]);
          }
          if (MODE == RIC_AFFINE) {
            const double vx = base[R24_VS + l];
            hacc[0] = fmad(fa, vx, hacc[0]);
            hacc[1] = fmad(fb, vx, hacc[1]);
          }
        }
      } else {
        // built-in dual model: slot A holds a column of vehicle 0 (x or u) in every lane, slot B one of vehicle 1;
        // rows of the other vehicle are structurally zero there except its three spring-damper rows, so 9 of the
        // 24 (l, slot) products are skipped by all lanes alike
#pragma unroll
        for (int l = 0; l < 24; ++l) {
          const bool nzA = l < 12 || (l >= 15 && l < 18), nzB = l >= 12 || (l >= 3 && l < 6);
          const double fa = nzA ? Fs[offA + l * strAB] : 0.0;
          const double fb = nzB ? Fs[offB + l * strAB] : 0.0;
#pragma unroll
          for (int r = 0; r < 24; r += 2) {
            const double2 v = lds128(Ps + l * 24 + r);
            if (nzA) { g[r][0] = fmad(v.x, fa, g[r][0]); g[r + 1][0] = fmad(v.y, fa, g[r + 1][0]); }
            if (nzB) { g[r][1] = fmad(v.x, fb, g[r][1]); g[r + 1][1] = fmad(v.y, fb, g[r + 1][1]); }
          }
          if (MODE == RIC_AFFINE) {
            const double vx = base[R24_VS + l];
            if (nzA) hacc[0] = fmad(fa, vx, hacc[0]);
            if (nzB) hacc[1] = fmad(fb, vx, hacc[1]);
          }
        }
      }

      // ---------------- phase 2a: the eight u-rows of H = W + F^T G for the own columns (Hux for x-lanes, Huu for
      // u-lanes).  A separate pass with its own 16 accumulators: together with the 36 Hxx accumulators and the 48
      // G values the single-pass version needed 200+ registers and spilled (ncu: long-scoreboard stalls on local
      // loads, profiles/ncu_r01_k_ric24_v1.txt).  The results go to shared memory at once and are re-read as the
      // right-hand sides of the gain solves.
      {
        double huA[8], huB[8];
        const double* wa = Ws + cA * R24_WLD + 24;
        const double* wb = Ws + cB * R24_WLD + 24;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const double2 va = lds128(wa + i);
          const double2 vb = lds128(wb + i);
          huA[i] = va.x; huA[i + 1] = va.y; huB[i] = vb.x; huB[i + 1] = vb.y;
        }
#pragma unroll
        for (int l = 0; l < 24; ++l) {
          const double ga = g[l][0], gb = g[l][1];
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const bool n0 = !FUSED || model::dual_jac_nz_B(l, c), n1 = !FUSED || model::dual_jac_nz_B(l, c + 1);
            if (!n0 && !n1) continue;
            const double2 v = lds128(Fs + r24_foff<LAYOUT>(l, 24 + c));
            if (n0) { huA[c] = fmad(v.x, ga, huA[c]); huB[c] = fmad(v.x, gb, huB[c]); }
            if (n1) { huA[c + 1] = fmad(v.y, ga, huA[c + 1]); huB[c + 1] = fmad(v.y, gb, huB[c + 1]); }
          }
        }
        if (MODE == RIC_AFFINE && !xl) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c == uA) huA[c] = huA[c] + mu;
            if (c == uB) huB[c] = huB[c] + mu;
          }
        }
        // exchange: Hux columns (x-lanes) and Huu columns (u-lanes) -> shared
        double* dA = xl ? Hs + jA : Us + uA;
        double* dB = xl ? Hs + jB : Us + uB;
        const int ds = xl ? 24 : 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) { dA[c * ds] = huA[c]; dB[c * ds] = huB[c]; }
        if (MODE == RIC_AFFINE && !xl) { base[R24_HU + uA] = hacc[0]; base[R24_HU + uB] = hacc[1]; }
      }

      // ---------------- phase 2b: Hxx rows of the own x-columns
      {
      double hA[12], hB[24];
      {
        const double* wa = Ws + cA * R24_WLD;
        const double* wb = Ws + cB * R24_WLD;
#pragma unroll
        for (int i = 0; i < 12; i += 2) { const double2 v = lds128(wa + i); hA[i] = v.x; hA[i + 1] = v.y; }
#pragma unroll
        for (int i = 0; i < 24; i += 2) { const double2 v = lds128(wb + i); hB[i] = v.x; hB[i + 1] = v.y; }
      }
#pragma unroll
      for (int l = 0; l < 24; ++l) {
        const double ga = g[l][0], gb = g[l][1];
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          const bool n0 = !FUSED || model::dual_jac_nz_A(l, i), n1 = !FUSED || model::dual_jac_nz_A(l, i + 1);
          if (!n0 && !n1) continue;
          const double2 v = lds128(Fs + r24_foff<LAYOUT>(l, i));
          if (n0) { hA[i] = fmad(v.x, ga, hA[i]); hB[i] = fmad(v.x, gb, hB[i]); }
          if (n1) { hA[i + 1] = fmad(v.y, ga, hA[i + 1]); hB[i + 1] = fmad(v.y, gb, hB[i + 1]); }
        }
#pragma unroll
        for (int i = 12; i < 24; i += 2) {
          const bool n0 = !FUSED || model::dual_jac_nz_A(l, i), n1 = !FUSED || model::dual_jac_nz_A(l, i + 1);
          if (!n0 && !n1) continue;
          const double2 v = lds128(Fs + r24_foff<LAYOUT>(l, i));
          if (n0) hB[i] = fmad(v.x, gb, hB[i]);
          if (n1) hB[i + 1] = fmad(v.y, gb, hB[i + 1]);
        }
      }
      __syncwarp();  // every lane is done with Fs (and Xs, Es, Vs) and with P; the exchanged Hux / Huu are visible
      // The 36 Hxx accumulators wait for phase 4 in shared memory, in the slot of P (dead between phase 1 and phase 4;
      // 36 x 16 lanes = its 576 doubles exactly), lane-contiguous: with them in registers next to the factorisation
      // (88), the solves and the affine / box operands phase 3 spilled ~200-550 B per thread to local memory, and with
      // the shared-memory carve-out at its maximum the seven warps' spill frames do not fit what is left of L1
      // (profiles/ncu_r01_k_ric24_affine_fused_sparse.txt: 9.5 M local loads, long-scoreboard stalls).
#pragma unroll
      for (int i = 0; i < 12; i += 2) sts128(Ps + (i >> 1) * 32 + 2 * t, hA[i], hA[i + 1]);
#pragma unroll
      for (int i = 0; i < 24; i += 2) sts128(Ps + (6 + (i >> 1)) * 32 + 2 * t, hB[i], hB[i + 1]);
      }
      if (!FUSED && it + 1 < N) fetch(k - 1);

      // ---------------- phase 3: redundant 8x8 L D L^T, gains of the own x-columns.  The right-hand sides are read
      // back from shared memory one at a time, just before their solve, to keep the live register set small
      // (x-lanes: their Hux columns; u-lanes solve dummies).
      LdlReg<8> ch;
      failed |= ch.factor(Us);
      bool cl[8] = {false, false, false, false, false, false, false, false};   // BOX: inputs held at a bound in this step
      bool boxed_step = false;
      const double* sA = xl ? Hs + jA : Us + uA;
      const double* sB = xl ? Hs + jB : Us + uB;
      const int ds = xl ? 24 : 8;
      if (MODE == RIC_AFFINE) {
        double hur[8], kf[8];
#pragma unroll
        for (int c = 0; c < 8; c += 2) { const double2 v = lds128(base + R24_HU + c); hur[c] = v.x; hur[c + 1] = v.y; }
        ch.solve_neg(hur, kf);
        if (BOX) {
          // input box (oracle: noc_oracle_backward_affine_box): ubar_k straight from global memory (the staged copy
          // has been overwritten by the prefetch of step k-1); a feed-forward that leaves the box is replaced by the
          // box QP's, and `ch` by the factorisation with the clamped inputs masked out, which the gain solves use
          double lo[8], hi[8];
          bool inside = true;
          const double* ub = a.ubar + ((size_t)b * N + k) * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double uc = ub[c];
            lo[c] = a.u_min - uc;
            hi[c] = a.u_max - uc;
            if (!(kf[c] > lo[c] && kf[c] < hi[c])) inside = false;
          }
          if (!inside) {
            static_assert(QP_WS_DOUBLES<8>() <= R24_STRIDE_AFF_BOX - R24_QP, "QP workspace");
            double* ws = base + R24_QP;
            box_qp_gen<8>(Us, hur, lo, hi, kf, cl, ws);
            ldl_mask<8>(Us, cl, ws);
            failed |= ch.factor(ws);
            boxed_step = true;
          }
        }
        if (xl) {
          double sa = hacc[0], sb = hacc[1];
#pragma unroll
          for (int c = 0; c < 8; ++c) { sa = fmad(sA[c * ds], kf[c], sa); sb = fmad(sB[c * ds], kf[c], sb); }
          base[R24_VS + jA] = sa;
          base[R24_VS + jB] = sb;
        }
        double t1 = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) t1 = fmad(kf[c], hur[c], t1);
        double t2 = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          double rr = 0.0;
#pragma unroll
          for (int c = 0; c < 8; ++c) rr = fmad(c <= r ? Us[r * 8 + c] : Us[c * 8 + r], kf[c], rr);
          t2 = fmad(kf[r], rr, t2);
        }
        dV1 = dV1 + t1;
        dV2 = dV2 + 0.5 * t2;
        if (t == 0 && valid) {
          double* fp = a.kff + ((size_t)b * N + k) * 8;
#pragma unroll
          for (int c = 0; c < 8; c += 2) stg128_cs(fp + c, kf[c], kf[c + 1]);
        }
      }
      double kA[8], kB[8];
      {
        double rhs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rhs[c] = (BOX && boxed_step && cl[c]) ? 0.0 : sA[c * ds];
        ch.solve_neg(rhs, kA);
#pragma unroll
        for (int c = 0; c < 8; ++c) rhs[c] = (BOX && boxed_step && cl[c]) ? 0.0 : sB[c * ds];
        ch.solve_neg(rhs, kB);
      }
      if (failed && kfail < 0) kfail = k;
      if (Kp) {
        if (xl) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { __stcs(Kp + c * 24 + jA, kA[c]); __stcs(Kp + c * 24 + jB, kB[c]); }
        }
        Kp -= 192;
      }
      if (k == 0 && a.K0 && valid && xl) {
        double* kp = a.K0 + (size_t)b * 192;
#pragma unroll
        for (int c = 0; c < 8; ++c) { kp[c * 24 + jA] = kA[c]; kp[c * 24 + jB] = kB[c]; }
      }

      // ---------------- phase 4: P'[i][j] = Hxx[i][j] + sum_a Hux[a][i] K[a][j], i <= j, own x-columns
      double hA[12], hB[24];
#pragma unroll
      for (int i = 0; i < 12; i += 2) { const double2 v = lds128(Ps + (i >> 1) * 32 + 2 * t); hA[i] = v.x; hA[i + 1] = v.y; }
#pragma unroll
      for (int i = 0; i < 24; i += 2) { const double2 v = lds128(Ps + (6 + (i >> 1)) * 32 + 2 * t); hB[i] = v.x; hB[i + 1] = v.y; }
      __syncwarp();  // every lane has its accumulators back before P' overwrites the slot
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double ka = kA[c], kb = kB[c];
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          const double2 v = lds128(Hs + c * 24 + i);
          hA[i] = fmad(v.x, ka, hA[i]); hA[i + 1] = fmad(v.y, ka, hA[i + 1]);
          hB[i] = fmad(v.x, kb, hB[i]); hB[i + 1] = fmad(v.y, kb, hB[i + 1]);
        }
#pragma unroll
        for (int i = 12; i < 24; i += 2) {
          const double2 v = lds128(Hs + c * 24 + i);
          hB[i] = fmad(v.x, kb, hB[i]); hB[i + 1] = fmad(v.y, kb, hB[i + 1]);
        }
      }
      r24_store_mirror<12>(Ps, hA, jA, xl);
      r24_store_mirror<24>(Ps, hB, jB, xl);
      __syncwarp();
      r24_store_direct<12>(Ps, hA, jA, xl);
      r24_store_direct<24>(Ps, hB, jB, xl);
    }  // steps
    __syncwarp();
    if (MODE == RIC_LQR && failed && valid && xl) {
      const double nanv = qnan();
      if (a.K)
        for (int k2 = 0; k2 <= kfail; ++k2) {
          double* kp = a.K + ((size_t)b * N + k2) * 192;
          for (int c = 0; c < 8; ++c) { kp[c * 24 + jA] = nanv; kp[c * 24 + jB] = nanv; }
        }
      if (a.K0) {
        double* kp = a.K0 + (size_t)b * 192;
        for (int c = 0; c < 8; ++c) { kp[c * 24 + jA] = nanv; kp[c * 24 + jB] = nanv; }
      }
    }
    if (MODE != RIC_AFFINE) break;
    // (the adaptive schedule's factor lives in global memory: it is touched once per sweep, not worth a register pair)
    double dl = (a.reg_schedule == 1) ? a.delta[b] : 1.0;
    __syncwarp();
    if (failed && !reg_fail) {
      if (a.reg_schedule == 1) {
        reg_increase(mu, dl);
        if (t == 0 && valid) a.delta[b] = dl;
      } else {
        mu = (mu * 10.0 > 1e-6) ? mu * 10.0 : 1e-6;
      }
      if (mu > 1e6) reg_fail = true;
    }
    if (!__any_sync(0xffffffffu, failed && !reg_fail)) break;
  }

  if (valid) {
    const double nanv = qnan();
    if (MODE == RIC_AFFINE) {
      if (t == 0) {
        a.mu[b] = mu;
        a.dV[(size_t)b * 2] = dV1;
        a.dV[(size_t)b * 2 + 1] = dV2;
        if (a.status) a.status[b] = reg_fail ? 4 : 0;
      }
      return;
    }
    if (a.P0) {
      double* po = a.P0 + (size_t)b * 576;
      for (int e = t; e < 576; e += 16) po[e] = failed ? nanv : Ps[e];
    }
    if (t == 0) {
      if (a.cost && a.x0) {
        const double* x = a.x0 + (size_t)b * 24;
        double c = 0.0;
        for (int i = 0; i < 24; ++i) {
          double r = 0.0;
          for (int j = 0; j < 24; ++j) r = fmad(Ps[i * 24 + j], x[j], r);
          c = fmad(x[i], r, c);
        }
        a.cost[b] = failed ? nanv : 0.5 * c;
      }
      if (a.status) a.status[b] = failed ? 1 : 0;
    }
  }
}

}  // namespace noc

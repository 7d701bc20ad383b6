// riccati12_wpi.cuh — the SLQ backward pass for n=12, m=4 with ONE instance per warp and its products on the FP64 tensor
// cores: the kernel of SHORT work lists (the tail of an SLQ solve, small MPC batches).  Same spec, same bits as
// k_ric12<RIC_AFFINE, FUSED> and the oracle; every SLQ parity test of
// tests/test_gpu_dare_slq_parity.py runs against both.
//
// Why: k_ric12 gives four lanes to an instance and eight instances to a warp — right for throughput, but a warp-step is
// 1,525 instructions (2.2 us) however few instances the list holds, and the last 30 of config 3's 50 iterations run on a
// few hundred stragglers: a chain of 200 dependent steps per sweep that nothing overlaps.  With the whole warp on one
// instance a step is 24 DMMAs plus a distributed 4x4 factorisation.
//
// Everything is indexed by POSITION in the permuted order [u3 u2 u1 u0 | x0 .. x11] (16 = two 8x8 tiles), the order in
// which every entry the spec needs lies in the upper triangle of H' (riccati12_mma.cuh).  P is kept as P~ = the 16 x 16
// matrix with P in positions 4..15 (upper tiles (0,0), (0,1), (1,1)); its u-rows and u-columns hold dead values that only
// ever meet dead rows of G~':
//   G~' = P~[:, 4..15] F'      M = 16 (two tiles, nothing padded), k = the twelve state columns: 12 DMMAs
//   H'  = W~' + F'^T G~'[4..15, :]   three upper tiles, 9 DMMAs; H' IS P~'s tile grid, so
//   P~' = H' + Hux^T K~        accumulates in place on the H' fragments (3 DMMAs, one k-step: the four inputs)
// and Huu / Hux are rows 0..3 of tiles (0,0) / (0,1).  L D L^T of Huu: m24_ldl_pivot<L, 4> on the C fragments of tile (0,0)
// (lanes g < 4, t < 2); gains: lanes 4..15 back-substitute their state column, lanes 0..3 the feed-forward.
// The step records (Jacobian entries, h0 = W e, ubar) come from k_prep12, as for n=24 (riccati24_mma.cuh).
#pragma once
#include "common.cuh"
#include "model.cuh"
#include "riccati12.cuh"
#include "riccati12_mma.cuh"   // rm_dmma, rm_elect_one, rm_perm
#include "riccati24_mma.cuh"   // m24_el, m24_ldl_pivot, m24_solve_neg

namespace noc {

constexpr int W12_JE = 37;                                        // state-dependent Jacobian entries of the quadrotor
constexpr int W12_PREP_H0 = W12_JE, W12_PREP_U = W12_PREP_H0 + 16, W12_PREP_DOUBLES = W12_PREP_U + 4 + 1;   // 58 doubles = 464 B
// per-warp (= per-instance) shared memory, in doubles
constexpr int W12_FS = 0;        // record A[12][12] | B[12][4]
constexpr int W12_PT = 192;      // P~: upper tiles (0,0) (0,1) (1,1)
constexpr int W12_HT = 384;      // H' tiles (0,0) (0,1) for the scalar phase (Huu, Hux: their rows 0..3)
constexpr int W12_GB = 512;      // G~': four tiles; scalar phase: K~[4][16]
constexpr int W12_US = 768;      // Huu[4][4] row-major (lower triangle valid)
constexpr int W12_LS = 784;      // columns of L: Ls[i*4 + l]
constexpr int W12_PR = 800;      // two step records
constexpr int W12_VS = W12_PR + 2 * W12_PREP_DOUBLES;   // Vx[12]
constexpr int W12_HV = W12_VS + 12, W12_KF = W12_HV + 4, W12_RR = W12_KF + 4, W12_DOUBLES = W12_RR + 4;   // 940
constexpr int W12_WT_DOUBLES = 192;   // per CTA: W~' tiles (0,0) (0,1) (1,1) in C-fragment order
constexpr int W12_BAR_DOUBLES = 8;
inline size_t ric12_wpi_smem_bytes(int warps) { return sizeof(double) * (size_t)(W12_BAR_DOUBLES + W12_WT_DOUBLES + warps * W12_DOUBLES); }

// k_prep12: two threads per (work-list slot w, step k) — one evaluates the Jacobian, the other h0 and ubar (different
// warps) — 64 records per CTA, assembled in shared memory and written out as one contiguous block (as k_prep24).
constexpr int PREP12_RECS = 64, PREP12_THREADS = 2 * PREP12_RECS;
constexpr size_t PREP12_SMEM = sizeof(double) * (256 + PREP12_RECS * W12_PREP_DOUBLES);
__global__ void __launch_bounds__(PREP12_THREADS) k_prep12(const Ric12Args a, double* __restrict__ prep) {
  extern __shared__ __align__(16) double psm12[];
  double* Ws = psm12;
  double* so = psm12 + 256;
  for (int e = threadIdx.x; e < 256; e += PREP12_THREADS) Ws[e] = a.W[e];
  __syncthreads();
  const int N = a.N;
  const long long total = a.batch * N;
  const long long wk0 = (long long)blockIdx.x * PREP12_RECS;
  const int role = threadIdx.x / PREP12_RECS, pr = threadIdx.x % PREP12_RECS;
  const long long wk = wk0 + pr;
  if (wk < total) {
    const long long w = wk / N;
    const int k = (int)(wk - w * N);
    const long long b = a.idx ? (long long)a.idx[w] : w;
    const size_t xsb = a.xbar_sb ? (size_t)a.xbar_sb : (size_t)(N + 1) * 12, xsk = a.xbar_sk ? (size_t)a.xbar_sk : 12;
    const double* xs = a.xbar + (size_t)b * xsb + (size_t)k * xsk;
    const double* us = a.ubar + ((size_t)b * N + k) * 4;
    double* out = so + pr * W12_PREP_DOUBLES;
    double xk[12], uk[4];
#pragma unroll
    for (int i = 0; i < 12; i += 2) { const double2 q = __ldg(reinterpret_cast<const double2*>(xs + i)); xk[i] = q.x; xk[i + 1] = q.y; }
#pragma unroll
    for (int c = 0; c < 4; c += 2) { const double2 q = __ldg(reinterpret_cast<const double2*>(us + c)); uk[c] = q.x; uk[c + 1] = q.y; }
    if (role == 0) {
      int e = 0;
      model::quad_jac_discrete(
          xk, uk, a.dt, [&](int i, int jj, double val) { if (!model::quad_jac_is_const_A(i, jj)) out[e++] = val; },
          [&](int i, int jj, double val) { if (!model::quad_jac_is_const_B(i, jj)) out[e++] = val; });
    } else {
      // h0 by position: x-column c at 4 + c, u-column a at 3 - a (W block diagonal: the chain runs over the own block)
      const double* xg = a.xgoal + (size_t)b * 12;
      double ex[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) ex[i] = xk[i] - xg[i];
#pragma unroll 2
      for (int c = 0; c < 12; ++c) {
        double h = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) h = fmad(Ws[i * 16 + c], ex[i], h);
        out[W12_PREP_H0 + 4 + c] = h;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double h = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) h = fmad(Ws[(12 + i) * 16 + 12 + c], uk[i] - model::kHover, h);
        out[W12_PREP_H0 + 3 - c] = h;
        out[W12_PREP_U + c] = uk[c];
      }
      out[W12_PREP_DOUBLES - 1] = 0.0;
    }
  }
  __syncthreads();
  const long long left = total - wk0;
  const int cnt = (int)(left < PREP12_RECS ? left : PREP12_RECS) * W12_PREP_DOUBLES;
  double* dst = prep + (size_t)wk0 * W12_PREP_DOUBLES;
  for (int e = 2 * threadIdx.x; e < cnt; e += 2 * PREP12_THREADS) __stcg(reinterpret_cast<double2*>(dst + e), lds128(so + e));
}

template <int WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) k_ric12_wpi(const Ric12Args a, const double* __restrict__ prep) {
  extern __shared__ __align__(128) double smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double* Wt = smem + W12_BAR_DOUBLES;
  double* base = Wt + W12_WT_DOUBLES + warp * W12_DOUBLES;
  double* Fs = base + W12_FS;
  double* Pt = base + W12_PT;
  double* Ht = base + W12_HT;
  double* Gb = base + W12_GB;
  double* Us = base + W12_US;
  double* Ls = base + W12_LS;
  const unsigned bar = smem_u32(smem) + 8u * warp;

  // W~' = W in the permuted order: tiles (0,0), (0,1), (1,1), plain row-major (a lane reads its own 16 bytes)
  for (int e = threadIdx.x; e < W12_WT_DOUBLES; e += WARPS * 32) {
    const int T = e >> 6, r = (e >> 3) & 7, c = e & 7;
    const int mi = T == 2 ? 1 : 0, nj = T >= 1 ? 1 : 0;
    Wt[e] = a.W[rm_perm(8 * mi + r) * 16 + rm_perm(8 * nj + c)];
  }
  if (lane == 0) mbar_init(bar, 1);
  mbar_fence_init();
  __syncthreads();

  const long long w = (long long)blockIdx.x * WARPS + warp;
  if (w >= a.batch) return;
  const long long b = a.idx ? (long long)a.idx[w] : w;
  const int N = a.N;

  // ---- lane constants (tile store of riccati24_mma.cuh: element (r, c) at r*8 + (c ^ 4*((r>>1)&1)))
  const int s_g = (g >> 1) & 1, s_t = (t >> 1) & 1;
  const int cs_off = g * 8 + ((2 * t) ^ (4 * s_g));
  const int dir0 = g * 8 + 4 * (0 ^ s_g) + t, dir1 = g * 8 + 4 * (1 ^ s_g) + t;
  const int mir0 = t * 8 + (g ^ (4 * s_t)), mir1 = (4 + t) * 8 + (g ^ (4 * s_t));
  const int dia0 = (g <= t) ? dir0 : mir0, dia1 = (g <= 4 + t) ? dir1 : mir1;
  // F'[4ki+t][8ni+g]: ni = 0: positions g < 4 are u-columns 3-g (B part: 144 + l*4 + c), g >= 4 x-columns g-4 (A part: l*12 + c)
  const int fo0 = (g < 4) ? 144 + t * 4 + (3 - g) : t * 12 + (g - 4), fs0 = (g < 4) ? 16 : 48;
  const int fo1 = t * 12 + 4 + g;
  // scalar phase: lane p (0..15; lanes 16..31 mirror them and never store) owns position p
  const int p = lane & 15;
  const bool act = lane < 16, xlane = p >= 4;
  const int j = xlane ? p - 4 : 0;
  const double* Hcol = Ht + ((4 + j) >> 3) * 64;     // Hux[a][j] = H'(3-a, 4+j) = Hcol[m24_el(3 - a, (4 + j) & 7)]
  const int jc = (4 + j) & 7;
  const int fcol = xlane ? j : 144 + (3 - p), fcs = xlane ? 12 : 4;   // F[l][own column] = Fs[fcol + l * fcs]
  const int ar = (3 - t) * 8 + (g ^ (4 * (((3 - t) >> 1) & 1)));       // Hux^T A fragment: H'(3-t, 8mi+g) inside tile (0, mi)

  const double* prep_w = prep + (size_t)w * N * W12_PREP_DOUBLES;
  unsigned phase = 0;
  auto fetch_prep = [&](int k) {
    if (rm_elect_one()) {
      mbar_expect_tx(bar, 8u * W12_PREP_DOUBLES);
      bulk_g2s(smem_u32(base + W12_PR + W12_PREP_DOUBLES * (k & 1)), prep_w + (size_t)k * W12_PREP_DOUBLES, 8u * W12_PREP_DOUBLES, bar);
    }
  };
  // destinations of the record's entries lane and lane+32 inside the staged A|B (found once, by replaying the emission order)
  int so0 = 0, so1 = -1;
  {
    const double zx[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, zu[4] = {0, 0, 0, 0};
    int e = 0;
    auto note = [&](int o) {
      if (e == lane) so0 = o;
      if (e == lane + 32) so1 = o;
      ++e;
    };
    model::quad_jac_discrete(
        zx, zu, a.dt, [&](int i, int jj, double) { if (!model::quad_jac_is_const_A(i, jj)) note(i * 12 + jj); },
        [&](int i, int jj, double) { if (!model::quad_jac_is_const_B(i, jj)) note(144 + i * 4 + jj); });
  }
  auto scatter_prep = [&](int k) {
    const double* pr = base + W12_PR + W12_PREP_DOUBLES * (k & 1);
    Fs[so0] = pr[lane];
    if (so1 >= 0) Fs[so1] = pr[lane + 32];
  };

  const size_t xsb = a.xbar_sb ? (size_t)a.xbar_sb : (size_t)(N + 1) * 12, xsk = a.xbar_sk ? (size_t)a.xbar_sk : 12;
  double mu = a.mu[b];
  bool failed = false, reg_fail = false;
  double dV1 = 0.0, dV2 = 0.0;

  for (;;) {   // (re)start of a sweep: again with a larger mu after a failed factorisation
    failed = false;
    dV1 = 0.0; dV2 = 0.0;
    double* Kp = a.K + ((size_t)b * N + (N - 1)) * 48;
    fetch_prep(N - 1);
    // P~ <- Qf in positions 4..15, zeros elsewhere
    {
      const int T3r[3] = {0, 0, 1}, T3c[3] = {0, 1, 1};
#pragma unroll
      for (int T = 0; T < 3; ++T) {
        const int r = 8 * T3r[T] + g, c = 8 * T3c[T] + 2 * t;
        const bool on = r >= 4 && c >= 4;
        const double* q = a.Qf + (on ? (r - 4) * 12 + (c - 4) : 0);
        sts128(Pt + T * 64 + cs_off, on ? q[0] : 0.0, on ? q[1] : 0.0);
      }
    }
    for (int e = lane; e < R12_REC; e += 32) Fs[e] = 0.0;
    __syncwarp();
    if (lane == 0)
      model::quad_jac_constants(a.dt, [&](int i, int jj, double v) { Fs[i * 12 + jj] = v; },
                                [&](int i, int jj, double v) { Fs[144 + i * 4 + jj] = v; });
    {
      // Vx <- Qf (xbar_N - xgoal), one row per lane (row chain, ascending)
      const double* xN = a.xbar + (size_t)b * xsb + (size_t)N * xsk;
      const double* xg = a.xgoal + (size_t)b * 12;
      if (lane < 12) {
        double s = 0.0;
        for (int c = 0; c < 12; ++c) s = fmad(a.Qf[lane * 12 + c], xN[c] - xg[c], s);
        base[W12_VS + lane] = s;
      }
    }
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1u;
    scatter_prep(N - 1);
    __syncwarp();

    for (int it = 0; it < N; ++it) {
      const int k = N - 1 - it;
      const bool more = it + 1 < N;
      // ---------------- fragments: F' (3 k-steps x 2 column tiles) and P~ (2 row tiles x 3 k-steps: columns 4..15)
      double fb[3][2], pa[2][3];
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) { fb[ki][0] = Fs[fo0 + ki * fs0]; fb[ki][1] = Fs[fo1 + ki * 48]; }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          const int tc = ki >= 1 ? 1 : 0, h = ki == 1 ? 0 : 1;   // columns 4+4ki .. : tile column, half
          const double* T = Pt + (mi <= tc ? mi + tc : 1) * 64;    // upper tile (min, max): (0,0)=0, (0,1)=1, (1,1)=2
          pa[mi][ki] = T[mi < tc ? (h ? dir1 : dir0) : (mi > tc ? (h ? mir1 : mir0) : (h ? dia1 : dia0))];
        }
      double hacc = base[W12_PR + W12_PREP_DOUBLES * (k & 1) + W12_PREP_H0 + p];
      __syncwarp();   // every lane has its fragments and h0: the record slot and the other step-record buffer are free
      if (more) fetch_prep(k - 1);
      auto hstep = [&](int l) { hacc = fmad(Fs[fcol + l * fcs], base[W12_VS + l], hacc); };

      // ---------------- G~' = P~[:, 4..15] F'
      double gacc[2][2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) { gacc[mi][ni][0] = 0.0; gacc[mi][ni][1] = 0.0; }
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) rm_dmma(gacc[mi][ni][0], gacc[mi][ni][1], pa[mi][ki], fb[ki][ni]);
        hstep(2 * ki); hstep(2 * ki + 1);          // the F^T Vx chain, ascending l, two steps per k-step: l = 0 .. 5 here
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) sts128(Gb + (mi * 2 + ni) * 64 + cs_off, gacc[mi][ni][0], gacc[mi][ni][1]);
      __syncwarp();
      // ---------------- H' = W~' + F'^T G~'[4..15, :]  (B fragments: rows 4+4ki+t of G~': tile row, half as for P~'s columns)
      double gb[3][2];
#pragma unroll
      for (int ki = 0; ki < 3; ++ki)
#pragma unroll
        for (int nj = 0; nj < 2; ++nj) gb[ki][nj] = Gb[((ki >= 1 ? 1 : 0) * 2 + nj) * 64 + (ki == 1 ? mir0 : mir1)];
      double hh[3][2];   // tiles (0,0), (0,1), (1,1)
#pragma unroll
      for (int T = 0; T < 3; ++T) { const double2 wv = lds128(Wt + T * 64 + g * 8 + 2 * t); hh[T][0] = wv.x; hh[T][1] = wv.y; }
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) {
        rm_dmma(hh[0][0], hh[0][1], fb[ki][0], gb[ki][0]);
        rm_dmma(hh[1][0], hh[1][1], fb[ki][0], gb[ki][1]);
        rm_dmma(hh[2][0], hh[2][1], fb[ki][1], gb[ki][1]);
        hstep(6 + 2 * ki); hstep(7 + 2 * ki);      // l = 6 .. 11
      }
      // Huu[a][b] = H'(3-a, 3-b), b <= a (+ mu on the diagonal): rows / columns 0..3 of tile (0,0), i.e. lanes g < 4, t < 2
      if (g < 4) {
        if (g == 2 * t) hh[0][0] = hh[0][0] + mu;
        if (g == 2 * t + 1) hh[0][1] = hh[0][1] + mu;
      }
      sts128(Ht + cs_off, hh[0][0], hh[0][1]);
      sts128(Ht + 64 + cs_off, hh[1][0], hh[1][1]);
      if (g < 4 && t < 2) {
        Us[(3 - g) * 4 + (3 - 2 * t)] = hh[0][0];
        Us[(3 - g) * 4 + (2 - 2 * t)] = hh[0][1];
      }
      if (act && !xlane) base[W12_HV + (3 - p)] = hacc;
      double ldr[4], lu0 = hh[0][0], lu1 = hh[0][1];
      __syncwarp();   // (the previous step's readers are done with Ls: they passed the barriers above)
      failed |= m24_ldl_pivot<0, 4>(lu0, lu1, ldr, Ls, lane);
      failed |= m24_ldl_pivot<1, 4>(lu0, lu1, ldr, Ls, lane);
      failed |= m24_ldl_pivot<2, 4>(lu0, lu1, ldr, Ls, lane);
      failed |= m24_ldl_pivot<3, 4>(lu0, lu1, ldr, Ls, lane);
      __syncwarp();

      // ---------------- gains of the own state column (lanes 4..15) and the feed-forward (lanes 0..3)
      double hux[4], hur[4], kk[4], kf[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) hux[c] = Hcol[m24_el(3 - c, 0) ^ jc];
#pragma unroll
      for (int c = 0; c < 4; c += 2) { const double2 v = lds128(base + W12_HV + c); hur[c] = v.x; hur[c + 1] = v.y; }
      {
        double rhs[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rhs[c] = xlane ? hux[c] : hur[c];
        m24_solve_neg<4>(Ls, ldr, rhs, kk);
      }
      if (lane == 0) {
        sts128(base + W12_KF, kk[0], kk[1]);
        sts128(base + W12_KF + 2, kk[2], kk[3]);
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; c += 2) { const double2 v = lds128(base + W12_KF + c); kf[c] = v.x; kf[c + 1] = v.y; }
      // Vx' = h_x + Hux^T k for the own state column
      if (act && xlane) {
        double s = hacc;
#pragma unroll
        for (int c = 0; c < 4; ++c) s = fmad(hux[c], kf[c], s);
        base[W12_VS + j] = s;
      }
      // dV1 += k' h_u ; dV2 += k' Huu k / 2
      double t1 = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) t1 = fmad(kf[c], hur[c], t1);
      {
        const int r = lane & 3;
        double rr = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) rr = fmad(c <= r ? Us[r * 4 + c] : Us[c * 4 + r], kf[c], rr);
        if (lane < 4) base[W12_RR + r] = rr;
      }
      __syncwarp();
      double t2 = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) t2 = fmad(kf[r], base[W12_RR + r], t2);
      dV1 = dV1 + t1;
      dV2 = dV2 + 0.5 * t2;
      if (lane < 4) __stcs(a.kff + ((size_t)b * N + k) * 4 + lane, base[W12_KF + lane]);
      // K_k: to global memory and, by position, to shared memory for the P~' product
      double* Kb = Gb;
      if (act && xlane) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          Kb[c * 16 + p] = kk[c];
          __stcs(Kp + c * 12 + j, kk[c]);
        }
      }
      Kp -= 48;
      __syncwarp();

      // ---------------- P~' = H' + Hux^T K~ on the H' fragments (one k-step: the four inputs)
      {
        const double af0 = Ht[ar], af1 = Ht[64 + ar];
        const double bf0 = Kb[t * 16 + g], bf1 = Kb[t * 16 + 8 + g];
        rm_dmma(hh[0][0], hh[0][1], af0, bf0);
        rm_dmma(hh[1][0], hh[1][1], af0, bf1);
        rm_dmma(hh[2][0], hh[2][1], af1, bf1);
        if (more) {
          mbar_wait(bar, phase);
          phase ^= 1u;
          __syncwarp();
          scatter_prep(k - 1);
        }
#pragma unroll
        for (int T = 0; T < 3; ++T) sts128(Pt + T * 64 + cs_off, hh[T][0], hh[T][1]);
      }
      __syncwarp();
    }  // steps
    double dl = (a.reg_schedule == 1) ? a.delta[b] : 1.0;
    __syncwarp();
    if (failed && !reg_fail) {
      if (a.reg_schedule == 1) {
        reg_increase(mu, dl);
        if (lane == 0) a.delta[b] = dl;
      } else {
        mu = (mu * 10.0 > 1e-6) ? mu * 10.0 : 1e-6;
      }
      if (mu > 1e6) reg_fail = true;
    }
    if (!(failed && !reg_fail)) break;
  }
  if (lane == 0) {
    a.mu[b] = mu;
    a.dV[(size_t)b * 2] = dV1;
    a.dV[(size_t)b * 2 + 1] = dV2;
    if (a.status) a.status[b] = reg_fail ? 4 : 0;
  }
}

}  // namespace noc

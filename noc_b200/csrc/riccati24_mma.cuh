// riccati24_mma.cuh — the n=24, m=8 backward Riccati recursion (dual-UAV model, BASELINE.json configs[4]) with its three
// matrix products on the FP64 TENSOR CORES, bit-identical to k_ric24 and the oracle.
//
// Why bitwise: on B200 mma.sync.m8n8k4.f64 (SASS: DMMA) evaluates every output element as the sequential chain
// fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) (tools/ubench/dmma_order.cu, profiles/ubench_dmma_order_r02.json), so a
// k-loop of DMMAs that feeds the accumulator back IS the spec's "fma chain over l ascending".  n=24 and n+m=32 are
// multiples of the 8x8x4 tile: nothing is padded, and unlike k_ric24 (which prunes the model's structural zeros to halve
// its DFMA count and still spills) the dense tiles need few registers and almost no issue slots.
//
// Mapping: ONE instance per warp, lane = fragment coordinates (g = lane>>2, t = lane&3).  Per step:
//   G' = P F'            24 x 24 x 32: 72 DMMAs.  F' = F = [A | B] with its columns in the order [u7 .. u0 | x0 .. x23]: in
//                        that order every entry the spec needs — Hxx[i][j] for i <= j, Hux[a][j], Huu[a][b] for b <= a —
//                        lies in the UPPER triangle of H'.
//   H' = W' + F'^T G'    the ten upper 8x8 tiles of 32 x 32: 60 DMMAs.  The A fragments of F'^T are the B fragments of F'
//                        (same registers); G' goes through a shared-memory buffer, column tile by column tile, because a
//                        DMMA's C layout is not its B layout.
//   L D L^T of Huu (8x8) distributed over the warp (m24_ldl_pivot: the lanes hold Huu as the C fragments of H' tile (0,0)),
//                        the 24 gain columns (lanes 8..31) and the feed-forward (lanes 0..7) back-substituted in ONE
//                        pass of the same code,
//   P' = Hxx + Hux^T K   the six upper tiles of 24 x 24, k = 8: 12 DMMAs (chain over the input index from Hxx: §2.3 step 5).
//   RIC_AFFINE (the SLQ backward pass): everything of a step that does not depend on the recursion — the Jacobian entries,
//   h0 = W e, ubar — arrives as a 912-byte step record from k_prep24 (PREP); h = h0 + F^T Vx, Vx', dV1 / dV2 are per-lane
//   DFMA chains over the lane's own column, issued between the DMMAs; the input box is a warp-cooperative QP (m24_box_qp).
//   RIC_DARE: the fixed-point iteration on one LTI record whose fragments stay in registers.
// Tile store: every 8x8 tile in shared memory keeps element (r, c) at r*8 + (c ^ 4*((r>>1)&1)) — C-fragment stores
// (16-byte pairs), A-fragment loads of either half (rows g, columns 4h+t) and transposed loads (rows 4h+t, column g) are
// all bank-conflict free.  P is kept as its six upper tiles; an A fragment below the diagonal reads its mirror image.
#pragma once
#include "common.cuh"
#include "model.cuh"
#include "riccati12_mma.cuh"   // rm_dmma, rm_elect_one
#include "riccati24.cuh"       // Ric24Args, LdlReg, box_qp_gen

namespace noc {

// per-warp (= per-instance) shared memory, in doubles
constexpr int M24_FS = 0;        // record A[24][24] | B[24][8]
constexpr int M24_PT = 768;      // P: upper tiles (0,0) (0,1) (0,2) (1,1) (1,2) (2,2); inside a step the Hxx tiles of H'
constexpr int M24_HU = 1152;     // u-row tiles of H': (0,1) (0,2) (0,3) at +64, +128, +192 (slot 0 unused)
constexpr int M24_GB = 1408;     // G' column tiles, two buffers of three tiles; scalar phase: K[8][24]; epilogue: row sums
constexpr int M24_US = 1792;     // Huu[8][8] row-major, lower triangle valid (LdlReg / box_qp_gen read it)
constexpr int M24_XS = 1856;     // RIC_LQR + FUSED: two buffers of xbar_k[24] | ubar_k[8]   (RIC_AFFINE: the M24P_* slots below)
constexpr int M24_XS_END = 1920;
constexpr int M24_WT_DOUBLES = 640;   // per CTA: the ten upper tiles of W' in C-fragment order
constexpr int M24_BAR_DOUBLES = 8;

// PREP (the SLQ backward pass): everything of a step that does not depend on the recursion is computed beforehand by
// k_prep24, time-parallel and at full lane efficiency, into one 912-byte record per (work-list slot, k):
//   [0, 74)    the state-dependent Jacobian entries of the two vehicles, in quad_jac_discrete's emission order
//   [74, 106)  h0[p] = (W e)[column at position p], e = [xbar_k - xgoal ; ubar_k - u_h]
//   [106, 114) ubar_k (the input box needs it)
// The sweep pulls the record of step k-1 in by one bulk copy during step k and scatters the 74 entries into the staged
// A|B record; evaluated inside the sweep (32 lanes, two distinct vehicles) the Jacobian was 24 % of its instructions.
constexpr int M24_JE = 37;                       // state-dependent entries per vehicle: 25 of A, 12 of B
constexpr int M24_PREP_H0 = 2 * M24_JE, M24_PREP_U = M24_PREP_H0 + 32, M24_PREP_DOUBLES = M24_PREP_U + 8;   // 114
constexpr int M24P_PR = 1856;                    // PREP: two buffers of M24_PREP_DOUBLES
constexpr int M24P_VS = M24P_PR + 2 * M24_PREP_DOUBLES,   // Vx[24]
              M24P_HV = M24P_VS + 24,                     // h_u[8]
              M24P_KF = M24P_HV + 8,                      // feed-forward k[8]
              M24P_RR = M24P_KF + 8,                      // Huu k [8]
              M24P_QP = M24P_RR + 8;                      // BOX: workspace of the 8x8 box QP (32 doubles used)

template <int MODE, bool FUSED, bool BOX, bool PREP = false>
__host__ __device__ constexpr int m24_warp_doubles() {
  return PREP ? (BOX ? M24P_QP + 32 : M24P_QP) : (FUSED ? M24_XS_END : M24_XS);
}
// PERW (per-instance weights, RIC_LQR on A_THEN_B records): every warp keeps the seven W' tiles its H' products start from
// — (0,0) and the six upper tiles of the x-x block — for its own instance; twelve warps still share an SM (221 KB)
constexpr int M24_WT_PERW = 448;
template <int MODE, bool FUSED, bool BOX, bool PREP = false, bool PERW = false>
inline size_t ric24_mma_smem_bytes(int warps) {
  return sizeof(double) * (size_t)(M24_BAR_DOUBLES + (PERW ? 0 : M24_WT_DOUBLES) +
                                   warps * (m24_warp_doubles<MODE, FUSED, BOX, PREP>() + (PERW ? M24_WT_PERW : 0)));
}

// Two threads per (work-list slot w, step k, vehicle v) — one evaluates the Jacobian, the other h0 and ubar (different warps:
// no divergence) — and a CTA's 64 records are assembled in shared memory and written out as one contiguous 58 KB block
// (every thread storing its own 456 bytes straight to global memory ran at 0.9 TB/s).
constexpr int PREP24_PAIRS = 32, PREP24_THREADS = 4 * PREP24_PAIRS;   // 37 KB per CTA: six CTAs of four warps per SM
constexpr size_t PREP24_SMEM = sizeof(double) * (1024 + PREP24_PAIRS * M24_PREP_DOUBLES);
__global__ void __launch_bounds__(PREP24_THREADS) k_prep24(const Ric24Args a, double* __restrict__ prep) {
  extern __shared__ __align__(16) double psm[];
  double* Ws = psm;
  double* so = psm + 1024;
  for (int e = threadIdx.x; e < 1024; e += PREP24_THREADS) Ws[e] = a.W[e];
  __syncthreads();
  const int N = a.N;
  const long long total = a.batch * N;
  const long long wk0 = (long long)blockIdx.x * PREP24_PAIRS;
  const int role = threadIdx.x / (2 * PREP24_PAIRS), pr = (threadIdx.x >> 1) % PREP24_PAIRS, v = threadIdx.x & 1;
  const long long wk = wk0 + pr;
  if (wk < total) {
    const long long w = wk / N;
    const int k = (int)(wk - w * N);
    const long long b = a.idx ? (long long)a.idx[w] : w;
    const size_t xsb = a.xbar_sb ? (size_t)a.xbar_sb : (size_t)(N + 1) * 24, xsk = a.xbar_sk ? (size_t)a.xbar_sk : 24;
    const double* xs = a.xbar + (size_t)b * xsb + (size_t)k * xsk;
    const double* us = a.ubar + ((size_t)b * N + k) * 8;
    double* out = so + pr * M24_PREP_DOUBLES;
    if (role == 0) {
      double xk[12], uk[4];   // (16-byte loads: a warp's 8-byte loads of 32 different records kept the LSU busy, not the FP64 pipe)
#pragma unroll
      for (int i = 0; i < 12; i += 2) { const double2 q = __ldg(reinterpret_cast<const double2*>(xs + 12 * v + i)); xk[i] = q.x; xk[i + 1] = q.y; }
#pragma unroll
      for (int c = 0; c < 4; c += 2) { const double2 q = __ldg(reinterpret_cast<const double2*>(us + 4 * v + c)); uk[c] = q.x; uk[c + 1] = q.y; }
      double* o = out + M24_JE * v;
      int e = 0;
      model::quad_jac_discrete(
          xk, uk, a.dt,
          [&](int i, int jj, double val) { if (!model::quad_jac_is_const_A(i, jj) && !model::dual_coupled_diag(i, jj)) o[e++] = val; },
          [&](int i, int jj, double val) { if (!model::quad_jac_is_const_B(i, jj)) o[e++] = val; });
    } else {
      // h0 for the vehicle's own columns (W block diagonal: the chain runs over the own block, as in k_ric24)
      const double* xg = a.xgoal + (size_t)b * 24;
      double ex[24];
#pragma unroll
      for (int i = 0; i < 24; i += 2) {
        const double2 q = __ldg(reinterpret_cast<const double2*>(xs + i)), r = __ldg(reinterpret_cast<const double2*>(xg + i));
        ex[i] = q.x - r.x; ex[i + 1] = q.y - r.y;
      }
#pragma unroll 2
      for (int c = 12 * v; c < 12 * v + 12; ++c) {
        double h = 0.0;
#pragma unroll
        for (int i = 0; i < 24; ++i) h = fmad(Ws[i * 32 + c], ex[i], h);
        out[M24_PREP_H0 + 8 + c] = h;
      }
      double eu[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 q = __ldg(reinterpret_cast<const double2*>(us + i));
        eu[i] = q.x - model::kHover; eu[i + 1] = q.y - model::kHover;
        if ((i >> 2) == v) { out[M24_PREP_U + i] = q.x; out[M24_PREP_U + i + 1] = q.y; }
      }
#pragma unroll 2
      for (int c = 4 * v; c < 4 * v + 4; ++c) {
        double h = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) h = fmad(Ws[(24 + i) * 32 + 24 + c], eu[i], h);
        out[M24_PREP_H0 + 7 - c] = h;
      }
    }
  }
  __syncthreads();
  const long long left = total - wk0;
  const int cnt = (int)(left < PREP24_PAIRS ? left : PREP24_PAIRS) * M24_PREP_DOUBLES;
  double* dst = prep + (size_t)wk0 * M24_PREP_DOUBLES;
  for (int e = 2 * threadIdx.x; e < cnt; e += 2 * PREP24_THREADS) {
    const double2 val = lds128(so + e);
    __stcg(reinterpret_cast<double2*>(dst + e), val);
  }
}

// position p of the permuted order [u7 .. u0 | x0 .. x23] -> index in the W = blockdiag(Q, R) order [x0 .. x23 | u0 .. u7]
__host__ __device__ constexpr int m24_perm(int p) { return p < 8 ? 31 - p : p - 8; }
// upper tile (mi, nj), mi <= nj, of a 3 x 3 / 4 x 4 tile grid -> running index
__host__ __device__ constexpr int m24_t3(int mi, int nj) { return mi * (7 - mi) / 2 + (nj - mi); }
__host__ __device__ constexpr int m24_t4(int mi, int nj) { return mi * (9 - mi) / 2 + (nj - mi); }
// element (r, c) of a stored tile
__host__ __device__ constexpr int m24_el(int r, int c) { return r * 8 + (c ^ (4 * ((r >> 1) & 1))); }

// Record built in the kernel (FUSED): the columns of row l are stored XOR 4*((l>>1)&1), so that the B-fragment loads
// (rows 4ki+t, eight consecutive columns) of a half-warp fall into four different bank groups; with the plain row
// strides of 24 (A) and 8 (B) doubles the rows t and t+2 collided (2 wavefronts per load instead of 1).  Records that
// arrive from HBM by bulk copy keep the caller's layout.
__host__ __device__ constexpr int m24_aoff(int l, int c, bool swz) { return l * 24 + (swz ? (c ^ (4 * ((l >> 1) & 1))) : c); }
__host__ __device__ constexpr int m24_boff(int l, int u, bool swz) { return 576 + l * 8 + (swz ? (u ^ (4 * ((l >> 1) & 1))) : u); }

// Built-in dual model (FUSED): is the 4 x 8 tile of F' = [B | A] (rows 4ki .., positions 8ni ..) structurally non-zero?
// Four of the 24 tiles are not (vehicle 1's attitude rows in vehicle 0's columns and the like): a DMMA with an all-zero
// operand tile leaves its accumulator as it is (fma(p, 0, acc) == acc for finite p — the identity k_ric24's entry-wise
// pruning rests on), so 12 of the 72 DMMAs of G' and 7 of the 60 of H' are not issued.
__host__ __device__ constexpr bool m24_tile_nz(int ki, int ni) {
  for (int l = 4 * ki; l < 4 * ki + 4; ++l)
    for (int pp = 8 * ni; pp < 8 * ni + 8; ++pp)
      if (pp < 8 ? model::dual_jac_nz_B(l, 7 - pp) : model::dual_jac_nz_A(l, pp - 8)) return true;
  return false;
}

// L D L^T of Huu (8 x 8) DISTRIBUTED over the warp, right-looking, one pivot per call.  After the DMMAs of H' tile (0,0)
// lane (g, t) holds Huu[a][b0] and Huu[a][b1], a = 7 - g, b0 = 7 - 2t, b1 = 6 - 2t (the C fragment in the permuted order);
// entries with b > a are dead slots.  Pivot l: D[l] and then column l of L are broadcast by shuffles and every lane applies
// update l to its own two entries, U[a][b] = fma(-L[a][l], L[b][l] D[l], U[a][b]) for b > l — the spec's left-looking chains
// term by term and in the same ascending order (§2.3 step 3), so the result is bit-identical to LdlReg<8>::factor, at 12
// FP64 instructions per pivot instead of 14-30 (every lane used to run the whole factorisation redundantly: 22 % of the
// sweep's instructions).  Column l of L goes to shared memory (Ls[i*8+l]) for the gain solves; r[] = 1/D stays in registers.
template <int L, int M = 8>
__device__ __forceinline__ bool m24_ldl_pivot(double& u0, double& u1, double (&r)[M], double* Ls, int lane) {
  // general M (k_ric12_wpi: M = 4): a = M-1-g, b0 = M-1-2t, b1 = M-2-2t; lanes outside the M x M block hold dead values
  constexpr int tl = (M - 1 - L) >> 1;                 // the lanes t == tl hold column L ...
  constexpr bool slot1 = ((M - 1 - L) & 1) == 1;       // ... in u1 or in u0
  const int g = lane >> 2, t = lane & 3, a = M - 1 - g;
  const double mine = slot1 ? u1 : u0;
  const double d = __shfl_sync(0xffffffffu, mine, (M - 1 - L) * 4 + tl);
  const bool bad = !(d > R12_PIVOT_MIN) || !(d < R12_PIVOT_MAX);
  r[L] = rcp_rn_normal(d);
  if (L < M - 1) {
    const double lv = mine * r[L];             // L[a][L] on the lanes t == tl
    if (t == tl && a > L) Ls[a * M + L] = lv;
    const double la = __shfl_sync(0xffffffffu, lv, (lane & ~3) | tl);
    const double lb0 = __shfl_sync(0xffffffffu, lv, (2 * t) * 4 + tl);        // row b0 = M-1-2t lives on g' = 2t
    const double lb1 = __shfl_sync(0xffffffffu, lv, (2 * t + 1) * 4 + tl);    // row b1 = M-2-2t on g' = 2t + 1
    const double n0 = fmad(-la, lb0 * d, u0), n1 = fmad(-la, lb1 * d, u1);
    if (M - 1 - 2 * t > L) u0 = n0;
    if (M - 2 - 2 * t > L) u1 = n1;
  }
  return bad;
}
// the spec's forward / backward substitution (LdlReg<8>::solve_neg) with L read from shared memory (broadcast loads)
template <int M = 8>
__device__ __forceinline__ void m24_solve_neg(const double* Ls, const double (&r)[M], const double (&rhs)[M], double (&out)[M]) {
  double Lr[M][M];
#pragma unroll
  for (int i = 1; i < M; ++i)
#pragma unroll
    for (int l = 0; l < i; ++l) Lr[i][l] = Ls[i * M + l];
  double y[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double sacc = rhs[i];
#pragma unroll
    for (int l = 0; l < i; ++l) sacc = fmad(-Lr[i][l], y[l], sacc);
    y[i] = sacc;
  }
  double z[M];   // negated at the end like the oracle (a chain on -z_i flips the sign of exact zeros)
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    double sacc = y[i] * r[i];
#pragma unroll
    for (int l = i + 1; l < M; ++l) sacc = fmad(-Lr[l][i], z[l], sacc);
    z[i] = sacc;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) out[i] = neg_bits(z[i]);
}

// ---- input box, warp-cooperative restatement of box_qp_gen<8> for one instance per warp: the same
// operations on the same values in the same order (bit-identical iterates, clamped sets and exits), but the two parts
// that every lane used to run redundantly are spread over the lanes — the matrix-vector product of the objective (row a on
// lane a) and the masked factorisation (m24_ldl_pivot) — and nothing lives in local memory.
// U: Quu row-major, lower triangle valid (shared); ws: 32 doubles x | r | xc | rc; Ls: column store of the masked factor.
__device__ __forceinline__ double m24_qp_eval(const double* U, const double (&g)[8], const double* x, double* r, int lane) {
  const int a = lane & 7;
  double sacc = 0.0;
#pragma unroll
  for (int b = 0; b < 8; ++b) sacc = fmad(b <= a ? U[a * 8 + b] : U[b * 8 + a], x[b], sacc);
  __syncwarp();   // (every lane is done with the previous contents of r)
  r[a] = sacc;
  __syncwarp();
  double f = 0.0;
#pragma unroll
  for (int b = 0; b < 8; ++b) f = fmad(x[b], fmad(0.5, r[b], g[b]), f);
  return f;
}
__device__ __forceinline__ bool m24_ldl_masked(const double* U, unsigned clmask, double (&r)[8], double* Ls, int lane) {
  const int g = lane >> 2, t = lane & 3, a = 7 - g, b0 = 7 - 2 * t, b1 = 6 - 2 * t;
  auto um = [&](int aa, int bb) {
    const bool c = (((clmask >> aa) | (clmask >> bb)) & 1u) != 0;
    return c ? (aa == bb ? 1.0 : 0.0) : U[aa * 8 + bb];
  };
  double u0 = b0 <= a ? um(a, b0) : 0.0, u1 = b1 <= a ? um(a, b1) : 0.0;
  __syncwarp();   // (the previous factor's readers are done with Ls)
  bool bad = false;
  bad |= m24_ldl_pivot<0>(u0, u1, r, Ls, lane); bad |= m24_ldl_pivot<1>(u0, u1, r, Ls, lane);
  bad |= m24_ldl_pivot<2>(u0, u1, r, Ls, lane); bad |= m24_ldl_pivot<3>(u0, u1, r, Ls, lane);
  bad |= m24_ldl_pivot<4>(u0, u1, r, Ls, lane); bad |= m24_ldl_pivot<5>(u0, u1, r, Ls, lane);
  bad |= m24_ldl_pivot<6>(u0, u1, r, Ls, lane); bad |= m24_ldl_pivot<7>(u0, u1, r, Ls, lane);
  __syncwarp();
  return bad;
}
// qr / fmask: 1/D of the last masked factorisation the QP computed (its L is in Ls) and the clamped set it belongs to
// (~0u: none, or it failed) — when the QP stops on that very set, which is the rule once the active set has settled, the
// caller's gain solves reuse it instead of factoring again.
__device__ __forceinline__ void m24_box_qp(const double* U, const double (&g)[8], const double (&lo)[8], const double (&hi)[8],
                                           double (&xio)[8], unsigned& clmask, double* ws, double* Ls, int lane,
                                           double (&qr)[8], unsigned& fmask) {
  fmask = ~0u;
  double* x = ws;
  double* r = ws + 8;
  double* xc = ws + 16;
  double* rc = ws + 24;
  double gmax = 0.0;
  __syncwarp();
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    if (lane == 0) x[a] = r12_clamp(xio[a], lo[a], hi[a]);
    if (fabs(g[a]) > gmax) gmax = fabs(g[a]);
  }
  __syncwarp();
  double f = m24_qp_eval(U, g, x, r, lane);
  auto finish = [&]() {
#pragma unroll
    for (int a = 0; a < 8; ++a) xio[a] = x[a];
  };
  auto clamped_set = [&](double (&grad)[8], int& nfree, double& gn) {
    unsigned m = 0;
    nfree = 0; gn = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double ga = r[a] + g[a], xa = x[a];
      grad[a] = ga;
      const bool c = (xa <= lo[a] && ga > 0.0) || (xa >= hi[a] && ga < 0.0);
      if (c) m |= 1u << a;
      else { ++nfree; if (fabs(ga) > gn) gn = fabs(ga); }
    }
    return m;
  };
  for (int it = 0; it < R12_QP_MAXIT; ++it) {
    double grad[8], gn;
    int nfree;
    clmask = clamped_set(grad, nfree, gn);
    if (nfree == 0 || gn < R12_QP_GTOL * (1.0 + gmax)) { finish(); return; }
    if (fmask != clmask) {   // (an iteration on the clamped set of the previous one keeps its factorisation)
      if (m24_ldl_masked(U, clmask, qr, Ls, lane)) { fmask = ~0u; finish(); return; }
      fmask = clmask;
    }
    double rhs[8], dx[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) rhs[a] = ((clmask >> a) & 1u) ? 0.0 : grad[a];
    m24_solve_neg(Ls, qr, rhs, dx);
    double sdotg = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) sdotg = fmad(grad[a], dx[a], sdotg);
    if (!(sdotg < 0.0)) { finish(); return; }
    double step = 1.0;
    bool ok = false;
    for (int ls = 0; ls < R12_QP_MAXLS; ++ls) {
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) xc[a] = r12_clamp(fmad(step, dx[a], x[a]), lo[a], hi[a]);
      }
      __syncwarp();
      const double fc = m24_qp_eval(U, g, xc, rc, lane);
      if (fc - f <= R12_QP_ARMIJO * (step * sdotg)) {
        __syncwarp();
        if (lane < 8) { x[lane] = xc[lane]; r[lane] = rc[lane]; }
        __syncwarp();
        f = fc;
        ok = true;
        break;
      }
      step = step * 0.5;
    }
    if (!ok) { finish(); return; }
  }
  {   // iteration cap: the clamped set of the last accepted point
    double grad[8], gn;
    int nfree;
    clmask = clamped_set(grad, nfree, gn);
  }
  finish();
}

template <int MODE, bool FUSED, int WARPS, int MIN_CTAS, bool BOX = false, bool PREP = false, bool PERW = false>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) k_ric24_mma(const Ric24Args a, const double* __restrict__ prep) {
  static_assert(!PERW || (MODE == RIC_LQR && !FUSED), "per-instance weights: the LTV sweep on records from HBM");
  static_assert(MODE == RIC_LQR || MODE == RIC_AFFINE || MODE == RIC_DARE, "LQR sweep, SLQ backward pass or DARE iteration");
  static_assert(MODE != RIC_DARE || !FUSED, "RIC_DARE iterates on one LTI record from HBM");
  constexpr bool DARE = MODE == RIC_DARE;
  static_assert(MODE != RIC_AFFINE || FUSED, "the SLQ backward pass linearises in the kernel");
  static_assert(!BOX || MODE == RIC_AFFINE, "input box: SLQ only");
  static_assert(PREP == (MODE == RIC_AFFINE), "the SLQ backward pass, and only it, takes its step records from k_prep24");
  extern __shared__ __align__(128) double smem[];
  constexpr int WD = m24_warp_doubles<MODE, FUSED, BOX, PREP>() + (PERW ? M24_WT_PERW : 0);
  constexpr bool AFF = MODE == RIC_AFFINE;
  constexpr bool BUILD = FUSED && !PREP;   // the record is evaluated in the sweep
  constexpr bool TMA = !FUSED || PREP;     // something arrives by bulk copy on the warp's mbarrier
  constexpr int O_VS = M24P_VS, O_HV = M24P_HV, O_KF = M24P_KF, O_RR = M24P_RR, O_QP = M24P_QP;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // shared weights: ten W' tiles per CTA; PERW: the warp's own seven tiles behind its slice (a.W = the caller's
  // [batch][Q 576 | R 64 | Qf 576] array, a.w_stride doubles apart)
  double* base = smem + M24_BAR_DOUBLES + (PERW ? 0 : M24_WT_DOUBLES) + warp * WD;
  double* Wt = PERW ? base + (WD - M24_WT_PERW) : smem + M24_BAR_DOUBLES;
  double* Fs = base + M24_FS;
  double* Pt = base + M24_PT;
  double* Hu = base + M24_HU;
  double* Gb = base + M24_GB;
  double* Us = base + M24_US;
  const unsigned bar = smem_u32(smem) + 8u * warp;

  // W' = W in the permuted order, upper tiles, C-fragment order (plain row-major tiles: each lane reads its own 16 bytes)
  for (int e = threadIdx.x; e < (PERW ? 0 : M24_WT_DOUBLES); e += WARPS * 32) {
    const int T = e >> 6, r = (e >> 3) & 7, c = e & 7;
    const int mi = T < 4 ? 0 : (T < 7 ? 1 : (T < 9 ? 2 : 3));
    const int nj = T - m24_t4(mi, mi) + mi;
    Wt[e] = a.W[m24_perm(8 * mi + r) * 32 + m24_perm(8 * nj + c)];
  }
  if (TMA) {
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const long long w = (long long)blockIdx.x * WARPS + warp;
  if (w >= a.batch) return;                  // a whole warp: nothing below synchronises across warps
  const long long b = (AFF && a.idx) ? (long long)a.idx[w] : w;
  const int N = a.N;
  const double* Wown = PERW ? a.W + (size_t)b * (size_t)a.w_stride : nullptr;
  if (PERW) {   // the spec reads Q[i][j] for i <= j and R[a][b] for b <= a (as k_build_w does for shared weights)
    for (int e = lane; e < M24_WT_PERW; e += 32) {
      const int T = e >> 6, r = (e >> 3) & 7, c = e & 7;
      const int mi = T == 0 ? 0 : (T < 4 ? 1 : (T < 6 ? 2 : 3));
      const int nj = T == 0 ? 0 : T + 3 - m24_t4(mi, mi) + mi;
      const int pr = m24_perm(8 * mi + r), pc = m24_perm(8 * nj + c);   // W order: [x0 .. x23 | u0 .. u7]
      const int lo = pr < pc ? pr : pc, hi = pr < pc ? pc : pr;
      double v = 0.0;
      if (hi < 24) v = Wown[lo * 24 + hi];
      else if (lo >= 24) v = Wown[576 + (hi - 24) * 8 + (lo - 24)];
      Wt[e] = v;
    }
    __syncwarp();
  }

  // ---- lane constants
  const int s_g = (g >> 1) & 1, s_t = (t >> 1) & 1;
  const int cs_off = g * 8 + ((2 * t) ^ (4 * s_g));                                   // C fragment {D[g][2t], D[g][2t+1]}
  const int dir0 = g * 8 + 4 * (0 ^ s_g) + t, dir1 = g * 8 + 4 * (1 ^ s_g) + t;          // element (g, 4h+t)
  const int mir0 = t * 8 + (g ^ (4 * s_t)), mir1 = (4 + t) * 8 + (g ^ (4 * s_t));        // element (4h+t, g)
  const int dia0 = (g <= t) ? dir0 : mir0, dia1 = (g <= 4 + t) ? dir1 : mir1;            // diagonal tile: upper part only
  constexpr bool SWZ = FUSED;               // m24_aoff / m24_boff: rows 4ki+t have (l>>1)&1 == s_t
  const int fo_u = 576 + t * 8 + (SWZ ? ((7 - g) ^ (4 * s_t)) : (7 - g));   // F'[4ki+t][g]      = B[4ki+t][7-g]:        + 32 ki
  const int fo_x = t * 24 + (SWZ ? (g ^ (4 * s_t)) : g);                    // F'[4ki+t][8ni+g]  = A[4ki+t][8(ni-1)+g]:  + 96 ki + 8 (ni-1)
  // Hux^T A fragments of the P' product: Hux[4ki+t][8mi+g] = H'(7-4ki-t, 8+8mi+g)
  const int ar0 = m24_el(0, 0) + (7 - t) * 8 + (g ^ (4 * (((7 - t) >> 1) & 1)));
  const int ar1 = (3 - t) * 8 + (g ^ (4 * (((3 - t) >> 1) & 1)));
  // scalar phase: lane p owns position p of the permuted order — lanes 8..31 the state column j = p - 8, lanes 0..7 the
  // input column 7 - p
  const int p = lane;
  const bool xlane = p >= 8;
  const int j = xlane ? p - 8 : 0;
  const double* Hcol = Hu + (1 + (j >> 3)) * 64;    // Hux[a][j] = Hcol[m24_el(7 - a, j & 7)]
  const int jc = j & 7;
  const int fcol = xlane ? (p - 8) : 576 + (7 - p);  // F[l][own column] = Fs[fcol + l * fcs] (rows with (l>>1)&1: fcol ^ 4)
  const int fcolx = SWZ ? (fcol ^ 4) : fcol;
  const int fcs = xlane ? 24 : 8;

  const double* rec = FUSED ? nullptr : a.AB + (size_t)b * (DARE ? 1 : N) * R24_REC;
  const unsigned fs_u32 = smem_u32(Fs);
  unsigned phase = 0;
  const size_t xsb = a.xbar_sb ? (size_t)a.xbar_sb : (size_t)(N + 1) * 24, xsk = a.xbar_sk ? (size_t)a.xbar_sk : 24;

  // xbar_k | ubar_k -> XS[k & 1] (sixteen 16-byte chunks)
  auto fetch_x = [&](int k) {
    double* dst = base + M24_XS + 32 * (k & 1);
    const double* xs = a.xbar + (size_t)b * xsb + (size_t)k * xsk;
    if (lane < 12) cp_async16(smem_u32(dst + 2 * lane), xs + 2 * lane);
    else if (lane < 16 && a.ubar) cp_async16(smem_u32(dst + 24 + 2 * (lane - 12)), a.ubar + ((size_t)b * N + k) * 8 + 2 * (lane - 12));
    cp_async_commit();
  };
  auto fetch_rec = [&](int k) {
    if (rm_elect_one()) {
      mbar_expect_tx(bar, 6144u);
      bulk_g2s(fs_u32, rec + (size_t)k * R24_REC, 6144u, bar);
    }
  };
  // PREP: the record of (w, k) -> PR[k & 1]; scatter of its Jacobian entries into the staged A|B record.  The three
  // destinations of a lane (entries lane, lane+32, lane+64 of the 74) are found once, by replaying the emission order.
  const double* prep_w = PREP ? prep + (size_t)w * N * M24_PREP_DOUBLES : nullptr;
  auto fetch_prep = [&](int k) {
    if (rm_elect_one()) {
      mbar_expect_tx(bar, 8u * M24_PREP_DOUBLES);
      bulk_g2s(smem_u32(base + M24P_PR + M24_PREP_DOUBLES * (k & 1)), prep_w + (size_t)k * M24_PREP_DOUBLES, 8u * M24_PREP_DOUBLES, bar);
    }
  };
  int so0 = 0, so1 = 0, so2 = -1;
  if (PREP) {
    const double zx[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, zu[4] = {0, 0, 0, 0};
    int e = 0;
    auto note = [&](int i, int jj, bool isB) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int id = M24_JE * v + e, o = isB ? m24_boff(12 * v + i, 4 * v + jj, SWZ) : m24_aoff(12 * v + i, 12 * v + jj, SWZ);
        if (id == lane) so0 = o;
        if (id == lane + 32) so1 = o;
        if (id == lane + 64) so2 = o;
      }
      ++e;
    };
    model::quad_jac_discrete(
        zx, zu, a.dt,
        [&](int i, int jj, double) { if (!model::quad_jac_is_const_A(i, jj) && !model::dual_coupled_diag(i, jj)) note(i, jj, false); },
        [&](int i, int jj, double) { if (!model::quad_jac_is_const_B(i, jj)) note(i, jj, true); });
  }
  auto scatter_prep = [&](int k) {
    const double* pr = base + M24P_PR + M24_PREP_DOUBLES * (k & 1);
    Fs[so0] = pr[lane];
    Fs[so1] = pr[lane + 32];
    if (so2 >= 0) Fs[so2] = pr[lane + 64];
  };
  // a1 in the sweep (as k_ric24<FUSED>): lanes 0..15 evaluate vehicle 0's Jacobian, lanes 16..31 vehicle 1's; each stores
  // every sixteenth state-dependent entry.  Zeros and the state-independent entries are written once per sweep.
  auto build_rec = [&](int k) {
    const int v = lane >> 4, t16 = lane & 15;
    const double* Xs = base + M24_XS + 32 * (k & 1);
    double xk[12], uk[4];
#pragma unroll
    for (int i = 0; i < 12; ++i) xk[i] = Xs[12 * v + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) uk[c] = a.ubar ? Xs[24 + 4 * v + c] : model::kHover;
    int e = 0;
    model::quad_jac_discrete(
        xk, uk, a.dt,
        [&](int i, int jj, double val) {
          if (!model::quad_jac_is_const_A(i, jj) && !model::dual_coupled_diag(i, jj) && ((e++) & 15) == t16) Fs[m24_aoff(12 * v + i, 12 * v + jj, SWZ)] = val;
        },
        [&](int i, int jj, double val) {
          if (!model::quad_jac_is_const_B(i, jj) && ((e++) & 15) == t16) Fs[m24_boff(12 * v + i, 4 * v + jj, SWZ)] = val;
        });
  };

  double mu = 0.0;
  if (AFF) mu = a.mu[b];
  bool failed = false, reg_fail = false;
  double dV1 = 0.0, dV2 = 0.0;
  int kfail = -1;
  int dare_iters = 0;
  bool converged = false;
  double kk_last[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

  for (;;) {   // (re)start of a sweep (RIC_AFFINE: again with a larger mu after a failed factorisation)
    failed = false;
    dV1 = 0.0; dV2 = 0.0;
    kfail = -1;
    double* Kp = a.K ? a.K + ((size_t)b * N + (N - 1)) * 192 : nullptr;
    // P <- Qf, as upper tiles
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int nj = mi; nj < 3; ++nj) {
        if (PERW) {   // Qf[i][j] for i <= j only
          const int i = 8 * mi + g, j0 = 8 * nj + 2 * t, j1 = j0 + 1;
          const double* qf = Wown + 640;
          sts128(Pt + m24_t3(mi, nj) * 64 + cs_off, i <= j0 ? qf[i * 24 + j0] : qf[j0 * 24 + i], i <= j1 ? qf[i * 24 + j1] : qf[j1 * 24 + i]);
        } else {
          const double* q = a.Qf + (8 * mi + g) * 24 + 8 * nj + 2 * t;
          sts128(Pt + m24_t3(mi, nj) * 64 + cs_off, q[0], q[1]);
        }
      }
    if (FUSED) {
      if (PREP) fetch_prep(N - 1);
      else fetch_x(N - 1);
      for (int e = lane; e < R24_REC; e += 32) Fs[e] = 0.0;
      __syncwarp();
      if (lane == 0) {   // state-independent entries: per-vehicle constants first, then the coupling (it overwrites (3+i,3+i))
        for (int v = 0; v < 2; ++v)
          model::quad_jac_constants(a.dt, [&](int i, int jj, double val) { Fs[m24_aoff(12 * v + i, 12 * v + jj, SWZ)] = val; },
                                    [&](int i, int jj, double val) { Fs[m24_boff(12 * v + i, 4 * v + jj, SWZ)] = val; });
        model::dual_coupling_jac(a.dt, [&](int i, int jj, double val) { Fs[m24_aoff(i, jj, SWZ)] = val; });
      }
    } else {
      fetch_rec(DARE ? 0 : N - 1);
    }
    if (AFF) {
      // xgoal -> shared; Vx <- Qf (xbar_N - xgoal), one row per lane (row chain, ascending)
      const double* xN = a.xbar + (size_t)b * xsb + (size_t)N * xsk;
      const double* xg = a.xgoal + (size_t)b * 24;
      if (lane < 24) {
        double s = 0.0;
        for (int c = 0; c < 24; ++c) s = fmad(a.Qf[lane * 24 + c], xN[c] - xg[c], s);
        base[O_VS + lane] = s;
      }
    }
    if (BUILD) {
      cp_async_wait_all();
      __syncwarp();
      build_rec(N - 1);
    }
    if (PREP) {
      __syncwarp();   // the zeros and constants are in place
      mbar_wait(bar, phase);
      phase ^= 1u;
      scatter_prep(N - 1);
    }
    __syncwarp();

    // RIC_DARE: the fixed-point iteration P <- step(P) on the one record, whose fragments stay in registers; stops when
    // max|P' - P| < tol max|P'| (oracle: noc_oracle_dare; same bits, hence the same iteration count)
    const int steps = DARE ? a.max_iter : N;
    double fb[6][4];
    for (int it = 0; it < steps; ++it) {
      const int k = DARE ? 0 : N - 1 - it;
      const bool more = !DARE && it + 1 < N;
      if (!FUSED && (!DARE || it == 0)) {
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
      __syncwarp();

      // ---------------- fragments: all of F' (24) and P (18) into registers
      double pa[3][6];
      if (!DARE || it == 0) {
#pragma unroll
        for (int ki = 0; ki < 6; ++ki) {
          fb[ki][0] = Fs[fo_u + 32 * ki];
#pragma unroll
          for (int ni = 1; ni < 4; ++ni) fb[ki][ni] = Fs[fo_x + 96 * ki + 8 * (ni - 1)];
        }
      }
#pragma unroll
      for (int mi = 0; mi < 3; ++mi)
#pragma unroll
        for (int ki = 0; ki < 6; ++ki) {
          const int nj = ki >> 1, h = ki & 1;
          const double* T = Pt + (mi <= nj ? m24_t3(mi, nj) : m24_t3(nj, mi)) * 64;
          pa[mi][ki] = T[mi < nj ? (h ? dir1 : dir0) : (mi > nj ? (h ? mir1 : mir0) : (h ? dia1 : dia0))];
        }
      if (BUILD && more) fetch_x(k - 1);   // lands under the tensor phase; the record is rebuilt in the scalar phase

      // ---------------- RIC_AFFINE: h for the own column = W[col][:] e + F[:][col]^T Vx, a 48-long (u-lanes: 32-long) DFMA
      // chain fed to the scheduler in pieces between the DMMAs
      double hacc = 0.0;
      if (PREP) hacc = base[M24P_PR + M24_PREP_DOUBLES * (k & 1) + M24_PREP_H0 + lane];
      __syncwarp();   // every lane has its fragments (the record and P may be overwritten); e is visible
      if (!FUSED && more) fetch_rec(k - 1);
      if (PREP && more) fetch_prep(k - 1);
      double vx_odd = 0.0;        // Vx is read in 16-byte pairs (broadcast): twelve wavefronts per step instead of 24
      auto hstep = [&](int i) {   // i = 0..47
        if (!AFF) return;
        if (PREP) {   // h0 = W e came with the record: one step of the F^T Vx chain per k-step of the G' products
          if (i & 1) return;
          const int l = i >> 1;
          double vx = vx_odd;
          if ((l & 1) == 0) { const double2 v = lds128(base + O_VS + l); vx = v.x; vx_odd = v.y; }
          hacc = fmad(Fs[(((l >> 1) & 1) ? fcolx : fcol) + l * fcs], vx, hacc);
        }
      };

      // ---------------- tensor phase
      double gacc[2][3][2];
      auto gprod = [&](int ni, int buf) {
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) { gacc[buf][mi][0] = 0.0; gacc[buf][mi][1] = 0.0; }
#pragma unroll
        for (int ki = 0; ki < 6; ++ki) {
#pragma unroll
          for (int mi = 0; mi < 3; ++mi)
            if (!FUSED || m24_tile_nz(ki, ni)) rm_dmma(gacc[buf][mi][0], gacc[buf][mi][1], pa[mi][ki], fb[ki][ni]);
          hstep(ni * 12 + 2 * ki);
          hstep(ni * 12 + 2 * ki + 1);
        }
      };
      auto gstore = [&](int buf) {
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) sts128(Gb + buf * 192 + mi * 64 + cs_off, gacc[buf][mi][0], gacc[buf][mi][1]);
      };
      // RIC_LQR: the Hxx tiles of H' stay in registers — they are the accumulators of P' (48 wavefronts less per step).  The
      // SLQ pass has no registers to spare for them (204 bytes of spills, 2.84 -> 3.01 ms): its tiles wait in the P slots.
      constexpr bool KEEP_HXX = !AFF;
      double pacc[6][2];
      double ldr[8], lu0 = 0.0, lu1 = 0.0;   // the distributed factorisation: 1 / D and the lane's two entries of Huu
      double* Ls = Hu;                         // column store of L: the unused slot of the u-row tiles
      gprod(0, 0);
      gstore(0);
      __syncwarp();
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        if (nj < 3) gprod(nj + 1, (nj + 1) & 1);
        double gb[6];
#pragma unroll
        for (int ki = 0; ki < 6; ++ki) gb[ki] = Gb[(nj & 1) * 192 + (ki >> 1) * 64 + ((ki & 1) ? mir1 : mir0)];
        double hacc2[4][2];
#pragma unroll
        for (int mi = 0; mi <= nj; ++mi) {
          if (mi == 0 && nj > 0) { hacc2[0][0] = 0.0; hacc2[0][1] = 0.0; continue; }   // W is block diagonal: Hux starts from 0
          const double2 wv = lds128(Wt + (PERW ? (mi == 0 ? 0 : m24_t4(mi, nj) - 3) : m24_t4(mi, nj)) * 64 + g * 8 + 2 * t);
          hacc2[mi][0] = wv.x; hacc2[mi][1] = wv.y;
        }
#pragma unroll
        for (int ki = 0; ki < 6; ++ki)
#pragma unroll
          for (int mi = 0; mi <= nj; ++mi)
            if (!FUSED || m24_tile_nz(ki, mi)) rm_dmma(hacc2[mi][0], hacc2[mi][1], fb[ki][mi], gb[ki]);
        if (nj < 3) gstore((nj + 1) & 1);
        if (nj == 0) {
          // Huu[a][b] = H'(7-a, 7-b) (+ mu on the diagonal), row-major for the factorisation
          double d0 = hacc2[0][0], d1 = hacc2[0][1];
          if (AFF) {
            if (g == 2 * t) d0 = d0 + mu;
            if (g == 2 * t + 1) d1 = d1 + mu;
          }
          Us[(7 - g) * 8 + (7 - 2 * t)] = d0;
          Us[(7 - g) * 8 + (6 - 2 * t)] = d1;
          lu0 = d0; lu1 = d1;
        } else {
          sts128(Hu + nj * 64 + cs_off, hacc2[0][0], hacc2[0][1]);
#pragma unroll
          for (int mi = 1; mi <= nj; ++mi) {
            if (KEEP_HXX) { pacc[m24_t3(mi - 1, nj - 1)][0] = hacc2[mi][0]; pacc[m24_t3(mi - 1, nj - 1)][1] = hacc2[mi][1]; }
            else sts128(Pt + m24_t3(mi - 1, nj - 1) * 64 + cs_off, hacc2[mi][0], hacc2[mi][1]);
          }
        }
        __syncwarp();
        // the factorisation, two columns per column tile of H', in the shadow of the DMMAs still to come
        if (nj == 0) { failed |= m24_ldl_pivot<0>(lu0, lu1, ldr, Ls, lane); failed |= m24_ldl_pivot<1>(lu0, lu1, ldr, Ls, lane); }
        if (nj == 1) {
          failed |= m24_ldl_pivot<2>(lu0, lu1, ldr, Ls, lane); failed |= m24_ldl_pivot<3>(lu0, lu1, ldr, Ls, lane);
          failed |= m24_ldl_pivot<4>(lu0, lu1, ldr, Ls, lane);
        }
        if (nj == 2) {
          failed |= m24_ldl_pivot<5>(lu0, lu1, ldr, Ls, lane); failed |= m24_ldl_pivot<6>(lu0, lu1, ldr, Ls, lane);
          failed |= m24_ldl_pivot<7>(lu0, lu1, ldr, Ls, lane);
        }
      }
      if (AFF && !xlane) base[O_HV + (7 - p)] = hacc;
      __syncwarp();

      // ---------------- scalar phase: gains of the own state column (lanes 8..31) and the feed-forward (lanes 0..7)
      double hux[8], hur[8], kk[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) hux[c] = Hcol[m24_el(7 - c, 0) ^ jc];   // (jc < 8 only touches the low three bits)
      if (AFF) {
#pragma unroll
        for (int c = 0; c < 8; c += 2) { const double2 v = lds128(base + O_HV + c); hur[c] = v.x; hur[c + 1] = v.y; }
      }
      {
        double rhs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rhs[c] = (AFF && !xlane) ? hur[c] : hux[c];
        m24_solve_neg(Ls, ldr, rhs, kk);
      }
      bool cl[8] = {false, false, false, false, false, false, false, false};
      if (AFF) {
        double kf[8];
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 8; c += 2) sts128(base + O_KF + c, kk[c], kk[c + 1]);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; c += 2) { const double2 v = lds128(base + O_KF + c); kf[c] = v.x; kf[c + 1] = v.y; }
        if (BOX) {
          // input box (oracle: noc_oracle_backward_affine_box): a feed-forward that leaves the box is replaced by the box
          // QP's, and `ch` by the factorisation with the clamped inputs masked out, which the gain solves then use.
          // The whole warp is one instance: no divergence between instances, every shared-memory access a broadcast.
          double lo[8], hi[8];
          bool inside = true;
          const double* ub = base + M24P_PR + M24_PREP_DOUBLES * (k & 1) + M24_PREP_U;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double uc = ub[c];
            lo[c] = a.u_min - uc;
            hi[c] = a.u_max - uc;
            if (!(kf[c] > lo[c] && kf[c] < hi[c])) inside = false;
          }
          if (!inside) {
            double* ws = base + O_QP;
            unsigned clmask = 0, fmask = ~0u;
            m24_box_qp(Us, hur, lo, hi, kf, clmask, ws, Ls, lane, ldr, fmask);
#pragma unroll
            for (int c = 0; c < 8; ++c) cl[c] = ((clmask >> c) & 1u) != 0;
            if (fmask != clmask) failed |= m24_ldl_masked(Us, clmask, ldr, Ls, lane);   // (else: Ls / ldr already hold this factorisation)
            double rhs[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) rhs[c] = cl[c] ? 0.0 : hux[c];
            m24_solve_neg(Ls, ldr, rhs, kk);
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < 8; c += 2) sts128(base + O_KF + c, kf[c], kf[c + 1]);
            }
          }
        }
        // Vx' = h_x + Hux^T k for the own state column
        if (xlane) {
          double s = hacc;
#pragma unroll
          for (int c = 0; c < 8; ++c) s = fmad(hux[c], kf[c], s);
          base[O_VS + j] = s;
        }
        // dV1 += k' h_u ; dV2 += k' Huu k / 2 (row r of Huu k on lane r, the chain over rows on every lane)
        double t1 = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) t1 = fmad(kf[c], hur[c], t1);
        {
          const int r = lane & 7;
          double rr = 0.0;
#pragma unroll
          for (int c = 0; c < 8; ++c) rr = fmad(c <= r ? Us[r * 8 + c] : Us[c * 8 + r], kf[c], rr);
          if (lane < 8) base[O_RR + r] = rr;
        }
        __syncwarp();
        double t2 = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t2 = fmad(kf[r], base[O_RR + r], t2);
        dV1 = dV1 + t1;
        dV2 = dV2 + 0.5 * t2;
        if (lane < 8) __stcs(a.kff + ((size_t)b * N + k) * 8 + lane, base[O_KF + lane]);
      }
      if (failed && kfail < 0) kfail = k;
      if (DARE) {
        dare_iters = it + 1;        // (the failing iteration counts: the oracle returns -it)
        if (failed) break;
#pragma unroll
        for (int c = 0; c < 8; ++c) kk_last[c] = kk[c];
      }
      // K_k: to global memory (row a of K is 24 consecutive doubles over the lanes) and, for the P' product, to shared
      double* Kb = Gb;
      if (xlane) {
#pragma unroll
        for (int c = 0; c < 8; ++c) Kb[c * 24 + j] = kk[c];
        if (Kp) {
#pragma unroll
          for (int c = 0; c < 8; ++c) __stcs(Kp + c * 24 + j, kk[c]);
        }
        if (!DARE && k == 0 && a.K0) {
          double* kp = a.K0 + (size_t)b * 192;
#pragma unroll
          for (int c = 0; c < 8; ++c) kp[c * 24 + j] = kk[c];
        }
      }
      if (Kp) Kp -= 192;
      __syncwarp();

      // ---------------- P' = Hxx + Hux^T K: A = Hux^T (from the u-row tiles), B = K, accumulators = the Hxx tiles in place
      {
        double af[3][2], bf[2][3];
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) { af[mi][0] = Hu[(1 + mi) * 64 + ar0]; af[mi][1] = Hu[(1 + mi) * 64 + ar1]; }
#pragma unroll
        for (int ki = 0; ki < 2; ++ki)
#pragma unroll
          for (int ni = 0; ni < 3; ++ni) bf[ki][ni] = Kb[(4 * ki + t) * 24 + 8 * ni + g];
        if (!KEEP_HXX) {
#pragma unroll
          for (int T = 0; T < 6; ++T) { const double2 v = lds128(Pt + T * 64 + cs_off); pacc[T][0] = v.x; pacc[T][1] = v.y; }
        }
#pragma unroll
        for (int ki = 0; ki < 2; ++ki)
#pragma unroll
          for (int mi = 0; mi < 3; ++mi)
#pragma unroll
            for (int nj = mi; nj < 3; ++nj) rm_dmma(pacc[m24_t3(mi, nj)][0], pacc[m24_t3(mi, nj)][1], af[mi][ki], bf[ki][nj]);
        // (the next record in the shadow of the product: its inputs arrived during the tensor phase)
        if (BUILD && more) {
          cp_async_wait_all();
          __syncwarp();
          build_rec(k - 1);
        }
        if (PREP && more) {
          mbar_wait(bar, phase);
          phase ^= 1u;
          __syncwarp();
          scatter_prep(k - 1);
        }
        if (DARE) {
          // max|P' - P| and max|P'| over the upper triangle (P, P' symmetric): the lane's entries of the six tiles, then the warp
          double dmax = 0.0, pmax = 0.0;
#pragma unroll
          for (int mi = 0; mi < 3; ++mi)
#pragma unroll
            for (int nj = mi; nj < 3; ++nj) {
              const int T = m24_t3(mi, nj);
              const double2 old = lds128(Pt + T * 64 + cs_off);
              if (mi < nj || g <= 2 * t) { dmax = fmax(dmax, fabs(pacc[T][0] - old.x)); pmax = fmax(pmax, fabs(pacc[T][0])); }
              if (mi < nj || g <= 2 * t + 1) { dmax = fmax(dmax, fabs(pacc[T][1] - old.y)); pmax = fmax(pmax, fabs(pacc[T][1])); }
            }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
            pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
          }
          converged = dmax < a.tol * pmax;
        }
#pragma unroll
        for (int T = 0; T < 6; ++T) sts128(Pt + T * 64 + cs_off, pacc[T][0], pacc[T][1]);
      }
      if (DARE && converged) break;
    }  // steps
    __syncwarp();
    if (!AFF) break;
    // (the adaptive schedule's factor lives in global memory: it is touched once per sweep)
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

  if (AFF) {
    if (lane == 0) {
      a.mu[b] = mu;
      a.dV[(size_t)b * 2] = dV1;
      a.dV[(size_t)b * 2 + 1] = dV2;
      if (a.status) a.status[b] = reg_fail ? 4 : 0;
    }
    return;
  }
  const double nanv = qnan();
  auto pel = [&](int i, int c) {   // P[i][c] from the upper tiles
    const int r = i <= c ? i : c, cc = i <= c ? c : i;
    return Pt[m24_t3(r >> 3, cc >> 3) * 64 + m24_el(r & 7, cc & 7)];
  };
  if (DARE) {   // outputs of k_ric_generic's DARE mode: P, the stationary gain, the iteration count, OK / NOT_PD / MAX_ITER
    if (a.K0 && xlane) {
      double* kp = a.K0 + (size_t)b * 192;
#pragma unroll
      for (int c = 0; c < 8; ++c) kp[c * 24 + j] = failed ? nanv : kk_last[c];
    }
    if (a.P0) {
      double* po = a.P0 + (size_t)b * 576;
      for (int e = lane; e < 576; e += 32) po[e] = failed ? nanv : pel(e / 24, e % 24);
    }
    if (lane == 0) {
      if (a.iters) a.iters[b] = dare_iters;
      if (a.status) a.status[b] = failed ? 1 : (converged ? 0 : 2);
    }
    return;
  }
  // ---------------- RIC_LQR outputs (the rules of k_ric24<RIC_LQR>)
  if (failed && xlane) {
    if (a.K)
      for (int k2 = 0; k2 <= kfail; ++k2) {
        double* kp = a.K + ((size_t)b * N + k2) * 192;
        for (int c = 0; c < 8; ++c) kp[c * 24 + j] = nanv;
      }
    if (a.K0) {
      double* kp = a.K0 + (size_t)b * 192;
      for (int c = 0; c < 8; ++c) kp[c * 24 + j] = nanv;
    }
  }
  if (a.P0) {
    double* po = a.P0 + (size_t)b * 576;
    for (int e = lane; e < 576; e += 32) po[e] = failed ? nanv : pel(e / 24, e % 24);
  }
  if (a.cost && a.x0) {
    // cost = 0.5 x0' P0 x0: one row chain per lane, the chain over rows on lane 0 (spec order)
    const double* x = a.x0 + (size_t)b * 24;
    double* rs = Gb;
    if (lane < 24) {
      double r = 0.0;
      for (int c = 0; c < 24; ++c) r = fmad(pel(lane, c), x[c], r);
      rs[lane] = r;
    }
    __syncwarp();
    if (lane == 0) {
      double c = 0.0;
      for (int i = 0; i < 24; ++i) c = fmad(x[i], rs[i], c);
      a.cost[b] = failed ? nanv : 0.5 * c;
    }
  }
  if (lane == 0 && a.status) a.status[b] = failed ? 1 : 0;
}

}  // namespace noc

// riccati_scan.cuh — TIME-PARALLEL Riccati sweep for n=12, m=4: the latency path for small batches
// (single-vehicle MPC).  The sequential sweep is a chain of N dependent steps; here one CTA owns one instance and its
// warps own CHUNKS of the horizon that are processed concurrently, glued together by the associative operator of
// Särkkä & García-Fernández, "Temporal parallelization of dynamic programming and linear quadratic control"
// (IEEE TAC 2023): an interval i -> j of the horizon is summarised by the conditional value function element
// (A, C, J); two adjacent intervals combine as
//     A_il = A_jl (I + C_ij J_jl)^-1 A_ij
//     C_il = A_jl (I + C_ij J_jl)^-1 C_ij A_jl' + C_jl
//     J_il = A_ij' (I + J_jl C_ij)^-1 J_jl A_ij + J_ij
// and the cost-to-go of step k is the J of the interval k -> N combined with the terminal element (0, 0, Qf).
//
// Two-level scan, chosen so that every inverse but C-1 of them is the m x m SPD system of an ordinary Riccati step:
//   phase A  (all chunks at once) chunk c folds its steps from the right, a_k (x) agg.  The LEFT operand is always a
//            single step (A_k, B_k R^-1 B_k', Q), for which Woodbury turns the n x n inverse into the step's own
//            Huu = R + B'JB:   J <- Q + A'J(A + BK)  (the Riccati step on J, started from 0),   A_agg <- A_agg (A + BK),
//            C_agg <- C_agg + (A_agg B) Huu^-1 (A_agg B)'.        The LAST chunk starts from Qf instead, i.e. it simply
//            runs the true recursion and emits its gains.
//   phase B  (a chain over the C-1 chunk boundaries, from the end) S_start(c) = J_c + A_c' (I + S C_c)^-1 S A_c with
//            S = S_start(c+1): one general 12 x 12 solve each (Gauss-Jordan with partial pivoting in shared memory).
//   phase C  (chunks again at once, each as soon as its boundary value exists) the ordinary Riccati steps of chunk c
//            from S_start(c+1): the gains K_k; chunk 0 also delivers P0 and the cost.
// Depth ~ 2.5 L + (C-1) boundary solves instead of N steps (L = chunk length).  The arithmetic differs from the
// sequential sweep (other association, a pivoted solve), so parity with the oracle is by TOLERANCE (1e-9 relative,
// north_star's bar; measured ~1e-13), not bitwise: the path is opt-in (NOC_FLAG_TIME_PARALLEL).
#pragma once
#include "common.cuh"
#include "riccati12.cuh"

namespace noc {

constexpr int TP_MAX_CHUNKS = 8;
// per-warp shared memory, in doubles
constexpr int TP_FS = 0;              // 2 x record [A 12x12 | B 12x4] (double buffer)
constexpr int TP_JS = 384;            // J / S of the running recursion, 12x12
constexpr int TP_GS = 528;            // G = J F, 12x16
constexpr int TP_HS = 720;            // H = W + F'G, 16x16 (phase B: the 12x24 augmented system lives in GS..HS)
constexpr int TP_KS = 976;            // K, 4x12
constexpr int TP_AA = 1024;           // A_agg, transposed
constexpr int TP_CA = 1168;           // C_agg
constexpr int TP_T1 = 1312;           // X = [A + BK | B] in record layout (192)
constexpr int TP_ES = 1504;           // the new A_agg (transposed) while the old one is still being read (144)
constexpr int TP_NS = 1648;           // Huu^-1 E', 4x12
constexpr int TP_WARP_DOUBLES = 1696;
constexpr int TP_CTA_DOUBLES = 16 * 16 + 144 + TP_MAX_CHUNKS * 144 + 16;   // W | Qf | boundary values | flags

inline size_t ric_tp_smem_bytes(int chunks) { return sizeof(double) * (size_t)(TP_CTA_DOUBLES + chunks * TP_WARP_DOUBLES); }

struct RicTpArgs {
  const double* AB;      // [batch][N][192], A_THEN_B records
  const double* W;       // [16][16] symmetric blockdiag(Q,R) (k_build_w)
  const double* Qf;      // [12][12]
  const double* x0;      // [batch][12] or nullptr (cost only)
  double* K;             // [batch][N][48] or nullptr
  double* K0;            // [batch][48] or nullptr
  double* P0;            // [batch][144] or nullptr
  double* cost;          // [batch] or nullptr
  int32_t* status;       // [batch] or nullptr
  long long batch;
  int N;
  int chunks;            // warps per CTA
  int len;               // steps per chunk, chunks 0 .. chunks-2; the last chunk takes the rest (>= 1)
};

// upper-triangle enumeration of a 12x12 matrix: entry e -> (i, j), i <= j
__device__ __forceinline__ void tp_tri(int e, int& i, int& j) {
  // row i starts at e_i = i*12 - i*(i-1)/2
  int r = 0, start = 0;
#pragma unroll
  for (int q = 1; q < 12; ++q) {
    const int s = q * 12 - (q * (q - 1)) / 2;
    if (e >= s) { r = q; start = s; }
  }
  i = r;
  j = r + (e - start);
}

// FP64 tensor-core tile: D(8x8) += A(8x4) B(4x8), mma.sync.m8n8k4 (SASS: DMMA).  Fragments of lane (g = lane>>2,
// t = lane&3): a = A[g][t], b = B[t][g], {d0, d1} = D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// DMMA = true: the 12x12x16 products of the step (G = J F, H = W + F'G, the fold's A_agg [A+BK | B]) run on the FP64
// tensor cores (24 + 12 DMMAs per step instead of ~240 DFMAs and their shared-memory operand loads).  This path is
// tolerance-parity by construction (the scan re-associates anyway), so the fragment accumulation order is no obstacle
// here — unlike in the bitwise sweeps of riccati12.cuh.
template <bool DMMA>
struct TpWarp {
  double* w;          // this warp's shared memory
  const double* Ws;   // W 16x16
  int lane;
  // loop-invariant index maps of this lane (integer divisions cost more than the arithmetic they index)
  int ti[3], tj[3];   // upper-triangle entries lane + 32 q  (ti < 0: none)
  int fi[5], fj[5];   // full 12x12 entries lane + 32 q      (fi < 0: none)
  __device__ __forceinline__ void init_maps() {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = lane + 32 * q;
      ti[q] = -1; tj[q] = 0;
      if (e < 78) tp_tri(e, ti[q], tj[q]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int e = lane + 32 * q;
      fi[q] = (e < 144) ? e / 12 : -1;
      fj[q] = e % 12;
    }
  }
  __device__ __forceinline__ const double* F(int buf) const { return w + TP_FS + buf * 192; }
  // element (l, c) of a 12x16 matrix stored as [12x12 | 12x4] (the record layout A_THEN_B)
  __device__ __forceinline__ static int foff(int l, int c) { return c < 12 ? l * 12 + c : 144 + l * 4 + (c - 12); }

  // out (12x16, row stride 16) = L X, where X is a 12x16 matrix in record layout and L is given by LT with
  // LT[l*12 + r] = L[r][l] (for a symmetric L, LT = L).  Lane (c = lane & 15, half = lane >> 4) owns column c, rows
  // 6 half .. 6 half + 5; if `transposed_out` the 12x12 part of the result is ALSO stored transposed at outT
  // (outT[c*12 + r] = out[r][c]) — how A_agg is kept.
  __device__ __forceinline__ void mul12x16(const double* LT, const double* X, double* out, double* outT) {
    if (DMMA) {
      const int g = lane >> 2, t = lane & 3;
      double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) {
        const double a0 = LT[(4 * ki + t) * 12 + g];                               // L[g][4ki+t]
        const double a1 = (g < 4) ? LT[(4 * ki + t) * 12 + 8 + g] : 0.0;           // L[8+g][4ki+t], rows 12..15 are padding
        const double b0 = X[foff(4 * ki + t, g)], b1 = X[foff(4 * ki + t, 8 + g)];
        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
      }
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        sts128(out + g * 16 + 8 * ni + 2 * t, acc[0][ni][0], acc[0][ni][1]);
        if (g < 4) sts128(out + (8 + g) * 16 + 8 * ni + 2 * t, acc[1][ni][0], acc[1][ni][1]);
      }
      if (outT) {   // out[r][c] -> outT[c*12 + r], c < 12
        outT[(2 * t) * 12 + g] = acc[0][0][0];
        outT[(2 * t + 1) * 12 + g] = acc[0][0][1];
        if (t < 2) { outT[(8 + 2 * t) * 12 + g] = acc[0][1][0]; outT[(9 + 2 * t) * 12 + g] = acc[0][1][1]; }
        if (g < 4) {
          outT[(2 * t) * 12 + 8 + g] = acc[1][0][0];
          outT[(2 * t + 1) * 12 + 8 + g] = acc[1][0][1];
          if (t < 2) { outT[(8 + 2 * t) * 12 + 8 + g] = acc[1][1][0]; outT[(9 + 2 * t) * 12 + 8 + g] = acc[1][1][1]; }
        }
      }
      return;
    }
    const int c = lane & 15, r0 = 6 * (lane >> 4);
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int l = 0; l < 12; ++l) {
      const double f = X[foff(l, c)];
      const double2 a = lds128(LT + l * 12 + r0), b = lds128(LT + l * 12 + r0 + 2), d = lds128(LT + l * 12 + r0 + 4);
      acc[0] = fmad(a.x, f, acc[0]); acc[1] = fmad(a.y, f, acc[1]);
      acc[2] = fmad(b.x, f, acc[2]); acc[3] = fmad(b.y, f, acc[3]);
      acc[4] = fmad(d.x, f, acc[4]); acc[5] = fmad(d.y, f, acc[5]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) out[(r0 + i) * 16 + c] = acc[i];
    if (outT && c < 12) {
      sts128(outT + c * 12 + r0, acc[0], acc[1]);
      sts128(outT + c * 12 + r0 + 2, acc[2], acc[3]);
      sts128(outT + c * 12 + r0 + 4, acc[4], acc[5]);
    }
  }

  // One Riccati step on the J held in JS with the record in F(buf): G = J F, H = W + F'G, the 4x4 LDL' of Huu,
  // K -> KS (and global memory).  Returns true if a pivot failed.  update_J() then overwrites J.
  __device__ __forceinline__ bool step(int buf, double* Kglobal, Ldl4& ch) {
    const double* Fs = F(buf);
    double* Gs = w + TP_GS;
    double* Hs = w + TP_HS;
    double* Ks = w + TP_KS;
    const int c = lane & 15, half = lane >> 4;
    mul12x16(w + TP_JS, Fs, Gs, nullptr);   // J symmetric: row l of J = column l
    __syncwarp();
    if (DMMA) {   // H = W + F'G, 16x16x12: A fragments F'[8mi+g][4ki+t] = F[4ki+t][8mi+g], B fragments G[4ki+t][8ni+g]
      const int g = lane >> 2, t = lane & 3;
      double acc[2][2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const double2 v = lds128(Ws + (8 * mi + g) * 16 + 8 * ni + 2 * t);
          acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
        }
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) {
        const double a0 = Fs[foff(4 * ki + t, g)], a1 = Fs[foff(4 * ki + t, 8 + g)];
        const double b0 = Gs[(4 * ki + t) * 16 + g], b1 = Gs[(4 * ki + t) * 16 + 8 + g];
        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) sts128(Hs + (8 * mi + g) * 16 + 8 * ni + 2 * t, acc[mi][ni][0], acc[mi][ni][1]);
    } else {   // H = W + F'G: column c, rows 8*half .. +7
      double acc[8];
      const int i0 = 8 * half;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = Ws[(i0 + i) * 16 + c];
#pragma unroll
      for (int l = 0; l < 12; ++l) {
        const double g = Gs[l * 16 + c];
        double fr[8];
        if (half == 0) {
#pragma unroll
          for (int i = 0; i < 8; i += 2) { const double2 v = lds128(Fs + l * 12 + i); fr[i] = v.x; fr[i + 1] = v.y; }
        } else {
          const double2 v0 = lds128(Fs + l * 12 + 8), v1 = lds128(Fs + l * 12 + 10);
          const double2 v2 = lds128(Fs + 144 + l * 4), v3 = lds128(Fs + 144 + l * 4 + 2);
          fr[0] = v0.x; fr[1] = v0.y; fr[2] = v1.x; fr[3] = v1.y; fr[4] = v2.x; fr[5] = v2.y; fr[6] = v3.x; fr[7] = v3.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmad(fr[i], g, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) Hs[(i0 + i) * 16 + c] = acc[i];
    }
    __syncwarp();
    // Huu = L D L' (every lane, redundantly), gains by the twelve lanes that own a state column
    const double* U = Hs + 12 * 16 + 12;
    const bool bad = ch.factor(U[0], U[16], U[17], U[32], U[33], U[34], U[48], U[49], U[50], U[51]);
    if (lane < 12) {
      double rhs[4], kk[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) rhs[a] = Hs[(12 + a) * 16 + lane];
      ch.solve_neg(rhs, kk);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        Ks[a * 12 + lane] = kk[a];
        if (Kglobal) Kglobal[a * 12 + lane] = kk[a];
      }
    }
    __syncwarp();
    return bad;
  }

  // J <- Hxx + Hux'K on the upper triangle, mirrored
  __device__ __forceinline__ void update_J() {
    double* Js = w + TP_JS;
    const double* Hs = w + TP_HS;
    const double* Ks = w + TP_KS;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = ti[q], j = tj[q];
      if (i >= 0) {
        double v = Hs[i * 16 + j];
#pragma unroll
        for (int a = 0; a < 4; ++a) v = fmad(Hs[(12 + a) * 16 + i], Ks[a * 12 + j], v);
        Js[i * 12 + j] = v;
        Js[j * 12 + i] = v;
      }
    }
    __syncwarp();
  }

  // phase A bookkeeping of the interval element after step(): A_agg <- A_agg (A + BK), C_agg += (A_agg B) Huu^-1 (A_agg B)'.
  // A_agg is kept TRANSPOSED (AA[l*12 + r] = A_agg[r][l]) so that [A_agg (A+BK) | A_agg B] is one mul12x16.
  __device__ __forceinline__ void fold(int buf, const Ldl4& ch) {
    const double* Fs = F(buf);
    const double* Ks = w + TP_KS;
    double *AaT = w + TP_AA, *Ca = w + TP_CA, *X = w + TP_T1, *P = w + TP_GS, *Ns = w + TP_NS;
    // X = [A + BK | B] in record layout (T1 and T2 are adjacent: 144 + 48 doubles)
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int i = fi[q], j = fj[q];
      if (i >= 0) {
        double v = Fs[i * 12 + j];
#pragma unroll
        for (int a = 0; a < 4; ++a) v = fmad(Fs[144 + i * 4 + a], Ks[a * 12 + j], v);
        X[i * 12 + j] = v;
      }
    }
    if (lane < 24) sts128(X + 144 + 2 * lane, Fs[144 + 2 * lane], Fs[144 + 2 * lane + 1]);
    __syncwarp();
    // P = A_agg X = [new A_agg | E]; the new A_agg goes (transposed) to T2's second half... to a scratch, since AaT is
    // still being read by other lanes
    double* AaT_new = w + TP_ES;   // 144 doubles: ES | NS' region is sized for it (see TP_ES)
    mul12x16(AaT, X, P, AaT_new);
    __syncwarp();
    if (lane < 12) {   // N = Huu^-1 E'  (column `lane`)
      double rhs[4], z[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) rhs[a] = P[lane * 16 + 12 + a];
      ch.solve_neg(rhs, z);
#pragma unroll
      for (int a = 0; a < 4; ++a) Ns[a * 12 + lane] = -z[a];
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int e = lane + 32 * q;
      if (e < 144) AaT[e] = AaT_new[e];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; ++q) {   // C_agg += E N  (symmetric)
      const int i = ti[q], j = tj[q];
      if (i >= 0) {
        double v = Ca[i * 12 + j];
#pragma unroll
        for (int a = 0; a < 4; ++a) v = fmad(P[i * 16 + 12 + a], Ns[a * 12 + j], v);
        Ca[i * 12 + j] = v;
        Ca[j * 12 + i] = v;
      }
    }
    __syncwarp();
  }

  // phase B: out = J + A_agg' (I + S C_agg)^-1 S A_agg, S given (12x12, shared memory), J = JS.
  // The 12 x 24 system [I + S C | S A] is formed by the whole warp in shared memory, then lane r < 12 takes row r into
  // registers and the warp runs Gauss-Jordan with partial pivoting on registers (pivot search and pivot-row broadcast by
  // shuffles; all indices static).  Returns false if the system is numerically singular.
  __device__ __forceinline__ bool boundary(const double* S, double* out) {
    const double *AaT = w + TP_AA, *Ca = w + TP_CA, *Js = w + TP_JS;
    double* M = w + TP_GS;   // 12 x 24, row stride 24 (GS..HS: 448 doubles available)
    {
      const int cl = lane & 7, rl = lane >> 3;
      // column c = cl + 8 b of [C_agg | A_agg]: element (l, c) at base[b] + l * stride[b]
      const double* base[3] = {Ca + cl, (cl < 4) ? Ca + 8 + cl : AaT + (cl - 4) * 12, AaT + (cl + 4) * 12};
      const int stride[3] = {12, (cl < 4) ? 12 : 1, 1};
      double acc[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b] = (rl + 4 * a == cl + 8 * b) ? 1.0 : 0.0;
#pragma unroll
      for (int l = 0; l < 12; ++l) {
        double sv[3], rv[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) sv[a] = S[(rl + 4 * a) * 12 + l];
#pragma unroll
        for (int b = 0; b < 3; ++b) rv[b] = base[b][l * stride[b]];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[a][b] = fmad(sv[a], rv[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) M[(rl + 4 * a) * 24 + cl + 8 * b] = acc[a][b];
    }
    __syncwarp();
    double m[24];
    const int row = lane < 12 ? lane : 0;
#pragma unroll
    for (int c = 0; c < 24; c += 2) { const double2 v = lds128(M + row * 24 + c); m[c] = v.x; m[c + 1] = v.y; }
    bool ok = true, used = false;
    int var = 0;   // the unknown whose solution row ends up in this lane
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      double v = (lane < 12 && !used) ? fabs(m[j]) : -1.0;
      int idx = lane;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
      }
      const int p = __shfl_sync(0xffffffffu, idx, 0);
      const double pv = __shfl_sync(0xffffffffu, m[j], p);
      if (!(fabs(pv) > 1e-300)) ok = false;
      const double inv = 1.0 / pv;
      const double f = m[j];
      const bool is_p = (lane == p);
#pragma unroll
      for (int c = j + 1; c < 24; ++c) {
        const double r = __shfl_sync(0xffffffffu, m[c], p) * inv;
        m[c] = is_p ? r : fmad(-f, r, m[c]);
      }
      if (is_p) { used = true; var = j; }
    }
    __syncwarp();
    double* Xs = M;   // solution rows, 12 x 12 (every lane has its row in registers: M can be overwritten)
    if (lane < 12) {
#pragma unroll
      for (int c = 0; c < 12; c += 2) sts128(Xs + var * 12 + c, m[12 + c], m[12 + c + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; ++q) {   // out = J + A_agg' X on the upper triangle, mirrored
      const int i = ti[q], j = tj[q];
      if (i >= 0) {
        double v = Js[i * 12 + j];
#pragma unroll
        for (int l = 0; l < 12; ++l) v = fmad(AaT[i * 12 + l], Xs[l * 12 + j], v);
        out[i * 12 + j] = v;
        out[j * 12 + i] = v;
      }
    }
    __syncwarp();
    return ok;
  }
};

template <bool DMMA>
__global__ void __launch_bounds__(TP_MAX_CHUNKS * 32, 1) k_ric12_tp(const RicTpArgs a) {
  extern __shared__ __align__(128) double smem[];
  double* Ws = smem;
  double* Qfs = smem + 256;
  double* Sb = smem + 256 + 144;                       // boundary values: Sb[c] = cost-to-go at the START of chunk c
  volatile int* ready = reinterpret_cast<volatile int*>(smem + 256 + 144 + TP_MAX_CHUNKS * 144);   // ready[c]: Sb[c] is valid; [8+c]: failed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.chunks, N = a.N;
  const long long b = blockIdx.x;
  TpWarp<DMMA> t{smem + TP_CTA_DOUBLES + warp * TP_WARP_DOUBLES, Ws, lane};
  t.init_maps();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) Ws[i] = a.W[i];
  for (int i = threadIdx.x; i < 144; i += blockDim.x) Qfs[i] = a.Qf[i];
  if (threadIdx.x < 2 * TP_MAX_CHUNKS) ready[threadIdx.x] = 0;
  __syncthreads();

  const int c = warp;
  const bool last = (c == C - 1);
  const int k_begin = c * a.len;
  const int k_end = last ? N : k_begin + a.len;        // chunk = steps [k_begin, k_end)
  const double* rec = a.AB + ((size_t)b * N) * R12_REC;
  double* Kg = a.K ? a.K + (size_t)b * N * 48 : nullptr;
  auto fetch = [&](int k, int buf) {
    const double* src = rec + (size_t)k * R12_REC;
    double* dst = t.w + TP_FS + buf * 192;
#pragma unroll
    for (int q = 0; q < 3; ++q) cp_async16(smem_u32(dst + 2 * (lane + 32 * q)), src + 2 * (lane + 32 * q));
    cp_async_commit();
  };
  double* Js = t.w + TP_JS;
  bool failed = false;
  int kfail = -1;   // one chunk: the step at which the factorisation first failed (the sequential sweep's NaN rule)
  Ldl4 ch;

  // ---------------- phase A
  fetch(k_end - 1, 0);
  for (int e = lane; e < 144; e += 32) {
    Js[e] = last ? Qfs[e] : 0.0;
    t.w[TP_AA + e] = (e / 12 == e % 12) ? 1.0 : 0.0;
    t.w[TP_CA + e] = 0.0;
  }
  int buf = 0;
  for (int k = k_end - 1; k >= k_begin; --k) {
    cp_async_wait_all();
    __syncwarp();
    if (k > k_begin) fetch(k - 1, buf ^ 1);
    // the last chunk runs the true recursion: its gains are final
    double* kout = (last && Kg) ? Kg + (size_t)k * 48 : nullptr;
    failed |= t.step(buf, kout, ch);
    if (failed && kfail < 0) kfail = k;
    if (last && k == 0 && a.K0 && lane < 12) {
#pragma unroll
      for (int q = 0; q < 4; ++q) a.K0[(size_t)b * 48 + q * 12 + lane] = t.w[TP_KS + q * 12 + lane];
    }
    if (!last) t.fold(buf, ch);
    t.update_J();
    buf ^= 1;
  }

  // ---------------- phase B: chain over the chunk boundaries, from the end of the horizon
  if (last) {
    for (int e = lane; e < 144; e += 32) Sb[c * 144 + e] = Js[e];
    __syncwarp();
    __threadfence_block();
    if (lane == 0) { ready[TP_MAX_CHUNKS + c] = failed ? 1 : 0; ready[c] = 1; }
  } else {
    while (ready[c + 1] == 0) __nanosleep(64);
    __threadfence_block();
    failed |= ready[TP_MAX_CHUNKS + c + 1] != 0;
    const double* S = Sb + (c + 1) * 144;
    if (!t.boundary(S, Sb + c * 144)) failed = true;
    __threadfence_block();
    if (lane == 0) { ready[TP_MAX_CHUNKS + c] = failed ? 1 : 0; ready[c] = 1; }
    // ---------------- phase C: the chunk's true recursion from the cost-to-go at its end
    for (int e = lane; e < 144; e += 32) Js[e] = S[e];
    fetch(k_end - 1, 0);
    buf = 0;
    for (int k = k_end - 1; k >= k_begin; --k) {
      cp_async_wait_all();
      __syncwarp();
      if (k > k_begin) fetch(k - 1, buf ^ 1);
      failed |= t.step(buf, Kg ? Kg + (size_t)k * 48 : nullptr, ch);
      if (k == 0 && a.K0 && lane < 12) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a.K0[(size_t)b * 48 + q * 12 + lane] = t.w[TP_KS + q * 12 + lane];
      }
      t.update_J();
      buf ^= 1;
    }
  }

  // ---------------- outputs of chunk 0: P0, cost; the instance's status once every chunk is done
  if (c == 0) {
    if (a.P0)
      for (int e = lane; e < 144; e += 32) a.P0[(size_t)b * 144 + e] = Js[e];
    if (a.cost && a.x0) {
      // the spec's order: row chains r_i = P0[i][:] x0, then the chain over i, then the half
      const double* x = a.x0 + (size_t)b * 12;
      double* rs = t.w + TP_KS;
      if (lane < 12) {
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) r = fmad(Js[lane * 12 + j], x[j], r);
        rs[lane] = r;
      }
      __syncwarp();
      if (lane == 0) {
        double cc = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) cc = fmad(x[i], rs[i], cc);
        a.cost[b] = 0.5 * cc;
      }
    }
  }
  if (lane == 0) ready[TP_MAX_CHUNKS + c] = failed ? 1 : 0;
  __syncthreads();
  bool any_failed = false;
  for (int q = 0; q < C; ++q) any_failed |= ready[TP_MAX_CHUNKS + q] != 0;
  if (any_failed) {   // NOT_PD anywhere (or a singular boundary system): every output of the instance is NaN
    const double nanv = qnan();
    // one chunk = the plain recursion: the gains of the steps after the failing one (k > kfail) are valid, as in k_ric12
    const int nan_steps = (C == 1) ? kfail + 1 : N;
    if (Kg)
      for (int e = threadIdx.x; e < nan_steps * 48; e += blockDim.x) Kg[e] = nanv;
    if (a.K0)
      for (int e = threadIdx.x; e < 48; e += blockDim.x) a.K0[(size_t)b * 48 + e] = nanv;
    if (a.P0)
      for (int e = threadIdx.x; e < 144; e += blockDim.x) a.P0[(size_t)b * 144 + e] = nanv;
    if (a.cost && threadIdx.x == 0) a.cost[b] = nanv;
  }
  if (a.status && threadIdx.x == 0) a.status[b] = any_failed ? 1 : 0;
}

}  // namespace noc

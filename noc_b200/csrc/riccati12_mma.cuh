// riccati12_mma.cuh — the n=12, m=4 LTV Riccati sweep (RIC_LQR, A_THEN_B records from HBM) with its two big products on
// the FP64 TENSOR CORES, bit-identical to k_ric12 and the oracle.
//
// Why this is allowed to be bitwise: tools/ubench/dmma_order.cu (profiles/ubench_dmma_order_r02.json) shows that on B200
// mma.sync.m8n8k4.f64 (SASS: DMMA) evaluates every output element as the SEQUENTIAL chain
//     d = fma(a3, b3, fma(a2, b2, fma(a1, b1, fma(a0, b0, c))))
// (8.4 M random elements, 100 % bit-equal; every other order of the four products matches <= 83 %).  A k-loop of DMMAs
// that feeds the accumulator back is therefore exactly DESIGN.md §2's "fma chain over l ascending" — G = P F from 0 and
// H = W + F'G from W, the two products that are 87 % of the sweep's flops — and multiplication commutes, so which factor
// is the A and which the B operand does not matter.
//
// Why it is faster than k_ric12's DFMA chains although DMMA and DFMA share one FP64 datapath (§3.1): k_ric12 issues 161
// FP64 warp-instructions and ~230 others per instance-step and keeps the pipe 55 % busy with two warps per scheduler
// (register-port conflicts of three-operand DFMAs, shared-memory operand loads, dependent chains).  Here the products of
// an instance are 21 DMMAs (each 256 FMAs, 16 pipe cycles) fed by 24 fragment loads: the warp has almost nothing else to
// issue, and the long latency chain of the factorisation runs while the scheduler's other warp streams DMMAs.
//
// Mapping.  A warp owns EIGHT instances, as in k_ric12.  Per step:
//   tensor phase, instance by instance (the whole warp works on one instance; lane = (g = lane>>2, t = lane&3)):
//     G' = P F'        12 x 12 x 16, M padded to 16: 12 DMMAs.   A fragments P[8mi+g][4ki+t] straight from the
//                      instance's P (row stride 12: conflict-free), B fragments F'[4ki+t][8ni+g] from the staged record.
//     H' = W' + F'^T G'  upper-triangle tiles (0,0), (0,1), (1,1) only: 9 DMMAs.  The A fragments F'^T[8mi+g][4ki+t] ARE
//                      the B fragments of the first product (same registers); the B fragments of G' come back from a
//                      swizzled per-warp shared-memory buffer (the C-fragment layout of a DMMA is not its B layout).
//     F' = F with its columns permuted to [u3 u2 u1 u0 | x0 .. x11]: in that order every entry the spec needs — Hxx[i][j]
//     for i <= j, Hux[a][j] (row u_a, column x_j) and Huu[a][b] for b <= a — lies in the UPPER triangle of H', so three
//     tiles suffice instead of four.  H' overwrites the instance's P (dead once G' exists).
//   scalar phase, all eight instances at once, four lanes per instance exactly as k_ric12's phases 3 and 4: redundant 4x4
//     L D L^T in registers, gain solves of the own three state columns, P' = Hxx + Hux^T K on the own columns (DFMA
//     chains, upper triangle exact: no padding), mirrored store of P' in the stride-12 layout the next step's A fragments
//     read.
//   The A_k,B_k records arrive by TMA bulk copies on a per-warp mbarrier; the copy of step k-1 is issued as soon as the
//   tensor phase of step k is over and lands under the scalar phase.
// Shared memory per instance: record 1536 B + P/H' 1536 B (+32 B skew) = 3104 B; per warp 8 instances + two G' buffers
// = 27,904 B; two CTAs of four warps per SM.
#pragma once
#include "common.cuh"
#include "riccati12.cuh"

namespace noc {

constexpr int RM_FS = 0;            // record A[12][12] | B[12][4]
constexpr int RM_PH = 192;          // P[12][12] (row stride 12) between steps; the three 8x8 tiles of H' inside a step
constexpr int RM_STRIDE = 388;      // per instance: 3104 B = 32 B modulo 128 B (the scalar phase touches all eight instances at once)
constexpr int RM_GBUF = 192;        // G' (12 x 16, swizzled), two buffers per warp
constexpr int RM_ZPAD = 16;         // zeros: the padding rows 12..15 of P in the second M tile
constexpr int RM_WARP_DOUBLES = 8 * RM_STRIDE + 2 * RM_GBUF + RM_ZPAD;
constexpr int RM_BAR_DOUBLES = 8;

// DIRECT variant: no staged records — the F' fragments are read straight from global memory (the records of step k-1 are
// pulled into L2 by bulk prefetches during step k), so an instance needs only its P / H' slot and twelve warps fit an SM
constexpr int RMD_STRIDE = 196;     // 1568 B = 32 B modulo 128 B
constexpr int RMD_WARP_DOUBLES = 8 * RMD_STRIDE + 2 * RM_GBUF + RM_ZPAD;

template <bool DIRECT>
inline size_t ric12_mma_smem_bytes(int warps) {
  return sizeof(double) * (size_t)(RM_BAR_DOUBLES + warps * (DIRECT ? RMD_WARP_DOUBLES : RM_WARP_DOUBLES));
}

__device__ __forceinline__ double rm_ldg_stream(const double* p) {   // read-once data: no L1 allocation
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void rm_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

// FP64 tensor-core tile D(8x8) += A(8x4) B(4x8).  Fragments of lane (g, t): a = A[g][t], b = B[t][g], {d0, d1} = D[g][2t], D[g][2t+1].
__device__ __forceinline__ void rm_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// one elected lane of the (converged) warp
__device__ __forceinline__ bool rm_elect_one() {
  unsigned pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// position p of the permuted order [u3 u2 u1 u0 | x0 .. x11] -> index in the W = blockdiag(Q, R) order [x0 .. x11 | u0 .. u3]
__host__ __device__ constexpr int rm_perm(int p) { return p < 4 ? 15 - p : p - 4; }

// H'(r, c), r <= c (positions of the permuted order), inside the instance's tile store: tiles (0,0), (0,1), (1,1)
__device__ __forceinline__ int rm_hoff(int r, int c) { return ((r >> 3) + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7); }

// column j of P' (rows 0..ROWS-1 held, valid for i <= j) -> the upper triangle of the stride-12 store
template <int ROWS>
__device__ __forceinline__ void rm_store_upper(double* Ps, const double* h, int j) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i)
    if (i <= j) Ps[i * 12 + j] = h[i];
}

// PERW: per-instance weights.  a.W is then the caller's QRQf array ([batch][Q 144 | R 16 | Qf 144], a.w_stride doubles
// apart) and every instance-step reads the C fragments of its own W' from L2 (three 16-byte loads per lane, requested two
// pipeline slots — about forty DMMAs — before the H' product that starts from them; the eight instances' fragments would
// take 96 registers, their tiles 12 KB of shared memory per warp).  Only the entries the spec reads are touched: Q[i][j]
// for i <= j, R[a][b] for b <= a, Qf[i][j] for i <= j, exactly like k_build_w for the shared weights.
template <int WARPS, int MIN_CTAS, bool DIRECT, bool PERW = false>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) k_ric12_mma(const Ric12Args a) {
  extern __shared__ __align__(128) double smem[];
  constexpr bool PAIRED = (WARPS == 8) && !DIRECT;
  constexpr int STRIDE = DIRECT ? RMD_STRIDE : RM_STRIDE;
  constexpr int PHOFF = DIRECT ? 0 : RM_PH;
  // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the TMA operands (shared-memory
  // destination, mbarrier, global source) in uniform registers — computed from threadIdx alone every bulk copy became an
  // ELECT / R2UR waterfall loop of 17 instructions (155 per warp-step, profiles/ncu_r02_k_ric12_mma_v1.txt)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;   // tensor phase: fragment coordinates; scalar phase: instance slot g, lane t of it
  double* wbase = smem + RM_BAR_DOUBLES + warp * (DIRECT ? RMD_WARP_DOUBLES : RM_WARP_DOUBLES);
  double* Gb = wbase + 8 * STRIDE;
  double* zpad = Gb + 2 * RM_GBUF + (lane & 3);
  double* base = wbase + g * STRIDE;    // own instance in the scalar phase
  double* Ps = base + PHOFF;
  const unsigned bar = smem_u32(smem) + 8u * warp;

  const long long w0 = ((long long)blockIdx.x * WARPS + warp) * 8;
  const long long w_raw = w0 + g;
  const bool valid = w_raw < a.batch;
  const long long b = valid ? w_raw : a.batch - 1;
  const int N = a.N;
  const int j0 = t, j1 = 7 - t, j2 = 11 - t;   // own state columns (scalar phase), as in k_ric12

  // the eight records of a step sit at rec0 + q * rstride (clamped at the end of the batch): lane 0 issues the copies
  const size_t rstride = (size_t)N * R12_REC;
  const long long wlast = (w0 < a.batch) ? w0 : a.batch - 1;
  const double* rec0 = a.AB + (size_t)wlast * rstride;
  const int nvalid = (int)((a.batch - wlast < 8) ? (a.batch - wlast) : 8);
  const unsigned fs0_u32 = smem_u32(wbase + RM_FS);
  unsigned phase = 0;
  auto fetch = [&](int k) {
    if (rm_elect_one()) {
      const double* src = rec0 + (size_t)k * R12_REC;
      if (!DIRECT) mbar_expect_tx(bar, 8u * 1536u);
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const double* sq = src + (size_t)(qq < nvalid ? qq : nvalid - 1) * rstride;
        if (DIRECT) rm_prefetch_l2(sq, 1536u);
        else bulk_g2s(fs0_u32 + (unsigned)(qq * STRIDE * 8), sq, 1536u, bar);
      }
    }
  };
  if (!DIRECT) {
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  fetch(N - 1);
  if (DIRECT && N > 1) fetch(N - 2);

  // ---- lane constants of the tensor phase
  // B fragments of F' (and A fragments of F'^T): F'[4ki+t][8ni+g] = F[4ki+t][perm(8ni+g)], element offset fo[ni] + ki*fs[ni]
  //   ni = 0: positions g < 4 are u-columns 3-g (B part of the record: 144 + l*4 + c), g >= 4 are x-columns g-4 (A part: l*12 + c)
  //   ni = 1: x-columns 4+g
  const int fo0 = (g < 4) ? 144 + t * 4 + (3 - g) : t * 12 + (g - 4);
  const int fs0 = (g < 4) ? 16 : 48;
  const int fo1 = t * 12 + 4 + g;
  // A fragments of P: P[8mi+g][4ki+t]; rows 12..15 are padding (zero).  Inside the sweep only the UPPER triangle of P is
  // kept in shared memory (the scalar phase stores P'[i][j], i <= j, once; the mirrored copy cost as many wavefronts
  // again, with bank conflicts): an element below the diagonal is read from its mirror image, which is the same double.
  auto ptri = [](int r, int c) { return r <= c ? r * 12 + c : c * 12 + r; };
  int po[2][3];
#pragma unroll
  for (int ki = 0; ki < 3; ++ki) { po[0][ki] = ptri(g, 4 * ki + t); po[1][ki] = ptri(8 + (g & 3), 4 * ki + t); }
  const bool p1on = g < 4;
  // G' buffer, element (l, c) at l*16 + (c ^ swz(l)), swz(l) = 8*(l&1) + 4*((l>>1)&1): C-fragment stores (rows g, 8+g;
  // 16-byte pairs) and B-fragment loads (rows 4ki+t) are both bank-conflict free
  const int swz_g = 8 * (g & 1) + 4 * ((g >> 1) & 1), swz_t = 8 * (t & 1) + 4 * ((t >> 1) & 1);
  const int gs0 = g * 16 + ((2 * t) ^ swz_g), gs1 = g * 16 + ((8 + 2 * t) ^ swz_g);          // + 128 for rows 8+g
  const int gl0 = t * 16 + (g ^ swz_t), gl1 = t * 16 + ((8 + g) ^ swz_t);                      // + 64*ki
  const int hs = g * 8 + 2 * t;                                                                // + 64*tile
  // C fragments of W' = W permuted (constant over the sweep): tiles (0,0), (0,1), (1,1)
  double wc[3][2];
  if (!PERW) {
    const int tr[3] = {0, 0, 8}, tc[3] = {0, 8, 8};
#pragma unroll
    for (int T = 0; T < 3; ++T)
#pragma unroll
      for (int e = 0; e < 2; ++e) wc[T][e] = a.W[rm_perm(tr[T] + g) * 16 + rm_perm(tc[T] + 2 * t + e)];
  }
  // PERW: the same fragments out of an instance's Q | R block, as 16-byte pairs.  Tile (0,0), element (g, 2t+e): both
  // positions u (g < 4, t < 2): R[3-g][3-2t-e], a descending pair; both x: Q[g-4][2t-4+e]; mixed: zero.  Tile (0,1),
  // element (g, 8+2t+e): Q[g-4][4+2t+e] for g >= 4, else zero.  Tile (1,1): Q[4+g][4+2t+e].
  const bool w_uu = g < 4 && t < 2, w_on0 = (g < 4) == (t < 2), w_on1 = g >= 4;
  const int w_off0 = w_uu ? 144 + (3 - g) * 4 + (2 - 2 * t) : (g - 4) * 12 + 2 * t - 4;
  const int w_off1 = (g - 4) * 12 + 4 + 2 * t, w_off2 = (4 + g) * 12 + 4 + 2 * t;
  double2 wq[2][3];        // ring over the instances of the tensor pipeline (raw pairs; the u-u pair is swapped on use)
  auto load_w = [&](int q) {
    const double* Wq = a.W + (size_t)(wlast + (q < nvalid ? q : nvalid - 1)) * (size_t)a.w_stride;
    wq[q & 1][0] = wq[q & 1][1] = make_double2(0.0, 0.0);
    if (w_on0) wq[q & 1][0] = __ldg(reinterpret_cast<const double2*>(Wq + w_off0));
    if (w_on1) wq[q & 1][1] = __ldg(reinterpret_cast<const double2*>(Wq + w_off1));
    wq[q & 1][2] = __ldg(reinterpret_cast<const double2*>(Wq + w_off2));
  };
  if (PERW) { load_w(0); load_w(1); }

  if (lane < RM_ZPAD) Gb[2 * RM_GBUF + lane] = 0.0;
  // P <- Qf
  const double* Qf_own = PERW ? a.W + (size_t)b * (size_t)a.w_stride + 160 : a.Qf;
  for (int e = t; e < 144; e += 4) Ps[e] = Qf_own[e];
  for (int e = 144 + t; e < 192; e += 4) Ps[e] = 0.0;

  bool failed = false;
  int kfail = -1;
  double* Kp = (a.K && valid) ? a.K + ((size_t)b * N + (N - 1)) * 48 : nullptr;
  double *Kp0 = Kp + j0, *Kp1 = Kp + j1, *Kp2 = Kp + j2;
  __syncwarp();

  // Phase interlock (one CTA of eight warps per SM): warps w and w+4 share a scheduler, and left alone they drift into
  // the same phase — both in the tensor phase (each DMMA stream at half speed), then both in the latency-bound scalar
  // phase with the FP64 pipe idle; that state is stable (measured: 8,900 cycles per pair of warp-steps, 70 % pipe).  A
  // named barrier per pair at every phase boundary, the upper warp starting half a period late, keeps one warp's
  // tensor phase under the other's scalar phase.
  const unsigned pair_bar = 1u + (unsigned)(warp & 3);
  auto pair_sync = [&]() {
    if (PAIRED) asm volatile("bar.sync %0, 64;\n" ::"r"(pair_bar) : "memory");
  };
  if (PAIRED && warp >= 4) pair_sync();

#ifdef RM_TIMING
  long long tm_wait = 0, tm_tensor = 0, tm_scalar = 0, tm_sync = 0;
#define RM_TICK(acc) { const long long now_ = clock64(); acc += now_ - tm_last; tm_last = now_; }
  long long tm_last = clock64();
#else
#define RM_TICK(acc)
#endif
  for (int it = 0; it < N; ++it) {
    const int k = N - 1 - it;
    if (!DIRECT) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    __syncwarp();
    RM_TICK(tm_wait)

    // ---------------- tensor phase: H' of the eight instances, software-pipelined over the instances.  A warp issues in
    // order, and a DMMA holds the FP64 pipe for 16 cycles (26 to its result: tools/ubench/dmma_issue.cu), so everything
    // that is not a DMMA is placed BETWEEN DMMAs whose operands are long ready — slot q of the loop issues
    //     G'(q)   [12 DMMAs]   with, in their shadow: the store of H'(q-2), the fragment loads of instance q+1 and the TMA
    //                          request for instance q's next record (its fragments are in registers by then), and
    //     H'(q-1) [ 9 DMMAs]   with, in their shadow: the store of G'(q), the warp barrier and the reload of G'(q) in the
    //                          B-fragment layout (consumed one slot later).
    // The first version issued block after block (loads, 12 DMMAs, stores, barrier, loads, 9 DMMAs, stores): 4,700 cycles
    // per warp-step for 2,688 cycles of DMMA issue.
    constexpr int FD = DIRECT ? 2 : 1;   // F' fragments are requested FD instances ahead (L2 latency vs shared memory)
    double fb[4][3][2];      // F' fragments, ring over instances q-1 .. q+FD
    double pa[2][3][2];      // P fragments, instances q, q+1
    double gb[2][3][2];      // G' fragments (B layout), instances q-1, q
    double hacc[3][2];
    const double* rec_k = rec0 + (size_t)k * R12_REC;
    auto load_f = [&](int q) {
      if (DIRECT) {
        const double* Fq = rec_k + (size_t)(q < nvalid ? q : nvalid - 1) * rstride;
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          fb[q % 4][ki][0] = rm_ldg_stream(Fq + fo0 + ki * fs0);
          fb[q % 4][ki][1] = rm_ldg_stream(Fq + fo1 + ki * 48);
        }
      } else {
        const double* Fq = wbase + q * STRIDE + RM_FS;
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          fb[q % 4][ki][0] = Fq[fo0 + ki * fs0];
          fb[q % 4][ki][1] = Fq[fo1 + ki * 48];
        }
      }
    };
    auto load_p = [&](int q) {
      const double* PHq = wbase + q * STRIDE + PHOFF;
#pragma unroll
      for (int ki = 0; ki < 3; ++ki) {
        pa[q & 1][ki][0] = PHq[po[0][ki]];
        double p1 = 0.0;                                          // padding rows of the second M tile are zeros
        if (p1on) p1 = PHq[po[1][ki]];                            // (a predicated load: half the wavefronts)
        pa[q & 1][ki][1] = p1;
      }
    };
    const bool more = it + 1 < N;
    if (!DIRECT && more && rm_elect_one()) mbar_expect_tx(bar, 8u * 1536u);
    const double* src_next = rec0 + (size_t)(k - 1) * R12_REC;
    if (DIRECT && it + 2 < N) fetch(k - 2);   // L2 prefetch two steps ahead (the first two were requested in the prologue)
#pragma unroll
    for (int q0 = 0; q0 < FD; ++q0) load_f(q0);
    load_p(0);
#pragma unroll
    for (int q = 0; q <= 8; ++q) {
      double gacc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};   // [mi][ni][e]
      if (q < 8) {
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          rm_dmma(gacc[0][0][0], gacc[0][0][1], pa[q & 1][ki][0], fb[q % 4][ki][0]);
          rm_dmma(gacc[0][1][0], gacc[0][1][1], pa[q & 1][ki][0], fb[q % 4][ki][1]);
          rm_dmma(gacc[1][0][0], gacc[1][0][1], pa[q & 1][ki][1], fb[q % 4][ki][0]);
          rm_dmma(gacc[1][1][0], gacc[1][1][1], pa[q & 1][ki][1], fb[q % 4][ki][1]);
          if (ki == 0 && q >= 2) {   // H'(q-2), issued a slot ago, has long landed
            double* PHp = wbase + (q - 2) * STRIDE + PHOFF;
#pragma unroll
            for (int T = 0; T < 3; ++T) sts128(PHp + 64 * T + hs, hacc[T][0], hacc[T][1]);
          }
          if (ki == 1 && q + 1 < 8) load_p(q + 1);
          if (ki == 1 && q + FD < 8) load_f(q + FD);
        }
        // every lane's fragments of instance q are in registers (the DMMAs above consumed them): its record slot is free
        if (!DIRECT && more && rm_elect_one())
          bulk_g2s(fs0_u32 + (unsigned)(q * STRIDE * 8), src_next + (size_t)(q < nvalid ? q : nvalid - 1) * rstride, 1536u, bar);
      }
      if (q >= 1) {   // H'(q-1) = W' + F'^T G'
        if (q == 8) {   // (no G' slot to hide the store of H'(6) in)
          double* PHp = wbase + 6 * STRIDE + PHOFF;
#pragma unroll
          for (int T = 0; T < 3; ++T) sts128(PHp + 64 * T + hs, hacc[T][0], hacc[T][1]);
        }
        if (PERW) {
          const double2 v0 = wq[(q - 1) & 1][0], v1 = wq[(q - 1) & 1][1], v2 = wq[(q - 1) & 1][2];
          hacc[0][0] = w_uu ? v0.y : v0.x; hacc[0][1] = w_uu ? v0.x : v0.y;
          hacc[1][0] = v1.x; hacc[1][1] = v1.y;
          hacc[2][0] = v2.x; hacc[2][1] = v2.y;
          load_w((q + 1) % 8);   // into the ring slot just read; slots 7 and 8 request instances 0 and 1 of the next step
        } else {
#pragma unroll
          for (int T = 0; T < 3; ++T) { hacc[T][0] = wc[T][0]; hacc[T][1] = wc[T][1]; }
        }
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          rm_dmma(hacc[0][0], hacc[0][1], fb[(q - 1) % 4][ki][0], gb[(q - 1) & 1][ki][0]);   // tile (0,0)
          rm_dmma(hacc[1][0], hacc[1][1], fb[(q - 1) % 4][ki][0], gb[(q - 1) & 1][ki][1]);   // tile (0,1)
          rm_dmma(hacc[2][0], hacc[2][1], fb[(q - 1) % 4][ki][1], gb[(q - 1) & 1][ki][1]);   // tile (1,1)
          if (ki == 0 && q < 8) {   // G'(q) is complete by now: C-fragment layout -> shared memory -> B-fragment layout
            double* Gq = Gb + (q & 1) * RM_GBUF;
            sts128(Gq + gs0, gacc[0][0][0], gacc[0][0][1]);
            sts128(Gq + gs1, gacc[0][1][0], gacc[0][1][1]);
            if (p1on) {
              sts128(Gq + 128 + gs0, gacc[1][0][0], gacc[1][0][1]);
              sts128(Gq + 128 + gs1, gacc[1][1][0], gacc[1][1][1]);
            }
            __syncwarp();
          }
          if (ki == 1 && q < 8) {
            const double* Gq = Gb + (q & 1) * RM_GBUF;
#pragma unroll
            for (int kj = 0; kj < 3; ++kj) { gb[q & 1][kj][0] = Gq[gl0 + 64 * kj]; gb[q & 1][kj][1] = Gq[gl1 + 64 * kj]; }
          }
        }
      } else {   // q == 0: nothing to overlap the round trip of G'(0) with
        double* Gq = Gb;
        sts128(Gq + gs0, gacc[0][0][0], gacc[0][0][1]);
        sts128(Gq + gs1, gacc[0][1][0], gacc[0][1][1]);
        if (p1on) {
          sts128(Gq + 128 + gs0, gacc[1][0][0], gacc[1][0][1]);
          sts128(Gq + 128 + gs1, gacc[1][1][0], gacc[1][1][1]);
        }
        __syncwarp();
#pragma unroll
        for (int kj = 0; kj < 3; ++kj) { gb[0][kj][0] = Gq[gl0 + 64 * kj]; gb[0][kj][1] = Gq[gl1 + 64 * kj]; }
      }
    }
    {
      double* PH7 = wbase + 7 * STRIDE + PHOFF;
#pragma unroll
      for (int T = 0; T < 3; ++T) sts128(PH7 + 64 * T + hs, hacc[T][0], hacc[T][1]);
    }
    __syncwarp();   // H' of all eight instances visible
    RM_TICK(tm_tensor)
    pair_sync();
    RM_TICK(tm_sync)

    // ---------------- scalar phase (four lanes per instance): L D L^T, gains, P'
    const double* Hq = Ps;
    double hu[3][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      hu[0][c] = Hq[rm_hoff(3 - c, 4) + j0];          // Hux[c][j] = H'(3-c, 4+j): row 3-c of tile 0 (j < 4) ...
      hu[1][c] = Hq[rm_hoff(3 - c, 4 + j1)];          // ... or tile 1 (j >= 4)
      hu[2][c] = Hq[rm_hoff(3 - c, 4 + j2)];
    }
    Ldl4 ch;
    {
      // Huu[a][b], b <= a, = H'(3-a, 3-b)
      const double u00 = Hq[rm_hoff(3, 3)], u10 = Hq[rm_hoff(2, 3)], u11 = Hq[rm_hoff(2, 2)], u20 = Hq[rm_hoff(1, 3)],
                   u21 = Hq[rm_hoff(1, 2)], u22 = Hq[rm_hoff(1, 1)], u30 = Hq[rm_hoff(0, 3)], u31 = Hq[rm_hoff(0, 2)],
                   u32 = Hq[rm_hoff(0, 1)], u33 = Hq[rm_hoff(0, 0)];
      failed |= ch.factor(u00, u10, u11, u20, u21, u22, u30, u31, u32, u33);
    }
    // Hxx rows of the own columns: Hxx[i][j] = H'(4+i, 4+j), i <= j (rows above the diagonal of a column are dead slots of
    // the uniform code: their addresses stay inside the tile store, their values are never stored)
    double hx0[4], hx1[8], hx2[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) hx0[i] = Hq[rm_hoff(4 + i, 4 + j0)];
#pragma unroll
    for (int i = 0; i < 8; ++i) hx1[i] = Hq[rm_hoff(4 + i, 4 + j1)];
#pragma unroll
    for (int i = 0; i < 12; ++i) hx2[i] = Hq[rm_hoff(4 + i, 4 + j2)];
    double kk[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s) ch.solve_neg(hu[s], kk[s]);

    if (failed && kfail < 0) kfail = k;
    if (Kp) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __stcs(Kp0 + c * 12, kk[0][c]);
        __stcs(Kp1 + c * 12, kk[1][c]);
        __stcs(Kp2 + c * 12, kk[2][c]);
      }
      Kp0 -= 48; Kp1 -= 48; Kp2 -= 48;
    }
    if (k == 0 && a.K0 && valid) {
      double* kp = a.K0 + (size_t)b * 48;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        kp[c * 12 + j0] = kk[0][c];
        kp[c * 12 + j1] = kk[1][c];
        kp[c * 12 + j2] = kk[2][c];
      }
    }
    // P'[i][j] = Hxx[i][j] + sum_a Hux[a][i] K[a][j], i <= j, own j: Hux row c = H'(3-c, 4 .. 15)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double hr[12];
      {
        const double* r0 = Hq + (3 - c) * 8 + 4;        // tile 0, columns 4..7  -> states 0..3
        const double* r1 = Hq + 64 + (3 - c) * 8;       // tile 1, columns 8..15 -> states 4..11
        const double2 v0 = lds128(r0), v1 = lds128(r0 + 2);
        hr[0] = v0.x; hr[1] = v0.y; hr[2] = v1.x; hr[3] = v1.y;
#pragma unroll
        for (int i = 0; i < 8; i += 2) { const double2 v = lds128(r1 + i); hr[4 + i] = v.x; hr[5 + i] = v.y; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) hx0[i] = fmad(hr[i], kk[0][c], hx0[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) hx1[i] = fmad(hr[i], kk[1][c], hx1[i]);
#pragma unroll
      for (int i = 0; i < 12; ++i) hx2[i] = fmad(hr[i], kk[2][c], hx2[i]);
    }
    __syncwarp();   // every lane has taken what it needs from H': P' may overwrite the slot
    rm_store_upper<4>(Ps, hx0, j0);
    rm_store_upper<8>(Ps, hx1, j1);
    rm_store_upper<12>(Ps, hx2, j2);
    RM_TICK(tm_scalar)
    pair_sync();
    RM_TICK(tm_sync)
  }
#ifdef RM_TIMING
  if (lane == 0 && a.iters) {   // (a.iters is unused by the LQR sweep: the timing build borrows it)
    long long* o = reinterpret_cast<long long*>(a.iters) + 4 * ((long long)blockIdx.x * WARPS + warp);
    o[0] = tm_wait; o[1] = tm_tensor; o[2] = tm_scalar; o[3] = tm_sync;
  }
#endif
  if (PAIRED && warp < 4) pair_sync();   // the partner's last scalar phase
  __syncwarp();
  for (int e = t; e < 144; e += 4) {      // P0 as a full symmetric matrix for the outputs
    const int i = e / 12, j = e % 12;
    if (i > j) Ps[e] = Ps[j * 12 + i];
  }
  __syncwarp();

  // ---------------- outputs (the rules of k_ric12<RIC_LQR>)
  const double nanv = qnan();
  if (failed && valid) {
    if (a.K)
      for (int k2 = 0; k2 <= kfail; ++k2) {
        double* kp = a.K + ((size_t)b * N + k2) * 48;
#pragma unroll
        for (int c = 0; c < 4; ++c) { kp[c * 12 + j0] = nanv; kp[c * 12 + j1] = nanv; kp[c * 12 + j2] = nanv; }
      }
    if (a.K0) {
      double* kp = a.K0 + (size_t)b * 48;
#pragma unroll
      for (int c = 0; c < 4; ++c) { kp[c * 12 + j0] = nanv; kp[c * 12 + j1] = nanv; kp[c * 12 + j2] = nanv; }
    }
  }
  if (valid && a.P0) {
    double* po = a.P0 + (size_t)b * 144;
    for (int e = 2 * t; e < 144; e += 8) {
      const double2 v = lds128(Ps + e);
      *reinterpret_cast<double2*>(po + e) = failed ? make_double2(nanv, nanv) : v;
    }
  }
  if (a.cost && a.x0) {
    // cost = 0.5 x0' P0 x0: row chains spread over the four lanes, the chain over rows on lane 0 (spec order)
    const double* x = a.x0 + (size_t)b * 12;
    double xv[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) xv[j] = x[j];
    double* rs = Gb + g * 12;    // the G' buffers are dead
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      const int i = t + 4 * rr;
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < 12; ++j) r = fmad(Ps[i * 12 + j], xv[j], r);
      rs[i] = r;
    }
    __syncwarp();
    if (t == 0 && valid) {
      double c = 0.0;
#pragma unroll
      for (int i = 0; i < 12; ++i) c = fmad(xv[i], rs[i], c);
      a.cost[b] = failed ? nanv : 0.5 * c;
    }
  }
  if (valid && t == 0 && a.status) a.status[b] = failed ? 1 : 0;
}

}  // namespace noc

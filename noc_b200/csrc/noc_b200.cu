// noc_b200.cu — C-ABI of libnoc_b200.so (include/lqr_control/noc_b200.h).
// Host-side batch scheduler: pointer classification, staging, chunking to the free HBM, kernel launches.
// There is no CPU fallback anywhere in this file: every entry point either launches CUDA kernels or fails.
#include "../../include/lqr_control/noc_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "pipeline_kernels.cuh"
#include "riccati12.cuh"
#include "riccati24.cuh"
#include "riccati_generic.cuh"
#include "slq_kernels.cuh"

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct PendingCopy {
  void* host;
  const void* dev;
  size_t bytes;
};

}  // namespace

struct noc_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // device->host copies of finished chunks, overlapped with the next chunk
  cudaStream_t aux_stream = nullptr;    // every other inner chunk of a from-x0 call: neighbouring sweeps overlap at their edges
  cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  std::string err;
  int64_t launches = 0;
  int sm_count = 0;
  std::vector<DevBuf> in_slots, out_slots, scratch;
  std::vector<PendingCopy> pending;
  // per-kernel device timing (noc_profile_*): event pairs recorded around launches of each kernel id
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<std::pair<int, size_t>> ev_used;  // (kernel id, pool index)
  size_t ev_next = 0;
  double prof_ms[NOC_KERNEL_COUNT] = {0};
  int64_t prof_launches[NOC_KERNEL_COUNT] = {0};
};

namespace {

using namespace noc;

#define NOC_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
      return (e_ == cudaErrorMemoryAllocation) ? NOC_ERR_OOM : NOC_ERR_CUDA;                     \
    }                                                                                            \
  } while (0)

// RAII bracket around one kernel launch: counts it and, when profiling, records an event pair on the
// launching stream (noc_profile_read turns the pairs into per-kernel device time).
struct LaunchScope {
  noc_ctx* c;
  cudaEvent_t stop = nullptr;
  LaunchScope(noc_ctx* ctx, int kid) : c(ctx) {
    c->launches++;
    c->prof_launches[kid]++;
    if (!c->profiling) return;
    if (c->ev_next == c->ev_pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      c->ev_pool.push_back({a, b});
    }
    auto& pr = c->ev_pool[c->ev_next];
    c->ev_used.push_back({kid, c->ev_next});
    c->ev_next++;
    cudaEventRecord(pr.first, c->stream);
    stop = pr.second;
  }
  ~LaunchScope() {
    if (stop) cudaEventRecord(stop, c->stream);
  }
};

int fail(noc_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int ensure(noc_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return NOC_OK;
  if (b.p) { NOC_CUDA(ctx, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  const size_t want = (bytes + 255) & ~size_t(255);
  NOC_CUDA(ctx, cudaMalloc(&b.p, want));
  b.cap = want;
  return NOC_OK;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// input: returns a device pointer holding `bytes` of `p` (p itself if it already is a device pointer)
int stage_in(noc_ctx* ctx, int slot, const void* p, size_t bytes, const void** out) {
  if (!p) { *out = nullptr; return NOC_OK; }
  if (is_device_ptr(p)) { *out = p; return NOC_OK; }
  if ((int)ctx->in_slots.size() <= slot) ctx->in_slots.resize(slot + 1);
  int rc = ensure(ctx, ctx->in_slots[slot], bytes);
  if (rc) return rc;
  NOC_CUDA(ctx, cudaMemcpyAsync(ctx->in_slots[slot].p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *out = ctx->in_slots[slot].p;
  return NOC_OK;
}

// output: returns a device pointer to write to; host destinations are copied back by flush_out()
int stage_out(noc_ctx* ctx, int slot, void* p, size_t bytes, void** out) {
  if (!p) { *out = nullptr; return NOC_OK; }
  if (is_device_ptr(p)) { *out = p; return NOC_OK; }
  if ((int)ctx->out_slots.size() <= slot) ctx->out_slots.resize(slot + 1);
  int rc = ensure(ctx, ctx->out_slots[slot], bytes);
  if (rc) return rc;
  *out = ctx->out_slots[slot].p;
  ctx->pending.push_back({p, ctx->out_slots[slot].p, bytes});
  return NOC_OK;
}

int flush_out(noc_ctx* ctx, bool async) {
  for (auto& c : ctx->pending)
    NOC_CUDA(ctx, cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  const bool had_host = !ctx->pending.empty();
  ctx->pending.clear();
  if (!async || had_host) NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NOC_OK;
}

int scratch(noc_ctx* ctx, int slot, size_t bytes, void** out) {
  if ((int)ctx->scratch.size() <= slot) ctx->scratch.resize(slot + 1);
  int rc = ensure(ctx, ctx->scratch[slot], bytes);
  if (rc) return rc;
  *out = ctx->scratch[slot].p;
  return NOC_OK;
}

enum { SCR_W = 0, SCR_XBAR, SCR_AB, SCR_ARGMIN, SCR_K, SCR_KFF, SCR_DV, SCR_MU, SCR_J, SCR_BSTATUS, SCR_STATUS,
       SCR_ITERS, SCR_IDX0, SCR_IDX1, SCR_COUNT, SCR_X, SCR_U, SCR_XN, SCR_UN, SCR_RETRY0, SCR_RETRY1, SCR_LC, SCR_FLAGS, SCR_XC, SCR_UC, SCR_LCC };

// W = symmetric blockdiag(Q, R) as a dense (n+m)x(n+m) matrix, followed by the symmetric terminal weight (n*n).
// Q / Qf are taken from their upper triangles and R from its lower triangle (the entries the spec reads,
// DESIGN.md §2.3); qf_off selects Qf (n*n+m*m) or, for the DARE iteration, Q itself (0).
__global__ void k_build_w(const double* __restrict__ qrqf, double* __restrict__ w, int n, int m, int qf_off) {
  const int nm = n + m;
  for (int e = threadIdx.x; e < nm * nm; e += blockDim.x) {
    const int i = e / nm, j = e % nm;
    double v = 0.0;
    if (i < n && j < n) v = qrqf[min(i, j) * n + max(i, j)];
    else if (i >= n && j >= n) v = qrqf[n * n + (max(i, j) - n) * m + (min(i, j) - n)];
    w[e] = v;
  }
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    w[nm * nm + e] = qrqf[qf_off + min(i, j) * n + max(i, j)];
  }
}

// diagnostics: the branch-free reciprocal of the factorisation next to the IEEE division
__global__ void k_selftest_rcp(const double* __restrict__ x, double* __restrict__ fast, double* __restrict__ ieee,
                               long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fast[i] = rcp_rn_normal(x[i]);
  ieee[i] = 1.0 / x[i];
}

int check_prob(noc_ctx* ctx, const noc_problem_t* p, int64_t batch) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!p) return fail(ctx, NOC_ERR_BAD_ARG, "problem descriptor is NULL");
  if (batch < 0) return fail(ctx, NOC_ERR_BAD_ARG, "negative batch");
  if (p->N < 1) return fail(ctx, NOC_ERR_BAD_ARG, "horizon N must be >= 1");
  if (p->layout != NOC_LAYOUT_A_THEN_B && p->layout != NOC_LAYOUT_ROWS_AB)
    return fail(ctx, NOC_ERR_BAD_ARG, "unknown AB layout");
  if (p->model == NOC_MODEL_QUAD12 && (p->n != 12 || p->m != 4))
    return fail(ctx, NOC_ERR_BAD_ARG, "NOC_MODEL_QUAD12 requires n=12, m=4");
  if (p->model == NOC_MODEL_DUAL24 && (p->n != 24 || p->m != 8))
    return fail(ctx, NOC_ERR_BAD_ARG, "NOC_MODEL_DUAL24 requires n=24, m=8");
  if (p->model < 0 || p->model > NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "unknown model");
  return NOC_OK;
}

bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

// kernel configurations: 8 instances per warp; LQR/DARE 2 CTAs x 4 warps per SM,
// SLQ backward pass 1 CTA x 7 warps (its per-instance shared-memory slice is larger)
template <int MODE> struct RicCfg { static constexpr int warps = 4, min_ctas = 2; };
template <> struct RicCfg<RIC_AFFINE> { static constexpr int warps = 7, min_ctas = 1; };

template <int LAYOUT, int MODE, bool FUSED, int W, int MIN_CTAS, bool BOX = false>
int launch_ric12_cfg(noc_ctx* ctx, const Ric12Args& a) {
  auto kern = k_ric12<LAYOUT, MODE, FUSED, W, MIN_CTAS, BOX>;
  const size_t smem = ric12_smem_bytes<MODE, FUSED>(W);
  // function attributes are per device and contexts may live on several devices / host threads: set them on every
  // launch (two cheap driver calls) instead of caching a process-wide flag
  NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long per_cta = 8LL * W;
  const unsigned grid = (unsigned)((a.batch + per_cta - 1) / per_cta);
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, W * 32, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

template <int LAYOUT, int MODE, bool FUSED = false, bool BOX = false>
int launch_ric12(noc_ctx* ctx, const Ric12Args& a) {
  // short work lists (the tail of an SLQ solve, small batches): one warp per CTA, so that the few warps spread over
  // all SMs instead of sharing a handful — the sweep is then latency-bound and a warp alone on an SM steps ~2x faster
#if defined(NOC_EXP_MASK)   // experiment build only: one 7/8-warp CTA per SM, warps selected by the mask
  return launch_ric12_cfg<LAYOUT, MODE, FUSED, (MODE == RIC_AFFINE ? 7 : 8), 1>(ctx, a);
#else
  if (a.batch <= 8LL * 2 * ctx->sm_count) return launch_ric12_cfg<LAYOUT, MODE, FUSED, 1, 1, BOX>(ctx, a);
  return launch_ric12_cfg<LAYOUT, MODE, FUSED, RicCfg<MODE>::warps, RicCfg<MODE>::min_ctas, BOX>(ctx, a);
#endif
}

template <int MODE>
int launch_ric12_layout(noc_ctx* ctx, int layout, const Ric12Args& a) {
  return layout == NOC_LAYOUT_A_THEN_B ? launch_ric12<0, MODE>(ctx, a) : launch_ric12<1, MODE>(ctx, a);
}

template <int MODE> struct Ric24Cfg { static constexpr int warps = 8; };
template <> struct Ric24Cfg<RIC_AFFINE> { static constexpr int warps = 7; };

template <int LAYOUT, int MODE, bool FUSED, int W, bool BOX = false>
int launch_ric24_cfg(noc_ctx* ctx, const Ric24Args& a) {
  auto kern = k_ric24<LAYOUT, MODE, FUSED, W, BOX>;
  const size_t smem = ric24_smem_bytes<MODE, FUSED>(W);
  NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long per_cta = 2LL * W;
  const unsigned grid = (unsigned)((a.batch + per_cta - 1) / per_cta);
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, W * 32, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

template <int LAYOUT, int MODE, bool FUSED = false, bool BOX = false>
int launch_ric24(noc_ctx* ctx, const Ric24Args& a) {
  // short work lists: one warp (two instances) per CTA so that the latency-bound tail spreads over all SMs
  if (a.batch <= 2LL * 2 * ctx->sm_count) return launch_ric24_cfg<LAYOUT, MODE, FUSED, 1, BOX>(ctx, a);
  return launch_ric24_cfg<LAYOUT, MODE, FUSED, Ric24Cfg<MODE>::warps, BOX>(ctx, a);
}

// packs the weights for the n=24 kernels into scratch: W[1024] | Qf[576]
int build_w24(noc_ctx* ctx, const double* dQRQf, const double** W, const double** Qf) {
  void* w = nullptr;
  int rc = scratch(ctx, SCR_W, sizeof(double) * (32 * 32 + 576), &w);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_SETUP);
    k_build_w<<<1, 256, 0, ctx->stream>>>(dQRQf, (double*)w, 24, 8, 24 * 24 + 8 * 8);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  *W = (const double*)w;
  *Qf = (const double*)w + 1024;
  return NOC_OK;
}

int lqr24_device(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* dAB, const double* dQRQf,
                 const double* dx0, double* dK, double* dK0, double* dP0, double* dcost, int32_t* dstatus) {
  Ric24Args a{};
  int rc = build_w24(ctx, dQRQf, &a.W, &a.Qf);
  if (rc) return rc;
  a.AB = dAB; a.x0 = dx0;
  a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
  a.batch = batch; a.N = p->N;
  return p->layout == NOC_LAYOUT_A_THEN_B ? launch_ric24<0, RIC_LQR>(ctx, a) : launch_ric24<1, RIC_LQR>(ctx, a);
}

// packs the weights for the n=12 kernels into scratch: W[256] | Qf[144]
int build_w12(noc_ctx* ctx, const double* dQRQf, bool dare, const double** W, const double** Qf) {
  void* w = nullptr;
  int rc = scratch(ctx, SCR_W, sizeof(double) * (16 * 16 + 144), &w);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_SETUP);
    k_build_w<<<1, 256, 0, ctx->stream>>>(dQRQf, (double*)w, 12, 4, dare ? 0 : 12 * 12 + 4 * 4);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  *W = (const double*)w;
  *Qf = (const double*)w + 256;
  return NOC_OK;
}

// device-pointer core of the n=12 sweep
int lqr12_device(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* dAB, const double* dQRQf,
                 const double* dx0, double* dK, double* dK0, double* dP0, double* dcost, int32_t* dstatus) {
  Ric12Args a{};
  int rc = build_w12(ctx, dQRQf, false, &a.W, &a.Qf);
  if (rc) return rc;
  a.AB = dAB; a.x0 = dx0;
  a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
  a.batch = batch; a.N = p->N;
  return launch_ric12_layout<RIC_LQR>(ctx, p->layout, a);
}

template <int MODEL>
int from_x0_impl(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                 const double* QRQf, double* xbar_out, double* K, double* K0, double* P0, double* cost,
                 int32_t* status) {
  int rc;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m, N = p->N, rec = n * n + n * m;
  const void *dx0, *dQ, *dub;
  void *dK, *dK0, *dP0, *dcost, *dstatus, *dxout;
  if ((rc = stage_in(ctx, 0, x0, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_in(ctx, 1, QRQf, sizeof(double) * (2 * n * n + m * m), &dQ))) return rc;
  if ((rc = stage_in(ctx, 2, ubar, sizeof(double) * B * N * m, &dub))) return rc;   // nominal inputs (NULL = hover thrust)
  if (dub && misaligned16(dub)) return fail(ctx, NOC_ERR_BAD_ARG, "device ubar must be 16-byte aligned");
  if ((rc = stage_out(ctx, 5, xbar_out, sizeof(double) * B * (N + 1) * n, &dxout))) return rc;
  if (dxout && misaligned16(dxout)) return fail(ctx, NOC_ERR_BAD_ARG, "device xbar must be 16-byte aligned");
  if ((rc = stage_out(ctx, 0, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 1, K0, sizeof(double) * B * m * n, &dK0))) return rc;
  if ((rc = stage_out(ctx, 2, P0, sizeof(double) * B * n * n, &dP0))) return rc;
  if ((rc = stage_out(ctx, 3, cost, sizeof(double) * B, &dcost))) return rc;
  if ((rc = stage_out(ctx, 4, status, sizeof(int32_t) * B, &dstatus))) return rc;
  // Both built-in models linearise inside the sweep (k_ric12 / k_ric24 <…, FUSED>): only the nominal trajectory is
  // staged in scratch, no A_k,B_k stream exists.  Batch scheduler: an OUTER chunk bounds the scratch to a share of
  // the free HBM and gets one rollout launch (latency-bound: one launch for everything costs as much as one for a
  // quarter); with host-resident outputs it is cut into INNER chunks so that the device->host copy of chunk i runs
  // on the copy stream under the sweep of chunk i+1.
  size_t free_b = 0, total_b = 0;
  NOC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  size_t have = 0;
  if (ctx->scratch.size() > SCR_XBAR) have = ctx->scratch[SCR_XBAR].cap;
  size_t budget = std::max<size_t>((size_t)((free_b + have) * 0.6), (size_t)64 << 20);
  if (const char* cap = std::getenv("NOC_B200_SCRATCH_MB"))   // test hook: force several outer chunks
    budget = std::max<size_t>((size_t)std::atoll(cap) << 20, (size_t)1 << 20);
  const size_t per_inst = sizeof(double) * (N + 1) * n;
  size_t outer = std::max<size_t>(1, std::min<size_t>(B, budget / per_inst));
  if (outer < B) outer = std::max<size_t>(32, outer & ~size_t(31));
  if (dxout) outer = B;   // the caller's xbar array (instance-major) holds the nominal: no scratch, no outer chunks
  const bool overlap = !ctx->pending.empty() && B >= 8192 && !(p->flags & NOC_FLAG_ASYNC);
  // inner chunks are whole waves of the sweep kernel (instances resident on all SMs at once: 64 per SM for n=12,
  // 16 for n=24), about eight per call: a chunk that ends mid-wave pays for the full wave, and the smaller the last
  // chunk, the shorter the copy that nothing hides (B = 65536: 4 x 16384 = 8 waves took 2.53 ms of sweeps, 7 chunks
  // of one wave take 2.2 ms)
  const size_t wave = (size_t)ctx->sm_count * (MODEL == 0 ? 64 : 16);
  size_t inner_pref = overlap ? std::max<size_t>(1, (B / wave + 4) / 8) * wave : outer;
  if (const char* e = std::getenv("NOC_B200_CHUNK_WAVES"))   // experiment hook
    if (overlap) inner_pref = std::max<size_t>(1, (size_t)std::atoll(e)) * wave;
  void* dxbar = nullptr;
  if (!dxout && (rc = scratch(ctx, SCR_XBAR, sizeof(double) * outer * (N + 1) * n, &dxbar))) return rc;
  (void)rec;
  typename std::conditional<MODEL == 0, Ric12Args, Ric24Args>::type fa{};
  if (MODEL == 0) rc = build_w12(ctx, (const double*)dQ, false, &fa.W, &fa.Qf);
  else rc = build_w24(ctx, (const double*)dQ, &fa.W, &fa.Qf);
  if (rc) { ctx->pending.clear(); return rc; }
  int chunk_no = 0;
  for (size_t o = 0; o < B; o += outer) {
    const size_t ob = std::min(outer, B - o);
    {
      LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
      const double* ub = dub ? (const double*)dub + o * N * m : nullptr;
      if (dxout)
        k_rollout_open_loop<MODEL><<<(unsigned)((ob + 127) / 128), 128, 0, ctx->stream>>>(
            (const double*)dx0 + o * n, ub, (double*)dxout + o * (N + 1) * n, (long long)ob, p->N, p->dt,
            (long long)((N + 1) * n), (long long)n);
      else
        k_rollout_open_loop<MODEL><<<(unsigned)((ob + 127) / 128), 128, 0, ctx->stream>>>(
            (const double*)dx0 + o * n, ub, (double*)dxbar, (long long)ob, p->N, p->dt, (long long)n, (long long)(ob * n));
    }
    NOC_CUDA(ctx, cudaGetLastError());
    const size_t inner = std::min(ob, inner_pref);
    // inner chunks alternate between the caller's stream and an auxiliary one: each is one wave of CTAs, so the
    // next chunk's CTAs move onto the SMs as the previous chunk's retire instead of waiting for its last warp
    cudaStream_t const main_stream = ctx->stream;
    const bool two_streams = overlap && inner < ob && !std::getenv("NOC_B200_ONE_STREAM");
    if (two_streams) {
      NOC_CUDA(ctx, cudaEventRecord(ctx->fork_ev, main_stream));            // rollout (and the staged inputs) done
      NOC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_ev, 0));
    }
    for (size_t i = 0; i < ob; i += inner, ++chunk_no) {
      const size_t off = o + i;
      const long long cb = (long long)std::min(inner, ob - i);
      if (dxout) { fa.xbar = (const double*)dxout + off * (N + 1) * n; fa.xbar_sb = 0; fa.xbar_sk = 0; }   // dense, instance-major
      else { fa.xbar = (const double*)dxbar + i * n; fa.xbar_sb = (long long)n; fa.xbar_sk = (long long)(ob * n); }   // step-major scratch
      fa.ubar = dub ? (const double*)dub + off * N * m : nullptr; fa.dt = p->dt;
      fa.x0 = (const double*)dx0 + off * n;
      fa.K = dK ? (double*)dK + off * N * m * n : nullptr;
      fa.K0 = dK0 ? (double*)dK0 + off * m * n : nullptr;
      fa.P0 = dP0 ? (double*)dP0 + off * n * n : nullptr;
      fa.cost = dcost ? (double*)dcost + off : nullptr;
      fa.status = dstatus ? (int32_t*)dstatus + off : nullptr;
      fa.batch = cb; fa.N = p->N;
      cudaStream_t const cs = (two_streams && (chunk_no & 1)) ? ctx->aux_stream : main_stream;
      ctx->stream = cs;
      if constexpr (MODEL == 0) rc = launch_ric12<0, RIC_LQR, true>(ctx, fa);
      else rc = launch_ric24<0, RIC_LQR, true>(ctx, fa);
      ctx->stream = main_stream;
      if (rc) { ctx->pending.clear(); return rc; }
      if (overlap) {
        cudaEvent_t ev = ctx->chunk_ev[chunk_no & 3];
        NOC_CUDA(ctx, cudaEventRecord(ev, cs));
        NOC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
        for (auto& c : ctx->pending) {
          const size_t per = c.bytes / B;   // every output is [batch][...]
          NOC_CUDA(ctx, cudaMemcpyAsync((char*)c.host + off * per, (const char*)c.dev + off * per, (size_t)cb * per,
                                        cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
      }
    }
    if (two_streams) {   // the next outer chunk's rollout rewrites the scratch the auxiliary stream still reads
      NOC_CUDA(ctx, cudaEventRecord(ctx->join_ev, ctx->aux_stream));
      NOC_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->join_ev, 0));
    }
  }
  if (overlap) {
    ctx->pending.clear();
    NOC_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NOC_OK;
  }
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

template <int MODEL>
int slq_impl(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* xgoal,
             const double* QRQf, const double* u_init, double* x, double* u, double* K, double* cost, int32_t* iters,
             int32_t* alpha_idx, int32_t* status) {
  int rc;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m, N = p->N, rec = n * n + n * m;
  const void *dx0, *dxg, *dQ, *dui;
  void *dx, *du, *dK, *dJ, *dit, *dai, *dst;
  if ((rc = stage_in(ctx, 0, x0, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_in(ctx, 1, xgoal, sizeof(double) * B * n, &dxg))) return rc;
  if ((rc = stage_in(ctx, 2, QRQf, sizeof(double) * (2 * n * n + m * m), &dQ))) return rc;
  if ((rc = stage_in(ctx, 3, u_init, sizeof(double) * B * N * m, &dui))) return rc;   // warm start (NULL = hover thrust)
  // outputs double as the solver's working state; the ones the caller does not want live in scratch
  auto out_or_scratch = [&](int oslot, int sslot, void* host, size_t bytes, void** dev) -> int {
    if (host) return stage_out(ctx, oslot, host, bytes, dev);
    return scratch(ctx, sslot, bytes, dev);
  };
  if ((rc = out_or_scratch(0, SCR_X, x, sizeof(double) * B * (N + 1) * n, &dx))) return rc;
  if ((rc = out_or_scratch(1, SCR_U, u, sizeof(double) * B * N * m, &du))) return rc;
  if ((rc = out_or_scratch(2, SCR_K, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = out_or_scratch(3, SCR_J, cost, sizeof(double) * B, &dJ))) return rc;
  if ((rc = out_or_scratch(4, SCR_ITERS, iters, sizeof(int32_t) * B, &dit))) return rc;
  if ((rc = stage_out(ctx, 5, alpha_idx, sizeof(int32_t) * B * p->max_iter, &dai))) return rc;
  if ((rc = out_or_scratch(6, SCR_STATUS, status, sizeof(int32_t) * B, &dst))) return rc;
  void *dAB, *dkff, *ddV, *dmu, *dbst, *didx0, *didx1, *dcount;
  dAB = nullptr;   // both models linearise inside the sweep (FUSED): no A_k,B_k stream
  (void)rec;
  if ((rc = scratch(ctx, SCR_KFF, sizeof(double) * B * N * m, &dkff))) return rc;
  if ((rc = scratch(ctx, SCR_DV, sizeof(double) * B * 2, &ddV))) return rc;
  if ((rc = scratch(ctx, SCR_MU, sizeof(double) * B, &dmu))) return rc;
  if ((rc = scratch(ctx, SCR_BSTATUS, sizeof(int32_t) * B, &dbst))) return rc;
  if ((rc = scratch(ctx, SCR_IDX0, sizeof(int32_t) * B, &didx0))) return rc;
  if ((rc = scratch(ctx, SCR_IDX1, sizeof(int32_t) * B, &didx1))) return rc;
  if ((rc = scratch(ctx, SCR_COUNT, 64, &dcount))) return rc;
  void *dxn, *dun, *dretry0, *dretry1;
  if ((rc = scratch(ctx, SCR_RETRY0, sizeof(int32_t) * B, &dretry0))) return rc;
  if ((rc = scratch(ctx, SCR_RETRY1, sizeof(int32_t) * B, &dretry1))) return rc;
  if ((rc = scratch(ctx, SCR_XN, sizeof(double) * B * (N + 1) * n, &dxn))) return rc;
  if ((rc = scratch(ctx, SCR_UN, sizeof(double) * B * N * m, &dun))) return rc;
  void *dlc, *dflags;
  if ((rc = scratch(ctx, SCR_LC, sizeof(double) * B * (N + 1), &dlc))) return rc;
  if ((rc = scratch(ctx, SCR_FLAGS, sizeof(int32_t) * B, &dflags))) return rc;
  NOC_CUDA(ctx, cudaMemsetAsync(dflags, 0, sizeof(int32_t) * B, ctx->stream));
  using FC = FwdCfg<MODEL>;
  const size_t fwd_smem = FC::ring_bytes;
  NOC_CUDA(ctx, cudaFuncSetAttribute(k_slq_rollout<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
  const size_t wbytes = sizeof(double) * (2 * n * n + m * m);
  {
    LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
    k_slq_init<MODEL><<<(unsigned)((batch + 127) / 128), 128, wbytes, ctx->stream>>>(
        (const double*)dx0, (const double*)dxg, (const double*)dQ, (const double*)dui, (double*)dx, (double*)du, (double*)dJ, (double*)dmu,
        (int32_t*)dst, (int32_t*)dit, (int32_t*)dai, (int32_t*)didx0, batch, p->N, p->dt, p->reg0, p->max_iter);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  typename std::conditional<MODEL == 0, Ric12Args, Ric24Args>::type ra{};
  if (MODEL == 0) rc = build_w12(ctx, (const double*)dQ, false, &ra.W, &ra.Qf);
  else rc = build_w24(ctx, (const double*)dQ, &ra.W, &ra.Qf);
  if (rc) { ctx->pending.clear(); return rc; }
  const bool boxed = std::isfinite(p->u_min) || std::isfinite(p->u_max);   // input box: control-limited backward pass
  int count = (int)batch;
  int32_t* cur = (int32_t*)didx0;
  int32_t* nxt = (int32_t*)didx1;
  for (int it = 0; it < p->max_iter && count > 0; ++it) {
    // a1 happens inside the sweep (k_ric12 / k_ric24 <…, RIC_AFFINE, FUSED>)
    // a2+a3: affine backward pass (regularisation retries in-kernel)
    ra.AB = (const double*)dAB; ra.idx = cur; ra.xbar = (const double*)dx; ra.ubar = (const double*)du;
    ra.xgoal = (const double*)dxg; ra.K = (double*)dK; ra.kff = (double*)dkff; ra.dV = (double*)ddV;
    ra.mu = (double*)dmu; ra.status = (int32_t*)dbst; ra.batch = count; ra.N = p->N;
    ra.dt = p->dt;
    if constexpr (MODEL == 0) {
      ra.u_min = p->u_min; ra.u_max = p->u_max;
      rc = boxed ? launch_ric12<0, RIC_AFFINE, true, true>(ctx, ra) : launch_ric12<0, RIC_AFFINE, true>(ctx, ra);
    }
    else {
      ra.u_min = p->u_min; ra.u_max = p->u_max;
      rc = boxed ? launch_ric24<0, RIC_AFFINE, true, true>(ctx, ra) : launch_ric24<0, RIC_AFFINE, true>(ctx, ra);
    }
    if (rc) { ctx->pending.clear(); return rc; }
    // a4: closed-loop rollouts with backtracking, acceptance, convergence, next work list.  One launch per step
    // length: trial 0 on the whole list, trial j+1 only on the instances whose trial j was rejected.
    SlqFwdArgs fa{};
    fa.flags = (int32_t*)dflags;
    fa.xgoal = (const double*)dxg;
    fa.qrqf = (const double*)dQ; fa.x = (double*)dx; fa.u = (double*)du; fa.xn = (double*)dxn; fa.un = (double*)dun;
    fa.K = (const double*)dK;
    fa.kff = (const double*)dkff; fa.dV = (const double*)ddV; fa.lc = (double*)dlc; fa.J = (double*)dJ;
    fa.bstatus = (const int32_t*)dbst;
    fa.status = (int32_t*)dst; fa.iters = (int32_t*)dit; fa.alpha_idx = (int32_t*)dai; fa.N = p->N;
    fa.iter = it; fa.max_iter = p->max_iter; fa.n_alpha = p->n_alpha; fa.dt = p->dt; fa.tol = p->tol;
    fa.armijo = p->armijo; fa.u_min = p->u_min; fa.u_max = p->u_max;
    int32_t hc[2] = {0, 0};
    const int32_t* list = cur;
    int list_count = count;
    int32_t* rin = (int32_t*)dretry0;
    int32_t* rout = (int32_t*)dretry1;
    for (int trial = 0; trial < p->n_alpha && list_count > 0; ++trial) {
      fa.idx = list; fa.count = list_count; fa.trial = trial;
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_rollout<MODEL><<<(unsigned)((list_count + FC::inst - 1) / FC::inst), FC::threads, fwd_smem, ctx->stream>>>(fa);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      {
        const long long total = (long long)list_count * (p->N + 1);
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_stage_cost<MODEL><<<(unsigned)((total + 127) / 128), 128, wbytes, ctx->stream>>>(fa);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_accept<MODEL><<<(unsigned)((list_count + 3) / 4), 128, 0, ctx->stream>>>(fa);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      {  // rejected trials -> sorted retry list
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_compact_flags<<<1, 1024, 0, ctx->stream>>>((int32_t*)dflags, 2, batch, rin, (int32_t*)dcount + 1);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      NOC_CUDA(ctx, cudaMemcpyAsync(hc + 1, (int32_t*)dcount + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
      NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      list = rin; list_count = hc[1];
      std::swap(rin, rout);
      // Speculative rounds: the remaining step lengths of the rejected instances as parallel candidates (one launch
      // set instead of up to n_alpha-1 sequential ones), when the candidate trajectories fit a 2 GB scratch.  All
      // of them at once while few instances backtrack (1 % without input limits); in rounds of four when many do
      // (two thirds of the instances under a tight input box, accepted anywhere down to 2^-6): an instance leaves
      // with the first round that holds a passing step length, which is the oracle's sequential search.
      const int J = p->n_alpha - 1 - trial;
      const size_t per_cand = sizeof(double) * ((N + 1) * n + N * m + (N + 1));
      if (trial == 0 && list_count > 0 && J > 0) {
        const int round_J = ((size_t)list_count * J > 2 * (size_t)count) ? std::min(J, 4) : J;
        if ((size_t)list_count * round_J * per_cand <= ((size_t)2 << 30)) {
          int first = 1;
          while (list_count > 0 && first < p->n_alpha) {
            const int Jr = std::min(round_J, p->n_alpha - first);
            void *dxc, *duc, *dlcc;
            const size_t rc_ = (size_t)list_count * Jr;
            int e1 = scratch(ctx, SCR_XC, sizeof(double) * rc_ * (N + 1) * n, &dxc);
            int e2 = e1 ? e1 : scratch(ctx, SCR_UC, sizeof(double) * rc_ * N * m, &duc);
            int e3 = e2 ? e2 : scratch(ctx, SCR_LCC, sizeof(double) * rc_ * (N + 1), &dlcc);
            if (e3) { ctx->pending.clear(); return e3; }
            fa.idx = list; fa.trial = first; fa.cand_J = Jr; fa.cand_first = first; fa.cand_last = (first + Jr >= p->n_alpha);
            fa.xc = (double*)dxc; fa.uc = (double*)duc; fa.lcc = (double*)dlcc;
            fa.count = (int)rc_;
            {
              LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
              k_slq_rollout<MODEL><<<(unsigned)((rc_ + FC::inst - 1) / FC::inst), FC::threads, fwd_smem, ctx->stream>>>(fa);
            }
            NOC_CUDA(ctx, cudaGetLastError());
            {
              const long long total = (long long)rc_ * (p->N + 1);
              LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
              k_slq_stage_cost<MODEL><<<(unsigned)((total + 127) / 128), 128, wbytes, ctx->stream>>>(fa);
            }
            NOC_CUDA(ctx, cudaGetLastError());
            fa.count = list_count;
            {
              LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
              k_slq_accept_cand<MODEL><<<(unsigned)((list_count + 3) / 4), 128, 0, ctx->stream>>>(fa);
            }
            NOC_CUDA(ctx, cudaGetLastError());
            first += Jr;
            if (fa.cand_last) { list_count = 0; break; }   // every rejected instance has been settled
            {  // instances without a winner so far -> sorted list of the next round
              LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
              k_compact_flags<<<1, 1024, 0, ctx->stream>>>((int32_t*)dflags, 2, batch, rin, (int32_t*)dcount + 1);
            }
            NOC_CUDA(ctx, cudaGetLastError());
            NOC_CUDA(ctx, cudaMemcpyAsync(hc + 1, (int32_t*)dcount + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
            NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            list = rin; list_count = hc[1];
            std::swap(rin, rout);
          }
          fa.cand_J = 0;
          list_count = 0;
        }
      }
    }
    {  // instances that go on to the next iteration -> sorted work list
      LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
      k_compact_flags<<<1, 1024, 0, ctx->stream>>>((int32_t*)dflags, 1, batch, nxt, (int32_t*)dcount);
    }
    NOC_CUDA(ctx, cudaGetLastError());
    NOC_CUDA(ctx, cudaMemcpyAsync(hc, dcount, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    count = hc[0];
    std::swap(cur, nxt);
  }
  return flush_out(ctx, false);
}

}  // namespace

extern "C" {

const char* noc_version(void) { return "noc_b200 0.1 (sm_100a)"; }

void noc_problem_defaults(noc_problem_t* p, int32_t model, int32_t N) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->model = model;
  p->n = (model == NOC_MODEL_DUAL24) ? 24 : 12;
  p->m = (model == NOC_MODEL_DUAL24) ? 8 : 4;
  p->N = N;
  p->layout = NOC_LAYOUT_A_THEN_B;
  p->max_iter = 50;
  p->n_alpha = 11;
  p->flags = NOC_FLAG_NONE;
  p->dt = 0.02;
  p->tol = 1e-8;
  p->reg0 = 0.0;
  p->armijo = 1e-4;
  p->u_min = -INFINITY;
  p->u_max = INFINITY;
}

int noc_create(noc_ctx** out, int device_id) {
  if (!out) return NOC_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return NOC_ERR_NO_DEVICE; }
  if (device_id < 0 || device_id >= count) return NOC_ERR_BAD_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return NOC_ERR_CUDA;
  if (prop.major != 10) return NOC_ERR_NO_DEVICE;  // sm_100a cubins only
  if (cudaSetDevice(device_id) != cudaSuccess) return NOC_ERR_CUDA;
  noc_ctx* c = new noc_ctx();
  c->device = device_id;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NOC_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NOC_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NOC_ERR_CUDA; }
  for (auto& e : c->chunk_ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete c; return NOC_ERR_CUDA; }
  if (cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) != cudaSuccess) { delete c; return NOC_ERR_CUDA; }
  c->stream = c->own_stream;
  *out = c;
  return NOC_OK;
}

void noc_destroy(noc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto* v : {&ctx->in_slots, &ctx->out_slots, &ctx->scratch})
    for (auto& b : *v)
      if (b.p) cudaFree(b.p);
  for (auto& e : ctx->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (auto& e : ctx->chunk_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* noc_last_error(const noc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int noc_set_stream(noc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return NOC_OK;
}

int noc_synchronize(noc_ctx* ctx) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NOC_OK;
}

int64_t noc_launch_count(const noc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int noc_profile_enable(noc_ctx* ctx, int on) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  ctx->profiling = on != 0;
  return NOC_OK;
}

static int profile_collect(noc_ctx* ctx) {
  if (ctx->ev_used.empty()) return NOC_OK;
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& u : ctx->ev_used) {
    float ms = 0.f;
    NOC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_pool[u.second].first, ctx->ev_pool[u.second].second));
    ctx->prof_ms[u.first] += ms;
  }
  ctx->ev_used.clear();
  ctx->ev_next = 0;
  return NOC_OK;
}

int noc_profile_read(noc_ctx* ctx, int kernel_id, double* total_ms, int64_t* launches) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (kernel_id < 0 || kernel_id >= NOC_KERNEL_COUNT) return fail(ctx, NOC_ERR_BAD_ARG, "unknown kernel id");
  int rc = profile_collect(ctx);
  if (rc) return rc;
  if (total_ms) *total_ms = ctx->prof_ms[kernel_id];
  if (launches) *launches = ctx->prof_launches[kernel_id];
  return NOC_OK;
}

int noc_profile_reset(noc_ctx* ctx) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  int rc = profile_collect(ctx);
  if (rc) return rc;
  for (int i = 0; i < NOC_KERNEL_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
  return NOC_OK;
}

int noc_lqr_solve_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* AB, const double* QRQf,
                        int qr_per_instance, const double* x0dev, double* K, double* K0, double* P0, double* cost,
                        int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (!AB || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "AB and QRQf are required");
  if (batch == 0) return NOC_OK;
  // per-instance weights run on the thread-per-instance kernel whatever the shape: the tuned kernels keep one W / Qf
  // per CTA in shared memory (a per-instance copy would halve their occupancy)
  const bool n24 = (p->n == 24 && p->m == 8);
  const bool generic = qr_per_instance || (!(p->n == 12 && p->m == 4) && !n24);
  if (generic && (p->n < 1 || p->n > RG_NMAX || p->m < 1 || p->m > RG_MMAX))
    return fail(ctx, NOC_ERR_UNSUPPORTED, "LTV sweep: 1 <= n <= 24 and 1 <= m <= 8 (tuned kernels: (12,4), (24,8))");
  if (misaligned16(AB) && is_device_ptr(AB)) return fail(ctx, NOC_ERR_BAD_ARG, "device AB must be 16-byte aligned (TMA bulk copies)");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = p->n, m = p->m, N = p->N;
  const size_t rec = (size_t)n * n + (size_t)n * m, B = (size_t)batch;
  const void *dAB, *dQ, *dx0;
  void *dK, *dK0, *dP0, *dcost, *dstatus;
  if ((rc = stage_in(ctx, 0, AB, sizeof(double) * B * N * rec, &dAB))) return rc;
  const size_t qr_len = (size_t)(2 * n * n + m * m);
  if ((rc = stage_in(ctx, 1, QRQf, sizeof(double) * qr_len * (qr_per_instance ? B : 1), &dQ))) return rc;
  if ((rc = stage_in(ctx, 2, x0dev, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_out(ctx, 0, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 1, K0, sizeof(double) * B * m * n, &dK0))) return rc;
  if ((rc = stage_out(ctx, 2, P0, sizeof(double) * B * n * n, &dP0))) return rc;
  if ((rc = stage_out(ctx, 3, cost, sizeof(double) * B, &dcost))) return rc;
  if ((rc = stage_out(ctx, 4, status, sizeof(int32_t) * B, &dstatus))) return rc;
  if (generic) {
    RicGenArgs g{};
    g.AB = (const double*)dAB; g.QRQf = (const double*)dQ; g.x0 = (const double*)dx0; g.K = (double*)dK;
    g.K0 = (double*)dK0; g.P0 = (double*)dP0; g.cost = (double*)dcost; g.status = (int32_t*)dstatus;
    g.batch = batch; g.n = n; g.m = m; g.N = N; g.layout = p->layout;
    g.qr_stride = qr_per_instance ? (long long)(2 * n * n + m * m) : 0;
    {
      LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
      k_ric_generic<<<(unsigned)((batch + 63) / 64), 64, 0, ctx->stream>>>(g);
    }
    rc = (cudaGetLastError() == cudaSuccess) ? NOC_OK : fail(ctx, NOC_ERR_CUDA, "k_ric_generic launch failed");
  } else {
    rc = (n24 ? lqr24_device : lqr12_device)(ctx, p, batch, (const double*)dAB, (const double*)dQ, (const double*)dx0,
                                             (double*)dK, (double*)dK0, (double*)dP0, (double*)dcost, (int32_t*)dstatus);
  }
  if (rc) { ctx->pending.clear(); return rc; }
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_rollout_open_loop(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                          double* xbar) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "rollout needs a built-in model");
  if (!x0 || !xbar) return fail(ctx, NOC_ERR_BAD_ARG, "x0 and xbar are required");
  if (batch == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = p->n, m = p->m, N = p->N;
  const void *dx0, *du;
  void* dx;
  if ((rc = stage_in(ctx, 0, x0, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_in(ctx, 1, ubar, sizeof(double) * B * N * m, &du))) return rc;
  if ((rc = stage_out(ctx, 0, xbar, sizeof(double) * B * (N + 1) * n, &dx))) return rc;
  const unsigned grid = (unsigned)((batch + 127) / 128);
  {
    LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
    if (p->model == NOC_MODEL_QUAD12)
      k_rollout_open_loop<0><<<grid, 128, 0, ctx->stream>>>((const double*)dx0, (const double*)du, (double*)dx, batch, p->N, p->dt,
                                                  (long long)(p->N + 1) * p->n, (long long)p->n);
    else
      k_rollout_open_loop<1><<<grid, 128, 0, ctx->stream>>>((const double*)dx0, (const double*)du, (double*)dx, batch, p->N, p->dt,
                                                  (long long)(p->N + 1) * p->n, (long long)p->n);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_linearize_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* xbar, const double* ubar,
                        double* AB) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "linearisation needs a built-in model");
  if (!xbar || !AB) return fail(ctx, NOC_ERR_BAD_ARG, "xbar and AB are required");
  if (batch == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = p->n, m = p->m, N = p->N, rec = n * n + n * m;
  const void *dx, *du;
  void* dAB;
  if ((rc = stage_in(ctx, 0, xbar, sizeof(double) * B * (N + 1) * n, &dx))) return rc;
  if ((rc = stage_in(ctx, 1, ubar, sizeof(double) * B * N * m, &du))) return rc;
  if ((rc = stage_out(ctx, 0, AB, sizeof(double) * B * N * rec, &dAB))) return rc;
  const long long total = (long long)batch * p->N;
  const unsigned grid = (unsigned)((total + 127) / 128);
  {
    LaunchScope ls(ctx, NOC_KERNEL_LINEARIZE);
    if (p->model == NOC_MODEL_QUAD12)
      k_linearize<0><<<grid, 128, 0, ctx->stream>>>((const double*)dx, (const double*)du, (double*)dAB, batch, p->N, p->dt, p->layout, nullptr);
    else
      k_linearize<1><<<grid, 128, 0, ctx->stream>>>((const double*)dx, (const double*)du, (double*)dAB, batch, p->N, p->dt, p->layout, nullptr);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_lqr_solve_from_x0(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* QRQf,
                          double* K, double* K0, double* P0, double* cost, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "from_x0 pipeline needs a built-in model");
  if (!x0 || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "x0 and QRQf are required");
  if (batch == 0) return NOC_OK;
  return p->model == NOC_MODEL_QUAD12 ? from_x0_impl<0>(ctx, p, batch, x0, nullptr, QRQf, nullptr, K, K0, P0, cost, status)
                                      : from_x0_impl<1>(ctx, p, batch, x0, nullptr, QRQf, nullptr, K, K0, P0, cost, status);
}

int noc_lqr_solve_along(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                        const double* QRQf, double* xbar, double* K, double* K0, double* P0, double* cost,
                        int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "noc_lqr_solve_along needs a built-in model");
  if (!x0 || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "x0 and QRQf are required");
  if (batch == 0) return NOC_OK;
  return p->model == NOC_MODEL_QUAD12 ? from_x0_impl<0>(ctx, p, batch, x0, ubar, QRQf, xbar, K, K0, P0, cost, status)
                                      : from_x0_impl<1>(ctx, p, batch, x0, ubar, QRQf, xbar, K, K0, P0, cost, status);
}

int noc_dare_solve_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* AB, const double* QR,
                         double* P, double* K, int32_t* iters, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (!AB || !QR) return fail(ctx, NOC_ERR_BAD_ARG, "AB and QR are required");
  if (p->max_iter < 1) return fail(ctx, NOC_ERR_BAD_ARG, "max_iter must be >= 1");
  if (batch == 0) return NOC_OK;
  const bool tuned = p->n == 12 && p->m == 4;
  if (tuned && misaligned16(AB) && is_device_ptr(AB)) return fail(ctx, NOC_ERR_BAD_ARG, "device AB must be 16-byte aligned");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = (size_t)p->n, m = (size_t)p->m;
  const void *dAB, *dQ;
  void *dP, *dK, *dit, *dst;
  if ((rc = stage_in(ctx, 0, AB, sizeof(double) * B * (n * n + n * m), &dAB))) return rc;
  if ((rc = stage_in(ctx, 1, QR, sizeof(double) * (n * n + m * m), &dQ))) return rc;
  if ((rc = stage_out(ctx, 0, P, sizeof(double) * B * n * n, &dP))) return rc;
  if ((rc = stage_out(ctx, 1, K, sizeof(double) * B * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 2, iters, sizeof(int32_t) * B, &dit))) return rc;
  if ((rc = stage_out(ctx, 3, status, sizeof(int32_t) * B, &dst))) return rc;
  if (!tuned) {
    // every other dimension (n=24, m=8 included): the thread-per-instance completeness kernel, same arithmetic
    RicGenArgs g{};
    g.AB = (const double*)dAB; g.QRQf = (const double*)dQ; g.K0 = (double*)dK; g.P0 = (double*)dP;
    g.iters = (int32_t*)dit; g.status = (int32_t*)dst; g.batch = batch; g.n = p->n; g.m = p->m; g.N = 1;
    g.layout = p->layout; g.dare = 1; g.max_iter = p->max_iter; g.tol = p->tol;
    {
      LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
      k_ric_generic<<<(unsigned)((batch + 63) / 64), 64, 0, ctx->stream>>>(g);
    }
    NOC_CUDA(ctx, cudaGetLastError());
    return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
  }
  Ric12Args a{};
  if ((rc = build_w12(ctx, (const double*)dQ, true, &a.W, &a.Qf))) { ctx->pending.clear(); return rc; }
  a.AB = (const double*)dAB; a.K = (double*)dK; a.P0 = (double*)dP; a.iters = (int32_t*)dit; a.status = (int32_t*)dst;
  a.tol = p->tol; a.max_iter = p->max_iter; a.batch = batch; a.N = 1;
  rc = launch_ric12_layout<RIC_DARE>(ctx, p->layout, a);
  if (rc) { ctx->pending.clear(); return rc; }
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_dare_solve_at(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* xbar, const double* ubar,
                      const double* QR, double* P, double* K, int32_t* iters, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (!xbar || !QR) return fail(ctx, NOC_ERR_BAD_ARG, "xbar and QR are required");
  if (p->max_iter < 1) return fail(ctx, NOC_ERR_BAD_ARG, "max_iter must be >= 1");
  if (p->model != NOC_MODEL_QUAD12) return fail(ctx, NOC_ERR_UNSUPPORTED, "noc_dare_solve_at: NOC_MODEL_QUAD12 only");
  if (batch == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = 12, m = 4;
  const void *dxb, *dub, *dQ;
  void *dP, *dK, *dit, *dst;
  if ((rc = stage_in(ctx, 0, xbar, sizeof(double) * B * n, &dxb))) return rc;
  if ((rc = stage_in(ctx, 1, QR, sizeof(double) * (n * n + m * m), &dQ))) return rc;
  if ((rc = stage_in(ctx, 2, ubar, sizeof(double) * B * m, &dub))) return rc;
  if (misaligned16(dxb) || (dub && misaligned16(dub))) return fail(ctx, NOC_ERR_BAD_ARG, "device xbar / ubar must be 16-byte aligned");
  if ((rc = stage_out(ctx, 0, P, sizeof(double) * B * n * n, &dP))) return rc;
  if ((rc = stage_out(ctx, 1, K, sizeof(double) * B * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 2, iters, sizeof(int32_t) * B, &dit))) return rc;
  if ((rc = stage_out(ctx, 3, status, sizeof(int32_t) * B, &dst))) return rc;
  Ric12Args a{};
  if ((rc = build_w12(ctx, (const double*)dQ, true, &a.W, &a.Qf))) { ctx->pending.clear(); return rc; }
  a.xbar = (const double*)dxb; a.xbar_sb = 12; a.xbar_sk = 12; a.ubar = (const double*)dub; a.dt = p->dt;
  a.K = (double*)dK; a.P0 = (double*)dP; a.iters = (int32_t*)dit; a.status = (int32_t*)dst;
  a.tol = p->tol; a.max_iter = p->max_iter; a.batch = batch; a.N = 1;
  rc = launch_ric12<0, RIC_DARE, true>(ctx, a);
  if (rc) { ctx->pending.clear(); return rc; }
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_slq_solve_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* xgoal,
                        const double* QRQf, double* x, double* u, double* K, double* cost, int32_t* iters,
                        int32_t* alpha_idx, int32_t* status) {
  return noc_slq_solve_warm(ctx, p, batch, x0, xgoal, QRQf, nullptr, x, u, K, cost, iters, alpha_idx, status);
}

int noc_slq_solve_warm(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* xgoal,
                       const double* QRQf, const double* u_init, double* x, double* u, double* K, double* cost,
                       int32_t* iters, int32_t* alpha_idx, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "SLQ needs a built-in nonlinear model");
  if (!x0 || !xgoal || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "x0, xgoal and QRQf are required");
  if (p->max_iter < 1 || p->n_alpha < 1 || p->n_alpha > 32)
    return fail(ctx, NOC_ERR_BAD_ARG, "need max_iter >= 1 and 1 <= n_alpha <= 32");
  if (!(p->u_min < p->u_max)) return fail(ctx, NOC_ERR_BAD_ARG, "need u_min < u_max (-INFINITY / INFINITY = no input box)");
  if (batch > 0x7fffffffLL) return fail(ctx, NOC_ERR_BAD_ARG, "batch too large for 32-bit work lists");
  if (batch == 0) return NOC_OK;
  return p->model == NOC_MODEL_QUAD12
             ? slq_impl<0>(ctx, p, batch, x0, xgoal, QRQf, u_init, x, u, K, cost, iters, alpha_idx, status)
             : slq_impl<1>(ctx, p, batch, x0, xgoal, QRQf, u_init, x, u, K, cost, iters, alpha_idx, status);
}

int noc_selftest_rcp(noc_ctx* ctx, const double* x, double* y_fast, double* y_ieee, int64_t n) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!x || !y_fast || !y_ieee || n < 1) return fail(ctx, NOC_ERR_BAD_ARG, "x, outputs and n>=1 required");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const void* dx;
  void *df, *di;
  int rc;
  if ((rc = stage_in(ctx, 0, x, sizeof(double) * (size_t)n, &dx))) return rc;
  if ((rc = stage_out(ctx, 0, y_fast, sizeof(double) * (size_t)n, &df))) return rc;
  if ((rc = stage_out(ctx, 1, y_ieee, sizeof(double) * (size_t)n, &di))) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_SETUP);
    k_selftest_rcp<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const double*)dx, (double*)df, (double*)di, n);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return flush_out(ctx, false);
}

int noc_argmin_cost(noc_ctx* ctx, const double* cost, int64_t batch, int64_t* best_index, double* best_cost) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!cost || !best_index || !best_cost || batch < 1) return fail(ctx, NOC_ERR_BAD_ARG, "cost, outputs and batch>=1 required");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const void* dc;
  int rc;
  if ((rc = stage_in(ctx, 0, cost, sizeof(double) * (size_t)batch, &dc))) return rc;
  void* s;
  if ((rc = scratch(ctx, SCR_ARGMIN, 16, &s))) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_ARGMIN);
    k_argmin_cost<<<1, 1024, 0, ctx->stream>>>((const double*)dc, batch, (long long*)s, (double*)s + 1);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  long long hi;
  double hv;
  NOC_CUDA(ctx, cudaMemcpyAsync(&hi, s, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NOC_CUDA(ctx, cudaMemcpyAsync(&hv, (double*)s + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *best_index = hi;
  *best_cost = hv;
  return NOC_OK;
}

}  // extern "C"

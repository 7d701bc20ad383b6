// noc_b200.cu — C-ABI of libnoc_b200.so (include/lqr_control/noc_b200.h).
// Host-side batch scheduler: pointer classification, staging, chunking to the free HBM, kernel launches.
// There is no CPU fallback anywhere in this file: every entry point either launches CUDA kernels or fails.
#include "../../include/lqr_control/noc_b200.h"

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is loaded with dlopen at run time

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "pipeline_kernels.cuh"
#include "riccati12.cuh"
#include "riccati24.cuh"
#include "riccati24_mma.cuh"
#include "riccati_generic.cuh"
#include "riccati_scan.cuh"
#include "riccati12_mma.cuh"
#include "riccati12_wpi.cuh"
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
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr, defer_ev = nullptr;
  cudaEvent_t pipe_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // host-AB pipeline: in / swept / out x 2 buffers
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
  // scheduler options (noc_set_option)
  int64_t opt[NOC_OPT_COUNT] = {0};
  // kernels whose function attributes (dynamic shared memory, carve-out) have been set on this context's device
  std::unordered_map<const void*, size_t> configured;   // kernel -> largest dynamic shared-memory size it was configured for
  // scratch budgets already granted, per call signature (scratch_budget)
  struct BudgetKey { int entry; int n, N, extra; int64_t batch; size_t budget; };
  std::vector<BudgetKey> budgets;
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

// Every entry point that stages host outputs holds one of these: whatever path the call leaves by (a failed
// allocation, a CUDA error, an argument check after staging), no host copy-back of THIS call survives into the
// next one (the caller may free its arrays after an error).  flush_out() empties the list itself on success.
struct PendingGuard {
  noc_ctx* c;
  explicit PendingGuard(noc_ctx* ctx) : c(ctx) { c->pending.clear(); }
  ~PendingGuard() { c->pending.clear(); }
};

// Calls that fan work out over the context's auxiliary / copy streams hold one of these: on ANY exit (errors included)
// the side streams are drained, so no copy into the caller's host arrays is still in flight when the call returns.
struct SideStreamDrain {
  noc_ctx* c;
  bool armed;
  ~SideStreamDrain() {
    if (!armed) return;
    cudaStreamSynchronize(c->aux_stream);
    cudaStreamSynchronize(c->copy_stream);
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
  if (cudaMalloc(&b.p, want) != cudaSuccess) {   // the cached scratch budgets were granted under other memory conditions
    b.p = nullptr;
    ctx->budgets.clear();
    cudaGetLastError();
    return fail(ctx, NOC_ERR_OOM, "cudaMalloc of " + std::to_string(want) + " bytes of scratch failed");
  }
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
  std::vector<PendingCopy> todo;
  todo.swap(ctx->pending);   // the list is empty from here on, also when a copy below fails
  for (auto& c : todo)
    NOC_CUDA(ctx, cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  const bool had_host = !todo.empty();
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

// Scratch budget of one call: NOC_OPT_SCRATCH_LIMIT_BYTES, else 60 % of (free HBM + what the context already owns).
// cudaMemGetInfo is an ioctl into the kernel driver and queues behind every other client of the GPU's resource-manager
// lock (a monitoring nvidia-smi, another process' allocation): measured as 5-90 ms stalls of an otherwise 0.1 ms
// enqueue (tools/probe_outliers2.py, profiles/e2e_outliers_r02.md).  The answer is therefore cached per call
// signature — a repeated call (an MPC loop) asks the driver nothing — and dropped when an allocation fails or an
// option changes (ensure / noc_set_option).
enum { BUDGET_FROM_X0 = 0, BUDGET_SLQ, BUDGET_HOST_AB };
int scratch_budget(noc_ctx* ctx, int entry, int n, int N, int extra, int64_t batch, size_t have, size_t* out) {
  if (ctx->opt[NOC_OPT_SCRATCH_LIMIT_BYTES] > 0) {   // caller's cap (tests use it to force several outer chunks)
    *out = std::max<size_t>((size_t)ctx->opt[NOC_OPT_SCRATCH_LIMIT_BYTES], (size_t)1 << 20);
    return NOC_OK;
  }
  for (auto& k : ctx->budgets)
    if (k.entry == entry && k.n == n && k.N == N && k.extra == extra && k.batch == batch) { *out = k.budget; return NOC_OK; }
  size_t free_b = 0, total_b = 0;
  NOC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  *out = std::max<size_t>((size_t)((free_b + have) * 0.6), (size_t)64 << 20);
  if (ctx->budgets.size() >= 16) ctx->budgets.erase(ctx->budgets.begin());
  ctx->budgets.push_back({entry, n, N, extra, batch, *out});
  return NOC_OK;
}

enum { SCR_W = 0, SCR_XBAR, SCR_AB, SCR_ARGMIN, SCR_K, SCR_KFF, SCR_DV, SCR_MU, SCR_J, SCR_BSTATUS, SCR_STATUS,
       SCR_ITERS, SCR_IDX0, SCR_IDX1, SCR_COUNT, SCR_X, SCR_U, SCR_XN, SCR_UN, SCR_RETRY0, SCR_RETRY1, SCR_LC, SCR_FLAGS, SCR_XC, SCR_UC, SCR_LCC,
       SCR_PIPE_AB0, SCR_PIPE_AB1, SCR_PIPE_X0, SCR_PIPE_X1, SCR_PIPE_QR0, SCR_PIPE_QR1, SCR_PIPE_OUT0,
       SCR_PIPE_OUT_LAST = SCR_PIPE_OUT0 + 9, SCR_ARGMIN_PART, SCR_COMM_PAIRS, SCR_COMM_COST, SCR_DELTA, SCR_PREP, SCR_SPEC_LIST, SCR_SPEC_X, SCR_SPEC_U, SCR_SPEC_LC, SCR_END };

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
  if (p->n < 1 || p->m < 1) return fail(ctx, NOC_ERR_BAD_ARG, "n and m must be >= 1");
  if (p->n > RG_NMAX || p->m > RG_MMAX)
    return fail(ctx, NOC_ERR_UNSUPPORTED, "supported dimensions: 1 <= n <= 24, 1 <= m <= 8");
  return NOC_OK;
}

// Function attributes are per device; a context is bound to one device, so each kernel is configured once per
// context (the SLQ loop launches the same instantiations hundreds of times per call).
int configure_kernel(noc_ctx* ctx, const void* kern, size_t smem, bool carveout) {
  auto it = ctx->configured.find(kern);
  if (it != ctx->configured.end() && it->second >= smem) return NOC_OK;
  NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (carveout) NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  ctx->configured[kern] = smem;
  return NOC_OK;
}

bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

// kernel configurations: 8 instances per warp; LQR/DARE 2 CTAs x 4 warps per SM,
// SLQ backward pass 1 CTA x 7 warps (its per-instance shared-memory slice is larger)
template <int MODE> struct RicCfg { static constexpr int warps = 4, min_ctas = 2; };
template <> struct RicCfg<RIC_AFFINE> { static constexpr int warps = 7, min_ctas = 1; };

// Wave plan of a latency-bound sweep over a work list (the SLQ backward pass, whose list shrinks from iteration to
// iteration).  Every warp runs the same N dependent steps, so a launch costs (number of rounds) x T(w), where w is the
// number of warps resident per SM and T(w) = T(1) (1 + slope (w - 1)) grows slowly with w (n=12: 0.43 -> 0.76 ms from one
// to seven warps per SM at N=200; profiles/launches_cfg3_slq_r01d.csv).  The wide configuration (one CTA of `wide`
// warps per SM) pays T(wide) per wave however few SMs the last wave fills; one-warp CTAs spread breadth-first over the
// SMs, and capping their residency at w per SM (by the size of the dynamic shared-memory request) turns e.g. 1.14 wide
// waves (2 x T(7)) into two balanced rounds of four warps per SM (2 x T(4)).  Returns 0 for the wide configuration,
// else the residency cap w of the one-warp configuration.
constexpr size_t kSmemPerSm = 233472, kSmemPerCtaReserved = 1024, kSmemMaxDynamic = 232448;
int plan_waves(int64_t nw, int sms, int wide, size_t narrow_smem, double slope) {
  auto T = [&](int w) { return 1.0 + slope * (w - 1); };
  const int narrow_max = (int)std::min<size_t>(kSmemPerSm / (narrow_smem + kSmemPerCtaReserved), 16);
  double best = (double)((nw + (int64_t)wide * sms - 1) / ((int64_t)wide * sms)) * T(wide);
  int pick = 0;
  for (int w = 1; w <= narrow_max; ++w) {
    const double c = (double)((nw + (int64_t)w * sms - 1) / ((int64_t)w * sms)) * T(w);
    if (c < best - 1e-9) { best = c; pick = w; }
  }
  return pick;
}
// The same plan from a measured table: narrow_cost[w-1] = T(w) / T(1) for w = 1 .. narrow_max one-warp CTAs per SM, wide_cost
// for a full wave of the wide configuration.  (k_ric24_mma, tools/probe_ric24.py: 1, 1.01, 1.03, 1.06 up to one warp per
// scheduler, 1.47 / 1.52 at six / eight, 2.2 for twelve in two six-warp CTAs — not a straight line; and the one-warp
// build's 255 registers allow eight CTAs per SM, not the ten its shared memory would.)
int plan_waves_table(int64_t nw, int sms, int wide, double wide_cost, const double* narrow_cost, int narrow_max) {
  double best = (double)((nw + (int64_t)wide * sms - 1) / ((int64_t)wide * sms)) * wide_cost;
  int pick = 0;
  for (int w = 1; w <= narrow_max; ++w) {
    const double c = (double)((nw + (int64_t)w * sms - 1) / ((int64_t)w * sms)) * narrow_cost[w - 1];
    if (c < best - 1e-9) { best = c; pick = w; }
  }
  return pick;
}
// dynamic shared-memory request that lets exactly `w` CTAs share an SM
size_t smem_for_residency(int w, size_t need) {
  size_t s = (kSmemPerSm / (size_t)w - kSmemPerCtaReserved) & ~(size_t)127;
  return std::min(std::max(s, need), kSmemMaxDynamic);
}

template <int LAYOUT, int MODE, bool FUSED, int W, int MIN_CTAS, bool BOX = false>
int launch_ric12_cfg(noc_ctx* ctx, const Ric12Args& a, int residency = 0) {
  auto kern = k_ric12<LAYOUT, MODE, FUSED, W, MIN_CTAS, BOX>;
  const size_t need = ric12_smem_bytes<MODE, FUSED>(W);
  const size_t smem = residency > 0 ? smem_for_residency(residency, need) : need;
  int rc = configure_kernel(ctx, (const void*)kern, residency > 0 ? kSmemMaxDynamic : need, true);
  if (rc) return rc;
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
  if (MODE == RIC_AFFINE && !ctx->opt[NOC_OPT_NO_WAVE_PLAN]) {
    const int w = plan_waves((a.batch + 7) / 8, ctx->sm_count, RicCfg<MODE>::warps, ric12_smem_bytes<MODE, FUSED>(1), 0.128);
    if (w > 0) return launch_ric12_cfg<LAYOUT, MODE, FUSED, 1, 1, BOX>(ctx, a, w);
    return launch_ric12_cfg<LAYOUT, MODE, FUSED, RicCfg<MODE>::warps, RicCfg<MODE>::min_ctas, BOX>(ctx, a);
  }
  // short work lists (the tail of an SLQ solve, small batches): one warp per CTA, so that the few warps spread over
  // all SMs instead of sharing a handful — the sweep is then latency-bound and a warp alone on an SM steps ~2x faster
  if (a.batch <= 8LL * 2 * ctx->sm_count) return launch_ric12_cfg<LAYOUT, MODE, FUSED, 1, 1, BOX>(ctx, a);
  return launch_ric12_cfg<LAYOUT, MODE, FUSED, RicCfg<MODE>::warps, RicCfg<MODE>::min_ctas, BOX>(ctx, a);
}

// The LTV sweep with its products on the FP64 tensor cores (riccati12_mma.cuh): A_THEN_B records, bit-identical to k_ric12
template <int W, int CTAS, bool DIRECT>
int launch_ric12_mma_cfg(noc_ctx* ctx, const Ric12Args& a) {
  auto kern = k_ric12_mma<W, CTAS, DIRECT>;
  const size_t smem = ric12_mma_smem_bytes<DIRECT>(W);
  int rc = configure_kernel(ctx, (const void*)kern, smem, true);
  if (rc) return rc;
  const unsigned grid = (unsigned)((a.batch + 8LL * W - 1) / (8LL * W));
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, W * 32, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}
// per-instance weights (a.W = the QRQf array, a.w_stride = 304): W' fragments from L2 per instance-step
int launch_ric12_mma_perw(noc_ctx* ctx, const Ric12Args& a) {
  auto kern = k_ric12_mma<4, 2, false, true>;
  const size_t smem = ric12_mma_smem_bytes<false>(4);
  int rc = configure_kernel(ctx, (const void*)kern, smem, true);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<(unsigned)((a.batch + 31) / 32), 128, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}
int launch_ric12_mma(noc_ctx* ctx, const Ric12Args& a) {
  switch (ctx->opt[NOC_OPT_TENSOR_SWEEP_VARIANT]) {
    case 1: return launch_ric12_mma_cfg<8, 1, false>(ctx, a);   // + phase interlock of scheduler mates (3.06 ms)
    case 2: return launch_ric12_mma_cfg<4, 3, true>(ctx, a);    // fragments straight from L2, 12 warps per SM (3.05 ms)
    default: return launch_ric12_mma_cfg<4, 2, false>(ctx, a);  // records staged by TMA, 2 CTAs x 4 warps per SM (2.87 ms)
  }
}

// Short SLQ work lists at n=12 (riccati12_wpi.cuh): one instance per warp on the tensor cores, step records by k_prep12;
// one-warp CTAs spread breadth-first over the SMs.  `prep_buf`: room for batch x N x W12_PREP_DOUBLES doubles.
constexpr int kRic12WpiMaxPerSm = 12;   // (tools/probe_slq12.py: 0.19 ms at one warp per SM, 0.36 + 0.05 at twelve, against 0.43-0.45 ms on k_ric12; two rounds beyond)
int launch_ric12_wpi(noc_ctx* ctx, const Ric12Args& a, double* prep_buf) {
  const long long recs = a.batch * a.N;
  int rc = configure_kernel(ctx, (const void*)k_prep12, PREP12_SMEM, false);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_LINEARIZE);
    k_prep12<<<(unsigned)((recs + PREP12_RECS - 1) / PREP12_RECS), PREP12_THREADS, PREP12_SMEM, ctx->stream>>>(a, prep_buf);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  auto kern = k_ric12_wpi<1, 8>;
  const size_t smem = ric12_wpi_smem_bytes(1);
  if ((rc = configure_kernel(ctx, (const void*)kern, smem, true))) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<(unsigned)a.batch, 32, smem, ctx->stream>>>(a, prep_buf);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

template <int MODE>
int launch_ric12_layout(noc_ctx* ctx, int layout, const Ric12Args& a) {
  return layout == NOC_LAYOUT_A_THEN_B ? launch_ric12<0, MODE>(ctx, a) : launch_ric12<1, MODE>(ctx, a);
}

template <int MODE> struct Ric24Cfg { static constexpr int warps = 8; };
template <> struct Ric24Cfg<RIC_AFFINE> { static constexpr int warps = 7; };

template <int LAYOUT, int MODE, bool FUSED, int W, bool BOX = false>
int launch_ric24_cfg(noc_ctx* ctx, const Ric24Args& a, int residency = 0) {
  auto kern = k_ric24<LAYOUT, MODE, FUSED, W, BOX>;
  const size_t need = ric24_smem_bytes<MODE, FUSED, BOX>(W);
  const size_t smem = residency > 0 ? smem_for_residency(residency, need) : need;
  int rc = configure_kernel(ctx, (const void*)kern, residency > 0 ? kSmemMaxDynamic : need, true);
  if (rc) return rc;
  const long long per_cta = 2LL * W;
  const unsigned grid = (unsigned)((a.batch + per_cta - 1) / per_cta);
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, W * 32, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

// The n=24 recursion with its products on the FP64 tensor cores (riccati24_mma.cuh): one instance per warp, A_THEN_B
// records or the fused linearisation, bit-identical to k_ric24.  Two CTAs of six warps per SM; work lists that do not
// fill the GPU run as one-warp CTAs (breadth-first over the SMs) with the residency cap of the wave plan.
constexpr int kRic24MmaWarps = 6, kRic24MmaCtas = 2;
template <int MODE, bool FUSED, int W, int CTAS, bool BOX, bool PREP, bool PERW = false>
int launch_ric24_mma_cfg(noc_ctx* ctx, const Ric24Args& a, const double* prep, int residency = 0) {
  auto kern = k_ric24_mma<MODE, FUSED, W, CTAS, BOX, PREP, PERW>;
  const size_t need = ric24_mma_smem_bytes<MODE, FUSED, BOX, PREP, PERW>(W);
  const size_t smem = residency > 0 ? smem_for_residency(residency, need) : need;
  int rc = configure_kernel(ctx, (const void*)kern, residency > 0 ? kSmemMaxDynamic : need, true);
  if (rc) return rc;
  const unsigned grid = (unsigned)((a.batch + W - 1) / W);
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, W * 32, smem, ctx->stream>>>(a, prep);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}
template <int MODE, bool FUSED, bool BOX, bool PREP>
int launch_ric24_mma_shape(noc_ctx* ctx, const Ric24Args& a, const double* prep) {
  const int wide = kRic24MmaWarps * kRic24MmaCtas;
  int w = 0;
  static const double narrow_cost[8] = {1.0, 1.01, 1.03, 1.06, 1.27, 1.47, 1.50, 1.52};
  if (MODE == RIC_AFFINE && !ctx->opt[NOC_OPT_NO_WAVE_PLAN])
    w = plan_waves_table(a.batch, ctx->sm_count, wide, BOX ? 1.9 : 2.2, narrow_cost, 8);
  else if (a.batch <= (long long)wide * ctx->sm_count)   // one launch of a small batch: spread it, at most 8 one-warp CTAs per SM
    w = (a.batch <= 8LL * ctx->sm_count) ? (int)((a.batch + ctx->sm_count - 1) / ctx->sm_count) : 0;
  if (w > 0) return launch_ric24_mma_cfg<MODE, FUSED, 1, 1, BOX, PREP>(ctx, a, prep, w);
  return launch_ric24_mma_cfg<MODE, FUSED, kRic24MmaWarps, kRic24MmaCtas, BOX, PREP>(ctx, a, prep);
}
// per-instance weights (a.W = the QRQf array, a.w_stride = 1216): each warp stages its instance's W' tiles
int launch_ric24_mma_perw(noc_ctx* ctx, const Ric24Args& a) {
  if (a.batch <= 8LL * ctx->sm_count)
    return launch_ric24_mma_cfg<RIC_LQR, false, 1, 1, false, false, true>(ctx, a, nullptr, (int)((a.batch + ctx->sm_count - 1) / ctx->sm_count));
  return launch_ric24_mma_cfg<RIC_LQR, false, kRic24MmaWarps, kRic24MmaCtas, false, false, true>(ctx, a, nullptr);
}
// `prep_buf` (RIC_AFFINE): room for batch x N x M24_PREP_DOUBLES doubles — the step records k_prep24 leaves for the sweep
template <int MODE, bool FUSED, bool BOX>
int launch_ric24_mma(noc_ctx* ctx, const Ric24Args& a, double* prep_buf = nullptr) {
  if constexpr (MODE == RIC_AFFINE) {
    {
      const long long pairs = a.batch * a.N;
      int rc = configure_kernel(ctx, (const void*)k_prep24, PREP24_SMEM, false);
      if (rc) return rc;
      {
        LaunchScope ls(ctx, NOC_KERNEL_LINEARIZE);
        k_prep24<<<(unsigned)((pairs + PREP24_PAIRS - 1) / PREP24_PAIRS), PREP24_THREADS, PREP24_SMEM, ctx->stream>>>(a, prep_buf);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      return launch_ric24_mma_shape<MODE, FUSED, BOX, true>(ctx, a, prep_buf);
    }
  } else {
    return launch_ric24_mma_shape<MODE, FUSED, BOX, false>(ctx, a, nullptr);
  }
}

template <int LAYOUT, int MODE, bool FUSED = false, bool BOX = false>
int launch_ric24(noc_ctx* ctx, const Ric24Args& a, double* prep_buf = nullptr) {
  if constexpr (LAYOUT == 0 && (MODE == RIC_LQR || FUSED)) {
    if (!ctx->opt[NOC_OPT_NO_TENSOR_SWEEP] && (MODE != RIC_AFFINE || prep_buf)) return launch_ric24_mma<MODE, FUSED, BOX>(ctx, a, prep_buf);
  }
  if (MODE == RIC_AFFINE && !ctx->opt[NOC_OPT_NO_WAVE_PLAN]) {   // (n=24: 0.62 -> 0.90 ms from one to seven warps per SM at N=100)
    const int w = plan_waves((a.batch + 1) / 2, ctx->sm_count, Ric24Cfg<MODE>::warps, ric24_smem_bytes<MODE, FUSED, BOX>(1), 0.075);
    if (w > 0) return launch_ric24_cfg<LAYOUT, MODE, FUSED, 1, BOX>(ctx, a, w);
    return launch_ric24_cfg<LAYOUT, MODE, FUSED, Ric24Cfg<MODE>::warps, BOX>(ctx, a);
  }
  // short work lists: one warp (two instances) per CTA so that the latency-bound tail spreads over all SMs
  if (a.batch <= 2LL * 2 * ctx->sm_count) return launch_ric24_cfg<LAYOUT, MODE, FUSED, 1, BOX>(ctx, a);
  return launch_ric24_cfg<LAYOUT, MODE, FUSED, Ric24Cfg<MODE>::warps, BOX>(ctx, a);
}

// packs the weights for the n=24 kernels into scratch: W[1024] | Qf[576]
int build_w24(noc_ctx* ctx, const double* dQRQf, const double** W, const double** Qf, bool dare = false) {
  void* w = nullptr;
  int rc = scratch(ctx, SCR_W, sizeof(double) * (32 * 32 + 576), &w);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_SETUP);
    k_build_w<<<1, 256, 0, ctx->stream>>>(dQRQf, (double*)w, 24, 8, dare ? 0 : 24 * 24 + 8 * 8);
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

// Time-parallel sweep (NOC_FLAG_TIME_PARALLEL, riccati_scan.cuh): one CTA per instance, its warps own chunks of the
// horizon.  The chunk plan minimises the depth of the dependent chain under a three-constant cost model (microseconds
// per phase-A step, phase-C step and boundary solve of a lone warp, measured on B200: tools/latency_probe.py): chunks
// 0..C-2 take `len` steps, the last one — which only runs plain steps while the others also fold their interval
// elements — takes the rest.
void plan_time_parallel(int N, int* chunks, int* len) {
  const double tA = 3.5, tC = 1.15, tB = 4.8;   // measured with clock64 in the kernel (eight warps of a CTA on one SM), microseconds
  double best = 1e300;
  *chunks = 1; *len = N;
  for (int C = 1; C <= TP_MAX_CHUNKS; ++C) {
    if (C == 1) { const double c1 = N * tC; if (c1 < best) { best = c1; *chunks = 1; *len = N; } continue; }
    for (int L = 1; (C - 1) * L < N; ++L) {
      const int last = N - (C - 1) * L;
      const double cost = std::max(L * tA, last * tC) + (C - 1) * tB + L * tC;
      if (cost < best) { best = cost; *chunks = C; *len = L; }
    }
  }
}

// `bitwise`: the plain warp-per-instance recursion (one chunk, DFMA chains) — every chain in the spec's order, so the
// results are bit-identical to k_ric12 and the oracle; used for small batches without the caller asking (lqr12_device).
int launch_ric12_tp(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* dAB, const double* W, const double* Qf,
                    const double* dx0, double* dK, double* dK0, double* dP0, double* dcost, int32_t* dstatus, bool bitwise = false) {
  RicTpArgs a{};
  a.AB = dAB; a.W = W; a.Qf = Qf; a.x0 = dx0; a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
  a.batch = batch; a.N = p->N;
  plan_time_parallel(p->N, &a.chunks, &a.len);
  bool dfma = ctx->opt[NOC_OPT_TIME_PARALLEL_DFMA] != 0;   // comparison build of the same kernel without tensor cores
  if (bitwise) { a.chunks = 1; a.len = p->N; }   // (DMMA evaluates the same ascending fma chains: tools/ubench/dmma_order.cu)
  else if (ctx->opt[NOC_OPT_TIME_PARALLEL_CHUNKS] > 0) {          // forced plan (measurements): C chunks of equal length
    a.chunks = (int)std::min<int64_t>(ctx->opt[NOC_OPT_TIME_PARALLEL_CHUNKS], std::min<int64_t>(TP_MAX_CHUNKS, p->N));
    a.len = a.chunks > 1 ? std::max(1, p->N / a.chunks) : p->N;
  }
  const size_t smem_c = ric_tp_smem_bytes(a.chunks);
  int rc = dfma ? configure_kernel(ctx, (const void*)k_ric12_tp<false>, ric_tp_smem_bytes(TP_MAX_CHUNKS), false)
                : configure_kernel(ctx, (const void*)k_ric12_tp<true>, ric_tp_smem_bytes(TP_MAX_CHUNKS), false);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    if (dfma) k_ric12_tp<false><<<(unsigned)batch, a.chunks * 32, smem_c, ctx->stream>>>(a);
    else k_ric12_tp<true><<<(unsigned)batch, a.chunks * 32, smem_c, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
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
  if (p->flags & NOC_FLAG_TIME_PARALLEL) {
    if (p->layout != NOC_LAYOUT_A_THEN_B) return fail(ctx, NOC_ERR_UNSUPPORTED, "NOC_FLAG_TIME_PARALLEL: A_THEN_B records only");
    return launch_ric12_tp(ctx, p, batch, dAB, a.W, a.Qf, dx0, dK, dK0, dP0, dcost, dstatus);
  }
  // Small batches: the sweep is a latency chain whatever the mapping, and a warp that owns ONE instance steps in about
  // half the time of a warp that owns eight (1,700 vs 3,600 clk).  Same chains, same order, same bits.
  if (p->layout == NOC_LAYOUT_A_THEN_B && batch <= 2LL * ctx->sm_count && !ctx->opt[NOC_OPT_NO_SMALL_BATCH_KERNEL])
    return launch_ric12_tp(ctx, p, batch, dAB, a.W, a.Qf, dx0, dK, dK0, dP0, dcost, dstatus, true);
  if (p->layout == NOC_LAYOUT_A_THEN_B && !ctx->opt[NOC_OPT_NO_TENSOR_SWEEP]) return launch_ric12_mma(ctx, a);
  return launch_ric12_layout<RIC_LQR>(ctx, p->layout, a);
}

// device-pointer core of noc_lqr_solve_batch: tuned kernels for (12,4) / (24,8) with shared weights, the
// thread-per-instance kernel for everything else
int lqr_batch_device(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, bool generic, const double* dAB,
                     const double* dQ, int qr_per_instance, const double* dx0, double* dK, double* dK0, double* dP0,
                     double* dcost, int32_t* dstatus) {
  if ((p->flags & NOC_FLAG_TIME_PARALLEL) && (generic || qr_per_instance || p->n != 12))
    return fail(ctx, NOC_ERR_UNSUPPORTED, "NOC_FLAG_TIME_PARALLEL: n=12, m=4 with shared weights only");
  if (qr_per_instance && !generic && p->n == 24) {   // (24,8), A_THEN_B records: each warp stages its instance's W' tiles
    Ric24Args a{};
    a.AB = dAB; a.W = dQ; a.w_stride = 2 * 24 * 24 + 8 * 8; a.x0 = dx0;
    a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
    a.batch = batch; a.N = p->N;
    return launch_ric24_mma_perw(ctx, a);
  }
  if (qr_per_instance && !generic) {   // (12,4), A_THEN_B records: the tensor-core sweep reads each instance's own Q | R | Qf
    Ric12Args a{};
    a.AB = dAB; a.W = dQ; a.w_stride = 2 * 12 * 12 + 4 * 4; a.x0 = dx0;
    a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
    a.batch = batch; a.N = p->N;
    return launch_ric12_mma_perw(ctx, a);
  }
  if (generic) {
    RicGenArgs g{};
    g.AB = dAB; g.QRQf = dQ; g.x0 = dx0; g.K = dK; g.K0 = dK0; g.P0 = dP0; g.cost = dcost; g.status = dstatus;
    g.batch = batch; g.n = p->n; g.m = p->m; g.N = p->N; g.layout = p->layout;
    g.qr_stride = qr_per_instance ? (long long)(2 * p->n * p->n + p->m * p->m) : 0;
    {
      LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
      k_ric_generic<<<(unsigned)((batch + 63) / 64), 64, 0, ctx->stream>>>(g);
    }
    return (cudaGetLastError() == cudaSuccess) ? NOC_OK : fail(ctx, NOC_ERR_CUDA, "k_ric_generic launch failed");
  }
  return (p->n == 24 ? lqr24_device : lqr12_device)(ctx, p, batch, dAB, dQ, dx0, dK, dK0, dP0, dcost, dstatus);
}

// Host-resident A_k,B_k (and outputs): two device buffers; chunk c+1 is copied in on the auxiliary stream while
// chunk c is swept on the caller's stream and chunk c-1's outputs are copied out on the copy stream.  The chunk
// size is bounded by the scratch budget (NOC_OPT_SCRATCH_LIMIT_BYTES or 60 % of the free HBM), so a stream larger
// than the GPU's memory is processed instead of failing with NOC_ERR_OOM.  Outputs that are DEVICE pointers are
// written in place.
int lqr_batch_host_pipeline(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, bool generic, const double* AB,
                            const double* QRQf, int qr_per_instance, const double* x0dev, double* K, double* K0,
                            double* P0, double* cost, int32_t* status) {
  int rc;
  const size_t n = p->n, m = p->m, N = p->N, rec = n * n + n * m, B = (size_t)batch;
  const size_t qr_len = 2 * n * n + m * m;
  struct Out { void* user; size_t per; bool host; void* dev[2]; };
  Out outs[5] = {{K, sizeof(double) * N * m * n, false, {nullptr, nullptr}}, {K0, sizeof(double) * m * n, false, {nullptr, nullptr}},
                 {P0, sizeof(double) * n * n, false, {nullptr, nullptr}}, {cost, sizeof(double), false, {nullptr, nullptr}},
                 {status, sizeof(int32_t), false, {nullptr, nullptr}}};
  size_t per_inst = sizeof(double) * N * rec;   // bytes of device scratch per instance and buffer
  const bool x0_host = x0dev && !is_device_ptr(x0dev), qr_host = !is_device_ptr(QRQf);
  if (x0_host) per_inst += sizeof(double) * n;
  if (qr_per_instance && qr_host) per_inst += sizeof(double) * qr_len;
  for (auto& o : outs) { o.host = o.user && !is_device_ptr(o.user); if (o.host) per_inst += o.per; }
  size_t have = 0, budget = 0;
  for (auto& b : ctx->scratch) have += b.cap;   // regrown below if too small: count what is already ours
  if ((rc = scratch_budget(ctx, BUDGET_HOST_AB, (int)n, (int)N, (int)(per_inst & 0x7fffffff), batch, have, &budget))) return rc;
  // about eight chunks per call so that the first copy-in and the last copy-out (which nothing hides) stay short
  size_t chunk = std::min<size_t>(std::max<size_t>(B / 8, 256), std::max<size_t>(1, budget / (2 * per_inst)));
  chunk = std::min(chunk, B);
  if (chunk >= 64) chunk &= ~size_t(31);
  void *dab[2], *dx0c[2] = {nullptr, nullptr}, *dqc[2] = {nullptr, nullptr};
  for (int s2 = 0; s2 < 2; ++s2) {
    if ((rc = scratch(ctx, SCR_PIPE_AB0 + s2, sizeof(double) * chunk * N * rec, &dab[s2]))) return rc;
    if (x0_host && (rc = scratch(ctx, SCR_PIPE_X0 + s2, sizeof(double) * chunk * n, &dx0c[s2]))) return rc;
    if (qr_per_instance && qr_host && (rc = scratch(ctx, SCR_PIPE_QR0 + s2, sizeof(double) * chunk * qr_len, &dqc[s2]))) return rc;
    for (int o = 0; o < 5; ++o)
      if (outs[o].host && (rc = scratch(ctx, SCR_PIPE_OUT0 + 2 * o + s2, outs[o].per * chunk, &outs[o].dev[s2]))) return rc;
  }
  const void* dQshared = nullptr;
  if (!qr_per_instance && (rc = stage_in(ctx, 1, QRQf, sizeof(double) * qr_len, &dQshared))) return rc;
  cudaStream_t const main_stream = ctx->stream, in_stream = ctx->aux_stream, out_stream = ctx->copy_stream;
  // the pipeline's events order three streams; whatever happens below, drain them before the buffers can be reused
  struct Drain { noc_ctx* c; ~Drain() { cudaStreamSynchronize(c->aux_stream); cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream); } } drain{ctx};
  NOC_CUDA(ctx, cudaEventRecord(ctx->fork_ev, main_stream));   // weights staged / earlier work of the caller's stream
  NOC_CUDA(ctx, cudaStreamWaitEvent(in_stream, ctx->fork_ev, 0));
  size_t c = 0;
  for (size_t off = 0; off < B; off += chunk, ++c) {
    const int s2 = (int)(c & 1);
    const size_t cb = std::min(chunk, B - off);
    cudaEvent_t ev_in = ctx->pipe_ev[s2], ev_swept = ctx->pipe_ev[2 + s2], ev_out = ctx->pipe_ev[4 + s2];
    if (c >= 2) NOC_CUDA(ctx, cudaStreamWaitEvent(in_stream, ev_swept, 0));   // the sweep of chunk c-2 has consumed buffer s2
    NOC_CUDA(ctx, cudaMemcpyAsync(dab[s2], AB + off * N * rec, sizeof(double) * cb * N * rec, cudaMemcpyHostToDevice, in_stream));
    if (x0_host) NOC_CUDA(ctx, cudaMemcpyAsync(dx0c[s2], x0dev + off * n, sizeof(double) * cb * n, cudaMemcpyHostToDevice, in_stream));
    if (dqc[s2]) NOC_CUDA(ctx, cudaMemcpyAsync(dqc[s2], QRQf + off * qr_len, sizeof(double) * cb * qr_len, cudaMemcpyHostToDevice, in_stream));
    NOC_CUDA(ctx, cudaEventRecord(ev_in, in_stream));
    NOC_CUDA(ctx, cudaStreamWaitEvent(main_stream, ev_in, 0));
    if (c >= 2) NOC_CUDA(ctx, cudaStreamWaitEvent(main_stream, ev_out, 0));   // chunk c-2's outputs have left buffer s2
    auto dst = [&](int o) -> void* {
      if (!outs[o].user) return nullptr;
      return outs[o].host ? outs[o].dev[s2] : (void*)((char*)outs[o].user + off * outs[o].per);
    };
    const double* dq = qr_per_instance ? (dqc[s2] ? (const double*)dqc[s2] : QRQf + off * qr_len) : (const double*)dQshared;
    const double* dx = x0dev ? (x0_host ? (const double*)dx0c[s2] : x0dev + off * n) : nullptr;
    rc = lqr_batch_device(ctx, p, (int64_t)cb, generic, (const double*)dab[s2], dq, qr_per_instance, dx, (double*)dst(0),
                          (double*)dst(1), (double*)dst(2), (double*)dst(3), (int32_t*)dst(4));
    if (rc) return rc;
    NOC_CUDA(ctx, cudaEventRecord(ev_swept, main_stream));
    NOC_CUDA(ctx, cudaStreamWaitEvent(out_stream, ev_swept, 0));
    for (auto& o : outs)
      if (o.host)
        NOC_CUDA(ctx, cudaMemcpyAsync((char*)o.user + off * o.per, o.dev[s2], cb * o.per, cudaMemcpyDeviceToHost, out_stream));
    NOC_CUDA(ctx, cudaEventRecord(ev_out, out_stream));
  }
  NOC_CUDA(ctx, cudaStreamSynchronize(in_stream));
  NOC_CUDA(ctx, cudaStreamSynchronize(out_stream));
  NOC_CUDA(ctx, cudaStreamSynchronize(main_stream));
  return NOC_OK;
}

// NOC_FLAG_TIME_PARALLEL on the from-x0 / along entry points (small batches): nominal rollout, time-parallel
// linearisation into a record scratch, time-parallel sweep — three launches, each a short dependent chain.
int from_x0_time_parallel(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                          const double* QRQf, double* xbar_out, double* K, double* K0, double* P0, double* cost,
                          int32_t* status, bool bitwise) {
  int rc;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = 12, m = 4, N = p->N;
  if (sizeof(double) * B * N * R12_REC > ((size_t)8 << 30))
    return fail(ctx, NOC_ERR_UNSUPPORTED, "NOC_FLAG_TIME_PARALLEL is the small-batch latency path (record scratch would exceed 8 GB)");
  const void *dx0, *dQ, *dub;
  void *dK, *dK0, *dP0, *dcost, *dstatus, *dxout, *dxbar, *dAB;
  if ((rc = stage_in(ctx, 0, x0, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_in(ctx, 1, QRQf, sizeof(double) * (2 * n * n + m * m), &dQ))) return rc;
  if ((rc = stage_in(ctx, 2, ubar, sizeof(double) * B * N * m, &dub))) return rc;
  if ((rc = stage_out(ctx, 5, xbar_out, sizeof(double) * B * (N + 1) * n, &dxout))) return rc;
  if ((rc = stage_out(ctx, 0, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 1, K0, sizeof(double) * B * m * n, &dK0))) return rc;
  if ((rc = stage_out(ctx, 2, P0, sizeof(double) * B * n * n, &dP0))) return rc;
  if ((rc = stage_out(ctx, 3, cost, sizeof(double) * B, &dcost))) return rc;
  if ((rc = stage_out(ctx, 4, status, sizeof(int32_t) * B, &dstatus))) return rc;
  dxbar = dxout;
  if (!dxbar && (rc = scratch(ctx, SCR_XBAR, sizeof(double) * B * (N + 1) * n, &dxbar))) return rc;
  if ((rc = scratch(ctx, SCR_AB, sizeof(double) * B * N * R12_REC, &dAB))) return rc;
  const double *W, *Qf;
  if ((rc = build_w12(ctx, (const double*)dQ, false, &W, &Qf))) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
    k_rollout_open_loop<0><<<(unsigned)((B + 31) / 32), 32, 0, ctx->stream>>>((const double*)dx0, (const double*)dub, (double*)dxbar, (long long)B, p->N,
                                                                             p->dt, (long long)((N + 1) * n), (long long)n);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  {
    const long long total = (long long)B * p->N;
    LaunchScope ls(ctx, NOC_KERNEL_LINEARIZE);
    k_linearize<0><<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>((const double*)dxbar, (const double*)dub, (double*)dAB, (long long)B, p->N,
                                                                           p->dt, NOC_LAYOUT_A_THEN_B, nullptr);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  if ((rc = launch_ric12_tp(ctx, p, batch, (const double*)dAB, W, Qf, (const double*)dx0, (double*)dK, (double*)dK0, (double*)dP0,
                            (double*)dcost, (int32_t*)dstatus, bitwise)))
    return rc;
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

template <int MODEL>
int from_x0_impl(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                 const double* QRQf, double* xbar_out, double* K, double* K0, double* P0, double* cost,
                 int32_t* status) {
  int rc;
  if (p->flags & NOC_FLAG_TIME_PARALLEL) {
    if (MODEL != 0) return fail(ctx, NOC_ERR_UNSUPPORTED, "NOC_FLAG_TIME_PARALLEL: n=12, m=4 only");
    return from_x0_time_parallel(ctx, p, batch, x0, ubar, QRQf, xbar_out, K, K0, P0, cost, status, false);
  }
  // small batches of the quadrotor: rollout -> time-parallel linearisation -> one instance per warp (bit-identical to the
  // fused 8-per-warp sweep: the records are the same bits and the pruned products are exact zeros)
  if (MODEL == 0 && batch <= 2LL * ctx->sm_count && !ctx->opt[NOC_OPT_NO_SMALL_BATCH_KERNEL])
    return from_x0_time_parallel(ctx, p, batch, x0, ubar, QRQf, xbar_out, K, K0, P0, cost, status, true);
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
  if (misaligned16(dK) || misaligned16(dK0) || misaligned16(dP0))
    return fail(ctx, NOC_ERR_BAD_ARG, "device K / K0 / P0 must be 16-byte aligned");
  // Both built-in models linearise inside the sweep (k_ric12 / k_ric24 <…, FUSED>): only the nominal trajectory is
  // staged in scratch, no A_k,B_k stream exists.  Batch scheduler: an OUTER chunk bounds the scratch to a share of
  // the free HBM and gets one rollout launch (latency-bound: one launch for everything costs as much as one for a
  // quarter); with host-resident outputs it is cut into INNER chunks so that the device->host copy of chunk i runs
  // on the copy stream under the sweep of chunk i+1.
  size_t have = 0, budget = 0;
  if (ctx->scratch.size() > SCR_XBAR) have = ctx->scratch[SCR_XBAR].cap;
  if ((rc = scratch_budget(ctx, BUDGET_FROM_X0, (int)n, (int)N, 0, batch, have, &budget))) return rc;
  const size_t per_inst = sizeof(double) * (N + 1) * n;
  size_t outer = std::max<size_t>(1, std::min<size_t>(B, budget / per_inst));
  if (outer < B) outer = std::max<size_t>(32, outer & ~size_t(31));
  if (dxout) outer = B;   // the caller's xbar array (instance-major) holds the nominal: no scratch, no outer chunks
  const bool overlap = !ctx->pending.empty() && B >= 8192 && !(p->flags & NOC_FLAG_ASYNC);
  SideStreamDrain drain{ctx, overlap};
  // inner chunks are whole waves of the sweep kernel (instances resident on all SMs at once: 64 per SM for n=12,
  // 16 for n=24), about eight per call: a chunk that ends mid-wave pays for the full wave, and the smaller the last
  // chunk, the shorter the copy that nothing hides (B = 65536: 4 x 16384 = 8 waves took 2.53 ms of sweeps, 7 chunks
  // of one wave take 2.2 ms)
  // instances resident per SM: n=12 eight per warp x eight warps; n=24 one per warp x twelve warps (k_ric24_mma) or two x eight
  const size_t wave = (size_t)ctx->sm_count * (MODEL == 0 ? 64 : (ctx->opt[NOC_OPT_NO_TENSOR_SWEEP] ? 16 : 12));
  size_t inner_pref = overlap ? std::max<size_t>(1, (B / wave + 4) / 8) * wave : outer;
  if (overlap && ctx->opt[NOC_OPT_CHUNK_WAVES] > 0) inner_pref = (size_t)ctx->opt[NOC_OPT_CHUNK_WAVES] * wave;
  void* dxbar = nullptr;
  if (!dxout && (rc = scratch(ctx, SCR_XBAR, sizeof(double) * outer * (N + 1) * n, &dxbar))) return rc;
  (void)rec;
  typename std::conditional<MODEL == 0, Ric12Args, Ric24Args>::type fa{};
  if (MODEL == 0) rc = build_w12(ctx, (const double*)dQ, false, &fa.W, &fa.Qf);
  else rc = build_w24(ctx, (const double*)dQ, &fa.W, &fa.Qf);
  if (rc) return rc;
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
    const bool two_streams = overlap && inner < ob && !ctx->opt[NOC_OPT_SINGLE_STREAM];
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
      if (rc) return rc;
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
int slq_chunk(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* xgoal,
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
  void* ddelta;
  if ((rc = scratch(ctx, SCR_DELTA, sizeof(double) * B, &ddelta))) return rc;
  if ((rc = scratch(ctx, SCR_BSTATUS, sizeof(int32_t) * B, &dbst))) return rc;
  if ((rc = scratch(ctx, SCR_IDX0, sizeof(int32_t) * B, &didx0))) return rc;
  if ((rc = scratch(ctx, SCR_IDX1, sizeof(int32_t) * B, &didx1))) return rc;
  if ((rc = scratch(ctx, SCR_COUNT, 64, &dcount))) return rc;
  void *dxn, *dun, *dretry0, *dretry1;
  if ((rc = scratch(ctx, SCR_RETRY0, sizeof(int32_t) * B, &dretry0))) return rc;
  if ((rc = scratch(ctx, SCR_RETRY1, sizeof(int32_t) * B, &dretry1))) return rc;
  if ((rc = scratch(ctx, SCR_XN, sizeof(double) * B * (N + 1) * n, &dxn))) return rc;
  if ((rc = scratch(ctx, SCR_UN, sizeof(double) * B * N * m, &dun))) return rc;
  void *dlc, *dflags, *dprep = nullptr;
  // n=24 on the tensor cores: the step records of k_prep24 (Jacobian entries, W e, ubar), one per (work-list slot, k)
  if (MODEL == 1 && !ctx->opt[NOC_OPT_NO_TENSOR_SWEEP] && (rc = scratch(ctx, SCR_PREP, sizeof(double) * B * N * M24_PREP_DOUBLES, &dprep))) return rc;
  // n=12: short work lists (at most twelve instances per SM) go to the warp-per-instance tensor kernel, which takes step records too
  const size_t prep12_cap = (MODEL == 0 && !ctx->opt[NOC_OPT_NO_TENSOR_SWEEP]) ? std::min<size_t>(B, (size_t)kRic12WpiMaxPerSm * ctx->sm_count) : 0;
  if (prep12_cap && (rc = scratch(ctx, SCR_PREP, sizeof(double) * prep12_cap * N * W12_PREP_DOUBLES, &dprep))) return rc;
  if ((rc = scratch(ctx, SCR_LC, sizeof(double) * B * (N + 1), &dlc))) return rc;
  if ((rc = scratch(ctx, SCR_FLAGS, sizeof(int32_t) * B, &dflags))) return rc;
  NOC_CUDA(ctx, cudaMemsetAsync(dflags, 0, sizeof(int32_t) * B, ctx->stream));
  using FC = FwdCfg<MODEL>;
  const size_t fwd_smem = FC::ring_bytes;
  if ((rc = configure_kernel(ctx, (const void*)k_slq_rollout<MODEL>, fwd_smem, false))) return rc;
  const size_t wbytes = sizeof(double) * (2 * n * n + m * m);
  {
    LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
    k_slq_init<MODEL><<<(unsigned)((batch + 127) / 128), 128, wbytes, ctx->stream>>>(
        (const double*)dx0, (const double*)dxg, (const double*)dQ, (const double*)dui, (double*)dx, (double*)du, (double*)dJ, (double*)dmu,
        (double*)ddelta, (int32_t*)dst, (int32_t*)dit, (int32_t*)dai, (int32_t*)didx0, batch, p->N, p->dt, p->reg0, p->max_iter);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  typename std::conditional<MODEL == 0, Ric12Args, Ric24Args>::type ra{};
  if (MODEL == 0) rc = build_w12(ctx, (const double*)dQ, false, &ra.W, &ra.Qf);
  else rc = build_w24(ctx, (const double*)dQ, &ra.W, &ra.Qf);
  if (rc) return rc;
  const bool boxed = std::isfinite(p->u_min) || std::isfinite(p->u_max);   // input box: control-limited backward pass
  int count = (int)batch;
  int32_t* cur = (int32_t*)didx0;
  int32_t* nxt = (int32_t*)didx1;
  // Deferred backtracking (opt-in, NOC_OPT_DEFERRED_BACKTRACK: measured SLOWER, see the end of this comment).  The instances whose full step was rejected (1-2 % per iteration without input limits) get
  // all their shorter step lengths as parallel candidates; that round is a latency-bound launch set (0.2-0.7 ms) the
  // whole batch used to wait for.  It now runs on the auxiliary stream while the main stream goes on with the next
  // backward sweep of everybody else; the settled instances rejoin the work list one sweep later (they carry their own
  // iteration index, iters[b], so records and limits are per instance and the results do not depend on the schedule).
  // `pending_bit`: the flag bit the round in flight marks its continuing instances with (alternating 4 / 8).
  // Measured on config 3 (16,384 x N=200): 70.1 ms synchronous; 87.5 ms deferring throughout (75 sweeps instead of 50: a
  // straggler that backtracks 25 times finishes 25 latency-bound sweeps later); 76.0 ms deferring only while the list
  // fills a wide wave (59 sweeps).  The hidden round is worth less than the sweeps it adds, so the default stays
  // synchronous; results are identical either way (tests/test_gpu_dare_slq_parity.py runs both).
  const bool defer_ok = ctx->opt[NOC_OPT_DEFERRED_BACKTRACK] != 0;
  // Speculative line search (default; NOC_OPT_NO_SPECULATIVE_BACKTRACK switches it off).  The instances that needed a shorter
  // step in their previous iteration — in the tail of a solve the same two or three stragglers, iteration after iteration —
  // used to cost every iteration a second, latency-bound launch set (candidate rollouts + costs + pick: 0.15-0.2 ms while the
  // GPU idles).  They are flagged instead (SLQ_FLAG_SPEC_NEXT), and in their next forward pass ALL their step lengths,
  // alpha = 1 included, run as candidates on the auxiliary stream beside everybody else's trial 0; the first passing
  // candidate in order is exactly the oracle's sequential search.  Only short lists are speculated (kSpecMax).
  constexpr int kSpecMax = 64;
  const bool spec_ok = !defer_ok && !ctx->opt[NOC_OPT_NO_SPECULATIVE_BACKTRACK] && p->n_alpha <= 32;
  int spec_count = 0;
  void *dslist = nullptr, *dsx = nullptr, *dsu = nullptr, *dslc = nullptr;
  if (spec_ok) {
    const size_t cands = (size_t)kSpecMax * p->n_alpha;
    if ((rc = scratch(ctx, SCR_SPEC_LIST, sizeof(int32_t) * B, &dslist))) return rc;
    if ((rc = scratch(ctx, SCR_SPEC_X, sizeof(double) * cands * (N + 1) * n, &dsx))) return rc;
    if ((rc = scratch(ctx, SCR_SPEC_U, sizeof(double) * cands * N * m, &dsu))) return rc;
    if ((rc = scratch(ctx, SCR_SPEC_LC, sizeof(double) * cands * (N + 1), &dslc))) return rc;
  }
  int pending_bit = 0;
  cudaEvent_t pending_ev = nullptr;
  cudaStream_t const main_stream = ctx->stream;
  struct RestoreStream { noc_ctx* c; cudaStream_t s; ~RestoreStream() { c->stream = s; cudaStreamSynchronize(c->aux_stream); } } restore{ctx, main_stream};
  const int sweep_limit = p->max_iter * (p->n_alpha + 1) + 2;   // (a bound, never reached: every sweep advances somebody)
  for (int it = 0; it < sweep_limit && (count > 0 || pending_bit); ++it) {
   int new_pending_bit = 0;
   cudaEvent_t new_pending_ev = nullptr;
   if (count > 0) {
    // a1 happens inside the sweep (k_ric12 / k_ric24 <…, RIC_AFFINE, FUSED>)
    // a2+a3: affine backward pass (regularisation retries in-kernel)
    ra.AB = (const double*)dAB; ra.idx = cur; ra.xbar = (const double*)dx; ra.ubar = (const double*)du;
    ra.xgoal = (const double*)dxg; ra.K = (double*)dK; ra.kff = (double*)dkff; ra.dV = (double*)ddV;
    ra.mu = (double*)dmu; ra.delta = (double*)ddelta; ra.reg_schedule = p->reg_schedule;
    ra.status = (int32_t*)dbst; ra.batch = count; ra.N = p->N;
    ra.dt = p->dt;
    if constexpr (MODEL == 0) {
      ra.u_min = p->u_min; ra.u_max = p->u_max;
      if (!boxed && dprep && (size_t)count <= prep12_cap) rc = launch_ric12_wpi(ctx, ra, (double*)dprep);
      else rc = boxed ? launch_ric12<0, RIC_AFFINE, true, true>(ctx, ra) : launch_ric12<0, RIC_AFFINE, true>(ctx, ra);
    }
    else {
      ra.u_min = p->u_min; ra.u_max = p->u_max;
      rc = boxed ? launch_ric24<0, RIC_AFFINE, true, true>(ctx, ra, (double*)dprep) : launch_ric24<0, RIC_AFFINE, true>(ctx, ra, (double*)dprep);
    }
    if (rc) return rc;
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
    fa.mu = (double*)dmu; fa.delta = (double*)ddelta; fa.reg_schedule = p->reg_schedule;
    fa.cont_bit = 1;
    fa.spec_mark = spec_ok ? 1 : 0;
    const bool spec_now = spec_ok && spec_count > 0 && spec_count <= kSpecMax;
    if (spec_now) {
      // fork: mark the speculated instances (the ordinary trial 0 skips them), then their candidate round on the auxiliary stream
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_mark_spec<<<(spec_count + 127) / 128, 128, 0, ctx->stream>>>((const int32_t*)dslist, spec_count, (int32_t*)dflags, 1);
      }
      NOC_CUDA(ctx, cudaEventRecord(ctx->fork_ev, main_stream));
      NOC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_ev, 0));
      ctx->stream = ctx->aux_stream;   // (the launches below and their profiling events go to the auxiliary stream)
      SlqFwdArgs fs = fa;
      const size_t rc_ = (size_t)spec_count * p->n_alpha;
      fs.spec = 1; fs.idx = (const int32_t*)dslist; fs.trial = 0; fs.cand_J = p->n_alpha; fs.cand_first = 0; fs.cand_last = 1;
      fs.xc = (double*)dsx; fs.uc = (double*)dsu; fs.lcc = (double*)dslc;
      fs.count = (int)rc_;
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_rollout<MODEL><<<(unsigned)((rc_ + FC::inst - 1) / FC::inst), FC::threads, fwd_smem, ctx->stream>>>(fs);
      }
      {
        const long long total = (long long)rc_ * (p->N + 1);
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_stage_cost<MODEL><<<(unsigned)std::min<long long>((total + 127) / 128, 16LL * ctx->sm_count), 128, wbytes, ctx->stream>>>(fs);
      }
      fs.count = spec_count;
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_slq_accept_cand<MODEL><<<(unsigned)((spec_count + 3) / 4), 128, 0, ctx->stream>>>(fs);
      }
      ctx->stream = main_stream;
      NOC_CUDA(ctx, cudaGetLastError());
      NOC_CUDA(ctx, cudaEventRecord(ctx->join_ev, ctx->aux_stream));
    }
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
        k_slq_stage_cost<MODEL><<<(unsigned)std::min<long long>((total + 127) / 128, 16LL * ctx->sm_count), 128, wbytes, ctx->stream>>>(fa);
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
        // (only while the sweeps are throughput-bound, i.e. the work list fills at least one wide wave: a deferred instance
        // finishes one sweep later, and in the latency-bound tail every extra sweep costs what the hidden round saves —
        // deferring throughout took config 3 from 70 to 88 ms, 75 sweeps instead of 50)
        const bool head = (size_t)count >= (size_t)ctx->sm_count * (MODEL == 0 ? 56 : 14);
        if (defer_ok && head && round_J == J && (size_t)list_count * J * per_cand <= ((size_t)2 << 30)) {
          // one round settles every rejected instance: run it beside the next sweep
          void *dxc, *duc, *dlcc;
          const size_t rc_ = (size_t)list_count * J;
          int e1 = scratch(ctx, SCR_XC, sizeof(double) * rc_ * (N + 1) * n, &dxc);
          int e2 = e1 ? e1 : scratch(ctx, SCR_UC, sizeof(double) * rc_ * N * m, &duc);
          int e3 = e2 ? e2 : scratch(ctx, SCR_LCC, sizeof(double) * rc_ * (N + 1), &dlcc);
          if (e3) return e3;
          new_pending_bit = (it & 1) ? 8 : 4;
          new_pending_ev = (it & 1) ? ctx->join_ev : ctx->fork_ev;
          NOC_CUDA(ctx, cudaEventRecord(ctx->defer_ev, main_stream));
          NOC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->defer_ev, 0));
          ctx->stream = ctx->aux_stream;   // the launches below (and their profiling events) go to the auxiliary stream
          fa.idx = list; fa.trial = 1; fa.cand_J = J; fa.cand_first = 1; fa.cand_last = 1;
          fa.xc = (double*)dxc; fa.uc = (double*)duc; fa.lcc = (double*)dlcc;
          fa.count = (int)rc_;
          fa.cont_bit = new_pending_bit;
          {
            LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
            k_slq_rollout<MODEL><<<(unsigned)((rc_ + FC::inst - 1) / FC::inst), FC::threads, fwd_smem, ctx->stream>>>(fa);
          }
          {
            const long long total = (long long)rc_ * (p->N + 1);
            LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
            k_slq_stage_cost<MODEL><<<(unsigned)std::min<long long>((total + 127) / 128, 16LL * ctx->sm_count), 128, wbytes, ctx->stream>>>(fa);
          }
          fa.count = list_count;
          {
            LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
            k_slq_accept_cand<MODEL><<<(unsigned)((list_count + 3) / 4), 128, 0, ctx->stream>>>(fa);
          }
          ctx->stream = main_stream;
          NOC_CUDA(ctx, cudaGetLastError());
          NOC_CUDA(ctx, cudaEventRecord(new_pending_ev, ctx->aux_stream));
          fa.cand_J = 0; fa.cont_bit = 1;
          list_count = 0;
        } else if ((size_t)list_count * round_J * per_cand <= ((size_t)2 << 30)) {
          int first = 1;
          while (list_count > 0 && first < p->n_alpha) {
            const int Jr = std::min(round_J, p->n_alpha - first);
            void *dxc, *duc, *dlcc;
            const size_t rc_ = (size_t)list_count * Jr;
            int e1 = scratch(ctx, SCR_XC, sizeof(double) * rc_ * (N + 1) * n, &dxc);
            int e2 = e1 ? e1 : scratch(ctx, SCR_UC, sizeof(double) * rc_ * N * m, &duc);
            int e3 = e2 ? e2 : scratch(ctx, SCR_LCC, sizeof(double) * rc_ * (N + 1), &dlcc);
            if (e3) return e3;
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
              k_slq_stage_cost<MODEL><<<(unsigned)std::min<long long>((total + 127) / 128, 16LL * ctx->sm_count), 128, wbytes, ctx->stream>>>(fa);
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
    if (spec_now) {   // join: the speculative round is complete; its instances lose their mark
      NOC_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->join_ev, 0));
      LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
      k_slq_mark_spec<<<(spec_count + 127) / 128, 128, 0, ctx->stream>>>((const int32_t*)dslist, spec_count, (int32_t*)dflags, 0);
    }
   }  // count > 0
    int32_t hn = 0;
    int mask = 1;
    if (pending_bit) {   // the round deferred one sweep ago has to be complete before its instances rejoin
      NOC_CUDA(ctx, cudaStreamWaitEvent(main_stream, pending_ev, 0));
      mask |= pending_bit;
    }
    {  // instances that go on to the next iteration -> sorted work list
      LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
      k_compact_flags<<<1, 1024, 0, ctx->stream>>>((int32_t*)dflags, mask, batch, nxt, (int32_t*)dcount);
    }
    NOC_CUDA(ctx, cudaGetLastError());
    int32_t hs = 0;
    if (spec_ok) {   // the instances that backtracked in this iteration and go on -> the next forward pass's speculative list
      {
        LaunchScope ls(ctx, NOC_KERNEL_FORWARD);
        k_compact_flags<<<1, 1024, 0, ctx->stream>>>((int32_t*)dflags, SLQ_FLAG_SPEC_NEXT, batch, (int32_t*)dslist, (int32_t*)dcount + 2);
      }
      NOC_CUDA(ctx, cudaGetLastError());
      NOC_CUDA(ctx, cudaMemcpyAsync(&hs, (int32_t*)dcount + 2, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    NOC_CUDA(ctx, cudaMemcpyAsync(&hn, dcount, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    count = hn;
    spec_count = hs;
    std::swap(cur, nxt);
    pending_bit = new_pending_bit;
    pending_ev = new_pending_ev;
  }
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->aux_stream));
  return flush_out(ctx, false);
}

// Device-resident from-x0 solve in wave-sized chunks with a callback after each chunk's sweep (the multi-GPU plane
// hangs its NVLink gathers there).  One rollout launch for the slice; `chunked` = false sweeps it in one launch.
template <int MODEL, class AfterChunk>
int from_x0_device_chunks(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* dx0, const double* dQ,
                          double* dK0, double* dcost, int32_t* dstatus, bool chunked, AfterChunk&& after_chunk) {
  int rc;
  const size_t B = (size_t)batch, n = model::Dims<MODEL>::n, N = p->N;
  void* dxbar = nullptr;
  if ((rc = scratch(ctx, SCR_XBAR, sizeof(double) * B * (N + 1) * n, &dxbar))) return rc;
  typename std::conditional<MODEL == 0, Ric12Args, Ric24Args>::type fa{};
  if (MODEL == 0) rc = build_w12(ctx, dQ, false, &fa.W, &fa.Qf);
  else rc = build_w24(ctx, dQ, &fa.W, &fa.Qf);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
    k_rollout_open_loop<MODEL><<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(dx0, nullptr, (double*)dxbar, (long long)B, p->N, p->dt,
                                                                                     (long long)n, (long long)(B * n));
  }
  NOC_CUDA(ctx, cudaGetLastError());
  // instances resident per SM: n=12 eight per warp x eight warps; n=24 one per warp x twelve warps (k_ric24_mma) or two x eight
  const size_t wave = (size_t)ctx->sm_count * (MODEL == 0 ? 64 : (ctx->opt[NOC_OPT_NO_TENSOR_SWEEP] ? 16 : 12));
  const size_t inner = chunked ? std::max<size_t>(1, (B / wave + 4) / 8) * wave : B;
  int chunk_no = 0;
  for (size_t off = 0; off < B; off += inner, ++chunk_no) {
    const size_t cb = std::min(inner, B - off);
    fa.xbar = (const double*)dxbar + off * n; fa.xbar_sb = (long long)n; fa.xbar_sk = (long long)(B * n);
    fa.ubar = nullptr; fa.dt = p->dt;
    fa.x0 = dx0 + off * n;
    fa.K = nullptr;
    fa.K0 = dK0 ? dK0 + off * model::Dims<MODEL>::m * n : nullptr;
    fa.P0 = nullptr;
    fa.cost = dcost ? dcost + off : nullptr;
    fa.status = dstatus ? dstatus + off : nullptr;
    fa.batch = (long long)cb; fa.N = p->N;
    if constexpr (MODEL == 0) rc = launch_ric12<0, RIC_LQR, true>(ctx, fa);
    else rc = launch_ric24<0, RIC_LQR, true>(ctx, fa);
    if (rc) return rc;
    if ((rc = after_chunk(off, cb, chunk_no))) return rc;
  }
  return NOC_OK;
}

// Outer chunks of an SLQ call: the solver's working set (trajectories, gains, the saved nominal, stage costs) is
// ~136 KB per instance at n=12, N=200, so a batch that would not fit the scratch budget is solved slice by slice.
template <int MODEL>
int slq_impl(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* xgoal,
             const double* QRQf, const double* u_init, double* x, double* u, double* K, double* cost, int32_t* iters,
             int32_t* alpha_idx, int32_t* status) {
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m, N = p->N, B = (size_t)batch;
  const size_t per_inst = sizeof(double) * (2 * (N + 1) * n + 3 * N * m + N * m * n + (N + 1) + 8 + (MODEL == 1 ? N * M24_PREP_DOUBLES : 0)) +
                          4 * (size_t)(8 + p->max_iter);
  size_t have = 0, budget = 0;
  for (auto* v : {&ctx->scratch, &ctx->out_slots, &ctx->in_slots})
    for (auto& b : *v) have += b.cap;
  if (int rc = scratch_budget(ctx, BUDGET_SLQ, (int)n, (int)N, p->max_iter, batch, have, &budget)) return rc;
  size_t outer = std::max<size_t>(1, std::min<size_t>(B, budget / per_inst));
  if (outer < B && outer >= 64) outer &= ~size_t(31);
  for (size_t o = 0; o < B; o += outer) {
    const size_t ob = std::min(outer, B - o);
    ctx->pending.clear();
    int rc = slq_chunk<MODEL>(ctx, p, (int64_t)ob, x0 + o * n, xgoal + o * n, QRQf, u_init ? u_init + o * N * m : nullptr,
                              x ? x + o * (N + 1) * n : nullptr, u ? u + o * N * m : nullptr, K ? K + o * N * m * n : nullptr,
                              cost ? cost + o : nullptr, iters ? iters + o : nullptr,
                              alpha_idx ? alpha_idx + o * (size_t)p->max_iter : nullptr, status ? status + o : nullptr);
    if (rc) return rc;
  }
  return NOC_OK;
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
  p->reg_schedule = NOC_REG_X10;
}

int noc_create(noc_ctx** out, int device_id) {
  if (!out) return NOC_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return NOC_ERR_NO_DEVICE; }
  if (device_id < 0 || device_id >= count) return NOC_ERR_BAD_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return NOC_ERR_CUDA;
  // the library carries an sm_100a cubin only (arch-specific, no PTX): exactly compute capability 10.0
  if (prop.major != 10 || prop.minor != 0) return NOC_ERR_NO_DEVICE;
  if (cudaSetDevice(device_id) != cudaSuccess) return NOC_ERR_CUDA;
  noc_ctx* c = new noc_ctx();
  c->device = device_id;
  c->sm_count = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  c->stream = c->own_stream;
  ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& e : c->chunk_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->defer_ev, cudaEventDisableTiming) == cudaSuccess;
  for (auto& e : c->pipe_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {   // noc_destroy releases whatever was created (every handle starts as nullptr)
    cudaGetLastError();
    noc_destroy(c);
    return NOC_ERR_CUDA;
  }
  *out = c;
  return NOC_OK;
}

void noc_destroy(noc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto* v : {&ctx->in_slots, &ctx->out_slots, &ctx->scratch})
    for (auto& b : *v)
      if (b.p) cudaFree(b.p);
  for (auto& e : ctx->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (auto& e : ctx->chunk_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ctx->pipe_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
  if (ctx->defer_ev) cudaEventDestroy(ctx->defer_ev);
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

int noc_set_option(noc_ctx* ctx, int option, int64_t value) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (option < 0 || option >= NOC_OPT_COUNT) return fail(ctx, NOC_ERR_BAD_ARG, "unknown option");
  if (value < 0) return fail(ctx, NOC_ERR_BAD_ARG, "option values are >= 0");
  ctx->opt[option] = value;
  ctx->budgets.clear();
  return NOC_OK;
}

int noc_synchronize(noc_ctx* ctx) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NOC_OK;
}

int64_t noc_launch_count(const noc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int noc_device_alloc(noc_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return NOC_ERR_BAD_ARG;
  *out = nullptr;
  if (bytes == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  NOC_CUDA(ctx, cudaMalloc(out, (size_t)bytes));
  return NOC_OK;
}

int noc_device_free(noc_ctx* ctx, void* p) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!p) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  NOC_CUDA(ctx, cudaFree(p));
  return NOC_OK;
}

int noc_memcpy(noc_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (bytes == 0) return NOC_OK;
  if (!dst || !src) return fail(ctx, NOC_ERR_BAD_ARG, "noc_memcpy: NULL pointer");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  NOC_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, ctx->stream));
  return NOC_OK;
}

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
  PendingGuard pending_guard(ctx);
  if (!AB || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "AB and QRQf are required");
  if (batch == 0) return NOC_OK;
  // per-instance weights: (12,4) and (24,8) with A_THEN_B records run on k_ric12_mma<PERW> (W' fragments from L2 per
  // instance-step) / k_ric24_mma<PERW> (W' tiles staged per warp); every other shape / layout on the thread-per-instance
  // kernel (the other tuned kernels keep one W / Qf per CTA)
  const bool n24 = (p->n == 24 && p->m == 8);
  const bool perw = qr_per_instance && ((p->n == 12 && p->m == 4) || n24) && p->layout == NOC_LAYOUT_A_THEN_B &&
                    !(is_device_ptr(QRQf) && misaligned16(QRQf)) && !ctx->opt[NOC_OPT_NO_TENSOR_SWEEP];
  const bool generic = (qr_per_instance && !perw) || (!(p->n == 12 && p->m == 4) && !n24);
  if (generic && (p->n < 1 || p->n > RG_NMAX || p->m < 1 || p->m > RG_MMAX))
    return fail(ctx, NOC_ERR_UNSUPPORTED, "LTV sweep: 1 <= n <= 24 and 1 <= m <= 8 (tuned kernels: (12,4), (24,8))");
  if (misaligned16(AB) && is_device_ptr(AB)) return fail(ctx, NOC_ERR_BAD_ARG, "device AB must be 16-byte aligned (TMA bulk copies)");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = p->n, m = p->m, N = p->N;
  const size_t rec = (size_t)n * n + (size_t)n * m, B = (size_t)batch;
  // A host-resident A_k,B_k stream is the largest thing this library ever stages (12.6 GB at config 2): it goes
  // through a two-buffer pipeline bounded to the scratch budget instead of one monolithic copy
  if (!is_device_ptr(AB) && sizeof(double) * B * N * rec > ((size_t)32 << 20) && !(p->flags & NOC_FLAG_ASYNC))
    return lqr_batch_host_pipeline(ctx, p, batch, generic, AB, QRQf, qr_per_instance, x0dev, K, K0, P0, cost, status);
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
  rc = lqr_batch_device(ctx, p, batch, generic, (const double*)dAB, (const double*)dQ, qr_per_instance, (const double*)dx0,
                        (double*)dK, (double*)dK0, (double*)dP0, (double*)dcost, (int32_t*)dstatus);
  if (rc) return rc;
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_rollout_open_loop(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* ubar,
                          double* xbar) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  PendingGuard pending_guard(ctx);
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
  PendingGuard pending_guard(ctx);
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
  PendingGuard pending_guard(ctx);
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
  PendingGuard pending_guard(ctx);
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
  PendingGuard pending_guard(ctx);
  if (!AB || !QR) return fail(ctx, NOC_ERR_BAD_ARG, "AB and QR are required");
  if (p->max_iter < 1) return fail(ctx, NOC_ERR_BAD_ARG, "max_iter must be >= 1");
  if (batch == 0) return NOC_OK;
  const bool tuned = p->n == 12 && p->m == 4;
  // n=24, m=8, A_THEN_B records: the warp-per-instance tensor-core kernel (riccati24_mma.cuh) in its fixed-point mode
  const bool tuned24 = p->n == 24 && p->m == 8 && p->layout == NOC_LAYOUT_A_THEN_B && !ctx->opt[NOC_OPT_NO_TENSOR_SWEEP];
  if ((tuned || tuned24) && misaligned16(AB) && is_device_ptr(AB)) return fail(ctx, NOC_ERR_BAD_ARG, "device AB must be 16-byte aligned");
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
  if (tuned24) {
    Ric24Args a{};
    if ((rc = build_w24(ctx, (const double*)dQ, &a.W, &a.Qf, true))) return rc;
    a.AB = (const double*)dAB; a.K0 = (double*)dK; a.P0 = (double*)dP; a.iters = (int32_t*)dit; a.status = (int32_t*)dst;
    a.tol = p->tol; a.max_iter = p->max_iter; a.batch = batch; a.N = 1;
    if ((rc = launch_ric24_mma_shape<RIC_DARE, false, false, false>(ctx, a, nullptr))) return rc;
    return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
  }
  if (!tuned) {
    // every other dimension (and n=24, m=8 with ROWS_AB records): the thread-per-instance completeness kernel, same arithmetic
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
  if ((rc = build_w12(ctx, (const double*)dQ, true, &a.W, &a.Qf))) return rc;
  a.AB = (const double*)dAB; a.K = (double*)dK; a.P0 = (double*)dP; a.iters = (int32_t*)dit; a.status = (int32_t*)dst;
  a.tol = p->tol; a.max_iter = p->max_iter; a.batch = batch; a.N = 1;
  rc = launch_ric12_layout<RIC_DARE>(ctx, p->layout, a);
  if (rc) return rc;
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_dare_solve_at(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* xbar, const double* ubar,
                      const double* QR, double* P, double* K, int32_t* iters, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  PendingGuard pending_guard(ctx);
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
  if ((rc = build_w12(ctx, (const double*)dQ, true, &a.W, &a.Qf))) return rc;
  a.xbar = (const double*)dxb; a.xbar_sb = 12; a.xbar_sk = 12; a.ubar = (const double*)dub; a.dt = p->dt;
  a.K = (double*)dK; a.P0 = (double*)dP; a.iters = (int32_t*)dit; a.status = (int32_t*)dst;
  a.tol = p->tol; a.max_iter = p->max_iter; a.batch = batch; a.N = 1;
  rc = launch_ric12<0, RIC_DARE, true>(ctx, a);
  if (rc) return rc;
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
  PendingGuard pending_guard(ctx);
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "SLQ needs a built-in nonlinear model");
  if (!x0 || !xgoal || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "x0, xgoal and QRQf are required");
  if (p->max_iter < 1 || p->n_alpha < 1 || p->n_alpha > 32)
    return fail(ctx, NOC_ERR_BAD_ARG, "need max_iter >= 1 and 1 <= n_alpha <= 32");
  if (!(p->u_min < p->u_max)) return fail(ctx, NOC_ERR_BAD_ARG, "need u_min < u_max (-INFINITY / INFINITY = no input box)");
  if (p->reg_schedule != NOC_REG_X10 && p->reg_schedule != NOC_REG_ADAPTIVE) return fail(ctx, NOC_ERR_BAD_ARG, "unknown reg_schedule");
  if (batch > 0x7fffffffLL) return fail(ctx, NOC_ERR_BAD_ARG, "batch too large for 32-bit work lists");
  if (batch == 0) return NOC_OK;
  // device buffers are used in place, by 16-byte accesses (cp.async, bulk copies, vector loads); host buffers are staged
  for (const void* q : {(const void*)xgoal, (const void*)u_init, (const void*)x, (const void*)u, (const void*)K})
    if (q && is_device_ptr(q) && misaligned16(q)) return fail(ctx, NOC_ERR_BAD_ARG, "device xgoal / u_init / x / u / K must be 16-byte aligned");
  return p->model == NOC_MODEL_QUAD12
             ? slq_impl<0>(ctx, p, batch, x0, xgoal, QRQf, u_init, x, u, K, cost, iters, alpha_idx, status)
             : slq_impl<1>(ctx, p, batch, x0, xgoal, QRQf, u_init, x, u, K, cost, iters, alpha_idx, status);
}

int noc_plan_time_parallel(int32_t N, int32_t* chunks, int32_t* len) {
  if (N < 1 || !chunks || !len) return NOC_ERR_BAD_ARG;
  int c = 1, l = N;
  plan_time_parallel(N, &c, &l);
  *chunks = c; *len = l;
  return NOC_OK;
}

int noc_selftest_rcp(noc_ctx* ctx, const double* x, double* y_fast, double* y_ieee, int64_t n) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!x || !y_fast || !y_ieee || n < 1) return fail(ctx, NOC_ERR_BAD_ARG, "x, outputs and n>=1 required");
  PendingGuard pending_guard(ctx);
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

// local pick: multi-CTA reduction, result {cost, index + index_offset} left at best_dev (device)
static int argmin_device(noc_ctx* ctx, const double* dcost, int64_t batch, int64_t index_offset, CostIndex* best_dev) {
  void* part;
  int rc;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((batch + 4095) / 4096, 2 * (int64_t)ctx->sm_count));
  // partial pairs | ticket (zeroed once at allocation; the kernel's last CTA resets it)
  const size_t need = sizeof(CostIndex) * 2 * (size_t)ctx->sm_count + 16;
  const bool fresh = (int)ctx->scratch.size() <= SCR_ARGMIN_PART || ctx->scratch[SCR_ARGMIN_PART].cap < need;
  if ((rc = scratch(ctx, SCR_ARGMIN_PART, need, &part))) return rc;
  unsigned* ticket = (unsigned*)((char*)part + sizeof(CostIndex) * 2 * (size_t)ctx->sm_count);
  if (fresh) NOC_CUDA(ctx, cudaMemsetAsync(ticket, 0, 16, ctx->stream));
  {
    LaunchScope ls(ctx, NOC_KERNEL_ARGMIN);
    k_argmin_cost<<<grid, 1024, 0, ctx->stream>>>(dcost, batch, index_offset, (CostIndex*)part, ticket, best_dev);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

int noc_argmin_cost_async(noc_ctx* ctx, const double* cost_dev, int64_t batch, int64_t index_offset,
                          noc_cost_index_t* best_dev) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!cost_dev || !best_dev || batch < 1) return fail(ctx, NOC_ERR_BAD_ARG, "cost, best_dev and batch>=1 required");
  if (!is_device_ptr(cost_dev) || !is_device_ptr(best_dev)) return fail(ctx, NOC_ERR_BAD_ARG, "noc_argmin_cost_async: device pointers only");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  return argmin_device(ctx, cost_dev, batch, index_offset, (CostIndex*)best_dev);
}

int noc_argmin_cost(noc_ctx* ctx, const double* cost, int64_t batch, int64_t* best_index, double* best_cost) {
  if (!ctx) return NOC_ERR_BAD_ARG;
  if (!cost || !best_index || !best_cost || batch < 1) return fail(ctx, NOC_ERR_BAD_ARG, "cost, outputs and batch>=1 required");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  PendingGuard pending_guard(ctx);
  const void* dc;
  int rc;
  if ((rc = stage_in(ctx, 0, cost, sizeof(double) * (size_t)batch, &dc))) return rc;
  void* s;
  if ((rc = scratch(ctx, SCR_ARGMIN, sizeof(CostIndex), &s))) return rc;
  if ((rc = argmin_device(ctx, (const double*)dc, batch, 0, (CostIndex*)s))) return rc;
  CostIndex h;
  NOC_CUDA(ctx, cudaMemcpyAsync(&h, s, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *best_index = h.index;
  *best_cost = h.cost;
  return NOC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU plane: NCCL over NVLink, loaded at run time.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct NcclApi;
NcclApi* nccl_api();
std::string nccl_api_error();
struct NcclApi {
  void* handle = nullptr;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // inside a PyTorch process this returns torch's already-loaded libnccl.so.2 (same soname); otherwise the system's
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) { api.err = std::string("dlopen libnccl.so.2: ") + dlerror(); return; }
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(api.handle, name); if (!p) { ok = false; api.err = std::string("missing NCCL symbol ") + name; } return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; }
  });
  return api.handle ? &api : nullptr;
}
std::string nccl_api_error() {
  // nccl_api() keeps the failure text of its one dlopen attempt in the static object; fetch it through a second handle
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (h) { dlclose(h); return "libnccl.so.2 is loaded but lacks a required symbol"; }
  return "libnccl.so.2 not found by the dynamic loader";
}
}  // namespace

struct noc_comm {
  noc_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  std::string err;
};

#define NOC_NCCL(cm, call)                                                                    \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) {                                                                  \
      (cm)->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r_);                 \
      (cm)->ctx->err = (cm)->err;                                                             \
      return NOC_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

int noc_comm_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) == NOC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclApi* api = nccl_api();
  if (!api || !id128) return api ? NOC_ERR_BAD_ARG : NOC_ERR_NCCL;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return NOC_ERR_NCCL;
  std::memcpy(id128, &id, sizeof(id));
  return NOC_OK;
}

int noc_comm_create_rank(noc_comm** out, noc_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!out || !ctx) return NOC_ERR_BAD_ARG;
  *out = nullptr;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, NOC_ERR_BAD_ARG, "noc_comm_create_rank: id, 0 <= rank < nranks");
  NcclApi* api = nccl_api();
  if (!api) return fail(ctx, NOC_ERR_NCCL, "NCCL is not available: " + nccl_api_error());
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  noc_comm* c = new noc_comm();
  c->ctx = ctx; c->rank = rank; c->nranks = nranks;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    ctx->err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
    delete c;
    return NOC_ERR_NCCL;
  }
  *out = c;
  return NOC_OK;
}

int noc_comm_create_all(noc_comm** out, noc_ctx* const* ctxs, int n) {
  if (!out || !ctxs || n < 1) return NOC_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) { out[i] = nullptr; if (!ctxs[i]) return NOC_ERR_BAD_ARG; }
  NcclApi* api = nccl_api();
  if (!api) return fail(ctxs[0], NOC_ERR_NCCL, "NCCL is not available: " + nccl_api_error());
  std::vector<int> devs(n);
  std::vector<ncclComm_t> comms(n);
  for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
  ncclResult_t r = api->CommInitAll(comms.data(), n, devs.data());
  if (r != ncclSuccess) return fail(ctxs[0], NOC_ERR_NCCL, std::string("ncclCommInitAll: ") + api->GetErrorString(r));
  for (int i = 0; i < n; ++i) {
    noc_comm* c = new noc_comm();
    c->ctx = ctxs[i]; c->comm = comms[i]; c->rank = i; c->nranks = n;
    out[i] = c;
  }
  return NOC_OK;
}

void noc_comm_destroy(noc_comm* comm) {
  if (!comm) return;
  if (comm->comm && nccl_api()) {
    cudaSetDevice(comm->ctx->device);
    cudaStreamSynchronize(comm->ctx->stream);
    cudaStreamSynchronize(comm->ctx->aux_stream);
    nccl_api()->CommDestroy(comm->comm);
  }
  delete comm;
}

int noc_comm_rank(const noc_comm* comm) { return comm ? comm->rank : -1; }
int noc_comm_size(const noc_comm* comm) { return comm ? comm->nranks : 0; }
const char* noc_comm_last_error(const noc_comm* comm) { return comm ? comm->err.c_str() : "null comm"; }
int noc_comm_group_begin(void) { return (nccl_api() && nccl_api()->GroupStart() == ncclSuccess) ? NOC_OK : NOC_ERR_NCCL; }
int noc_comm_group_end(void) { return (nccl_api() && nccl_api()->GroupEnd() == ncclSuccess) ? NOC_OK : NOC_ERR_NCCL; }

int noc_comm_allgather(noc_comm* comm, const void* src_dev, void* dst_dev, int64_t count, int64_t elem_bytes) {
  if (!comm) return NOC_ERR_BAD_ARG;
  noc_ctx* ctx = comm->ctx;
  if (!src_dev || !dst_dev || count < 0 || elem_bytes < 1) return fail(ctx, NOC_ERR_BAD_ARG, "noc_comm_allgather: src, dst, count >= 0, elem_bytes >= 1");
  if (count == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  NOC_NCCL(comm, nccl_api()->AllGather(src_dev, dst_dev, (size_t)(count * elem_bytes), ncclUint8, comm->comm, ctx->stream));
  ctx->launches++;
  return NOC_OK;
}

// pairs of all ranks -> final pick, on the context's stream
static int best_from_local(noc_comm* comm, const CostIndex* local_dev, noc_cost_index_t* best_dev, noc_cost_index_t* best_host) {
  noc_ctx* ctx = comm->ctx;
  void* pairs;
  int rc;
  if ((rc = scratch(ctx, SCR_COMM_PAIRS, sizeof(CostIndex) * (size_t)(comm->nranks + 1), &pairs))) return rc;
  NOC_NCCL(comm, nccl_api()->AllGather(local_dev, pairs, sizeof(CostIndex), ncclUint8, comm->comm, ctx->stream));
  ctx->launches++;
  {
    LaunchScope ls(ctx, NOC_KERNEL_ARGMIN);
    k_argmin_pairs<<<1, 32, 0, ctx->stream>>>((const CostIndex*)pairs, comm->nranks, (CostIndex*)best_dev);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  if (best_host) {
    NOC_CUDA(ctx, cudaMemcpyAsync(best_host, best_dev, sizeof(CostIndex), cudaMemcpyDeviceToHost, ctx->stream));
    NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return NOC_OK;
}

int noc_comm_best_rollout(noc_comm* comm, const double* cost_dev, int64_t count, int64_t index_offset,
                          noc_cost_index_t* best_dev, noc_cost_index_t* best_host) {
  if (!comm) return NOC_ERR_BAD_ARG;
  noc_ctx* ctx = comm->ctx;
  if (!cost_dev || !best_dev || count < 1) return fail(ctx, NOC_ERR_BAD_ARG, "noc_comm_best_rollout: cost, best_dev, count >= 1");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  void* pairs;
  int rc;
  if ((rc = scratch(ctx, SCR_COMM_PAIRS, sizeof(CostIndex) * (size_t)(comm->nranks + 1), &pairs))) return rc;
  CostIndex* local = (CostIndex*)pairs + comm->nranks;   // own pair, after the gathered ones
  if ((rc = argmin_device(ctx, cost_dev, count, index_offset, local))) return rc;
  return best_from_local(comm, local, best_dev, best_host);
}

int noc_comm_lqr_solve_from_x0(noc_comm* comm, const noc_problem_t* p, int64_t count, const double* x0_dev,
                               const double* QRQf_dev, double* cost_all_dev, double* K0_all_dev, int32_t* status_dev,
                               noc_cost_index_t* best_dev, noc_cost_index_t* best_host) {
  if (!comm) return NOC_ERR_BAD_ARG;
  noc_ctx* ctx = comm->ctx;
  int rc = check_prob(ctx, p, count);
  if (rc) return rc;
  if (p->model == NOC_MODEL_LTV_USER) return fail(ctx, NOC_ERR_BAD_ARG, "noc_comm_lqr_solve_from_x0 needs a built-in model");
  if (!x0_dev || !QRQf_dev || !best_dev || count < 1) return fail(ctx, NOC_ERR_BAD_ARG, "x0, QRQf, best_dev and count >= 1 are required");
  for (const void* q : {(const void*)x0_dev, (const void*)QRQf_dev, (const void*)best_dev})
    if (!is_device_ptr(q)) return fail(ctx, NOC_ERR_BAD_ARG, "noc_comm_lqr_solve_from_x0: device pointers only");
  PendingGuard pending_guard(ctx);
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = p->n, m = p->m;
  const int G = comm->nranks, me = comm->rank;
  double* cost_local;
  if (cost_all_dev) cost_local = cost_all_dev + (size_t)me * count;
  else { void* q; if ((rc = scratch(ctx, SCR_COMM_COST, sizeof(double) * (size_t)count, &q))) return rc; cost_local = (double*)q; }
  double* K0_local = K0_all_dev ? K0_all_dev + (size_t)me * count * m * n : nullptr;
  const bool gather = G > 1 && (cost_all_dev || K0_all_dev);
  SideStreamDrain drain{ctx, gather};
  cudaStream_t const gstream = ctx->aux_stream;
  // after every swept chunk: its cost / K0 slices go to every rank (a broadcast per root, grouped: an all-gather with a
  // strided destination) on the gather stream, under the sweep of the next chunk
  auto after_chunk = [&](size_t off, size_t cb, int chunk_no) -> int {
    if (!gather) return NOC_OK;
    cudaEvent_t ev = ctx->chunk_ev[chunk_no & 3];
    NOC_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    NOC_CUDA(ctx, cudaStreamWaitEvent(gstream, ev, 0));
    NcclApi* api = nccl_api();
    NOC_NCCL(comm, api->GroupStart());
    for (int r = 0; r < G; ++r) {
      if (cost_all_dev) {
        double* q = cost_all_dev + (size_t)r * count + off;
        NOC_NCCL(comm, api->Broadcast(q, q, cb, ncclFloat64, r, comm->comm, gstream));
      }
      if (K0_all_dev) {
        double* q = K0_all_dev + ((size_t)r * count + off) * m * n;
        NOC_NCCL(comm, api->Broadcast(q, q, cb * m * n, ncclFloat64, r, comm->comm, gstream));
      }
    }
    NOC_NCCL(comm, api->GroupEnd());
    ctx->launches++;
    return NOC_OK;
  };
  rc = p->model == NOC_MODEL_QUAD12
           ? from_x0_device_chunks<0>(ctx, p, count, x0_dev, QRQf_dev, K0_local, cost_local, status_dev, gather, after_chunk)
           : from_x0_device_chunks<1>(ctx, p, count, x0_dev, QRQf_dev, K0_local, cost_local, status_dev, gather, after_chunk);
  if (rc) return rc;
  if (gather) {
    NOC_CUDA(ctx, cudaEventRecord(ctx->join_ev, gstream));
    NOC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0));
  }
  void* pairs;
  if ((rc = scratch(ctx, SCR_COMM_PAIRS, sizeof(CostIndex) * (size_t)(G + 1), &pairs))) return rc;
  CostIndex* local = (CostIndex*)pairs + G;
  if ((rc = argmin_device(ctx, cost_local, count, (int64_t)me * count, local))) return rc;
  if ((rc = best_from_local(comm, local, best_dev, best_host))) return rc;
  if (!(p->flags & NOC_FLAG_ASYNC) && !best_host) NOC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NOC_OK;
}

}  // extern "C"

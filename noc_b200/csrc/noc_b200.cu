// noc_b200.cu — C-ABI of libnoc_b200.so (include/lqr_control/noc_b200.h).
// Host-side batch scheduler: pointer classification, staging, chunking to the free HBM, kernel launches.
// There is no CPU fallback anywhere in this file: every entry point either launches CUDA kernels or fails.
#include "../../include/lqr_control/noc_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pipeline_kernels.cuh"
#include "riccati12.cuh"

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

enum { SCR_W = 0, SCR_XBAR = 1, SCR_AB = 2, SCR_ARGMIN = 3 };

// W = blockdiag(Q, R) as a dense (n+m)x(n+m) matrix, followed by Qf (n*n)
__global__ void k_build_w(const double* __restrict__ qrqf, double* __restrict__ w, int n, int m) {
  const int nm = n + m;
  for (int e = threadIdx.x; e < nm * nm; e += blockDim.x) {
    const int i = e / nm, j = e % nm;
    double v = 0.0;
    if (i < n && j < n) v = qrqf[i * n + j];
    else if (i >= n && j >= n) v = qrqf[n * n + (i - n) * m + (j - n)];
    w[e] = v;
  }
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) w[nm * nm + e] = qrqf[n * n + m * m + e];
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

constexpr int LQR12_WARPS = 4;
constexpr int LQR12_MIN_CTAS = 5;

template <int LAYOUT>
int launch_lqr12(noc_ctx* ctx, const Lqr12Args& a) {
  auto kern = k_lqr12<LAYOUT, LQR12_WARPS, LQR12_MIN_CTAS>;
  const size_t smem = lqr12_smem_bytes(LQR12_WARPS);
  static bool attr_set[2] = {false, false};
  if (!attr_set[LAYOUT]) {
    NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NOC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set[LAYOUT] = true;
  }
  const long long per_cta = 2 * LQR12_WARPS;
  const unsigned grid = (unsigned)((a.batch + per_cta - 1) / per_cta);
  {
    LaunchScope ls(ctx, NOC_KERNEL_RICCATI);
    kern<<<grid, LQR12_WARPS * 32, smem, ctx->stream>>>(a);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return NOC_OK;
}

// device-pointer core of the n=12 sweep
int lqr12_device(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* dAB, const double* dQRQf,
                 const double* dx0, double* dK, double* dK0, double* dP0, double* dcost, int32_t* dstatus) {
  void* w = nullptr;
  int rc = scratch(ctx, SCR_W, sizeof(double) * (16 * 16 + 144), &w);
  if (rc) return rc;
  {
    LaunchScope ls(ctx, NOC_KERNEL_SETUP);
    k_build_w<<<1, 256, 0, ctx->stream>>>(dQRQf, (double*)w, 12, 4);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  Lqr12Args a;
  a.AB = dAB; a.W = (const double*)w; a.Qf = (const double*)w + 256; a.x0 = dx0;
  a.K = dK; a.K0 = dK0; a.P0 = dP0; a.cost = dcost; a.status = dstatus;
  a.batch = batch; a.N = p->N;
  return p->layout == NOC_LAYOUT_A_THEN_B ? launch_lqr12<0>(ctx, a) : launch_lqr12<1>(ctx, a);
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
  if (qr_per_instance) return fail(ctx, NOC_ERR_UNSUPPORTED, "per-instance Q/R/Qf not implemented yet");
  if (!(p->n == 12 && p->m == 4)) return fail(ctx, NOC_ERR_UNSUPPORTED, "only n=12, m=4 has a sweep kernel");
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = p->n, m = p->m, N = p->N;
  const size_t rec = (size_t)n * n + (size_t)n * m, B = (size_t)batch;
  const void *dAB, *dQ, *dx0;
  void *dK, *dK0, *dP0, *dcost, *dstatus;
  if ((rc = stage_in(ctx, 0, AB, sizeof(double) * B * N * rec, &dAB))) return rc;
  if ((rc = stage_in(ctx, 1, QRQf, sizeof(double) * (2 * n * n + m * m), &dQ))) return rc;
  if ((rc = stage_in(ctx, 2, x0dev, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_out(ctx, 0, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 1, K0, sizeof(double) * B * m * n, &dK0))) return rc;
  if ((rc = stage_out(ctx, 2, P0, sizeof(double) * B * n * n, &dP0))) return rc;
  if ((rc = stage_out(ctx, 3, cost, sizeof(double) * B, &dcost))) return rc;
  if ((rc = stage_out(ctx, 4, status, sizeof(int32_t) * B, &dstatus))) return rc;
  rc = lqr12_device(ctx, p, batch, (const double*)dAB, (const double*)dQ, (const double*)dx0, (double*)dK,
                    (double*)dK0, (double*)dP0, (double*)dcost, (int32_t*)dstatus);
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
      k_rollout_open_loop<0><<<grid, 128, 0, ctx->stream>>>((const double*)dx0, (const double*)du, (double*)dx, batch, p->N, p->dt);
    else
      k_rollout_open_loop<1><<<grid, 128, 0, ctx->stream>>>((const double*)dx0, (const double*)du, (double*)dx, batch, p->N, p->dt);
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
      k_linearize<0><<<grid, 128, 0, ctx->stream>>>((const double*)dx, (const double*)du, (double*)dAB, batch, p->N, p->dt, p->layout);
    else
      k_linearize<1><<<grid, 128, 0, ctx->stream>>>((const double*)dx, (const double*)du, (double*)dAB, batch, p->N, p->dt, p->layout);
  }
  NOC_CUDA(ctx, cudaGetLastError());
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_lqr_solve_from_x0(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double* x0, const double* QRQf,
                          double* K, double* K0, double* P0, double* cost, int32_t* status) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  if (p->model != NOC_MODEL_QUAD12) return fail(ctx, NOC_ERR_UNSUPPORTED, "from_x0 pipeline: only NOC_MODEL_QUAD12 so far");
  if (!x0 || !QRQf) return fail(ctx, NOC_ERR_BAD_ARG, "x0 and QRQf are required");
  if (batch == 0) return NOC_OK;
  NOC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t B = (size_t)batch, n = 12, m = 4, N = p->N, rec = 192;
  const void *dx0, *dQ;
  void *dK, *dK0, *dP0, *dcost, *dstatus;
  if ((rc = stage_in(ctx, 0, x0, sizeof(double) * B * n, &dx0))) return rc;
  if ((rc = stage_in(ctx, 1, QRQf, sizeof(double) * (2 * n * n + m * m), &dQ))) return rc;
  if ((rc = stage_out(ctx, 0, K, sizeof(double) * B * N * m * n, &dK))) return rc;
  if ((rc = stage_out(ctx, 1, K0, sizeof(double) * B * m * n, &dK0))) return rc;
  if ((rc = stage_out(ctx, 2, P0, sizeof(double) * B * n * n, &dP0))) return rc;
  if ((rc = stage_out(ctx, 3, cost, sizeof(double) * B, &dcost))) return rc;
  if ((rc = stage_out(ctx, 4, status, sizeof(int32_t) * B, &dstatus))) return rc;
  // chunk so that the A_k,B_k scratch stays within a share of the free HBM
  size_t free_b = 0, total_b = 0;
  NOC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  size_t have = 0;
  if (ctx->scratch.size() > SCR_AB) have = ctx->scratch[SCR_AB].cap;
  if (ctx->scratch.size() > SCR_XBAR) have += ctx->scratch[SCR_XBAR].cap;
  const size_t budget = std::max<size_t>((size_t)((free_b + have) * 0.6), (size_t)64 << 20);
  const size_t per_inst = sizeof(double) * (N * rec + (N + 1) * n);
  size_t chunk = std::max<size_t>(1, std::min<size_t>(B, budget / per_inst));
  if (chunk < B) chunk = std::max<size_t>(8, chunk & ~size_t(7));
  void *dxbar, *dAB;
  if ((rc = scratch(ctx, SCR_XBAR, sizeof(double) * chunk * (N + 1) * n, &dxbar))) return rc;
  if ((rc = scratch(ctx, SCR_AB, sizeof(double) * chunk * N * rec, &dAB))) return rc;
  noc_problem_t q = *p;
  q.layout = NOC_LAYOUT_ROWS_AB;  // internal stream: one contiguous [A|B] row per F row
  for (size_t off = 0; off < B; off += chunk) {
    const long long cb = (long long)std::min(chunk, B - off);
    const double* x0c = (const double*)dx0 + off * n;
    {
      LaunchScope ls(ctx, NOC_KERNEL_ROLLOUT);
      k_rollout_open_loop<0><<<(unsigned)((cb + 127) / 128), 128, 0, ctx->stream>>>(x0c, nullptr, (double*)dxbar, cb, p->N, p->dt);
    }
    NOC_CUDA(ctx, cudaGetLastError());
    const long long total = cb * p->N;
    {
      LaunchScope ls(ctx, NOC_KERNEL_LINEARIZE);
      k_linearize<0><<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>((const double*)dxbar, nullptr, (double*)dAB, cb, p->N, p->dt, q.layout);
    }
    NOC_CUDA(ctx, cudaGetLastError());
    rc = lqr12_device(ctx, &q, cb, (const double*)dAB, (const double*)dQ, x0c,
                      dK ? (double*)dK + off * N * m * n : nullptr, dK0 ? (double*)dK0 + off * m * n : nullptr,
                      dP0 ? (double*)dP0 + off * n * n : nullptr, dcost ? (double*)dcost + off : nullptr,
                      dstatus ? (int32_t*)dstatus + off : nullptr);
    if (rc) { ctx->pending.clear(); return rc; }
  }
  return flush_out(ctx, p->flags & NOC_FLAG_ASYNC);
}

int noc_dare_solve_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double*, const double*, double*,
                         double*, int32_t*, int32_t*) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  return fail(ctx, NOC_ERR_UNSUPPORTED, "noc_dare_solve_batch: not implemented yet");
}

int noc_slq_solve_batch(noc_ctx* ctx, const noc_problem_t* p, int64_t batch, const double*, const double*,
                        const double*, double*, double*, double*, double*, int32_t*, int32_t*, int32_t*) {
  int rc = check_prob(ctx, p, batch);
  if (rc) return rc;
  return fail(ctx, NOC_ERR_UNSUPPORTED, "noc_slq_solve_batch: not implemented yet");
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

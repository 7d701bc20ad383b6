// pipeline_kernels.cuh — the kernels either side of the Riccati sweep:
//   k_rollout_open_loop  nominal trajectory from x0                 (sequential in k, thread per instance)
//   k_linearize          A_k, B_k for all (instance, k) at once     (time-parallel)
//   k_argmin_cost        smallest finite cost of a batch            ("pick the best rollout", §8e)
// Arithmetic: model.cuh; bit-identical to oracle/noc_oracle.c.
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace noc {

template <int MODEL>
__global__ void __launch_bounds__(128) k_rollout_open_loop(const double* __restrict__ x0,
                                                           const double* __restrict__ ubar,  // may be null
                                                           double* __restrict__ xbar, long long batch, int N,
                                                           double dt, long long xsb, long long xsk) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m;
  constexpr int LD = n + 1;                         // padded row of the transpose tile (odd stride: no bank conflicts)
  __shared__ __align__(16) double tile[4][32 * LD];
  const long long b_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = b_raw < batch;
  const long long b = active ? b_raw : batch - 1;
  const int lane = threadIdx.x & 31;
  double* tl = tile[threadIdx.x >> 5];
  double x[n], xn[n], u[m];
#pragma unroll
  for (int i = 0; i < n; ++i) x[i] = x0[b * n + i];
#pragma unroll
  for (int j = 0; j < m; ++j) u[j] = model::kHover;
  // x_k of instance b goes to xbar + b*xsb + k*xsk (doubles).  Dense instance-major: xsb=(N+1)*n, xsk=n: every
  // thread stores its own segments.  The library's internal scratch is step-major (xsb=n, xsk=batch*n): the 32
  // states of a warp are contiguous, so they are transposed through shared memory and written as 512-byte rows
  // (thread-private stores touched 32 sectors per instruction and made the LSU the rollout's bottleneck).
  const bool step_major = (xsb == n);
  const long long wb0 = b_raw - lane;              // first instance of this warp
  const int wcount = (int)((batch - wb0 < 32) ? (batch - wb0 > 0 ? batch - wb0 : 0) : 32);
  auto emit = [&](int k, const double* v) {
    if (!step_major) {
      if (active) {
        double* o = xbar + (size_t)b * xsb + (size_t)k * xsk;
#pragma unroll
        for (int i = 0; i < n; ++i) o[i] = v[i];
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) tl[lane * LD + i] = v[i];
    __syncwarp();
    double* o = xbar + (size_t)wb0 * n + (size_t)k * xsk;   // wcount * n contiguous doubles
    for (int e = lane; e < wcount * n; e += 32) o[e] = tl[(e / n) * LD + (e % n)];
    __syncwarp();
  };
  emit(0, x);
  for (int k = 0; k < N; ++k) {
    if (ubar) {
#pragma unroll
      for (int j = 0; j < m; ++j) u[j] = ubar[((size_t)b * N + k) * m + j];
    }
    model::model_step<MODEL>(x, u, dt, xn);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    emit(k + 1, x);
  }
}

// One thread per (work item, k); `batch` counts work items (= instances unless idx is given).  The record is
// ~85 % zeros, so a warp first zero-fills the 32 records of its lanes cooperatively — every store instruction
// writes 512 contiguous bytes — and only then does each lane scatter the structural non-zeros of its own record
// (a first version let every thread fill its own record: 32 lines per store instruction, 0.9 TB/s on the 6-KB
// records of the dual-UAV model).  Layout 0: A then B; layout 1: rows of [A row | B row].
template <int MODEL>
__global__ void __launch_bounds__(128) k_linearize(const double* __restrict__ xbar,
                                                   const double* __restrict__ ubar,  // may be null
                                                   double* __restrict__ AB, long long batch, int N, double dt,
                                                   int layout, const int32_t* __restrict__ idx) {
  constexpr int n = model::Dims<MODEL>::n, m = model::Dims<MODEL>::m, rec = n * n + n * m;
  const long long tw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = tw < batch * N;
  const long long twc = active ? tw : batch * N - 1;
  const long long item = twc / N;
  const int k = (int)(twc - item * N);
  const long long b = idx ? (long long)idx[item] : item;   // SLQ: only the instances still iterating
  const long long t = b * N + k;
  double* r = AB + (size_t)t * rec;
  {
    const int lane = threadIdx.x & 31;
    const double2 z = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      const unsigned long long pj = __shfl_sync(0xffffffffu, (unsigned long long)r, j);
      const bool on = __shfl_sync(0xffffffffu, active ? 1 : 0, j) != 0;
      if (on) {
        double2* dst = reinterpret_cast<double2*>(pj);
#pragma unroll
        for (int i = lane; i < rec / 2; i += 32) dst[i] = z;
      }
    }
    __syncwarp();
  }
  if (!active) return;
  double x[n], u[m];
  const double* xs = xbar + ((size_t)b * (N + 1) + k) * n;
#pragma unroll
  for (int i = 0; i < n; ++i) x[i] = xs[i];
#pragma unroll
  for (int j = 0; j < m; ++j) u[j] = ubar ? ubar[((size_t)b * N + k) * m + j] : model::kHover;
  if (layout == 0) {
    model::model_jac_discrete<MODEL>(
        x, u, dt, [&](int i, int j, double v) { r[i * n + j] = v; },
        [&](int i, int j, double v) { r[n * n + i * m + j] = v; });
  } else {
    model::model_jac_discrete<MODEL>(
        x, u, dt, [&](int i, int j, double v) { r[i * (n + m) + j] = v; },
        [&](int i, int j, double v) { r[i * (n + m) + n + j] = v; });
  }
}

// "Pick the best rollout": argmin over finite costs (NaN = failed instance, skipped), ties -> smallest index.
// Multi-CTA: every CTA reduces a grid-strided share to one (cost, index) pair; the last CTA to finish (atomic
// ticket) reduces the pairs and leaves the result ON THE DEVICE as {cost, index + index_offset}: no host round trip
// unless the caller asks for one.  (Round 1 ran a single 1024-thread CTA over up to 8 x 65,536 gathered costs.)
struct CostIndex {
  double cost;       // NaN when no finite cost exists
  long long index;   // -1 when no finite cost exists
};

__device__ __forceinline__ void argmin_merge(double& v, long long& idx, double ov, long long oi) {
  if (oi >= 0 && (idx < 0 || ov < v || (ov == v && oi < idx))) { v = ov; idx = oi; }
}

__device__ __forceinline__ void argmin_block_reduce(double& v, long long& idx, double* sv, long long* si) {
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, v, o);
    const long long oi = __shfl_down_sync(0xffffffffu, idx, o);
    argmin_merge(v, idx, ov, oi);
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool on = threadIdx.x < ((blockDim.x + 31) >> 5);
    v = on ? sv[threadIdx.x] : 0.0;
    idx = on ? si[threadIdx.x] : -1;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const long long oi = __shfl_down_sync(0xffffffffu, idx, o);
      argmin_merge(v, idx, ov, oi);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_argmin_cost(const double* __restrict__ cost, long long batch,
                                                      long long index_offset, CostIndex* __restrict__ partial,
                                                      unsigned* __restrict__ ticket, CostIndex* __restrict__ best) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ bool last;
  double v = 0.0;
  long long idx = -1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < batch; i += (long long)gridDim.x * blockDim.x) {
    const double c = cost[i];
    if (c < __longlong_as_double(0x7ff0000000000000LL) && (idx < 0 || c < v)) { v = c; idx = i; }   // finite only (false for NaN, +inf); ascending i: the first minimum wins
  }
  argmin_block_reduce(v, idx, sv, si);
  if (threadIdx.x == 0) {
    partial[blockIdx.x].cost = v;
    partial[blockIdx.x].index = idx;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  v = 0.0; idx = -1;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    const volatile CostIndex* p = partial + i;   // written by other CTAs of this launch: read past the cache
    argmin_merge(v, idx, p->cost, p->index);
  }
  argmin_block_reduce(v, idx, sv, si);
  if (threadIdx.x == 0) {
    best->cost = idx >= 0 ? v : qnan();
    best->index = idx >= 0 ? idx + index_offset : -1;
    *ticket = 0;   // ready for the next launch
  }
}

// final step of the cross-GPU pick: the smallest of the G gathered (cost, global index) pairs
__global__ void __launch_bounds__(32) k_argmin_pairs(const CostIndex* __restrict__ pairs, int count,
                                                     CostIndex* __restrict__ best) {
  double v = 0.0;
  long long idx = -1;
  for (int i = threadIdx.x; i < count; i += 32) argmin_merge(v, idx, pairs[i].cost, pairs[i].index);
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, v, o);
    const long long oi = __shfl_down_sync(0xffffffffu, idx, o);
    argmin_merge(v, idx, ov, oi);
  }
  if (threadIdx.x == 0) {
    best->cost = idx >= 0 ? v : qnan();
    best->index = idx;
  }
}

}  // namespace noc

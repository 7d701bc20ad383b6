// Microbenchmarks that fix the design constants of the Riccati kernel:
//   1. FP64 FMA peak (DFMA stream)           -> the FP64 roofline denominator
//   2. shared-memory wavefront packing for the broadcast LDS.128 / LDS.64 patterns a
//      lanes-cooperate small-matrix product generates
//   3. DFMA + LDS overlap at the ratio the kernel runs at
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
// Run on the GPU box only (gpurun). Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// Per-block (start, end, smid); host takes per-SM max(end)-min(start) and then the max over SMs.
// (Block 0 alone is misleading: the warp arbiter is not fair, early blocks finish first.)
#define MAXB 4096
__device__ long long g_t0[MAXB], g_t1[MAXB];
__device__ int g_sm[MAXB];
__device__ __forceinline__ int smid() { int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define CYC_BEGIN long long c0_ = clock64();
#define CYC_END   __syncthreads(); if (threadIdx.x == 0) { g_t0[blockIdx.x] = c0_; g_t1[blockIdx.x] = clock64(); g_sm[blockIdx.x] = smid(); }
static int g_last_blocks = 0;
static long long get_cyc() {
  static long long t0[MAXB], t1[MAXB]; static int sm[MAXB];
  cudaMemcpyFromSymbol(t0, g_t0, sizeof(t0)); cudaMemcpyFromSymbol(t1, g_t1, sizeof(t1));
  cudaMemcpyFromSymbol(sm, g_sm, sizeof(sm));
  long long lo[256], hi[256]; for (int i = 0; i < 256; ++i) { lo[i] = (1ll << 62); hi[i] = 0; }
  for (int b = 0; b < g_last_blocks && b < MAXB; ++b) {
    int s_ = sm[b] & 255; if (t0[b] < lo[s_]) lo[s_] = t0[b]; if (t1[b] > hi[s_]) hi[s_] = t1[b]; }
  long long best = 0; for (int i = 0; i < 256; ++i) if (hi[i] && hi[i] - lo[i] > best) best = hi[i] - lo[i];
  return best;
}

// ---------------------------------------------------------------- DFMA peak
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (double)(threadIdx.x + c);
  CYC_BEGIN
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = __fma_rn(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  CYC_END
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- LDS patterns
// Every lane loads from  base + off(lane) + it*stride  (stride keeps alignment + bank pattern).
// MODE selects off(lane). VEC = 2 (LDS.128) or 1 (LDS.64).
__device__ __forceinline__ int lane_off_bytes(int mode, int l) {
  switch (mode) {
    case 0: return l * 16;                                   // all distinct, contiguous (512 B for .128)
    case 1: return 0;                                        // full broadcast
    case 2: return (l & 3) * 32 + (l >> 4) * 16;             // "by bj": 8 distinct chunks / 128 B, quarter-warp holds 4
    case 3: return ((l >> 2) & 3) * 32 + (l >> 4) * 16;      // "by bi": 8 distinct chunks, quarter-warp holds 2
    case 4: return (l & 7) * 16;                             // each quarter-warp: the same 8 distinct chunks
    case 5: return (l >> 2) * 16;                            // 8 distinct chunks, each quarter holds 2
    case 6: return (l & 15) * 16;                            // 16 distinct chunks (256 B), halves identical
    case 7: return (l & 3) * 16 + (l >> 4) * 64;             // 8 distinct, contiguous 128 B
    case 8: return l * 8;                                    // LDS.64 all distinct (256 B)
    case 9: return (l & 15) * 8;                             // LDS.64 16 distinct (128 B)
    case 10: return ((l & 3) * 4 + (l >> 4) * 2) * 8;        // LDS.64 8 distinct 8-B words, conflict-free
    case 11: return (l & 1) * 16;                            // 2 distinct chunks
    default: return 0;
  }
}

// LDS-only loop. The address depends on the iteration (so nothing can be hoisted) and only one
// 32-bit word of every load is consumed (1 IADD per LDS keeps the issue slots free).
template <int VEC>
__global__ void __launch_bounds__(256) k_lds_raw(unsigned long long* out, int iters, int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int i = threadIdx.x; i < 8192 / 8; i += blockDim.x) ((unsigned long long*)smem)[i] = i * 0x9E3779B97F4A7C15ull;
  __syncthreads();
  const int l = threadIdx.x & 31;
  const unsigned addr0 = (unsigned)__cvta_generic_to_shared(smem) + lane_off_bytes(mode, l);
  unsigned acc0 = 0, acc1 = 0;
  CYC_BEGIN
  for (int i = 0; i < iters; ++i) {
    const unsigned base = addr0 + ((i & 1) << 12);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned addr = base + u * 512;
      unsigned a, b, c, d;
      if (VEC == 2) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
      } else {
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
      }
      if (VEC == 2) { acc0 = acc0 + a + b; acc1 = acc1 + c + d; }
      else { acc0 = acc0 + a + b; }
    }
  }
  unsigned long long s = acc0 + acc1;
  CYC_END
  if (s == 0x1234567ull) out[0] = s;
}

// ---------------------------------------------------------------- DFMA + LDS.128 mix
// Per inner iteration: NF DFMAs on 16 accumulators, NL LDS.128 (pattern mode 2).
template <int NF, int NL>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int i = threadIdx.x; i < 8192 / 8; i += blockDim.x) ((double*)smem)[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int l = threadIdx.x & 31;
  const unsigned char* p = smem + lane_off_bytes(mode, l);
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = (double)c;
  CYC_BEGIN
  for (int i = 0; i < iters; ++i) {
    double2 v[NL];
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      unsigned addr = (unsigned)__cvta_generic_to_shared(p) + u * 512;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "r"(addr));
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double2 w = v[f % NL];
      acc[f % 16] = __fma_rn(acc[f % 16], (f & 1) ? w.x : w.y, w.x);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += acc[c];
  CYC_END
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- SHFL throughput (64-bit = 2 SHFL.32)
__global__ void __launch_bounds__(256) k_shfl(double* out, int iters) {
  double v = threadIdx.x;
  double s = 0;
  CYC_BEGIN
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) { v = __shfl_xor_sync(0xffffffffu, v, 1 + u); s += v; }
  }
  CYC_END
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- HBM copy (sanity vs MEASURED_PEAKS)
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}
__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ a, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += stride) { double2 v = a[i]; s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F&& launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  double* d_out; CK(cudaMalloc(&d_out, 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_khz\": %d,\n", prop.name, sms, clk_khz);

  // 1. DFMA peak: 8 CTAs of 256 threads per SM (one resident wave), 8/16 chains
  {
    const int iters = 20000, blocks = sms * 8;
    float ms = time_ms([&] { g_last_blocks = blocks; k_dfma<8><<<blocks, 256>>>(d_out, iters, 1.0000001, 1e-9); });
    long long cyc = get_cyc();
    printf(" \"dfma_tflops_8chains\": %.3f, \"dfma_lanes_per_clk_per_sm_8chains\": %.2f, \"eff_mhz_8chains\": %.0f,\n",
           2.0 * 8 * iters * 256.0 * blocks / ms * 1e-9, 8.0 * iters * 256.0 * 8 / (double)cyc, cyc / (ms * 1e3));
    float ms2 = time_ms([&] { g_last_blocks = blocks; k_dfma<16><<<blocks, 256>>>(d_out, iters, 1.0000001, 1e-9); });
    cyc = get_cyc();
    printf(" \"dfma_tflops_16chains\": %.3f, \"dfma_lanes_per_clk_per_sm_16chains\": %.2f,\n",
           2.0 * 16 * iters * 256.0 * blocks / ms2 * 1e-9, 16.0 * iters * 256.0 * 8 / (double)cyc);
    // occupancy sweep: how many warps/SMSP does the pipe need with 8 / 16 chains
    for (int cta = 1; cta <= 4; cta *= 2) {
      float m = time_ms([&] { g_last_blocks = sms * cta; k_dfma<16><<<sms * cta, 128>>>(d_out, iters, 1.0000001, 1e-9); });
      cyc = get_cyc();
      printf(" \"dfma16_warps_per_smsp_%d_lanes_per_clk\": %.2f,\n", cta, 16.0 * iters * 128.0 * cta / (double)cyc);
      (void)m;
    }
    // long run (about 2 s) for the sustained figure under the power cap
    const int iters_l = 400000;
    float ms3 = time_ms([&] { g_last_blocks = blocks; k_dfma<8><<<blocks, 256>>>(d_out, iters_l, 1.0000001, 1e-9); }, 2);
    cyc = get_cyc();
    printf(" \"dfma_tflops_sustained\": %.3f, \"dfma_sustained_ms\": %.1f, \"eff_mhz_sustained\": %.0f,\n",
           2.0 * 8 * iters_l * 256.0 * blocks / ms3 * 1e-9, ms3, cyc / (ms3 * 1e3));
  }
  // 2. LDS patterns: SM cycles per warp-instruction (all 4 SMSPs issuing; 32 warps/SM)
  {
    const int iters = 4000, blocks = sms * 4;
    printf(" \"lds_clk_per_warp_instr_per_sm\": {\n");
    for (int mode = 0; mode <= 11; ++mode) {
      const bool is64 = (mode >= 8 && mode <= 10);
      float ms = is64 ? time_ms([&] { g_last_blocks = blocks; k_lds_raw<1><<<blocks, 256, 8192>>>((unsigned long long*)d_out, iters, mode); })
                      : time_ms([&] { g_last_blocks = blocks; k_lds_raw<2><<<blocks, 256, 8192>>>((unsigned long long*)d_out, iters, mode); });
      long long cyc = get_cyc();
      double warp_instrs_per_sm = 8.0 * iters * (256 / 32) * 4;   // per SM
      printf("  \"mode%d_%s\": %.3f%s\n", mode, is64 ? "lds64" : "lds128", (double)cyc / warp_instrs_per_sm,
             mode == 11 ? "" : ",");
      (void)ms;
    }
    printf(" },\n");
  }
  // 3. mix: DFMA warp-instr per clk per SM (peak = 2.0) with NL LDS.128 per NF DFMA
  {
    const int iters = 4000, blocks = sms * 4;
    auto rep = [&](const char* name, float ms, int nf) {
      long long cyc = get_cyc();
      double fma_instr = (double)nf * iters * (256 / 32) * 4;     // per SM
      printf(" \"%s\": %.3f,\n", name, fma_instr / (double)cyc);
      (void)ms;
    };
    rep("mix_16f_4l_mode2", time_ms([&] { g_last_blocks = blocks; k_mix<16, 4><<<blocks, 256, 8192>>>(d_out, iters, 2); }), 16);
    rep("mix_16f_4l_mode1", time_ms([&] { g_last_blocks = blocks; k_mix<16, 4><<<blocks, 256, 8192>>>(d_out, iters, 1); }), 16);
    rep("mix_16f_4l_mode0", time_ms([&] { g_last_blocks = blocks; k_mix<16, 4><<<blocks, 256, 8192>>>(d_out, iters, 0); }), 16);
    rep("mix_12f_4l_mode2", time_ms([&] { g_last_blocks = blocks; k_mix<12, 4><<<blocks, 256, 8192>>>(d_out, iters, 2); }), 12);
    rep("mix_16f_8l_mode2", time_ms([&] { g_last_blocks = blocks; k_mix<16, 8><<<blocks, 256, 8192>>>(d_out, iters, 2); }), 16);
    rep("mix_16f_2l_mode2", time_ms([&] { g_last_blocks = blocks; k_mix<16, 2><<<blocks, 256, 8192>>>(d_out, iters, 2); }), 16);
    rep("mix_16f_4l_mode2_half_occ_x2", time_ms([&] { g_last_blocks = sms * 2; k_mix<16, 4><<<sms * 2, 256, 8192>>>(d_out, iters, 2); }), 8);
  }
  // 4. shuffle (64-bit value = 2 SHFL.32)
  {
    const int iters = 4000, blocks = sms * 4;
    float ms = time_ms([&] { g_last_blocks = blocks; k_shfl<<<blocks, 256>>>(d_out, iters); });
    long long cyc = get_cyc();
    double instr = 8.0 * iters * (256 / 32) * 4;
    printf(" \"shfl64_clk_per_warp_op_per_sm\": %.3f,\n", (double)cyc / instr);
    (void)ms;
  }
  // 5. HBM
  {
    const size_t bytes = (size_t)4 << 30;
    double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    const size_t n = bytes / sizeof(double2);
    float ms = time_ms([&] { k_copy<<<sms * 16, 256>>>(a, b, n); });
    printf(" \"hbm_copy_gbs\": %.1f,\n", 2.0 * bytes / ms * 1e-6);
    float ms2 = time_ms([&] { k_read<<<sms * 16, 256>>>(a, d_out, n); });
    printf(" \"hbm_read_gbs\": %.1f\n", 1.0 * bytes / ms2 * 1e-6);
    cudaFree(a); cudaFree(b);
  }
  printf("}\n");
  return 0;
}

// DFMA throughput of a register-blocked outer product (the shape of the sweep's GEMM phases): 12x4 accumulators,
// acc[r][s] += p[r] * f[s], operands changing every iteration, vs the plain dependent-chain stream of ubench.cu.
// Prints DFMA lanes per clock per SM for 1, 2 and 8 warps per scheduler.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int SNAKE>
__global__ void __launch_bounds__(128) k_outer(double* out, int iters, double a, double b) {
  double acc[12][4], p[12], f[4];
#pragma unroll
  for (int r = 0; r < 12; ++r) { p[r] = threadIdx.x + r; for (int s = 0; s < 4; ++s) acc[r][s] = r + s; }
#pragma unroll
  for (int s = 0; s < 4; ++s) f[s] = 1.0 + s;
  for (int it = 0; it < iters; ++it) {
    if (SNAKE == 2) {   // column-major: f[s] stays in its operand slot for twelve consecutive DFMAs
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < 12; ++r) acc[r][s] = __fma_rn(p[r], f[s], acc[r][s]);
    } else {
#pragma unroll
      for (int r = 0; r < 12; ++r)
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          const int s = (SNAKE && (r & 1)) ? 3 - ss : ss;
          acc[r][s] = __fma_rn(p[r], f[s], acc[r][s]);
        }
    }
    // new operands every iteration (cheap, non-FP64): flip sign bits
#pragma unroll
    for (int r = 0; r < 12; ++r) p[r] = __longlong_as_double(__double_as_longlong(p[r]) ^ ((long long)it << 63));
#pragma unroll
    for (int s = 0; s < 4; ++s) f[s] = __longlong_as_double(__double_as_longlong(f[s]) ^ ((long long)(it & 1) << 62 >> 62 << 63));
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < 12; ++r) for (int c = 0; c < 4; ++c) s += acc[r][c];
  if (s == 123.456) out[0] = s;
}

template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  double* d; CK(cudaMalloc(&d, 64));
  const int iters = 20000;
  printf("{\"gpu\": \"%s\"", p.name);
  for (int snake = 0; snake < 3; ++snake)
    for (int ctas = 1; ctas <= 8; ctas *= 2) {   // CTAs of 128 threads per SM = warps per scheduler
      const int blocks = p.multiProcessorCount * ctas;
      float ms = snake == 2 ? time_ms([&] { k_outer<2><<<blocks, 128>>>(d, iters, 1.0, 1.0); })
                 : snake  ? time_ms([&] { k_outer<1><<<blocks, 128>>>(d, iters, 1.0, 1.0); })
                          : time_ms([&] { k_outer<0><<<blocks, 128>>>(d, iters, 1.0, 1.0); });
      const double fma = 48.0 * iters * 128.0 * blocks;
      printf(", \"%s_w%d_tflops\": %.2f, \"%s_w%d_lanes_per_clk_per_sm\": %.1f", snake == 2 ? "colmajor" : snake ? "snake" : "rowmajor", ctas,
             2 * fma / ms * 1e-9, snake == 2 ? "colmajor" : snake ? "snake" : "rowmajor", ctas, fma / (ms * 1e-3) / (clk_khz * 1e3) / p.multiProcessorCount);
    }
  printf("}\n");
  return 0;
}

// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) microbenchmark for B200: is it worth using for the 12x12x16 products
// of the Riccati sweep?  Measures (1) DMMA alone, (2) DFMA alone, (3) both interleaved in one warp, (4) DMMA + LDS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu ; prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NM) dmma(c[i][0], c[i][1], a, b);
      if (i < NF) {
#pragma unroll
        for (int r = 0; r < 8; ++r) f[r] = __fma_rn(f[r], a, b);   // 8 DFMA per slot = the FMA count of one DMMA per lane
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
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
  double* d; CK(cudaMalloc(&d, 64));
  const int iters = 20000, blocks = p.multiProcessorCount * 8;
  const double lanes = 256.0 * blocks;
  // per lane per iteration: NM DMMA = NM*8 FMA-equivalents (256 FMAs / 32 lanes), NF*8 DFMA
  float m_dmma = time_ms([&] { k_mix<8, 0><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); });
  float m_dfma = time_ms([&] { k_mix<0, 8><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); });
  float m_both = time_ms([&] { k_mix<8, 8><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); });
  float m_half = time_ms([&] { k_mix<4, 4><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); });
  auto tf = [&](double fma_per_lane_iter, float ms) { return 2.0 * fma_per_lane_iter * iters * lanes / ms * 1e-9; };
  printf("{\"gpu\": \"%s\", \"dmma_only_tflops\": %.2f, \"dfma_only_tflops\": %.2f, \"dmma_plus_dfma_tflops\": %.2f, "
         "\"half_half_tflops\": %.2f, \"ms\": [%.3f, %.3f, %.3f, %.3f]}\n",
         p.name, tf(64, m_dmma), tf(64, m_dfma), tf(128, m_both), tf(64, m_half), m_dmma, m_dfma, m_both, m_half);
  return 0;
}

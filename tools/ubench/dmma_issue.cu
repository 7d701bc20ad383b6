// DMMA (mma.sync.m8n8k4.f64) issue rate and latency on B200 as a function of warps per scheduler and of the number of
// independent accumulator chains per warp — the numbers k_ric12_mma's tensor phase is designed around.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue dmma_issue.cu ; prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
double run(int warps_per_sm, int sms, double* d, long long* dc) {
  const int iters = 4096;
  k<CHAINS><<<sms, warps_per_sm * 32>>>(d, dc, iters, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
  return (double)c / ((double)iters * CHAINS);   // cycles per DMMA per warp
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* d; long long* dc; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&dc, 8));
  printf("{\"gpu\": \"%s\", \"cycles_per_dmma_per_warp\": {", p.name);
  const int wps[4] = {4, 8, 12, 16};
  for (int wi = 0; wi < 4; ++wi) {
    const int w = wps[wi];
    printf("%s\"%d_warps_per_sm\": {\"chains_1\": %.1f, \"chains_2\": %.1f, \"chains_3\": %.1f, \"chains_4\": %.1f, \"chains_6\": %.1f, \"chains_8\": %.1f}",
           wi ? ", " : "", w, run<1>(w, p.multiProcessorCount, d, dc), run<2>(w, p.multiProcessorCount, d, dc),
           run<3>(w, p.multiProcessorCount, d, dc), run<4>(w, p.multiProcessorCount, d, dc), run<6>(w, p.multiProcessorCount, d, dc),
           run<8>(w, p.multiProcessorCount, d, dc));
  }
  printf("}, \"note\": \"chains_1 = dependent-issue latency; one warp per scheduler at 4 warps per SM\"}\n");
  return 0;
}

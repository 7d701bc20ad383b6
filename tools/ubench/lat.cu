// lat.cu — dependent-chain latencies that bound the sequential rollouts (one warp, one CTA, clock64 around the chain):
// DFMA, noc_sincos, and one full Euler step of the quadrotor model (x <- x + dt f(x, u)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I ../../noc_b200/csrc -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#include "model.cuh"
using namespace noc;
__global__ void k(double* out, long long* cyc, double seed, int iters) {
  double a = seed, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __fma_rn(a, b, c);
  long long t1 = clock64();
  cyc[0] = t1 - t0;
  double s, co, ang = seed;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { model::noc_sincos(ang, s, co); ang = ang + 0.25 * s; }
  t1 = clock64();
  cyc[1] = t1 - t0;
  double x[12], xn[12], u[4] = {2.5, 2.4, 2.45, 2.47};
  for (int i = 0; i < 12; ++i) x[i] = 0.01 * (i + 1) * seed;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    model::model_step<0>(x, u, 0.02, xn);
    for (int j = 0; j < 12; ++j) x[j] = xn[j];
  }
  t1 = clock64();
  cyc[2] = t1 - t0;
  out[threadIdx.x] = a + ang + x[0] + x[6] + x[11];
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 3 * 8);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(d, c, 0.37, iters); cudaDeviceSynchronize(); }
  long long h[3]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"dfma_dependent_cycles\": %.2f, \"sincos_plus_add_cycles\": %.1f, \"quad_model_step_cycles\": %.1f}\n",
         (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters);
  return cudaGetLastError() != cudaSuccess;
}

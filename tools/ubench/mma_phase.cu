// Where does a warp-step of k_ric12_mma go?  Timing build of the product kernel (-DRM_TIMING: clock64 around the TMA
// wait, the tensor phase, the scalar phase and the pair barrier), run on synthetic stable LTV records at the headline
// size, with and without the phase interlock (8 warps x 1 CTA vs 4 warps x 2 CTAs per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -DRM_TIMING -I../../noc_b200/csrc -I../../include -o mma_phase mma_phase.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "riccati12_mma.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
using namespace noc;

__global__ void k_fill(double* AB, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % 192);
    unsigned h = (unsigned)(i * 2654435761u); h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const double r = ((h & 0xffff) / 65536.0 - 0.5) * 0.02;
    AB[i] = (e < 144 && e / 12 == e % 12) ? 1.0 + r : r;
  }
}

template <int W, int C, bool D>
void run(const Ric12Args& a0, long long* dtm, int batch, int N, const char* name) {
  Ric12Args a = a0;
  a.iters = reinterpret_cast<int32_t*>(dtm);
  auto kern = k_ric12_mma<W, C, D>;
  const size_t smem = ric12_mma_smem_bytes<D>(W);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const unsigned grid = (batch + 8 * W - 1) / (8 * W);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<grid, W * 32, smem>>>(a); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  kern<<<grid, W * 32, smem>>>(a);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const int nw = grid * W;
  std::vector<long long> tm(4 * (size_t)nw);
  CK(cudaMemcpy(tm.data(), dtm, tm.size() * 8, cudaMemcpyDeviceToHost));
  double s[4] = {0, 0, 0, 0};
  for (int w = 0; w < nw; ++w) for (int j = 0; j < 4; ++j) s[j] += (double)tm[4 * w + j];
  printf("{\"config\": \"%s\", \"ms\": %.3f, \"cycles_per_warp_step\": {\"tma_wait\": %.0f, \"tensor\": %.0f, \"scalar\": %.0f, \"pair_barrier\": %.0f}}\n",
         name, ms, s[0] / nw / N, s[1] / nw / N, s[2] / nw / N, s[3] / nw / N);
}

int main(int argc, char** argv) {
  const int batch = argc > 1 ? atoi(argv[1]) : 65536, N = argc > 2 ? atoi(argv[2]) : 100;
  double *AB, *W, *K, *P0; long long* dtm;
  CK(cudaMalloc(&AB, (size_t)batch * N * 192 * 8)); CK(cudaMalloc(&K, (size_t)batch * N * 48 * 8));
  CK(cudaMalloc(&P0, (size_t)batch * 144 * 8)); CK(cudaMalloc(&W, (256 + 144) * 8)); CK(cudaMalloc(&dtm, (size_t)(batch / 8 + 64) * 32));
  k_fill<<<148 * 8, 256>>>(AB, (long long)batch * N * 192);
  std::vector<double> w(400, 0.0);
  for (int i = 0; i < 12; ++i) { w[i * 16 + i] = 1.0 + 0.1 * i; w[256 + i * 12 + i] = 10.0; }
  for (int i = 12; i < 16; ++i) w[i * 16 + i] = 0.1;
  CK(cudaMemcpy(W, w.data(), 400 * 8, cudaMemcpyHostToDevice));
  Ric12Args a{};
  a.AB = AB; a.W = W; a.Qf = W + 256; a.K = K; a.P0 = P0; a.batch = batch; a.N = N;
  run<8, 1, false>(a, dtm, batch, N, "TMA-staged records, 8 warps x 1 CTA (phase interlock)");
  run<4, 2, false>(a, dtm, batch, N, "TMA-staged records, 4 warps x 2 CTAs (free-running)");
  run<4, 3, true>(a, dtm, batch, N, "fragments from L2, 4 warps x 3 CTAs");
  run<4, 2, true>(a, dtm, batch, N, "fragments from L2, 4 warps x 2 CTAs");
  return 0;
}

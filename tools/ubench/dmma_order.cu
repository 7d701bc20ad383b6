// What does mma.sync.m8n8k4.f64 (SASS: DMMA) compute, bit for bit?  For every output element d = c + sum_{k<4} a_k b_k
// the candidates are: a chain of four IEEE fmas in any of the 24 orders of k (c first), and the pairwise / "products
// summed first" forms.  Random operands with wide exponent spread and heavy cancellation make the candidates differ in
// the last bits almost always, so a candidate that matches every one of the trials is the hardware's evaluation order.
// If the ascending chain matches, a k-loop of DMMAs reproduces the oracle's `fma chain over l ascending`
// EXACTLY and the bitwise sweeps may use the FP64 tensor cores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o dmma_order dmma_order.cu ; prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

// one warp per trial: A 8x4 (row-major), B 4x8 (row-major: B[k][n]), C 8x8 -> D 8x8
__global__ void k_dmma(const double* A, const double* B, const double* C, double* D, int trials, int ksteps) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= trials) return;
  const int g = lane >> 2, t = lane & 3;
  double d0 = C[(size_t)w * 64 + g * 8 + 2 * t], d1 = C[(size_t)w * 64 + g * 8 + 2 * t + 1];
  for (int s = 0; s < ksteps; ++s) {   // chained k-steps: operands of step s
    const double a = A[((size_t)w * ksteps + s) * 32 + g * 4 + t];        // A[g][t]
    const double b = B[((size_t)w * ksteps + s) * 32 + t * 8 + g];        // B[t][g]
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
  }
  D[(size_t)w * 64 + g * 8 + 2 * t] = d0;
  D[(size_t)w * 64 + g * 8 + 2 * t + 1] = d1;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double rnd_double(int spread) {
  const double m = 1.0 + (double)(rnd() >> 11) * (1.0 / 9007199254740992.0);
  const int e = (int)(rnd() % (2 * spread + 1)) - spread;
  return ((rnd() & 1) ? -1.0 : 1.0) * std::ldexp(m, e);
}

int main() {
  const int trials = 1 << 16;
  int all_ok = 1;
  printf("{");
  for (int ksteps = 1; ksteps <= 6; ksteps += 5) {
    const size_t na = (size_t)trials * ksteps * 32;
    std::vector<double> A(na), B(na), C((size_t)trials * 64), D((size_t)trials * 64);
    for (auto& v : A) v = rnd_double(6);
    for (auto& v : B) v = rnd_double(6);
    for (auto& v : C) v = rnd_double(8);
    double *dA, *dB, *dC, *dD;
    CK(cudaMalloc(&dA, na * 8)); CK(cudaMalloc(&dB, na * 8)); CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dD, D.size() * 8));
    CK(cudaMemcpy(dA, A.data(), na * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), na * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    k_dmma<<<(trials * 32 + 255) / 256, 256>>>(dA, dB, dC, dD, trials, ksteps);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
    // candidates: the 24 chain orders, applied identically in every k-step
    int perm[4] = {0, 1, 2, 3};
    long long match[24] = {0}, match_sumfirst = 0, match_pair = 0, elems = 0;
    std::vector<std::vector<int>> perms;
    do { perms.push_back({perm[0], perm[1], perm[2], perm[3]}); } while (std::next_permutation(perm, perm + 4));
    for (int w = 0; w < trials; ++w)
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
          ++elems;
          uint64_t got; std::memcpy(&got, &D[(size_t)w * 64 + i * 8 + j], 8);
          for (size_t p = 0; p < perms.size(); ++p) {
            double acc = C[(size_t)w * 64 + i * 8 + j];
            for (int s = 0; s < ksteps; ++s) {
              const double* a = &A[((size_t)w * ksteps + s) * 32 + i * 4];
              const double* b = &B[((size_t)w * ksteps + s) * 32 + j];
              for (int q = 0; q < 4; ++q) acc = std::fma(a[perms[p][q]], b[perms[p][q] * 8], acc);
            }
            uint64_t ref; std::memcpy(&ref, &acc, 8);
            match[p] += (ref == got);
          }
          {   // products summed first (ascending), then added to c; and the pairwise tree
            double acc = C[(size_t)w * 64 + i * 8 + j], acc2 = acc;
            for (int s = 0; s < ksteps; ++s) {
              const double* a = &A[((size_t)w * ksteps + s) * 32 + i * 4];
              const double* b = &B[((size_t)w * ksteps + s) * 32 + j];
              double sum = a[0] * b[0];
              for (int q = 1; q < 4; ++q) sum = std::fma(a[q], b[q * 8], sum);
              acc = acc + sum;
              acc2 = std::fma(a[1], b[8], std::fma(a[0], b[0], acc2)) + std::fma(a[3], b[24], a[2] * b[16]);
            }
            uint64_t r1, r2; std::memcpy(&r1, &acc, 8); std::memcpy(&r2, &acc2, 8);
            match_sumfirst += (r1 == got); match_pair += (r2 == got);
          }
        }
    printf("\"ksteps_%d\": {\"elements\": %lld, \"ascending_chain_matches\": %lld, \"descending_chain_matches\": %lld, "
           "\"products_summed_first_matches\": %lld, \"pairwise_matches\": %lld, \"per_order\": [",
           ksteps, elems, match[0], match[23], match_sumfirst, match_pair);
    for (int p = 0; p < 24; ++p) printf("%s%lld", p ? ", " : "", match[p]);
    printf("]}, ");
    if (match[0] != elems) all_ok = 0;
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD);
  }
  printf("\"dmma_is_ascending_fma_chain\": %s}\n", all_ok ? "true" : "false");
  return 0;
}

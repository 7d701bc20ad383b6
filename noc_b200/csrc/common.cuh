// common.cuh — small device helpers shared by the kernels of libnoc_b200.so (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace noc {

// Every fused multiply-add the spec calls an "fma chain" goes through this; the translation unit is
// compiled with --fmad=false so nothing else is contracted (bitwise parity with the oracle).
__device__ __forceinline__ double fmad(double a, double b, double c) { return __fma_rn(a, b, c); }

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}

// 16-byte asynchronous global->shared copy that bypasses L1 (the A_k,B_k stream is read exactly once).
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ double2 lds128(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts128(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
// streaming 16-byte global store (gains are written once and not re-read by this kernel)
__device__ __forceinline__ void stg128_cs(double* p, double a, double b) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}

// -x by flipping the sign bit on the integer pipe (a DADD with a negated operand would take an FP64 issue slot; the gain
// solves negate their results last)
__device__ __forceinline__ double neg_bits(double x) {
  return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Adaptive Levenberg-Marquardt schedule of SLQ (oracle: reg_increase / reg_decrease): delta0 = 2,
// mu_min = 1e-6.  Multiplications and comparisons only, so both sides agree bit for bit.
__device__ __forceinline__ void reg_increase(double& mu, double& delta) {
  const double d = delta * 2.0;
  delta = (d > 2.0) ? d : 2.0;
  const double v = mu * delta;
  mu = (v > 1e-6) ? v : 1e-6;
}
__device__ __forceinline__ void reg_decrease(double& mu, double& delta) {
  const double d = delta * 0.5;
  delta = (d < 0.5) ? d : 0.5;
  const double v = mu * delta;
  mu = (v >= 1e-6) ? v : 0.0;
}

}  // namespace noc

// vic_math.cuh -- bit-level helpers and the explicit fused multiply-add shared by the elementary functions of vic_glibm.cuh.
// dl::fma is DFMA on the device and __builtin_fma on the host (the FMA instruction, or glibc's exact fma() where the CPU has
// none): both are the correctly rounded a*b+c.  Everything else in the physics headers is compiled WITHOUT contraction
// (nvcc --fmad=false, g++ -ffp-contract=off), as the reference is on x86-64.
#ifndef VIC_MATH_CUH
#define VIC_MATH_CUH
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VM_FN inline __host__ __device__ __noinline__
#define VM_IN __host__ __device__ __forceinline__
#else
#define VM_FN inline
#define VM_IN inline
#endif

namespace vic {
namespace dl {

VM_IN uint64_t bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
VM_IN double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
VM_IN double fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
VM_IN double qnan() { return from_bits(0x7ff8000000000000ULL); }
VM_IN double pinf() { return from_bits(0x7ff0000000000000ULL); }

}  // namespace dl
}  // namespace vic
#endif

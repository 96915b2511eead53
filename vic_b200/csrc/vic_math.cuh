// vic_math.cuh -- the elementary functions of the hot path (exp, log, log10, pow, sin, cos, acos) written in
// plain IEEE-754 double arithmetic (+, -, *, /, sqrt and EXPLICIT fused multiply-adds, no tables), so that the SAME
// sequence of roundings runs on the device (nvcc --fmad=false: no contraction; dl::fma is the DFMA instruction) and on the
// host (g++ -ffp-contract=off; dl::fma is __builtin_fma: the FMA instruction, or glibc's exact fma() where the CPU has
// none -- both are the correctly rounded a*b+c, so the results do not depend on which is used).
//
// Why: the reference model is full of exact comparisons on computed values (surface temperature == 0, snow
// store < threshold, first sunlit 30-second slot of a day, Brent branch tests).  A last-ulp difference between
// two math libraries flips such a test once per ~1e5..1e6 HRU-steps, and from then on the two trajectories of
// that HRU differ macroscopically -- between CUDA's libdevice and glibc exactly as between any two CPUs' libms.
// With one shared implementation the GPU result is BIT-IDENTICAL to the reference physics linked against the
// same functions (oracle/_ref/vic_ref_harness_dl), for any run length; against the glibc-linked reference build
// it agrees to ~1e-13 until the first such flip (tests/).  Accuracy: exp, log, sin, cos <= 2 ulp; acos, log10
// <= 4 ulp; pow(x, y) = exp(y log x) with log x and the product carried in two doubles: <= 1.5 ulp
// (oracle/mathcheck.cpp measures all of them against glibc in long double).
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

// 2^k for k in [-1022, 1023]
VM_IN double pow2i(int k) { return from_bits((uint64_t)(k + 1023) << 52); }

// x * 2^k with gradual underflow / overflow handled in two steps
VM_IN double scale2(double x, int k) {
  if (k > 1023) {
    x *= pow2i(1023);
    k -= 1023;
    if (k > 1023) k = 1023;
  } else if (k < -1022) {
    x *= pow2i(-969);  // keep x normal while scaling down
    k += 969;
    if (k < -1022) k = -1022;
  }
  return x * pow2i(k);
}

// round to nearest integer (|x| < 2^51), ties to even, without relying on the rounding-mode intrinsics
VM_IN double rint52(double x) {
  const double big = 6755399441055744.0;  // 1.5 * 2^52
  return (x + big) - big;
}

// e^(x + xl), |xl| << |x|: the second argument carries the low part of an extended-precision exponent (pow)
VM_IN double exp_ext(double x, double xl) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, inv_ln2 = 1.44269504088896338700e+00;
  if (x != x) return x;
  if (x > 709.782712893384) return pinf();
  if (x < -745.1332191019412) return 0.0;
  const double kd = rint52(x * inv_ln2);
  const int k = (int)kd;
  const double hi = dl::fma(-kd, ln2_hi, x);  // exact: kd * ln2_hi has <= 43 significant bits and cancels the leading bits of x
  const double r = dl::fma(-kd, ln2_lo, hi);  // |r| <= 0.3466
  const double rl = dl::fma(-kd, ln2_lo, hi - r) + xl;  // what the rounding of r dropped, plus the caller's low part
  // e^r = 1 + r + r^2/2! + ... + r^13/13!
  double p = 1.0 / 6227020800.0;
  p = dl::fma(p, r, 1.0 / 479001600.0);
  p = dl::fma(p, r, 1.0 / 39916800.0);
  p = dl::fma(p, r, 1.0 / 3628800.0);
  p = dl::fma(p, r, 1.0 / 362880.0);
  p = dl::fma(p, r, 1.0 / 40320.0);
  p = dl::fma(p, r, 1.0 / 5040.0);
  p = dl::fma(p, r, 1.0 / 720.0);
  p = dl::fma(p, r, 1.0 / 120.0);
  p = dl::fma(p, r, 1.0 / 24.0);
  p = dl::fma(p, r, 1.0 / 6.0);
  p = dl::fma(p, r, 0.5);
  const double q = (r * r) * p;
  const double e = 1.0 + (r + (q + dl::fma(rl, r + q, rl)));  // e^(r + rl) = e^r (1 + rl)
  return scale2(e, k);
}
VM_FN double exp(double x) { return exp_ext(x, 0.0); }

// log(x) = k ln2 + log(1+f), x = 2^k (1+f), sqrt(1/2) <= 1+f < sqrt(2); log(1+f) = f - (hfsq - s (hfsq + R)),
// s = f/(2+f), R = sum_{n>=1} 2 s^(2n) / (2n+1)
VM_FN double log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  if (x != x) return x;
  if (x < 0.0) return qnan();
  if (x == 0.0) return -pinf();
  if (x == pinf()) return x;
  int k = 0;
  uint64_t u = bits(x);
  if ((u >> 52) == 0) {  // subnormal
    x *= 18014398509481984.0;  // 2^54
    k -= 54;
    u = bits(x);
  }
  k += (int)(u >> 52) - 1023;
  u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;  // mantissa in [1, 2)
  double m = from_bits(u);
  if (m > 1.4142135623730951) {
    m *= 0.5;
    k += 1;
  }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double R = 2.0 / 23.0;
  R = dl::fma(R, z, 2.0 / 21.0);
  R = dl::fma(R, z, 2.0 / 19.0);
  R = dl::fma(R, z, 2.0 / 17.0);
  R = dl::fma(R, z, 2.0 / 15.0);
  R = dl::fma(R, z, 2.0 / 13.0);
  R = dl::fma(R, z, 2.0 / 11.0);
  R = dl::fma(R, z, 2.0 / 9.0);
  R = dl::fma(R, z, 2.0 / 7.0);
  R = dl::fma(R, z, 2.0 / 5.0);
  R = dl::fma(R, z, 2.0 / 3.0);
  R = R * z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  return dl::fma(dk, ln2_hi, -((hfsq - dl::fma(s, hfsq + R, dk * ln2_lo)) - f));
}

VM_FN double log10(double x) {
  const double inv_ln10 = 4.34294481903251816668e-01;
  return dl::log(x) * inv_ln10;
}

// ---- extended-precision helpers for pow ----
VM_IN void two_sum(double a, double b, double* s, double* e) {
  const double t = a + b;
  const double bb = t - a;
  *e = (a - (t - bb)) + (b - bb);
  *s = t;
}
VM_IN void two_prod(double a, double b, double* p, double* e) {
  const double t = a * b;
  *e = dl::fma(a, b, -t);
  *p = t;
}
// log(x) = *hi + *lo to ~2^-60 relative, x positive and finite
VM_IN void log_ext(double x, double* hi, double* lo) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  int k = 0;
  uint64_t u = bits(x);
  if ((u >> 52) == 0) {
    x *= 18014398509481984.0;
    k -= 54;
    u = bits(x);
  }
  k += (int)(u >> 52) - 1023;
  u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double m = from_bits(u);
  if (m > 1.4142135623730951) {
    m *= 0.5;
    k += 1;
  }
  const double f = m - 1.0;  // exact
  double dh, dl_;
  two_sum(2.0, f, &dh, &dl_);  // 2 + f = dh + dl_
  const double sh = f / dh;
  const double sl = dl::fma(-sh, dl_, dl::fma(-sh, dh, f)) / dh;  // s = f / (2 + f) = sh + sl
  const double z = sh * sh;
  // 2 atanh(s) = 2 s + s R,  R = 2 z/3 + 2 z^2/5 + ...
  double R = 2.0 / 25.0;
  R = dl::fma(R, z, 2.0 / 23.0);
  R = dl::fma(R, z, 2.0 / 21.0);
  R = dl::fma(R, z, 2.0 / 19.0);
  R = dl::fma(R, z, 2.0 / 17.0);
  R = dl::fma(R, z, 2.0 / 15.0);
  R = dl::fma(R, z, 2.0 / 13.0);
  R = dl::fma(R, z, 2.0 / 11.0);
  R = dl::fma(R, z, 2.0 / 9.0);
  R = dl::fma(R, z, 2.0 / 7.0);
  R = dl::fma(R, z, 2.0 / 5.0);
  R = dl::fma(R, z, 2.0 / 3.0);
  R = R * z;
  const double dk = (double)k;
  // k ln2_hi is exact (ln2_hi has 33 significant bits); sum the two leading terms exactly, the rest in double
  double h, e;
  two_sum(dk * ln2_hi, 2.0 * sh, &h, &e);
  const double tail = e + dl::fma(dk, ln2_lo, dl::fma(sh, R, 2.0 * sl));
  const double hh = h + tail;
  *lo = (h - hh) + tail;
  *hi = hh;
}

// pow(x, y) = exp(y log x) with the logarithm and the product carried in two doubles: ~1 ulp
VM_FN double pow_pos(double x, double y) {
  double lh, ll;
  log_ext(x, &lh, &ll);
  double ph, pl;
  two_prod(y, lh, &ph, &pl);
  if (!(ph > -1.0e300 && ph < 1.0e300)) return dl::exp(ph);  // overflow / underflow / NaN: plain path
  return exp_ext(ph, dl::fma(y, ll, pl));
}

VM_FN double pow(double x, double y) {
  // exact cases first (a compiler may strength-reduce these on the reference side; the results coincide)
  if (y == 0.0) return 1.0;
  if (x == 1.0) return 1.0;
  if (y == 1.0) return x;
  if (y == 2.0) return x * x;
  if (y == -1.0) return 1.0 / x;
  if (x != x || y != y) return qnan();
  const double yi = rint52(y);
  const bool y_int = (y == yi) && (y > -4503599627370496.0 && y < 4503599627370496.0);
  const bool y_odd = y_int && ((double)(2.0 * rint52(0.5 * yi)) != yi);
  if (x == 0.0) {
    if (y > 0.0) return y_odd ? x : 0.0;
    return (y_odd && bits(x) >> 63) ? -pinf() : pinf();
  }
  if (x == pinf()) return y > 0.0 ? pinf() : 0.0;
  if (x < 0.0) {
    if (!y_int && y > -4503599627370496.0 && y < 4503599627370496.0) return qnan();
    if (x == -pinf()) return y > 0.0 ? (y_odd ? -pinf() : pinf()) : 0.0;
    const double r = pow_pos(-x, y);
    return y_odd ? -r : r;
  }
  return pow_pos(x, y);
}

// Cody-Waite reduction by pi/2 in three pieces; valid for |x| up to ~1e5 (arguments on this path are < 10)
VM_IN double reduce_pio2(double x, int* q) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double p2 = 6.07710050630396597660e-11;  // next 33 bits
  const double p3 = 2.02226624871116645580e-21;  // next
  const double p3t = 8.47842766036889956997e-32;
  const double n = rint52(x * two_over_pi);
  *q = (int)((long long)n & 3);
  double r = dl::fma(-n, p1, x);
  r = dl::fma(-n, p2, r);
  r = dl::fma(-n, p3, r);
  r = dl::fma(-n, p3t, r);
  return r;
}
VM_IN double sin_kernel(double r) {  // |r| <= pi/4: r - r^3/3! + ... - r^19/19!
  const double z = r * r;
  double p = -1.0 / 121645100408832000.0;
  p = dl::fma(p, z, 1.0 / 355687428096000.0);
  p = dl::fma(p, z, -(1.0 / 1307674368000.0));
  p = dl::fma(p, z, 1.0 / 6227020800.0);
  p = dl::fma(p, z, -(1.0 / 39916800.0));
  p = dl::fma(p, z, 1.0 / 362880.0);
  p = dl::fma(p, z, -(1.0 / 5040.0));
  p = dl::fma(p, z, 1.0 / 120.0);
  p = dl::fma(p, z, -(1.0 / 6.0));
  return dl::fma(r, z * p, r);
}
VM_IN double cos_kernel(double r) {  // 1 - r^2/2! + ... + r^20/20!
  const double z = r * r;
  double p = 1.0 / 2432902008176640000.0;
  p = dl::fma(p, z, -(1.0 / 6402373705728000.0));
  p = dl::fma(p, z, 1.0 / 20922789888000.0);
  p = dl::fma(p, z, -(1.0 / 87178291200.0));
  p = dl::fma(p, z, 1.0 / 479001600.0);
  p = dl::fma(p, z, -(1.0 / 3628800.0));
  p = dl::fma(p, z, 1.0 / 40320.0);
  p = dl::fma(p, z, -(1.0 / 720.0));
  p = dl::fma(p, z, 1.0 / 24.0);
  const double hz = 0.5 * z;
  return dl::fma(z, z * p, 1.0 - hz);
}
VM_FN double sin(double x) {
  if (x != x || x == pinf() || x == -pinf()) return qnan();
  int q;
  const double r = reduce_pio2(x, &q);
  switch (q) {
    case 0: return sin_kernel(r);
    case 1: return cos_kernel(r);
    case 2: return -sin_kernel(r);
    default: return -cos_kernel(r);
  }
}
VM_FN double cos(double x) {
  if (x != x || x == pinf() || x == -pinf()) return qnan();
  int q;
  const double r = reduce_pio2(x, &q);
  switch (q) {
    case 0: return cos_kernel(r);
    case 1: return -sin_kernel(r);
    case 2: return -cos_kernel(r);
    default: return sin_kernel(r);
  }
}

// asin on |x| <= 0.5 by its Maclaurin series: x * (1 + sum_n c_n z^n), z = x^2, c_n = (2n)! / (4^n (n!)^2 (2n+1))
VM_IN double asin_small(double x) {
  const double z = x * x;
  // c_n / c_{n-1} = (2n-1)^2 / (2n (2n+1)); evaluate by Horner from n = 26 down (0.25^27 c_27 < 1e-18)
  double p = 0.0;
  for (int n = 26; n >= 1; n--) {
    const double a = (double)(2 * n - 1), b = (double)(2 * n);
    p = (1.0 + p) * z * ((a * a) / (b * (b + 1.0)));
  }
  return x + x * p;
}
VM_FN double acos(double x) {
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  if (x != x) return x;
  const double ax = fabs(x);
  if (ax > 1.0) return qnan();
  if (ax <= 0.5) return (pio2_hi - (asin_small(x) - pio2_lo));
  // acos(x) = 2 asin(sqrt((1-x)/2)) for x > 0.5;  pi - that for x < -0.5
  const double t = sqrt((1.0 - ax) * 0.5);
  const double a = 2.0 * asin_small(t);
  if (x > 0.0) return a;
  return (2.0 * pio2_hi - (a - 2.0 * pio2_lo));
}

}  // namespace dl
}  // namespace vic
#endif

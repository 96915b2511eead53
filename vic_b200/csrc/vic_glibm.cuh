// vic_glibm.cuh -- exp, log, log10, pow, sin, cos, acos with EXACTLY the results of glibc 2.39's libm (x86-64,
// the FMA variants its ifunc resolvers select on every CPU with FMA + AVX2: __exp_fma, __log_fma, __pow_fma,
// __sin_fma, __cos_fma, __ieee754_acos_fma; log10 is the generic __ieee754_log10 on top of __log_fma).
//
// Why this exists: the reference model is linked against the platform libm and is full of exact comparisons on
// computed values; its trajectories bifurcate on a last-bit difference in pow / exp / cos.
// Being "another accurate libm" therefore cannot meet a 1e-6 annual parity bar over a year of hourly steps -- being
// THE SAME libm can.  glibc's routines are plain IEEE-754 double arithmetic (+, -, *, /, fused multiply-add,
// integer bit manipulation, table lookups): the same sequence of roundings gives the same bits on the device
// (nvcc --fmad=false: no implicit contraction; dl::fma is DFMA) and on the host (g++ -ffp-contract=off).
//
// Source of truth: glibc 2.39 is not in /root/reference (it is the third-party dependency the reference's arithmetic
// lives in: SURVEY.md section 8(c).4).  The algorithms are the published ones (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c,
// e_pow.c by Szabolcs Nagy / ARM optimized-routines; s_sin.c, e_asin.c from the IBM Accurate Mathematical Library,
// e_log10.c from fdlibm); WHICH multiply-adds are fused is a property of the distribution's build (gcc contracts at
// will in the -mfma -mavx2 variants) and was read off the instruction stream of the image's libm.so.6 (Ubuntu GLIBC
// 2.39-0ubuntu8.5).  Tables: vic_glibm_tables.h (tools/gen_glibm_tables.py).  oracle/mathcheck.cpp demands ZERO
// differing bits against the platform's libm on 1e8+ arguments per function (tests/test_cpu.py::test_glibm_*).
//
// Not covered (returns NaN instead of a silently different value): sin / cos of |x| >= 105414350 (glibc switches to
// the __branred big-argument reduction there; the model's arguments are hour angles and day-of-year phases, |x| < 10).
#ifndef VIC_GLIBM_CUH
#define VIC_GLIBM_CUH
#include "vic_math.cuh"
#include "vic_glibm_tables.h"

#if defined(__CUDA_ARCH__)
#define GLT(name) ::vic::gl::tab::name##_d
#else
#define GLT(name) ::vic::gl::tab::name##_h
#endif

namespace vic {
namespace gl {

using dl::bits;
using dl::from_bits;
VM_IN double fma(double a, double b, double c) { return dl::fma(a, b, c); }
VM_IN double fnma(double a, double b, double c) { return dl::fma(-a, b, c); }  // c - a*b, one rounding (vfnmadd)
VM_IN double fms(double a, double b, double c) { return dl::fma(a, b, -c); }   // a*b - c, one rounding (vfmsub)
VM_IN double gabs(double x) { return from_bits(bits(x) & 0x7fffffffffffffffULL); }
VM_IN double gcopysign(double m, double s) { return from_bits((bits(m) & 0x7fffffffffffffffULL) | (bits(s) & 0x8000000000000000ULL)); }
VM_IN uint32_t top12(double x) { return (uint32_t)(bits(x) >> 52); }

// ---- exp (e_exp.c): x = k ln2/128 + r, exp(x) = 2^(k/128) (1 + tail + r + r^2 C2 + ...) ------------------------------
namespace ec {
constexpr double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p52, NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
constexpr double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
}  // namespace ec

// results that over- or underflow the exponent of `scale` (e_exp.c specialcase(); here nothing is fused except where noted)
VM_IN double exp_special(double tmp, uint64_t sbits, uint64_t ki) {
  if ((ki & 0x80000000ULL) == 0) {  // k > 0: the exponent of scale may have overflowed by <= 460
    sbits -= 1009ULL << 52;
    const double scale = from_bits(sbits);
    return 0x1p1009 * fma(scale, tmp, scale);
  }
  sbits += 1022ULL << 52;  // k < 0: care in the subnormal range
  const double scale = from_bits(sbits);
  const double st = scale * tmp;
  double y = scale + st;
  if (gabs(y) < 1.0) {  // scale (and y) is negative for pow(negative x, odd y)
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = (scale - y) + st;
    const double hi = one + y;
    lo = ((one - hi) + y) + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = from_bits(sbits & 0x8000000000000000ULL);  // the sign of 0
  }
  return 0x1p-1022 * y;
}

// the polynomial and scaling shared by exp and pow: r reduced argument, ki the shifted integer part (+ sign bias for pow)
VM_IN double exp_core(double r, uint64_t ki, uint64_t kbias, uint32_t abstop) {
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = (ki + kbias) << 45;
  const double tail = from_bits(GLT(exp_tab)[idx]);
  const uint64_t sbits = GLT(exp_tab)[idx + 1] + top;
  const double r2 = r * r;
  const double p23 = fma(ec::C3, r, ec::C2);
  const double p45 = fma(r, ec::C5, ec::C4);
  const double lin = r + tail;
  double tmp = fma(p23, r2, lin);
  tmp = fma(r2 * r2, p45, tmp);
  if (abstop == 0) return exp_special(tmp, sbits, ki);
  const double scale = from_bits(sbits);
  return fma(scale, tmp, scale);
}

VM_FN double exp(double x) {
  uint32_t abstop = top12(x) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {  // |x| < 2^-54, |x| >= 512, inf, nan
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;
    if (abstop >= 0x409u) {  // |x| >= 1024
      if (bits(x) == 0xfff0000000000000ULL) return 0.0;
      if (abstop >= 0x7ffu) return 1.0 + x;
      return (bits(x) >> 63) ? 0.0 : dl::pinf();
    }
    abstop = 0;
  }
  const double z = fma(x, ec::InvLn2N, ec::Shift);
  const uint64_t ki = bits(z);
  const double kd = z - ec::Shift;
  double r = fma(kd, ec::NegLn2hiN, x);
  r = fma(kd, ec::NegLn2loN, r);
  return exp_core(r, ki, 0, abstop);
}

// ---- log (e_log.c) ----------------------------------------------------------------------------------------------------
namespace lc {
constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
constexpr double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3, A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
constexpr double B0 = -0x1p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3, B3 = 0x1.999999995dd0cp-3, B4 = -0x1.55555556745a7p-3,
                 B5 = 0x1.24924a344de30p-3, B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4,
                 B9 = 0x1.78182f7afd085p-4, B10 = -0x1.5521375d145cdp-4;
}  // namespace lc

VM_FN double log(double x) {
  uint64_t ix = bits(x);
  const uint32_t top = (uint32_t)(ix >> 48);
  if (ix - 0x3fee000000000000ULL < 0x3090000000000ULL) {  // 1 - 0x1p-4 <= x < 1 + 0x1.09p-4
    if (ix == 0x3ff0000000000000ULL) return 0.0;
    const double r = x - 1.0;
    const double r2 = r * r;
    const double r3 = r * r2;
    const double q1 = fma(r2, lc::B3, fma(r, lc::B2, lc::B1));
    const double q2 = fma(r2, lc::B6, fma(r, lc::B5, lc::B4));
    double q3 = fma(r2, lc::B9, fma(r, lc::B8, lc::B7));
    q3 = fma(r3, lc::B10, q3);
    double y = fma(q3, r3, q2);
    y = fma(y, r3, q1);
    const double rw = fma(r, 0x1p27, r);          // r + r * 2^27
    const double rhi = fnma(0x1p27, r, rw);       // (r + w) - w
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = fma(rhi2, lc::B0, r);       // r + w', w' = rhi^2 * B0
    const double rmh = r - hi;
    const double rs = r + rhi;
    double lo = fma(rhi2, lc::B0, rmh);
    lo = fma(lc::B0 * rlo, rs, lo);
    y = fma(y, r3, lo);
    return hi + y;
  }
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {  // x < 2^-1022, inf, nan
    if (ix * 2 == 0) return -dl::pinf();
    if (ix == 0x7ff0000000000000ULL) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return (x - x) / (x - x);
    ix = bits(x * 0x1p52);
    ix -= 52ULL << 52;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const int i = (int)((tmp >> 45) % 128);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  const double invc = GLT(log_tab)[2 * i], logc = GLT(log_tab)[2 * i + 1];
  const double z = from_bits(iz);
  const double kd = (double)k;
  const double w = fma(kd, lc::Ln2hi, logc);
  const double r = fma(z, invc, -1.0);
  const double p12 = fma(r, lc::A2, lc::A1);
  const double hi = r + w;
  const double r2 = r * r;
  double lo = (w - hi) + r;
  lo = fma(kd, lc::Ln2lo, lo);
  const double r3 = r * r2;
  const double p34 = fma(r, lc::A4, lc::A3);
  lo = fma(r2, lc::A0, lo);
  const double p = fma(p34, r2, p12);
  return fma(r3, p, lo) + hi;
}

// ---- log10 (e_log10.c, generic build: nothing fused) ------------------------------------------------------------------
VM_FN double log10(double x) {
  const double two54 = 0x1p54, ivln10 = 0x1.bcb7b1526e50ep-2, log10_2hi = 0x1.34413509f6000p-2, log10_2lo = 0x1.9fef311f12b36p-42;
  int64_t hx = (int64_t)bits(x);
  int32_t k = 0;
  if (hx < 0x0010000000000000LL) {  // x < 2^-1022
    if ((hx & 0x7fffffffffffffffLL) == 0) return -two54 / gabs(x);
    if (hx < 0) return (x - x) / (x - x);
    k -= 54;
    x *= two54;
    hx = (int64_t)bits(x);
  }
  if ((uint64_t)hx > 0x7fefffffffffffffULL) return x + x;
  k += (int32_t)(hx >> 52) - 1023;
  const int64_t i = ((uint64_t)(int64_t)k) >> 63;
  hx = (hx & 0x000fffffffffffffLL) | ((0x3ff - i) << 52);
  const double y = (double)(k + i);
  const double z = y * log10_2lo + ivln10 * gl::log(from_bits((uint64_t)hx));
  return z + y * log10_2hi;
}

// ---- pow (e_pow.c) ------------------------------------------------------------------------------------------------------
namespace pc {
constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
constexpr double A0 = -0x1p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1, A3 = 0x1.999999959554ep-1, A4 = -0x1.555555529a47ap-1,
                 A5 = -0x1.2495b9b4845e9p+0, A6 = 0x1.0002b8b263fc3p+0;
}  // namespace pc

// 0: not an integer, 1: odd integer, 2: even integer
VM_IN int pow_checkint(uint64_t iy) {
  const int e = (int)((iy >> 52) & 0x7ff);
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
  return 2;
}
VM_IN bool pow_zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * 0x7ff0000000000000ULL - 1; }

VM_FN double pow(double x, double y) {
  uint64_t sign_bias = 0;
  uint64_t ix = bits(x);
  const uint64_t iy = bits(y);
  uint32_t topx = top12(x);
  const uint32_t topy = top12(y);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    if (pow_zeroinfnan(iy)) {
      if (2 * iy == 0) return 1.0;
      if (ix == 0x3ff0000000000000ULL) return 1.0;
      if (2 * ix > 2 * 0x7ff0000000000000ULL || 2 * iy > 2 * 0x7ff0000000000000ULL) return x + y;
      if (2 * ix == 2 * 0x3ff0000000000000ULL) return 1.0;
      if ((2 * ix < 2 * 0x3ff0000000000000ULL) == !(iy >> 63)) return 0.0;
      return y * y;
    }
    if (pow_zeroinfnan(ix)) {
      double x2 = x * x;
      if ((ix >> 63) && pow_checkint(iy) == 1) x2 = -x2;
      return (iy >> 63) ? 1 / x2 : x2;
    }
    if (ix >> 63) {  // finite x < 0
      const int yint = pow_checkint(iy);
      if (yint == 0) return (x - x) / (x - x);
      if (yint == 1) sign_bias = 0x800ULL << 7;
      ix &= 0x7fffffffffffffffULL;
      topx &= 0x7ff;
    }
    if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
      if (ix == 0x3ff0000000000000ULL) return 1.0;
      if ((topy & 0x7ff) < 0x3beu) return ix > 0x3ff0000000000000ULL ? 1.0 + y : 1.0 - y;
      return (ix > 0x3ff0000000000000ULL) == (topy < 0x800u) ? dl::pinf() : 0.0;
    }
    if (topx == 0) {  // subnormal x
      ix = bits(x * 0x1p52);
      ix &= 0x7fffffffffffffffULL;
      ix -= 52ULL << 52;
    }
  }
  // log_inline: hi + lo = log(x) to ~68 bits
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int)((tmp >> 45) % 128);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  const double z = from_bits(iz);
  const double kd = (double)k;
  const double invc = GLT(powlog_tab)[4 * i], logc = GLT(powlog_tab)[4 * i + 2], logctail = GLT(powlog_tab)[4 * i + 3];
  const double t1 = fma(kd, pc::Ln2hi, logc);
  const double lo1 = fma(kd, pc::Ln2lo, logctail);
  const double r = fma(z, invc, -1.0);
  const double ar = r * pc::A0;
  const double p12 = fma(r, pc::A2, pc::A1);
  const double p34 = fma(r, pc::A4, pc::A3);
  const double t2 = r + t1;
  const double lo2 = (t1 - t2) + r;
  const double ar2 = r * ar;
  const double ar3 = r * ar2;
  const double lo3 = fms(ar, r, ar2);
  const double lhi = t2 + ar2;
  const double p56 = fma(r, pc::A6, pc::A5);
  const double lo4 = (t2 - lhi) + ar2;
  double p = fma(p56, ar2, p34);
  p = fma(ar2, p, p12);
  double lo = lo1 + lo2;
  lo = lo + lo3;
  lo = lo + lo4;
  lo = fma(ar3, p, lo);
  const double lg = lhi + lo;
  const double lgtail = (lhi - lg) + lo;
  const double ehi = y * lg;
  const double elo = fma(y, lgtail, fms(lg, y, ehi));
  // exp_inline(ehi, elo, sign_bias)
  uint32_t abstop = top12(ehi) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) {
      const double one = 1.0 + ehi;
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) {
      const double s = sign_bias ? -1.0 : 1.0;
      return (bits(ehi) >> 63) ? s * 0.0 : s * dl::pinf();
    }
    abstop = 0;
  }
  const double zz = fma(ehi, ec::InvLn2N, ec::Shift);
  const uint64_t ki = bits(zz);
  const double kd2 = zz - ec::Shift;
  double rr = fma(kd2, ec::NegLn2hiN, ehi);
  rr = fma(kd2, ec::NegLn2loN, rr);
  rr = elo + rr;
  return exp_core(rr, ki, sign_bias, abstop);
}

// ---- sin / cos (s_sin.c) ------------------------------------------------------------------------------------------------
namespace sc {
constexpr double big = 0x1.8p45, t126 = 0x1.020c49ba5e354p-3;
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7, cs2 = 0x1p-1, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13, s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
constexpr double hpinv = 0x1.45f306dc9c883p-1, toint = 0x1.8p52;
constexpr double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27, pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
}  // namespace sc

VM_IN double sc_taylor_sin(double xx, double x, double dx) {
  double p = fma(xx, sc::s5, sc::s4);
  p = fma(xx, p, sc::s3);
  p = fma(xx, p, sc::s2);
  p = fma(xx, p, sc::s1);
  const double t = fma(xx, fms(p, x, 0.5 * dx), dx);
  return x + t;
}

// cos(x + dx) from the tables: x = xi + r, xi = k/128
VM_IN double sc_do_cos(double x, double dx) {
  if (x < 0) dx = -dx;
  const double ax = gabs(x);
  const double u = sc::big + ax;
  x = (ax - (u - sc::big)) + dx;
  const double xx = x * x;
  const double s = fma(x * xx, fma(xx, sc::sn5, sc::sn3), x);
  const double c = xx * fma(xx, fma(xx, sc::cs6, sc::cs4), sc::cs2);
  const int k = (int)((uint32_t)bits(u) << 2);
  const double sn = GLT(sincos_tab)[k], ssn = GLT(sincos_tab)[k + 1], cs = GLT(sincos_tab)[k + 2], ccs = GLT(sincos_tab)[k + 3];
  double cor = fnma(s, ssn, ccs);
  cor = fnma(c, cs, cor);
  cor = fnma(s, sn, cor);
  return cs + cor;
}

// sin(x + dx)
VM_IN double sc_do_sin(double x, double dx) {
  const double xold = x;
  const double ax = gabs(x);
  if (ax < sc::t126) return sc_taylor_sin(x * x, x, dx);
  if (x <= 0) dx = -dx;
  const double u = sc::big + ax;
  x = ax - (u - sc::big);
  const double xx = x * x;
  const double s = x + fma(x * xx, fma(xx, sc::sn5, sc::sn3), dx);
  const double c = fma(x, dx, xx * fma(xx, fma(xx, sc::cs6, sc::cs4), sc::cs2));
  const int k = (int)((uint32_t)bits(u) << 2);
  const double sn = GLT(sincos_tab)[k], ssn = GLT(sincos_tab)[k + 1], cs = GLT(sincos_tab)[k + 2], ccs = GLT(sincos_tab)[k + 3];
  double cor = fma(s, ccs, ssn);
  cor = fnma(c, sn, cor);
  cor = fma(s, cs, cor);
  return gcopysign(sn + cor, xold);
}

// x = n pi/2 + (a + da), |x| < 105414350
VM_IN int sc_reduce(double x, double* a, double* da) {
  const double t = fma(x, sc::hpinv, sc::toint);
  const double xn = t - sc::toint;
  const int n = (int)(bits(t) & 3);
  double y = fnma(xn, sc::mp1, x);
  y = fnma(xn, sc::mp2, y);
  const double t2 = fnma(xn, sc::pp3, y);
  double d = fnma(sc::pp3, xn, y - t2);
  const double b = fnma(xn, sc::pp4, t2);
  d = d + fnma(xn, sc::pp4, t2 - b);
  *a = b;
  *da = d;
  return n;
}

VM_IN double sc_do_sincos(double a, double da, int n) {
  const double r = (n & 1) ? sc_do_cos(a, da) : sc_do_sin(a, da);
  return (n & 2) ? -r : r;
}

VM_FN double sin(double x) {
  const uint32_t k = (uint32_t)(bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                       // |x| < 2^-26
  if (k < 0x3feb6000u) return sc_do_sin(x, 0.0);       // |x| < 0.855469
  if (k < 0x400368fdu) {                               // |x| < 2.426265
    const double t = sc::hp0 - gabs(x);
    return gcopysign(sc_do_cos(t, sc::hp1), x);
  }
  if (k < 0x419921fbu) {                               // |x| < 105414350
    double a, da;
    const int n = sc_reduce(x, &a, &da);
    return sc_do_sincos(a, da, n);
  }
  return dl::qnan();  // big-argument reduction not restated (see header); inf / nan -> nan as in glibc
}

VM_FN double cos(double x) {
  const uint32_t k = (uint32_t)(bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                     // |x| < 2^-27
  if (k < 0x3feb6000u) return sc_do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = sc::hp0 - gabs(x);
    const double a = y + sc::hp1;
    const double da = (y - a) + sc::hp1;
    return sc_do_sin(a, da);
  }
  if (k < 0x419921fbu) {
    double a, da;
    const int n = sc_reduce(x, &a, &da);
    return sc_do_sincos(a, da, n + 1);
  }
  return dl::qnan();
}

// ---- acos (e_asin.c) ----------------------------------------------------------------------------------------------------
namespace ac {
constexpr double f1 = 0x1.55555555554f9p-3, f2 = 0x1.333333336127dp-4, f3 = 0x1.6db6dae42c0e4p-5, f4 = 0x1.f1c7e04f4ad99p-6, f5 = 0x1.6e442c822d419p-6,
                 f6 = 0x1.292d80f453c72p-6;
constexpr double rt0 = 0x1.fffffffecc1ddp-1, rt1 = 0x1.fffffff757304p-2, rt2 = 0x1.800496769c91ap-2, rt3 = 0x1.4006318d1dab9p-2;
constexpr double pi = 0x1.921fb54442d18p+1;
}  // namespace ac

// Taylor expansion of asin around the tabulated point asncs[n]: NP coefficients in Horner form, then the xx^2 and xx
// terms; returns t, *e = the tabulated asin value
template <int NP>
VM_IN double ac_table(int n, double x, double* e) {
  const double xx = x - GLT(asncs_tab)[n];
  double p = GLT(asncs_tab)[n + NP + 1];
#pragma unroll
  for (int j = NP; j >= 2; --j) p = fma(xx, p, GLT(asncs_tab)[n + j]);
  p = fma(xx * xx, p, GLT(asncs_tab)[n + NP + 2]);
  *e = GLT(asncs_tab)[n + NP + 3];
  return fma(xx, GLT(asncs_tab)[n + 1], p);
}

VM_FN double acos(double x) {
  const int32_t m = (int32_t)(bits(x) >> 32);
  const int32_t k = m & 0x7fffffff;
  if (k < 0x3c880000) return sc::hp0;  // |x| < 2^-55
  if (k < 0x3fc00000) {                // |x| < 0.125
    const double x2 = x * x;
    double p = fma(x2, ac::f6, ac::f5);
    p = fma(x2, p, ac::f4);
    p = fma(x2, p, ac::f3);
    p = fma(x2, p, ac::f2);
    p = fma(x2, p, ac::f1);
    const double r = sc::hp0 - x;
    const double x3 = x * x2;
    double cor = ((sc::hp0 - r) - x) + sc::hp1;
    cor = fnma(p, x3, cor);
    return r + cor;
  }
  if (k < 0x3fef0000) {  // 0.125 <= |x| < 0.96875: expansions around tabulated points
    const double xa = (m > 0) ? x : -x;
    double t, e;
    if (k < 0x3fe00000) {
      const int n = (k < 0x3fd00000) ? 11 * ((k & 0x000fffff) >> 15) : 11 * ((k & 0x000fffff) >> 14) + 352;
      t = ac_table<5>(n, xa, &e);
    } else if (k < 0x3fe80000) {
      t = ac_table<6>(1056 + 12 * ((k & 0x000fe000) >> 13), xa, &e);
    } else if (k < 0x3fed8000) {
      t = ac_table<7>(992 + 13 * ((k & 0x000fe000) >> 13), xa, &e);
    } else if (k < 0x3fee8000) {
      t = ac_table<8>(884 + 14 * ((k & 0x000fe000) >> 13), xa, &e);
    } else {
      t = ac_table<9>(768 + 15 * ((k & 0x000fe000) >> 13), xa, &e);
    }
    if (m > 0) return (sc::hp1 - t) + (sc::hp0 - e);
    return (t + sc::hp1) + (e + sc::hp0);
  }
  if (k < 0x3ff00000) {  // 0.96875 <= |x| < 1: acos(x) = 2 asin(sqrt((1-|x|)/2))
    const double z = 0.5 * ((m > 0) ? (1.0 - x) : (1.0 + x));
    const int64_t zb = (int64_t)bits(z);
    double t = GLT(inroot_tab)[(zb >> 46) & 0x7f] * GLT(powtwo_tab)[511 - (int)(zb >> 53)];
    const double r = fnma(t * t, z, 1.0);
    double q = fma(r, ac::rt3, ac::rt2);
    q = fma(r, q, ac::rt1);
    q = fma(r, q, ac::rt0);
    t = q * t;
    const double c = z * t;
    const double h = fnma(c, t * 0.5, 1.5);
    const double cw = fma(c, 0x1p27, c);
    const double y = fnma(0x1p27, c, cw);
    const double ty = fma(h, c, y);
    const double cc = fnma(y, y, z) / ty;
    double p = fma(z, ac::f6, ac::f5);
    p = fma(z, p, ac::f4);
    p = fma(z, p, ac::f3);
    p = fma(z, p, ac::f2);
    p = fma(z, p, ac::f1);
    p = p * z;
    const double pc = p * (y + cc);
    if (m < 0) {
      const double a = (sc::hp1 - cc) - pc;
      const double res = a + (sc::hp0 - y);
      return res + res;
    }
    const double res = (cc + pc) + y;
    return res + res;
  }
  if (k == 0x3ff00000 && (uint32_t)bits(x) == 0) return (m > 0) ? 0.0 : ac::pi;
  return (x - x) / (x - x);
}

}  // namespace gl
}  // namespace vic
#endif

// oracle/glibmcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Compares vic_b200/csrc/vic_glibm.cuh (the restatement of glibc 2.39's exp/log/log10/pow/sin/cos/acos) with the
// platform's libm BIT FOR BIT.  Usage: glibmcheck [samples per sweep, default 2e7].  Prints one line per function:
// "<function> <samples> <mismatches>" and up to 5 offending arguments; exit status 1 if any bit differs.
// Only meaningful on an x86-64 CPU with FMA + AVX2 and glibc 2.39 (prints "skip" otherwise): the restatement is of the
// variant glibc's ifunc resolver picks there.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <gnu/libc-version.h>
#include "vic_glibm.cuh"

static inline uint64_t mix(uint64_t z) {  // splitmix64
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t i) { return (double)(mix(i) >> 11) / 9007199254740992.0; }
static inline uint64_t B(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline bool same(double a, double b) { return B(a) == B(b) || (a != a && b != b); }

template <class F, class G, class X>
static long sweep1(const char* name, F f, G g, X gen, long n, long& total_bad) {
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long i = 0; i < n; i++) {
    const double x = gen((uint64_t)i);
    const double a = f(x), b = g(x);
    if (!same(a, b)) {
      bad++;
      if (bad <= 5) {
#pragma omp critical
        fprintf(stderr, "  %s(%a = %.17g): here %a, libm %a\n", name, x, x, a, b);
      }
    }
  }
  printf("%s %ld %ld\n", name, n, bad);
  total_bad += bad;
  return bad;
}

int main(int argc, char** argv) {
  using namespace vic;
  if (!__builtin_cpu_supports("fma") || !__builtin_cpu_supports("avx2") || strcmp(gnu_get_libc_version(), "2.39") != 0) {
    printf("skip: needs glibc 2.39 on an FMA+AVX2 CPU (found glibc %s)\n", gnu_get_libc_version());
    return 0;
  }
  const long N = argc > 1 ? atol(argv[1]) : 20000000L;
  long bad = 0;
  auto lin = [](double lo, double hi, uint64_t salt) { return [=](uint64_t i) { return lo + (hi - lo) * u01(i * 0x9E37ULL + salt); }; };
  auto lg = [](double lo, double hi, uint64_t salt) { return [=](uint64_t i) { return lo * std::exp(std::log(hi / lo) * u01(i * 0x9E37ULL + salt)); }; };
  auto anybits = [](uint64_t salt) { return [=](uint64_t i) { uint64_t u = mix(i + salt); double x; memcpy(&x, &u, 8); return x; }; };
  volatile double (*pexp)(double) = (volatile double (*)(double))0; (void)pexp;
  // exp: the hot path's range, the whole finite range, and raw bit patterns (subnormal results, overflow, nan, inf)
  sweep1("exp", [](double x) { return gl::exp(x); }, [](double x) { return ::exp(x); }, lin(-60, 60, 1), N, bad);
  sweep1("exp_wide", [](double x) { return gl::exp(x); }, [](double x) { return ::exp(x); }, lin(-760, 720, 2), N / 4, bad);
  sweep1("exp_bits", [](double x) { return gl::exp(x); }, [](double x) { return ::exp(x); }, anybits(3), N / 4, bad);
  sweep1("exp_tiny", [](double x) { return gl::exp(x); }, [](double x) { return ::exp(x); }, lg(1e-300, 1e-3, 4), N / 4, bad);
  // log
  sweep1("log", [](double x) { return gl::log(x); }, [](double x) { return ::log(x); }, lg(1e-12, 1e12, 5), N, bad);
  sweep1("log_near1", [](double x) { return gl::log(x); }, [](double x) { return ::log(x); }, lin(0.9, 1.1, 6), N, bad);
  sweep1("log_bits", [](double x) { return gl::log(x); }, [](double x) { return ::log(x); }, anybits(7), N / 4, bad);
  sweep1("log10", [](double x) { return gl::log10(x); }, [](double x) { return ::log10(x); }, lg(1e-12, 1e12, 8), N, bad);
  sweep1("log10_bits", [](double x) { return gl::log10(x); }, [](double x) { return ::log10(x); }, anybits(9), N / 4, bad);
  // sin / cos: every branch below the big-argument reduction
  sweep1("sin", [](double x) { return gl::sin(x); }, [](double x) { return ::sin(x); }, lin(-10, 10, 10), N, bad);
  sweep1("sin_small", [](double x) { return gl::sin(x); }, [](double x) { return ::sin(x); }, lg(1e-10, 3.0, 11), N / 2, bad);
  sweep1("sin_large", [](double x) { return gl::sin(x); }, [](double x) { return ::sin(x); }, lin(-1.0e8, 1.0e8, 12), N / 2, bad);
  sweep1("cos", [](double x) { return gl::cos(x); }, [](double x) { return ::cos(x); }, lin(-10, 10, 13), N, bad);
  sweep1("cos_small", [](double x) { return gl::cos(x); }, [](double x) { return ::cos(x); }, lg(1e-10, 3.0, 14), N / 2, bad);
  sweep1("cos_large", [](double x) { return gl::cos(x); }, [](double x) { return ::cos(x); }, lin(-1.0e8, 1.0e8, 15), N / 2, bad);
  // acos
  sweep1("acos", [](double x) { return gl::acos(x); }, [](double x) { return ::acos(x); }, lin(-1, 1, 16), N, bad);
  sweep1("acos_near1", [](double x) { return gl::acos(x); }, [](double x) { return ::acos(x); },
         [](uint64_t i) { double s = (mix(i + 17) & 1) ? 1.0 : -1.0; return s * (1.0 - 0.04 * u01(i * 3 + 18) * u01(i * 5 + 19)); }, N / 2, bad);
  sweep1("acos_small", [](double x) { return gl::acos(x); }, [](double x) { return ::acos(x); }, lg(1e-20, 0.2, 20), N / 4, bad);
  sweep1("acos_bits", [](double x) { return gl::acos(x); }, [](double x) { return ::acos(x); }, anybits(21), N / 8, bad);
  // pow: two arguments
  {
    struct Case { const char* name; double xlo, xhi; bool xlog; double ylo, yhi; long n; };
    const Case cases[] = {
        {"pow", 1e-6, 1e3, true, -12, 12, N},            // Brooks-Corey / Clapp-Hornberger / albedo decay / stability terms
        {"pow_frac", 1e-4, 1.0, false, 0, 40, N},         // (moist / max_moist)^expt
        {"pow_wide", 1e-300, 1e300, true, -3, 3, N / 4},  // includes subnormal and overflowing results
        {"pow_neg", -100, 100, false, -8, 8, N / 4},      // negative bases: integer exponents are drawn below
    };
    for (const Case& c : cases) {
      long b = 0;
#pragma omp parallel for reduction(+ : b) schedule(static)
      for (long i = 0; i < c.n; i++) {
        const double ux = u01((uint64_t)i * 7 + 101), uy = u01((uint64_t)i * 11 + 202);
        double x = c.xlog ? c.xlo * std::exp(std::log(c.xhi / c.xlo) * ux) : c.xlo + (c.xhi - c.xlo) * ux;
        double y = c.ylo + (c.yhi - c.ylo) * uy;
        if (c.xlo < 0 && (i & 1)) y = std::floor(y);
        if ((i & 1023) == 7) y = 2.0;
        if ((i & 1023) == 8) y = 0.5;
        const double a = gl::pow(x, y), r = ::pow(x, y);
        if (!same(a, r)) {
          b++;
          if (b <= 5) {
#pragma omp critical
            fprintf(stderr, "  pow(%a, %a) = (%.17g, %.17g): here %a, libm %a\n", x, y, x, y, a, r);
          }
        }
      }
      printf("%s %ld %ld\n", c.name, c.n, b);
      bad += b;
    }
    // raw bit patterns in both arguments: every special case
    long b = 0;
    const long n = N / 4;
#pragma omp parallel for reduction(+ : b) schedule(static)
    for (long i = 0; i < n; i++) {
      uint64_t ux = mix((uint64_t)i + 303), uy = mix((uint64_t)i * 3 + 404);
      if ((i & 3) == 1) uy = (uy & 0x800fffffffffffffULL) | ((uint64_t)(0x3f0 + (uy >> 52) % 0x30) << 52);  // |y| in [2^-15, 2^33)
      if ((i & 7) == 2) ux = (ux & 0x800fffffffffffffULL) | ((uint64_t)(0x3f0 + (ux >> 52) % 0x20) << 52);
      double x, y;
      memcpy(&x, &ux, 8);
      memcpy(&y, &uy, 8);
      const double a = gl::pow(x, y), r = ::pow(x, y);
      if (!same(a, r)) {
        b++;
        if (b <= 5) {
#pragma omp critical
          fprintf(stderr, "  pow(%a, %a): here %a, libm %a\n", x, y, a, r);
        }
      }
    }
    printf("pow_bits %ld %ld\n", n, b);
    bad += b;
  }
  return bad ? 1 : 0;
}

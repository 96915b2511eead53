// oracle/mathcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Measures the error of the portable elementary functions (vic_b200/csrc/vic_math.cuh) against glibc evaluated in
// long double, in units in the last place of the double result, over the argument ranges the hot path uses.
// Prints "<function> <max ulp error>" per line (tests/test_cpu.py::test_portable_math_accuracy).
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "vic_math.cuh"

static uint64_t rng_state = 0x9e3779b97f4a7c15ULL;
static double urand() {  // xorshift64*, uniform in [0, 1)
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 0x2545F4914F6CDD1DULL) >> 11) / 9007199254740992.0;
}
static double ulp_err(double got, long double want) {
  if (want == 0.0L) return got == 0.0 ? 0.0 : 1e30;
  int e;
  frexpl(want, &e);
  const long double ulp = ldexpl(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}
template <class F, class G>
static double sweep(F f, G g, double lo, double hi, int n, bool logspace = false) {
  double worst = 0;
  for (int i = 0; i < n; i++) {
    const double x = logspace ? lo * std::pow(hi / lo, urand()) : lo + (hi - lo) * urand();
    const double e = ulp_err(f(x), g((long double)x));
    if (e > worst) worst = e;
  }
  return worst;
}
int main() {
  using namespace vic;
  const int N = 2000000;
  double w;
  w = sweep([](double x) { return dl::exp(x); }, [](long double x) { return expl(x); }, -60.0, 60.0, N);
  printf("exp %.3f\n", w);
  w = sweep([](double x) { return dl::log(x); }, [](long double x) { return logl(x); }, 1e-12, 1e12, N, true);
  printf("log %.3f\n", w);
  w = sweep([](double x) { return dl::log10(x); }, [](long double x) { return log10l(x); }, 1e-12, 1e12, N, true);
  printf("log10 %.3f\n", w);
  w = sweep([](double x) { return dl::sin(x); }, [](long double x) { return sinl(x); }, -10.0, 10.0, N);
  printf("sin %.3f\n", w);
  w = sweep([](double x) { return dl::cos(x); }, [](long double x) { return cosl(x); }, -10.0, 10.0, N);
  printf("cos %.3f\n", w);
  w = sweep([](double x) { return dl::acos(x); }, [](long double x) { return acosl(x); }, -1.0, 1.0, N);
  printf("acos %.3f\n", w);
  // pow: bases in (1e-6, 1e3), exponents in (-12, 12) with |y log x| < 40: Brooks-Corey / Clapp-Hornberger / svp / stability terms
  double worst = 0;
  for (int i = 0; i < N; i++) {
    const double x = 1e-6 * std::pow(1e9, urand());
    const double y = -12.0 + 24.0 * urand();
    if (fabs(y * std::log(x)) > 40.0) continue;
    const double e = ulp_err(dl::pow(x, y), powl((long double)x, (long double)y));
    if (e > worst) worst = e;
  }
  printf("pow %.3f\n", worst);
  return 0;
}

// oracle/brentcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives root_brent (the restatement of root_brent.c:97-335 with its ~15 residual call sites), root_brent_ss (the same control
// flow as a state machine around one call site) and the resumable brent_begin / brent_advance with random residuals -- monotone and non-monotone, with and without regions where
// the residual is undefined (ERROR), with and without a sign change in the first bracket -- and requires the two to evaluate the
// residual at exactly the same points in the same order and to return the same value.  Prints "cases N mismatches M ..." (tests/test_cpu.py).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include "vic_brent.cuh"

using namespace vic;

static uint64_t rs = 0x243f6a8885a308d3ULL;
static double urand() {
  rs ^= rs >> 12; rs ^= rs << 25; rs ^= rs >> 27;
  return (double)((rs * 0x2545F4914F6CDD1DULL) >> 11) / 9007199254740992.0;
}
struct Residual {
  double root, k3, k1, bump, err_lo, err_hi, err_lo2, err_hi2;  // f = k1 (x - root) + k3 (x - root)^3 + bump sin(3x); ERROR inside the intervals
  std::vector<double>* trace;
  double operator()(double x) {
    trace->push_back(x);
    if ((x > err_lo && x < err_hi) || (x > err_lo2 && x < err_hi2)) return ERROR_D;
    const double u = x - root;
    return k1 * u + k3 * u * u * u + bump * sin(3 * x);
  }
};
int main() {
  long cases = 0, mism = 0, errs = 0, expansions = 0, errpaths = 0, one_bound_undefined = 0, both_undefined = 0;
  for (int t = 0; t < 400000; t++) {
    Residual r;
    r.root = -40 + 80 * urand();
    r.k1 = (urand() < 0.5 ? 1 : -1) * (0.01 + 5 * urand());
    r.k3 = (urand() < 0.3) ? 0.0 : r.k1 * urand() * 0.01;
    r.bump = (urand() < 0.5) ? 0.0 : 3 * urand();
    r.err_lo = r.err_hi = r.err_lo2 = r.err_hi2 = 1e300;
    const double lo = r.root - 30 * urand() + 10 * urand(), hi = lo + 0.5 + 20 * urand();
    if (urand() < 0.35) { r.err_lo = lo - 5 + 12 * urand(); r.err_hi = r.err_lo + 8 * urand(); }
    if (urand() < 0.15) { r.err_lo2 = hi - 6 + 12 * urand(); r.err_hi2 = r.err_lo2 + 8 * urand(); }
    std::vector<double> ta, tb;
    {
      std::vector<double> tmp;
      r.trace = &tmp;
      const bool ea = r(lo) == ERROR_D, eb = r(hi) == ERROR_D;
      if (ea != eb) one_bound_undefined++;
      if (ea && eb) both_undefined++;
    }
    r.trace = &ta;
    const double xa = root_brent(lo, hi, r);
    r.trace = &tb;
    const double xb = root_brent_ss(lo, hi, r);
    // the resumable machine (brent_begin / brent_advance), driven the way vic_frozen.cuh drives it
    std::vector<double> tc;
    r.trace = &tc;
    BrentStep bs;
    brent_begin(bs, lo, hi);
    while (!brent_advance(bs, r(bs.x))) {}
    const double xc = bs.res;
    if (!((memcmp(&xa, &xc, 8) == 0) && ta.size() == tc.size() && (ta.empty() || memcmp(ta.data(), tc.data(), ta.size() * 8) == 0))) {
      if (mism < 5) fprintf(stderr, "stepper mismatch: case %d lo %.17g hi %.17g  ret %.17g vs %.17g  evals %zu vs %zu\n", t, lo, hi, xa, xc, ta.size(), tc.size());
      mism++;
    }
    cases++;
    if (xa == ERROR_D) errs++;
    if (ta.size() > 2 && (ta[2] == lo - 10 || ta[2] == 0.5 * (lo + hi))) expansions++;
    if (r.err_lo < 1e299) errpaths++;
    const bool same = (memcmp(&xa, &xb, 8) == 0) && ta.size() == tb.size() && (ta.empty() || memcmp(ta.data(), tb.data(), ta.size() * 8) == 0);
    if (!same) {
      if (mism < 5) fprintf(stderr, "mismatch: case %d lo %.17g hi %.17g  ret %.17g vs %.17g  evals %zu vs %zu\n", t, lo, hi, xa, xb, ta.size(), tb.size());
      mism++;
    }
  }
  printf("cases %ld mismatches %ld failed_solves %ld bracket_moves %ld with_error_regions %ld one_bound_undefined %ld both_undefined %ld\n", cases, mism, errs,
         expansions, errpaths, one_bound_undefined, both_undefined);
  return mism ? 1 : 0;
}

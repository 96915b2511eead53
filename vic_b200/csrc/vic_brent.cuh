// vic_brent.cuh -- bracketing root finder with the exact iteration sequence of the
// reference's RootBrent::root_brent (root_brent.c:97-335): same bracket expansion
// (TSTEP 10 C, MAXTRIES 5), same recovery when one bound returns ERROR, same tolerance
// 2*MACHEPS*|b| + T with MACHEPS 3e-8 / T 1e-7, same MAXITER 1000, and the same order of
// residual evaluations -- the residual functors have side effects whose LAST evaluation
// is the result (SURVEY.md section 7 "hard parts").
//
// The residual is a functor `double F::operator()(double)`; there is no virtual call and
// no heap object (the reference allocates a C++ functor object per solve).
#ifndef VIC_BRENT_CUH
#define VIC_BRENT_CUH
#include "vic_common.cuh"

namespace vic {

template <class F>
VIC_HDI double root_brent(double LowerBound, double UpperBound, F& f) {
  const int MAXTRIES = 5, MAXITER = 1000;
  const double MACHEPS = 3e-8, TSTEP = 10, T = 1e-7;
  double a = LowerBound, b = UpperBound, c = 0, d = 0, e = 0, fc, m, p, q, r, s, tol;
  double last_bad = 0, last_good = 0;
  int which_err = 0, i, j;
  double fa = f(a);
  double fb = f(b);
  if (fa == ERROR_D && fb == ERROR_D) return ERROR_D;
  if (fa == ERROR_D || fb == ERROR_D) {
    if (fa == ERROR_D) { which_err = -1; last_bad = a; last_good = b; }
    else { which_err = 1; last_good = a; last_bad = b; }
    c = 0.5 * (last_bad + last_good);
    fc = f(c);
    j = 0;
    while (fc == ERROR_D && j < MAXITER) {
      last_bad = c;
      c = 0.5 * (last_bad + last_good);
      fc = f(c);
      j++;
    }
    if (fc == ERROR_D) return ERROR_D;
    if (which_err == -1) { a = c; fa = fc; } else { b = c; fb = fc; }
  }
  j = 0;
  while ((fa * fb) >= 0 && j < MAXTRIES) {
    if (which_err == 0) {
      a -= TSTEP;
      b += TSTEP;
      fa = f(a);
      fb = f(b);
    } else {
      if (which_err == -1) {
        b += TSTEP;
        fb = f(b);
        if (fb == ERROR_D) return ERROR_D;
        last_good = a;
      } else {
        a -= TSTEP;
        fa = f(a);
        if (fa == ERROR_D) return ERROR_D;
        last_good = b;
      }
      c = 0.5 * (last_good + last_bad);
      fc = f(c);
      i = 0;
      while (fc == ERROR_D && i < MAXITER) {
        last_bad = c;
        c = 0.5 * (last_bad + last_good);
        fc = f(c);
        i++;
      }
      if (fc == ERROR_D) return ERROR_D;
      if (which_err == -1) { a = c; fa = fc; } else { b = c; fb = fc; }
    }
    j++;
  }
  if ((fa * fb) >= 0) return ERROR_D;
  fc = fb;
  for (i = 0; i < MAXITER; i++) {
    if (fb * fc > 0) {
      c = a;
      fc = fa;
      d = b - a;
      e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    tol = 2 * MACHEPS * fabs(b) + T;
    m = 0.5 * (c - b);
    if (fabs(m) <= tol || fb == 0) return b;
    if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
      d = m;
      e = d;
    } else {
      s = fb / fa;
      if (a == c) {
        p = 2 * m * s;
        q = 1 - s;
      } else {
        q = fa / fc;
        r = fb / fc;
        p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
        q = (q - 1) * (r - 1) * (s - 1);
      }
      if (p > 0) q = -q; else p = -p;
      s = e;
      e = d;
      if ((2 * p) < (3 * m * q - fabs(tol * q)) && p < fabs(0.5 * s * q)) d = p / q;
      else { d = m; e = d; }
    }
    a = b;
    fa = fb;
    b += (fabs(d) > tol) ? d : ((m > 0) ? tol : -tol);
    fb = f(b);
    if (fb == ERROR_D) return ERROR_D;
  }
  return ERROR_D;
}

}  // namespace vic
#endif

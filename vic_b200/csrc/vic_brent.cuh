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

// The same solver with ONE call site of the residual: the control flow above turned into a state machine around a single
// `fx = f(x)`.  It evaluates the residual at exactly the same points in exactly the same order and returns the same value
// (oracle/brentcheck.cpp drives both with random residuals, including ERROR regions, and compares the evaluation sequences).  With a
// single call site the residual can be inlined into its caller, which lets the compiler keep the solve's constants in registers
// instead of re-reading a functor object from thread-local memory at every evaluation.
// What a caller that needs the residual once more at the accepted root passes along (calc_surf_energy_bal.c:346-400 evaluates the
// energy balance a last time at Tsurf, or at the old temperature when the solve failed and TFALLBACK is on, or at the air
// temperature when the energy balance is not solved at all): that evaluation goes through the same single call site.
// A caller whose bracket depends on the residual at a first point (snow_melt.c:284-379 evaluates the pack's balance at 0 C and only
// solves when that is not zero; snow_intercept.c:391-430 does the same for the canopy) passes PRE: the machine evaluates x_pre, hands
// the value to `decide`, which fills in do_solve / do_final / lo / hi, and carries on from there -- still one residual call site.
struct BrentFinal {
  bool do_solve, allow_fallback;
  bool do_final = true;            // evaluate once more at the accepted point
  bool final_needs_valid = false;  // ... but only when that point is a number
  double fallback_x, nosolve_x;
  double lo = 0, hi = 0;  // PRE: the bracket, set by `decide`
  double f_final;  // out: the residual at the returned point
  int fell_back;   // out: how many solves failed and took fallback_x (0 or 1; up to 2 with RESOLVE)
};

struct BrentNoDecide {
  VIC_HD void operator()(double) const {}
};

// RESOLVE: after a solve (and its fallback) the functor is asked whether it wants the same bracket solved once more with whatever it has
// changed about itself in the meantime (f.resolve(x): QUICK_SOLVE's second pass over the full node profile, calc_surf_energy_bal.c:400-475).
template <bool FINAL, bool PRE = false, bool RESOLVE = false, class F, class D = BrentNoDecide>
VIC_HD double root_brent_ss_impl(double LowerBound, double UpperBound, F& f, BrentFinal* fin, bool pre = false, double x_pre = 0,
                                 D decide = D()) {
  const int MAXTRIES = 5, MAXITER = 1000;
  const double MACHEPS = 3e-8, TSTEP = 10, T = 1e-7;
  double a = LowerBound, b = UpperBound, c = 0, d = 0, e = 0, fa = 0, fb = 0, fc = 0, m, p, q, r, s, tol;
  double last_bad = 0, last_good = 0;
  int which_err = 0, i = 0, j = 0, st = 97;
  double x = a, res = 0;
  if constexpr (PRE) {
    if (pre) {
      x = x_pre;
      st = 98;
    }
  }
  for (;;) {
    if (st == 97) {  // start (after the pre-evaluation, if there was one)
      if constexpr (PRE) {
        a = fin->lo;
        b = fin->hi;
      }
      x = a;
      st = 0;
      if constexpr (FINAL) {
        if (!fin->do_solve) {
          res = fin->nosolve_x;
          if (!fin->do_final) return res;
          f.before_final();
          x = res;
          st = 99;
        }
      }
    }
    const double fx = f(x);
    switch (st) {
      case 98:
        if constexpr (PRE) decide(fx);
        st = 97;
        continue;
      case 0:
        fa = fx;
        x = b;
        st = 1;
        continue;
      case 1:
        fb = fx;
        if (fa == ERROR_D && fb == ERROR_D) { res = (ERROR_D); goto done; }
        if (fa == ERROR_D || fb == ERROR_D) {
          if (fa == ERROR_D) { which_err = -1; last_bad = a; last_good = b; }
          else { which_err = 1; last_good = a; last_bad = b; }
          c = 0.5 * (last_bad + last_good);
          x = c;
          j = 0;
          st = 2;
          continue;
        }
        j = 0;
        goto expand_check;
      case 2:  // bisection towards the good bound while the residual is undefined
        fc = fx;
        if (fc == ERROR_D && j < MAXITER) {
          last_bad = c;
          c = 0.5 * (last_bad + last_good);
          x = c;
          j++;
          continue;
        }
        if (fc == ERROR_D) { res = (ERROR_D); goto done; }
        if (which_err == -1) { a = c; fa = fc; } else { b = c; fb = fc; }
        j = 0;
        goto expand_check;
      case 4:
        fa = fx;
        x = b;
        st = 5;
        continue;
      case 5:
        fb = fx;
        j++;
        goto expand_check;
      case 6:
        fb = fx;
        if (fb == ERROR_D) { res = (ERROR_D); goto done; }
        last_good = a;
        c = 0.5 * (last_good + last_bad);
        x = c;
        i = 0;
        st = 8;
        continue;
      case 7:
        fa = fx;
        if (fa == ERROR_D) { res = (ERROR_D); goto done; }
        last_good = b;
        c = 0.5 * (last_good + last_bad);
        x = c;
        i = 0;
        st = 8;
        continue;
      case 8:
        fc = fx;
        if (fc == ERROR_D && i < MAXITER) {
          last_bad = c;
          c = 0.5 * (last_bad + last_good);
          x = c;
          i++;
          continue;
        }
        if (fc == ERROR_D) { res = (ERROR_D); goto done; }
        if (which_err == -1) { a = c; fa = fc; } else { b = c; fb = fc; }
        j++;
        goto expand_check;
      case 99:  // the extra evaluation at the accepted point (FINAL only)
        if constexpr (FINAL) fin->f_final = fx;
        return res;
      default:  // 10: an iteration of the main loop has evaluated f(b)
        fb = fx;
        if (fb == ERROR_D) { res = (ERROR_D); goto done; }
        i++;
        goto main_top;
    }
  expand_check:
    if ((fa * fb) >= 0 && j < MAXTRIES) {
      if (which_err == 0) {
        a -= TSTEP;
        b += TSTEP;
        x = a;
        st = 4;
      } else if (which_err == -1) {
        b += TSTEP;
        x = b;
        st = 6;
      } else {
        a -= TSTEP;
        x = a;
        st = 7;
      }
      continue;
    }
    if ((fa * fb) >= 0) { res = (ERROR_D); goto done; }
    fc = fb;
    i = 0;
  main_top:
    if (i >= MAXITER) { res = (ERROR_D); goto done; }
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
    if (fabs(m) <= tol || fb == 0) { res = (b); goto done; }
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
    x = b;
    st = 10;
    continue;
  done:
    if constexpr (!FINAL) {
      return res;
    } else {
      if (result_is_error(res)) {
        if (!fin->allow_fallback) return res;
        res = fin->fallback_x;
        fin->fell_back += 1;
      }
      if constexpr (RESOLVE) {
        if (f.resolve(res)) {
          a = LowerBound; b = UpperBound; c = 0; d = 0; e = 0; fa = 0; fb = 0; fc = 0; last_bad = 0; last_good = 0;
          which_err = 0; i = 0; j = 0;
          x = a;
          st = 0;
          continue;
        }
      }
      if (!fin->do_final || (fin->final_needs_valid && !is_valid(res))) return res;
      f.before_final();
      x = res;
      st = 99;
    }
  }
}

template <class F>
VIC_HD double root_brent_ss(double LowerBound, double UpperBound, F& f) {
  return root_brent_ss_impl<false>(LowerBound, UpperBound, f, (BrentFinal*)nullptr);
}


// The same solver once more, as a RESUMABLE machine: the caller owns the residual evaluations.  brent_begin() names the first point,
// the caller evaluates the residual there and hands the value to brent_advance(), which either names the next point (returns false)
// or finishes (returns true, BrentStep::res is what root_brent would have returned).  Same points, same order, same result as
// root_brent / root_brent_ss (oracle/brentcheck.cpp compares all three).  It exists for the soil thermal profile (vic_frozen.cuh):
// there a lane of a warp works through hundreds of small solves -- one per frozen node and sweep -- and with a resumable machine
// every lane can be at a different node, sweep and iteration while the warp still shares the one expensive call site, the residual.
struct BrentStep {
  double a, b, c, d, e, fa, fb, fc, last_bad, last_good, x, res;
  int which_err, i, j, st;
};
VIC_HD void brent_begin(BrentStep& s, double LowerBound, double UpperBound) {
  s.a = LowerBound; s.b = UpperBound; s.c = 0; s.d = 0; s.e = 0; s.fa = 0; s.fb = 0; s.fc = 0; s.last_bad = 0; s.last_good = 0;
  s.which_err = 0; s.i = 0; s.j = 0; s.st = 0; s.res = 0;
  s.x = s.a;
}
VIC_HD bool brent_advance(BrentStep& s, double fx) {
  const int MAXTRIES = 5, MAXITER = 1000;
  const double MACHEPS = 3e-8, TSTEP = 10, T = 1e-7;
  double m, p, q, r, sv, tol;
  switch (s.st) {
    case 0:
      s.fa = fx;
      s.x = s.b;
      s.st = 1;
      return false;
    case 1:
      s.fb = fx;
      if (s.fa == ERROR_D && s.fb == ERROR_D) { s.res = (ERROR_D); return true; }
      if (s.fa == ERROR_D || s.fb == ERROR_D) {
        if (s.fa == ERROR_D) { s.which_err = -1; s.last_bad = s.a; s.last_good = s.b; }
        else { s.which_err = 1; s.last_good = s.a; s.last_bad = s.b; }
        s.c = 0.5 * (s.last_bad + s.last_good);
        s.x = s.c;
        s.j = 0;
        s.st = 2;
        return false;
      }
      s.j = 0;
      goto expand_check;
    case 2:  // bisection towards the good bound while the residual is undefined
      s.fc = fx;
      if (s.fc == ERROR_D && s.j < MAXITER) {
        s.last_bad = s.c;
        s.c = 0.5 * (s.last_bad + s.last_good);
        s.x = s.c;
        s.j++;
        return false;
      }
      if (s.fc == ERROR_D) { s.res = (ERROR_D); return true; }
      if (s.which_err == -1) { s.a = s.c; s.fa = s.fc; } else { s.b = s.c; s.fb = s.fc; }
      s.j = 0;
      goto expand_check;
    case 4:
      s.fa = fx;
      s.x = s.b;
      s.st = 5;
      return false;
    case 5:
      s.fb = fx;
      s.j++;
      goto expand_check;
    case 6:
      s.fb = fx;
      if (s.fb == ERROR_D) { s.res = (ERROR_D); return true; }
      s.last_good = s.a;
      s.c = 0.5 * (s.last_good + s.last_bad);
      s.x = s.c;
      s.i = 0;
      s.st = 8;
      return false;
    case 7:
      s.fa = fx;
      if (s.fa == ERROR_D) { s.res = (ERROR_D); return true; }
      s.last_good = s.b;
      s.c = 0.5 * (s.last_good + s.last_bad);
      s.x = s.c;
      s.i = 0;
      s.st = 8;
      return false;
    case 8:
      s.fc = fx;
      if (s.fc == ERROR_D && s.i < MAXITER) {
        s.last_bad = s.c;
        s.c = 0.5 * (s.last_bad + s.last_good);
        s.x = s.c;
        s.i++;
        return false;
      }
      if (s.fc == ERROR_D) { s.res = (ERROR_D); return true; }
      if (s.which_err == -1) { s.a = s.c; s.fa = s.fc; } else { s.b = s.c; s.fb = s.fc; }
      s.j++;
      goto expand_check;
    default:  // 10: an iteration of the main loop has evaluated f(b)
      s.fb = fx;
      if (s.fb == ERROR_D) { s.res = (ERROR_D); return true; }
      s.i++;
      goto main_top;
  }
expand_check:
  if ((s.fa * s.fb) >= 0 && s.j < MAXTRIES) {
    if (s.which_err == 0) {
      s.a -= TSTEP;
      s.b += TSTEP;
      s.x = s.a;
      s.st = 4;
    } else if (s.which_err == -1) {
      s.b += TSTEP;
      s.x = s.b;
      s.st = 6;
    } else {
      s.a -= TSTEP;
      s.x = s.a;
      s.st = 7;
    }
    return false;
  }
  if ((s.fa * s.fb) >= 0) { s.res = (ERROR_D); return true; }
  s.fc = s.fb;
  s.i = 0;
main_top:
  if (s.i >= MAXITER) { s.res = (ERROR_D); return true; }
  if (s.fb * s.fc > 0) {
    s.c = s.a;
    s.fc = s.fa;
    s.d = s.b - s.a;
    s.e = s.d;
  }
  if (fabs(s.fc) < fabs(s.fb)) {
    s.a = s.b; s.b = s.c; s.c = s.a;
    s.fa = s.fb; s.fb = s.fc; s.fc = s.fa;
  }
  tol = 2 * MACHEPS * fabs(s.b) + T;
  m = 0.5 * (s.c - s.b);
  if (fabs(m) <= tol || s.fb == 0) { s.res = (s.b); return true; }
  if (fabs(s.e) < tol || fabs(s.fa) <= fabs(s.fb)) {
    s.d = m;
    s.e = s.d;
  } else {
    sv = s.fb / s.fa;
    if (s.a == s.c) {
      p = 2 * m * sv;
      q = 1 - sv;
    } else {
      q = s.fa / s.fc;
      r = s.fb / s.fc;
      p = sv * (2 * m * q * (q - r) - (s.b - s.a) * (r - 1));
      q = (q - 1) * (r - 1) * (sv - 1);
    }
    if (p > 0) q = -q; else p = -p;
    sv = s.e;
    s.e = s.d;
    if ((2 * p) < (3 * m * q - fabs(tol * q)) && p < fabs(0.5 * sv * q)) s.d = p / q;
    else { s.d = m; s.e = s.d; }
  }
  s.a = s.b;
  s.fa = s.fb;
  s.b += (fabs(s.d) > tol) ? s.d : ((m > 0) ? tol : -tol);
  s.x = s.b;
  s.st = 10;
  return false;
}

}  // namespace vic
#endif

// vic_blowing.cuh -- sublimation and transport of blowing snow (options.BLOWING): CalcBlowingSnow.c:101-799, Bowling et al. 2004.
// The grid-cell wind distribution (Laplace, ten equal-probability intervals), the probability of occurrence and the threshold shear
// stress after Li & Pomeroy 1997, the shear stress during saltation (Newton iteration on the Liston & Sturm roughness relation), and
// the sublimation / transport fluxes integrated over the suspension layer by Romberg integration (Numerical Recipes qromb / trapzd /
// polint, order 5, relative accuracy 1e-6) -- the reference's operations in the reference's order, in plain double arithmetic on the
// glibc-exact elementary functions, so the flux is the reference's to the bit.
//
// Where the reference ends the PROCESS (exit() when the shear-stress root is not bracketed, nrerror() when the integration does not
// converge in 100 halvings) the device cannot: those cases return ERROR and invalidate the cell like any other failed solve.
#ifndef VIC_BLOWING_CUH
#define VIC_BLOWING_CUH
#include "vic_leaf.cuh"

namespace vic {
namespace blow {

constexpr double Ka = .0245187, CSALT = 0.68, KIN_VIS = 1.3e-5, SETTLING = 0.3, MACHEPS_BS = 1.0e-6, G_STD = 9.80665;
constexpr double PI_BS = 3.1415927;  // vicNl_def.h:276 (mtclim_constants_vic.h:53 only defines PI when it is not defined yet)
constexpr int MAX_ITER_BS = 100, K_BS = 5, NUMINCS = 10;

// what the two integrands share (the reference passes eleven arguments through a function pointer)
struct Layer {
  double es, Wind, ZO, EactAir, F, hsalt, phi_r, ushear;
};

// sublimation rate [kg/m3 s] at height z, CalcBlowingSnow.c:508-563
VIC_HD double sub_with_height(double z, const Layer& a) {
  const double Rrz = 4.6e-5 * vpow(z, -.258);
  const double ALPHAz = 4.08 + 12.6 * z;
  const double Mz = (4. / 3.) * PI_BS * ice_density * Rrz * Rrz * Rrz * (1. + (3. / ALPHAz) + (2. / (ALPHAz * ALPHAz)));
  const double Rmean = vpow((3. * Mz) / (4. * PI_BS * ice_density), 1. / 3.);
  const double terminal_v = 1.1e7 * vpow(Rmean, 1.8);
  const double fluctuat_v = 0.005 * vpow(a.Wind, 1.36);
  const double Vtz = terminal_v + 3. * fluctuat_v * vcos(PI_BS / 4.);
  const double Re = 2. * Rmean * Vtz / KIN_VIS;
  const double Nu = 1.79 + 0.606 * vpow(Re, 0.5);
  const double sigz = ((a.EactAir / a.es) - 1.) * (1.019 + .027 * vlog(z));
  const double dMdt = 2 * PI_BS * Rmean * sigz * Nu / a.F;
  const double psi_t = dMdt / Mz;
  const double temp = (0.5 * a.ushear * a.ushear) / (a.Wind * SETTLING);
  const double phi_t = a.phi_r * ((temp + 1.) * vpow((z / a.hsalt), (-1. * SETTLING) / (von_K * a.ushear)) - temp);
  return psi_t * phi_t;
}

// transport rate [kg/m2 s] at height z, CalcBlowingSnow.c:771-799
VIC_HD double transport_with_height(double z, const Layer& a) {
  const double u_z = a.ushear * vlog(z / a.ZO) / von_K;
  const double temp = (0.5 * a.ushear * a.ushear) / (a.Wind * SETTLING);
  const double phi_t = a.phi_r * ((temp + 1.) * vpow((z / a.hsalt), (-1. * SETTLING) / (von_K * a.ushear)) - temp);
  return u_z * phi_t;
}

template <bool TRANSPORT>
VIC_HD double integrand(double z, const Layer& a) {
  return TRANSPORT ? transport_with_height(z, a) : sub_with_height(z, a);
}

// polynomial extrapolation to x through n points (Numerical Recipes polint; xa / ya are 1-based like the reference's)
VIC_HD void polint(const double* xa, const double* ya, int n, double x, double* y, double* dy) {
  double c[K_BS + 1], d[K_BS + 1];
  int ns = 1;
  double dif = fabs(x - xa[1]);
  for (int i = 1; i <= n; i++) {
    const double dift = fabs(x - xa[i]);
    if (dift < dif) {
      ns = i;
      dif = dift;
    }
    c[i] = ya[i];
    d[i] = ya[i];
  }
  *y = ya[ns--];
  for (int m = 1; m < n; m++) {
    for (int i = 1; i <= n - m; i++) {
      const double ho = xa[i] - x, hp = xa[i + m] - x;
      const double w = c[i + 1] - d[i];
      double den = ho - hp;  // (never 0: the abscissae are 1, 1/4, 1/16, ...)
      den = w / den;
      d[i] = hp * den;
      c[i] = ho * den;
    }
    *dy = (2 * ns < (n - m)) ? c[ns + 1] : d[ns--];
    *y += *dy;
  }
}

// Romberg integration of the integrand over [a, b], CalcBlowingSnow.c:331-424.  ok = false: no convergence in MAX_ITER halvings (nrerror).
// Only the last K estimates are kept (the reference keeps all hundred).
template <bool TRANSPORT>
VIC_HDI double qromb(const Layer& L, double a, double b, bool* ok) {
  double s[K_BS + 1], h[K_BS + 2];  // 1-based windows: s[1..K], h[1..K] hold steps j-K+1 .. j
  double hj = 1.0, lastS = 0.0;
  *ok = true;
  for (int j = 1; j <= MAX_ITER_BS; j++) {
    double sj;
    if (j == 1) sj = 0.5 * (b - a) * (integrand<TRANSPORT>(a, L) + integrand<TRANSPORT>(b, L));
    else {
      int it = 1;
      for (int q = 1; q < j - 1; q++) it <<= 1;
      const double tnm = it;
      const double del = (b - a) / tnm;
      double x = a + 0.5 * del, sum = 0.0;
      for (int q = 1; q <= it; q++, x += del) sum += integrand<TRANSPORT>(x, L);
      sj = 0.5 * (lastS + (b - a) * sum / tnm);
    }
    lastS = sj;
    // slide the window
    if (j <= K_BS) {
      s[j] = sj;
      h[j] = hj;
    } else {
      for (int q = 1; q < K_BS; q++) {
        s[q] = s[q + 1];
        h[q] = h[q + 1];
      }
      s[K_BS] = sj;
      h[K_BS] = hj;
    }
    if (j >= K_BS) {
      double ss, dss;
      polint(h, s, K_BS, 0.0, &ss, &dss);
      if (fabs(dss) <= MACHEPS_BS * fabs(ss)) return ss;
    }
    hj = 0.25 * hj;
    if (j >= 31) break;  // 2^29 evaluations per step from here on: the reference would run for days before giving up
  }
  *ok = false;
  return 0.0;
}

// f(u*) = 0 is the shear velocity for which the Liston & Sturm saltation roughness is consistent with the log profile; CalcBlowingSnow.c:477-480
VIC_HD void get_shear(double x, double* f, double* df, double Ur, double Zr) {
  *f = vlog(2. * G_STD * Zr / .12) + vlog(1 / (x * x)) - von_K * Ur / x;
  *df = von_K * Ur / (x * x) - 2. / x;
}

// Newton-Raphson with bisection safeguard, CalcBlowingSnow.c:424-475; ok = false where the reference exits the process
VIC_HD double rtnewt(double x1, double x2, double acc, double Ur, double Zr, bool* ok) {
  double df, dx, dxold, f, fh, fl, temp, xh, xl, rts;
  *ok = true;
  get_shear(x1, &fl, &df, Ur, Zr);
  get_shear(x2, &fh, &df, Ur, Zr);
  if ((fl > 0.0 && fh > 0.0) || (fl < 0.0 && fh < 0.0)) {
    *ok = false;
    return 0.0;
  }
  if (fl == 0.0) return x1;
  if (fh == 0.0) return x2;
  if (fl < 0.0) {
    xl = x1;
    xh = x2;
  } else {
    xh = x1;
    xl = x2;
  }
  rts = 0.5 * (x1 + x2);
  dxold = fabs(x2 - x1);
  dx = dxold;
  get_shear(rts, &f, &df, Ur, Zr);
  for (int j = 1; j <= MAX_ITER_BS; j++) {
    if ((((rts - xh) * df - f) * ((rts - x1) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
      dxold = dx;
      dx = 0.5 * (xh - xl);
      rts = xl + dx;
      if (xl == rts) return rts;
    } else {
      dxold = dx;
      dx = f / df;
      temp = rts;
      rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < acc) return rts;
    get_shear(rts, &f, &df, Ur, Zr);
    if (f < 0.0) xl = rts;
    else xh = rts;
  }
  return 0.0;  // "Maximum number of iterations exceeded in rtnewt"
}

// probability of blowing snow, Li & Pomeroy 1997; CalcBlowingSnow.c:571-600
VIC_HD double get_prob(double Tair, double Age, double SurfaceLiquidWater, double U10) {
  double mean_u, sigma;
  if (SurfaceLiquidWater < 0.001) {
    mean_u = 11.2 + 0.365 * Tair + 0.00706 * Tair * Tair + 0.9 * vlog(Age);
    sigma = 4.3 + 0.145 * Tair + 0.00196 * Tair * Tair;
  } else {
    mean_u = 21.;
    sigma = 7.;
  }
  double p = 1. / (1. + vexp(sqrt(PI_BS) * (mean_u - U10) / sigma));
  if (p < 0.0) p = 0.0;
  if (p > 1.0) p = 1.0;
  return p;
}

// threshold shear velocity (variable threshold), CalcBlowingSnow.c:602-627
VIC_HD double get_thresh(double Tair, double SurfaceLiquidWater, double Zo_salt) {
  const double ut10 = (SurfaceLiquidWater < 0.001) ? 9.43 + .18 * Tair + .0033 * Tair * Tair : 9.9;
  return von_K * ut10 / vlog(10. / Zo_salt);
}

// shear velocity and roughness during saltation, CalcBlowingSnow.c:629-662
VIC_HD bool shear_stress(double U10, double ZO, double* ushear, double* Zo_salt, double utshear) {
  const double umin = utshear, umax = von_K * U10, xacc = 0.10 * umin;
  double fl, fh, df;
  get_shear(umin, &fl, &df, U10, 10.);
  get_shear(umax, &fh, &df, U10, 10.);
  if (fl < 0.0 && fh < 0.0) return false;  // "Solution in rtnewt surpasses upper boundary": exit(0) in the reference
  if (fl > 0.0 && fh > 0.0) {
    *Zo_salt = ZO;
    *ushear = von_K * U10 / vlog(10. / ZO);
  } else {
    bool ok;
    *ushear = rtnewt(umin, umax, xacc, U10, 10., &ok);
    if (!ok) return false;
    *Zo_salt = 0.12 * (*ushear) * (*ushear) / (2. * G_STD);
  }
  return true;
}

// sublimation flux of one wind interval (Liston & Sturm mass flux, with fetch), CalcBlowingSnow.c:664-769
VIC_HDI double calc_sub_flux(double EactAir, double es, double AirDens, double utshear, double ushear, float fe, double U10, double Zo_salt, double F,
                             double* Transport, bool* ok) {
  double SubFlux = 0.0;
  *ok = true;
  const double particle = utshear * 2.8;
  double Qsalt = (CSALT * AirDens / G_STD) * (utshear / ushear) * (ushear * ushear - utshear * utshear);
  Qsalt *= (1. + (500. / (3. * fe)) * (vexp(-3. * fe / 500.) - 1.));
  const double hsalt = 0.08436 * vpow(ushear, 1.27);
  const double phi_s = Qsalt / (hsalt * particle);
  const double T = 0.5 * (ushear * ushear) / (U10 * SETTLING);
  const double ztop = hsalt * vpow(T / (T + 1.), (von_K * ushear) / (-1. * SETTLING));
  const Layer L{es, U10, Zo_salt, EactAir, F, hsalt, phi_s, ushear};
  if (EactAir >= es) SubFlux = 0.0;
  else {
    const double psi_s = sub_with_height(hsalt / 2., L);
    SubFlux = phi_s * psi_s * hsalt;
    SubFlux += qromb<false>(L, hsalt, ztop, ok);
    if (!*ok) return 0.0;
  }
  const double saltation_transport = Qsalt * (1 - vexp(-3. * fe / 500.));
  const double suspension_transport = qromb<true>(L, hsalt, ztop, ok);
  if (!*ok) return 0.0;
  *Transport = (suspension_transport + saltation_transport);
  *Transport /= fe;
  return SubFlux;
}

// Mass flux of sublimating blowing snow [kg/m2 s] (negative: a loss) and the transported mass; ERROR_D where the reference returns ERROR
// or ends the process.  Wind: 2 m above the snow; ZO: snow roughness; displacement / roughness: of the tile's canopy.
VIC_HDI double calc_blowing_snow(double Dt, double Tair, int LastSnow, double SurfaceLiquidWater, double Wind, double Ls, double AirDens, double EactAir,
                                 double ZO, double snowdepth, float lag_one, float sigma_slope, bool isArtificialBareSoil, float fe, double displacement,
                                 double roughness, double* TotalTransport) {
  const double MW = 18.0148e-3, Rgas = 8.3143;
  const double Age = LastSnow * (Dt);
  const double es = svp(Tair);
  const double Tk = Tair + KELVIN;
  const double Ros = 0.622 * es / (287 * Tk);
  const double Diffusivity = (2.06e-5) * vpow(Tk / 273., 1.75);
  double F = (Ls / (Ka * Tk)) * (Ls * MW / (Rgas * Tk) - 1.);
  F += 1. / (Diffusivity * Ros);
  const double wind10 = Wind * vlog(10. / ZO) / vlog((2 + ZO) / ZO);
  if (isArtificialBareSoil) {
    fe = 1500;
    sigma_slope = .0002;
  }
  const double ratio = (2.44 - (0.43) * lag_one) * sigma_slope;
  const double sigma_w = wind10 * ratio;
  const double Uo = wind10;
  const double hv = (3. / 2.) * displacement;
  const double Nd = (4. / 3.) * (roughness / displacement);
  double Total = 0.0;
  *TotalTransport = 0.0;
  const double area = 1. / NUMINCS;
  if (snowdepth > 0.0) {
    const int nint = (sigma_w != 0.) ? NUMINCS : 1;  // constant wind when the spread is zero (CalcBlowingSnow.c:285-312)
    for (int p = 0; p < nint; p++) {
      double U10 = Uo;
      if (sigma_w != 0.) {
        double lower = 0.0, upper = 0.0;
        if (p == 0) {
          lower = -9999;
          upper = Uo + sigma_w * vlog(2. * (p + 1) * area);
        } else if (p > 0 && p < NUMINCS / 2) {
          lower = Uo + sigma_w * vlog(2. * (p)*area);
          upper = Uo + sigma_w * vlog(2. * (p + 1) * area);
        } else if (p < (NUMINCS - 1) && p >= NUMINCS / 2) {
          lower = Uo - sigma_w * vlog(2. - 2. * (p * area));
          upper = Uo - sigma_w * vlog(2. - 2. * ((p + 1.) * area));
        } else if (p == NUMINCS - 1) {
          lower = Uo - sigma_w * vlog(2. - 2. * (p * area));
          upper = 9999;
        }
        if (lower > upper) lower = upper;
        if (lower >= Uo)
          U10 = -0.5 * ((upper + sigma_w) * vexp((-1. / sigma_w) * (upper - Uo)) - (lower + sigma_w) * vexp((-1. / sigma_w) * (lower - Uo))) / area;
        else if (upper <= Uo)
          U10 = 0.5 * ((upper - sigma_w) * vexp((1. / sigma_w) * (upper - Uo)) - (lower - sigma_w) * vexp((1. / sigma_w) * (lower - Uo))) / area;
        else return ERROR_D;
        if (U10 < 0.4) U10 = .4;
        if (U10 > 25.) U10 = 25.;
      }
      double Uveg;
      if (snowdepth < hv) Uveg = U10 / sqrt(1. + 170 * Nd * (hv - snowdepth));
      else Uveg = U10;
      const double prob_occurence = get_prob(Tair, Age, SurfaceLiquidWater, Uveg);
      const double utshear = get_thresh(Tair, SurfaceLiquidWater, ZO);
      double ushear, Zo_salt, SubFlux, Transport;
      if (!shear_stress(U10, ZO, &ushear, &Zo_salt, utshear)) return ERROR_D;
      if (ushear > utshear) {
        bool ok;
        Transport = 0.0;
        SubFlux = calc_sub_flux(EactAir, es, AirDens, utshear, ushear, fe, U10, Zo_salt, F, &Transport, &ok);
        if (!ok) return ERROR_D;
      } else {
        SubFlux = 0.0;
        Transport = 0.0;
      }
      if (sigma_w != 0.) {
        Total += (1. / NUMINCS) * SubFlux * prob_occurence;
        *TotalTransport += (1. / NUMINCS) * Transport * prob_occurence;
      } else {
        Total = SubFlux * prob_occurence;
        *TotalTransport = Transport * prob_occurence;
      }
    }
  }
  if (Total < -.00005) Total = -.00005;
  return Total;
}

}  // namespace blow
}  // namespace vic
#endif

// vic_disagg.cuh -- daily -> sub-daily forcing disaggregation (MTCLIM 4.3 as embedded in VIC):
//   initialize_atmos()                 initialize_atmos.c:7-1349   (daily PREC / TMAX / TMIN / WIND supplied)
//   mtclim_wrapper / mtclim_to_vic     mtclim_wrapper.c:64-260
//   calc_tair, calc_prcp, snowpack, calc_srad_humidity_iterative, compute_srad_humidity_onetime,
//   calc_pet, atm_pres, pulled_boxcar  mtclim_vic.c:394-522, 1063-1747, 1779-1918
//   set_max_min_hour, HourlyT (hermite / hermint)   calc_air_temperature.c:20-198
//   calc_longwave                      calc_longwave.c:8-73
//
// The work of one cell is cut into stages whose work items are independent, so that the CUDA
// library can run each stage as one kernel over (cell x item) and the host port as plain loops:
//   S1 solar geometry      item = day of year (the 30 s hour-angle loop, ~2880 steps, per day)
//   S2 daily chain         item = cell (sequential in time: snow pack, smoothing windows, Tdew iteration)
//   S3 hourly radiation, Tmax/Tmin hours, spline knots     item = local day
//   S4 spline coefficients item = knot;   S5 hourly air temperature / vapour pressure   item = local day
//   S6 model records       item = record (aggregation to the model step, pressure, density, vpd, longwave, snow flag)
// All per-cell scratch is column-major with the cell index fastest ([k][ncell]).
//
// tiny_radfract: the reference keeps a 366 x 2880 table per cell (253 MB allocated, mtclim_wrapper.c:77-86)
// and later sums 120 entries per hour (mtclim_wrapper.c:238-250).  Here the hour-angle loop is run twice per
// day of year -- once for the day's sums, once more to bin the normalised 30 s values straight into the 24
// hourly fractions -- so the table never exists.  Within an hour the values are added in the reference's order;
// only in the hour that contains local midnight could the order differ, and only under a midnight sun.
#ifndef VIC_DISAGG_CUH
#define VIC_DISAGG_CUH
#include "vic_leaf.cuh"

namespace vic {

enum { VP_ITER_NEVER = 0, VP_ITER_ALWAYS, VP_ITER_ANNUAL, VP_ITER_CONVERGE };                         // vicNl_def.h
enum { LW_TVA = 0, LW_ANDERSON, LW_BRUTSAERT, LW_SATTERLUND, LW_IDSO, LW_PRATA };
enum { LW_CLOUD_BRAS = 0, LW_CLOUD_DEARDORFF };

struct DisaggOpts {
  int dt, SNOW_STEP, NF, NR, nrecs, starthour, startyear, startmonth, startday, Ndays;
  int PLAPSE, MTCLIM_SWE_CORR, VP_ITER, VP_INTERP, LW_TYPE, LW_CLOUD, TEMP_TH_TYPE, Nbands, OUTPUT_FORCE;
  double SW_PREC_THRESH, MIN_WIND_SPEED;
  int f_nslot;
};

// scratch layout (all [k][ncell]); offsets in doubles per cell
struct DisaggScratch {
  double* base;
  size_t ncell;   // cells in this chunk (stride of the scratch columns)
  size_t ntotal;  // cells of the whole domain (stride of the daily input and of the forcing records)
  int cell0;      // first cell of the chunk
  int Ndl;        // maximum number of local days (Ndays + 1)
  // solar tables
  VIC_HD size_t o_ttmax0() const { return 0; }
  VIC_HD size_t o_flat() const { return 366; }
  VIC_HD size_t o_slope() const { return 2 * 366; }
  VIC_HD size_t o_dayl() const { return 3 * 366; }
  VIC_HD size_t o_hourfrac() const { return 4 * 366; }  // [366][24]
  VIC_HD size_t o_daily() const { return 4 * 366 + 366 * 24; }
  enum { D_yday = 0, D_prec, D_tmax, D_tmin, D_swe, D_dtr, D_smdtr, D_parray, D_tfmax, D_tdew, D_pva, D_pet, D_srad, D_sdayl, D_tskc,
         D_tmaxhour, D_tminhour, D_tday, D_sprcp, D_tdew_save, D_N };
  VIC_HD size_t o_knots() const { return o_daily() + (size_t)D_N * Ndl; }       // x, y, c3, c4: 4 x (2*Ndl+2)
  VIC_HD size_t o_hourly() const { return o_knots() + 4 * (size_t)(2 * Ndl + 2); }  // hourlyrad, tair, vp: 3 x Ndl*24
  VIC_HD size_t per_cell() const { return o_hourly() + 3 * (size_t)Ndl * 24; }
  VIC_HD double& at(size_t off, int cell) const { return base[off * ncell + (size_t)(cell - cell0)]; }
  VIC_HD double& daily(int f, int day, int cell) const { return at(o_daily() + (size_t)f * Ndl + day, cell); }
  VIC_HD double& knot(int f, int i, int cell) const { return at(o_knots() + (size_t)f * (2 * Ndl + 2) + i, cell); }
  VIC_HD double& hourly(int f, int idx, int cell) const { return at(o_hourly() + (size_t)f * Ndl * 24 + idx, cell); }
};

// per-cell time-zone bookkeeping (initialize_atmos.c:125-176); integer arithmetic, bit-exact
struct LocalTime {
  double hour_offset;
  int hour_offset_int, Ndays_local, local_starthour, local_startday, local_startmonth, local_startyear;
};

VIC_HD LocalTime local_time(const CellPar& cp, const DisaggOpts& d) {
  LocalTime t;
  // time_zone_lng and lng are float members of soil_con_struct
  t.hour_offset = ((double)(float)cp(CP_time_zone_lng) - (double)(float)cp(CP_lng)) * 24 / 360;
  if (t.hour_offset < 0) t.hour_offset_int = (int)(t.hour_offset - 0.5);
  else t.hour_offset_int = (int)(t.hour_offset + 0.5);
  t.Ndays_local = d.Ndays;
  if (t.hour_offset_int != 0) t.Ndays_local = d.Ndays + 1;
  const int month_days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  t.local_starthour = d.starthour - t.hour_offset_int;
  t.local_startday = d.startday;
  t.local_startmonth = d.startmonth;
  t.local_startyear = d.startyear;
  if (t.local_starthour < 0) {
    t.local_starthour += 24;
    t.local_startday--;
    if (t.local_startday < 1) {
      t.local_startmonth--;
      if (t.local_startmonth < 1) {
        t.local_startmonth = 12;
        t.local_startyear--;
      }
      t.local_startday = month_days[t.local_startmonth - 1];
      if (t.local_startyear % 4 == 0 && t.local_startmonth == 2) t.local_startday++;
    }
  }
  return t;
}

// index of the first local hour of model record `rec`, sub-step `i` (the expression repeated all over initialize_atmos.c)
VIC_HD int local_hour(const DisaggOpts& d, const LocalTime& t, int rec, int i) {
  int hour = rec * d.dt + i * d.SNOW_STEP + d.starthour - t.hour_offset_int;
  if (d.starthour - t.hour_offset_int < 0) hour += 24;
  return hour;
}

// ---- S1: solar geometry of one day of year (mtclim_vic.c:1283-1450) ---------------------------
VIC_HDI void disagg_solar(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, int cell, int i /* 0..364 */) {
  (void)d;
  const double SECPERRAD = 13750.9871, RADPERDAY = 0.017214, RADPERDEG = 0.01745329, MINDECL = -0.4092797, DAYSOFF = 11.25, SRADDT = 30.0;
  const double MA = 28.9644e-3, R = 8.3143, G_STD = 9.80665, T_STD = 288.15, LR_STD = 0.0065, PI_M = 3.1415927 /* vicNl_def.h:276 wins over mtclim_constants_vic.h:53 (#ifndef PI) */, TBASE = 0.870;
  const double optam[21] = {2.90, 3.05, 3.21, 3.39, 3.69, 3.82, 4.07, 4.37, 4.72, 5.12, 5.60, 6.18, 6.88, 7.77, 8.90, 10.39, 12.44, 15.36, 19.79, 26.96, 30.00};
  const double site_elev = (double)(float)cp(CP_elevation);
  const LocalTime lt = local_time(cp, d);
  // The reference fills the tables for the whole year whatever the length of the run (mtclim_vic.c:1345); what is read afterwards are
  // the entries of the run's local days only (disagg_daily: D_yday).  A window of a few days -- the continental bench disaggregates
  // two weeks at a time -- therefore skips the other days of year: 30-second stepping through ~350 unused days was the whole cost
  // of a short call.  The skipped entries of the scratch are never read.
  {
    const int month_days_[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int day_in_year = lt.local_startday;
    for (int m = 1; m < lt.local_startmonth; m++) {
      int dim = month_days_[m - 1];
      if (lt.local_startyear % 4 == 0 && m == 2) dim++;
      day_in_year += dim;
    }
    // indices used: (day_in_year - 1 + k) for k < Ndays_local, with the calendar's wrap after day 365 / 366; two days of margin
    const int first = day_in_year - 1, span = lt.Ndays_local + 2;
    if (span < 365) {
      const int dist = ((i - first) % 365 + 365) % 365;
      if (dist >= span && dist < 365 - 2) return;
    }
  }
  const double t1 = 1.0 - (LR_STD * site_elev) / T_STD;
  const double t2 = G_STD / (LR_STD * (R / MA));
  const double pratio = vpow(t1, t2);
  const double trans1 = vpow(TBASE, pratio);
  double lat = (double)(float)cp(CP_lat);
  lat *= RADPERDEG;
  if (lat > 1.5707) lat = 1.5707;
  if (lat < -1.5707) lat = -1.5707;
  const double coslat = vcos(lat), sinlat = vsin(lat);
  const double slp = cp(CP_slope), asp = cp(CP_aspect);
  const double cosslp = vcos(slp * RADPERDEG), sinslp = vsin(slp * RADPERDEG), cosasp = vcos(asp * RADPERDEG), sinasp = vsin(asp * RADPERDEG);
  const double coszeh = vcos(1.570796 - (cp(CP_ehoriz) * RADPERDEG));
  const double coszwh = vcos(1.570796 - (cp(CP_whoriz) * RADPERDEG));
  const double dt = SRADDT;
  const double dh = dt / SECPERRAD;
  const int tinystepspday = (int)(86400 / SRADDT);
  const int tinystepsphour = (int)(3600 / SRADDT);
  const int tiny_offset = (int)((float)tinystepsphour * lt.hour_offset);

  const double decl = MINDECL * vcos(((double)i + DAYSOFF) * RADPERDAY);
  const double cosdecl = vcos(decl), sindecl = vsin(decl);
  const double bsg1 = -sinslp * sinasp * cosdecl;
  const double bsg2 = (-cosasp * sinslp * sinlat + cosslp * coslat) * cosdecl;
  const double bsg3 = (cosasp * sinslp * coslat + cosslp * sinlat) * sindecl;
  const double cosegeom = coslat * cosdecl;
  const double sinegeom = sinlat * sindecl;
  double coshss = -(sinegeom) / cosegeom;
  if (coshss < -1.0) coshss = -1.0;
  if (coshss > 1.0) coshss = 1.0;
  const double hss = vacos(coshss);
  double daylength = 2.0 * hss * SECPERRAD;
  if (daylength > 86400) daylength = 86400;
  const double sc = 1368.0 + 45.5 * vsin((2.0 * PI_M * (double)i / 365.25) + 1.7);
  const double dir_beam_topa = sc * dt;
  double sum_trans = 0.0, sum_flat_potrad = 0.0, sum_slope_potrad = 0.0;
  // pass 1: the day's sums
  for (double h = -hss; h < hss; h += dh) {
    const double cosh_ = vcos(h), sinh_ = vsin(h);
    const double cza = cosegeom * cosh_ + sinegeom;
    const double cbsa = sinh_ * bsg1 + cosh_ * bsg2 + bsg3;
    if (cza > 0.0) {
      const double dir_flat_topa = dir_beam_topa * cza;
      double am = 1.0 / (cza + 0.0000001);
      if (am > 2.9) {
        int ami = (int)(vacos(cza) / RADPERDEG) - 69;
        if (ami < 0) ami = 0;
        if (ami > 20) ami = 20;
        am = optam[ami];
      }
      const double trans2 = vpow(trans1, am);
      sum_trans += trans2 * dir_flat_topa;
      sum_flat_potrad += dir_flat_topa;
      if ((h < 0.0 && cza > coszeh && cbsa > 0.0) || (h >= 0.0 && cza > coszwh && cbsa > 0.0)) sum_slope_potrad += dir_beam_topa * cbsa;
    }
  }
  double ttmax0, flat, slope;
  if (daylength) {
    ttmax0 = sum_trans / sum_flat_potrad;
    flat = sum_flat_potrad / daylength;
    slope = sum_slope_potrad / daylength;
  } else ttmax0 = flat = slope = 0.0;
  // pass 2: bin the (normalised) 30 s fractions into local-standard-time hours
  double hf[24];
  for (int j = 0; j < 24; j++) hf[j] = 0;
  const bool norm = (daylength != 0) && sum_flat_potrad > 0;
  // The reference sums an hour's 30 s slots in the order of the hour's own index k (mtclim_wrapper.c:240-249: tinystep = j*120 + k -
  // tiny_offset, wrapped into the day).  When tiny_offset is not a whole number of hours, ONE hour holds both the last slots of the solar
  // day and its first ones, and the reference adds the last ones first.  This loop meets the first slots of the day first, so the ones of
  // that hour are set aside and added after the loop.  They are non-zero only when the sun is up at solar midnight (polar day).
  const int off_in_hour = ((tiny_offset % tinystepsphour) + tinystepsphour) % tinystepsphour;
  const int n_early = off_in_hour ? tinystepsphour - off_in_hour : 0;  // slots [0, n_early) of the solar day belong to the straddling hour
  double early[120];
  int ne = 0;
  int cur = -1;
  double curv = 0;
  for (double h = -hss; h < hss; h += dh) {
    const double cza = cosegeom * vcos(h) + sinegeom;
    const double dir_flat_topa = (cza > 0.0) ? dir_beam_topa * cza : -1;
    int tinystep = (int)((12L * 3600L + h * SECPERRAD) / SRADDT);
    if (tinystep < 0) tinystep = 0;
    if (tinystep > tinystepspday - 1) tinystep = tinystepspday - 1;
    const double v = (dir_flat_topa > 0) ? dir_flat_topa : 0;
    if (tinystep != cur) {
      if (cur >= 0) {
        int tp = cur + tiny_offset;  // inverse of tinystep = j*120 + k - tiny_offset (with wrap)
        tp %= tinystepspday;
        if (tp < 0) tp += tinystepspday;
        const double frac = norm ? curv / sum_flat_potrad : curv;
        if (cur < n_early && ne < 120) early[ne++] = frac;
        else hf[tp / tinystepsphour] += frac;
      }
      cur = tinystep;
    }
    curv = v;  // a later hour angle that lands on the same 30 s slot overwrites it, as in the reference
  }
  if (cur >= 0) {
    int tp = (cur + tiny_offset) % tinystepspday;
    if (tp < 0) tp += tinystepspday;
    const double frac = norm ? curv / sum_flat_potrad : curv;
    if (cur < n_early && ne < 120) early[ne++] = frac;
    else hf[tp / tinystepsphour] += frac;
  }
  if (ne) {
    int tp = tiny_offset % tinystepspday;  // the hour that holds slot 0 of the solar day
    if (tp < 0) tp += tinystepspday;
    for (int q = 0; q < ne; q++) hf[tp / tinystepsphour] += early[q];
  }
  const int last = (i == 364) ? 2 : 1;  // day 366 repeats day 365 (mtclim_vic.c:1453-1460)
  for (int r = 0; r < last; r++) {
    const int ii = i + r;
    s.at(s.o_ttmax0() + ii, cell) = ttmax0;
    s.at(s.o_flat() + ii, cell) = flat;
    s.at(s.o_slope() + ii, cell) = slope;
    s.at(s.o_dayl() + ii, cell) = daylength;
    for (int j = 0; j < 24; j++) s.at(s.o_hourfrac() + (size_t)ii * 24 + j, cell) = hf[j];
  }
}

VIC_HD double mt_calc_pet(double rad, double ta, double pa, double dayl) {  // mtclim_vic.c:1779-1837
  const double CP_MT = 1010.0;
  const double rnet = rad * 0.72;
  const double lhvap = 2.5023e6 - 2430.54 * ta;
  const double gamma = CP_MT * pa / (lhvap * EPS);
  const double dt = 0.2;
  const double t1 = ta + dt, t2 = ta - dt;
  const double s = (svp(t1) - svp(t2)) / (t1 - t2);
  const double pet = (1.26 * (s / (s + gamma)) * rnet * dayl) / lhvap;
  return (pet / 10.0);
}

// mtclim_vic.c:1631-1747
VIC_HDI void disagg_srad_humidity_onetime(const DisaggOpts& d, const DisaggScratch& s, int cell, int ndays, double sky_prop, double pa) {
  typedef DisaggScratch S;
  const double ABASE = -6.1e-5, DIF_ALB = 0.6;
  for (int i = 0; i < ndays; i++) {
    const int yday = (int)s.daily(S::D_yday, i, cell) - 1;
    const double pva = s.daily(S::D_pva, i, cell);
    const double tfmax = s.daily(S::D_tfmax, i, cell);
    const double t_tmax = s.at(s.o_ttmax0() + yday, cell) + ABASE * pva;
    const double t_final = t_tmax * tfmax;
    s.daily(S::D_tskc, i, cell) = sqrt((1. - tfmax) / 0.65);
    double pdif = -1.25 * t_final + 1.25;
    if (pdif > 1.0) pdif = 1.0;
    if (pdif < 0.0) pdif = 0.0;
    const double pdir = 1.0 - pdif;
    const double srad1 = s.at(s.o_slope() + yday, cell) * t_final * pdir;
    const double srad2 = s.at(s.o_flat() + yday, cell) * t_final * pdif * (sky_prop + DIF_ALB * (1.0 - sky_prop));
    double sc;
    const double swe = s.daily(S::D_swe, i, cell);
    if (d.MTCLIM_SWE_CORR && swe > 0.0) {
      sc = (1.32 + 0.096 * swe) * 1e6;
      const double dl = s.at(s.o_dayl() + yday, cell);
      if (dl > 0.0) sc /= dl;
      else sc = 0.0;
      if (sc > 100.0) sc = 100.0;
    } else sc = 0.0;
    s.daily(S::D_srad, i, cell) = srad1 + srad2 + sc;
  }
  for (int i = 0; i < ndays; i++) {
    const double tmink = s.daily(S::D_tmin, i, cell) + KELVIN;  // s_tmin == tmin (site == base elevation)
    const double pet = mt_calc_pet(s.daily(S::D_srad, i, cell), s.daily(S::D_tday, i, cell), pa, s.daily(S::D_sdayl, i, cell));
    s.daily(S::D_pet, i, cell) = pet;
    const double ratio = pet / s.daily(S::D_parray, i, cell);
    const double ratio2 = ratio * ratio;
    const double ratio3 = ratio2 * ratio;
    const double tdewk = tmink * (-0.127 + 1.121 * (1.003 - 1.444 * ratio + 12.312 * ratio2 - 32.766 * ratio3) + 0.0006 * (s.daily(S::D_dtr, i, cell)));
    const double tdew = tdewk - KELVIN;
    s.daily(S::D_tdew, i, cell) = tdew;
    s.daily(S::D_pva, i, cell) = svp(tdew);
  }
}

// ---- S2: the daily chain of one cell ------------------------------------------------------------
// daily: [Ndays*4][ncell] column-major input (prec, tmax, tmin, wind per day)
VIC_HDI void disagg_daily(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, const double* daily, int cell) {
  typedef DisaggScratch S;
  const double TDAYCOEF = 0.45, SNOW_TCRIT = -6.0, SNOW_TRATE = 0.042, B0 = 0.031, B1 = 0.201, B2 = 0.185, C_MT = 1.5, RAIN_SCALAR = 0.75;
  const double MA = 28.9644e-3, R = 8.3143, G_STD = 9.80665, P_STD = 101325.0, T_STD = 288.15, LR_STD = 0.0065, RADPERDEG = 0.01745329;
  const LocalTime lt = local_time(cp, d);
  const int nd = lt.Ndays_local;
  const size_t nc = s.ntotal;
  const int month_days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  // local calendar: day of year of every local day (initialize_atmos.c:183-224)
  {
    int day_in_year = lt.local_startday;
    for (int m = 1; m < lt.local_startmonth; m++) {
      int dim = month_days[m - 1];
      if (lt.local_startyear % 4 == 0 && m == 2) dim++;
      day_in_year += dim;
    }
    int year = lt.local_startyear, month = lt.local_startmonth, day = lt.local_startday;
    for (int i = 0; i < nd; i++) {
      s.daily(S::D_yday, i, cell) = day_in_year;
      day_in_year++;
      day++;
      int dim = month_days[month - 1];
      if (year % 4 == 0 && month == 2) dim++;
      if (day > dim) {
        day = 1;
        month++;
        if (month > 12) {
          day_in_year = 1;
          month = 1;
          year++;
        }
      }
    }
  }
  // daily inputs shifted to local days (initialize_atmos.c:359-365)
  for (int idx = 0; idx < nd; idx++) {
    int i = idx;
    if (lt.hour_offset_int > 0) i--;
    if (i < 0) i = 0;
    if (i >= d.Ndays) i = d.Ndays - 1;
    s.daily(S::D_prec, idx, cell) = daily[((size_t)i * 4 + 0) * nc + cell];
    s.daily(S::D_tmax, idx, cell) = daily[((size_t)i * 4 + 1) * nc + cell];
    s.daily(S::D_tmin, idx, cell) = daily[((size_t)i * 4 + 2) * nc + cell];
  }
  // calc_tair / calc_prcp (site and base station coincide: dz == 0, isohyet ratio == 1)
  const double site_elev = (double)(float)cp(CP_elevation);
  const double dz = (site_elev - site_elev) / 1000.0;
  const double lr = cp(CP_T_LAPSE);
  const double isoh = cp(CP_annual_prec) / 10.;
  double ratio;
  if (isoh < 1e-10 && isoh < 1e-10) ratio = 1.;
  else ratio = isoh / isoh;
  for (int i = 0; i < nd; i++) {
    const double tmax = s.daily(S::D_tmax, i, cell) + (dz * lr);
    const double tmin = s.daily(S::D_tmin, i, cell) + (dz * lr);
    const double tmean = (tmax + tmin) / 2.0;
    s.daily(S::D_tday, i, cell) = ((tmax - tmean) * TDAYCOEF) + tmean;
    s.daily(S::D_sprcp, i, cell) = (s.daily(S::D_prec, i, cell) / 10.) * ratio;
  }
  // snowpack (mtclim_vic.c:463-522): two passes, the second starts from the mean pack at the turn of the year
  {
    double snowpack = 0.0;
    for (int pass = 0; pass < 2; pass++) {
      for (int i = 0; i < nd; i++) {
        double newsnow = 0.0, snowmelt = 0.0;
        const double tmin = s.daily(S::D_tmin, i, cell);
        if (tmin <= SNOW_TCRIT) newsnow = s.daily(S::D_sprcp, i, cell);
        else snowmelt = SNOW_TRATE * (tmin - SNOW_TCRIT);
        snowpack += newsnow - snowmelt;
        if (snowpack < 0.0) snowpack = 0.0;
        s.daily(S::D_swe, i, cell) = snowpack;
      }
      if (pass == 0) {
        const int start_yday = (int)s.daily(S::D_yday, 0, cell);
        const int prev_yday = (start_yday == 1) ? 365 : start_yday - 1;
        int count = 0;
        double sum = 0.0;
        for (int i = 1; i < nd; i++) {
          const int y = (int)s.daily(S::D_yday, i, cell);
          if (y == start_yday || y == prev_yday) {
            count++;
            sum += s.daily(S::D_swe, i, cell);
          }
        }
        if (!count) break;
        snowpack = sum / (double)count;
      }
    }
  }
  // diurnal temperature range and its 30-day trailing mean (pulled_boxcar, unweighted)
  for (int i = 0; i < nd; i++) {
    double tmax = s.daily(S::D_tmax, i, cell);
    const double tmin = s.daily(S::D_tmin, i, cell);
    if (tmax < tmin) tmax = tmin;
    s.daily(S::D_dtr, i, cell) = tmax - tmin;
  }
  {
    const int w = (nd >= 30) ? 30 : nd;
    double sum_wt = 0.0;
    for (int i = 0; i < w; i++) sum_wt += 1.0;
    for (int i = w - 1; i < nd; i++) {
      double total = 0.0;
      for (int j = 0; j < w; j++) total += s.daily(S::D_dtr, i - w + j + 1, cell) * 1.0;
      s.daily(S::D_smdtr, i, cell) = total / sum_wt;
    }
    for (int i = 0; i < w - 1; i++) s.daily(S::D_smdtr, i, cell) = s.daily(S::D_smdtr, w - 1, cell);
  }
  // annual and 90-day effective annual precipitation
  double sum_prcp = 0.0;
  for (int i = 0; i < nd; i++) sum_prcp += s.daily(S::D_sprcp, i, cell);
  double ann_prcp = (sum_prcp / (double)nd) * 365.25;
  if (ann_prcp == 0.0) ann_prcp = 1.0;
  if (nd < 90) {
    double eff = (sum_prcp / (double)nd) * 365.25;
    if (eff < 8.0) eff = 8.0;
    for (int i = 0; i < nd; i++) s.daily(S::D_parray, i, cell) = eff;
  } else {
    const int start_yday = (int)s.daily(S::D_yday, 0, cell), end_yday = (int)s.daily(S::D_yday, nd - 1, cell);
    int isloop;
    if (start_yday != 1) isloop = (end_yday == start_yday - 1) ? 1 : 0;
    else isloop = (end_yday == 365 || end_yday == 366) ? 1 : 0;
    // window[k] = k < 90 ? (isloop ? s_prcp[nd-90+k] : s_prcp[k]) : s_prcp[k-90]
    for (int i = 0; i < nd; i++) {
      double sp = 0.0;
      for (int j = 0; j < 90; j++) {
        const int k = i + j;
        const int src = (k < 90) ? (isloop ? nd - 90 + k : k) : k - 90;
        sp += s.daily(S::D_sprcp, src, cell);
      }
      sp = (sp / 90.0) * 365.25;
      s.daily(S::D_parray, i, cell) = (sp < 8.0) ? 8.0 : sp;
    }
  }
  // sky view
  const double slp = cp(CP_slope), eh = cp(CP_ehoriz), wh = cp(CP_whoriz);
  const double avg_horizon = (eh + wh) / 2.0;
  const double horizon_scalar = 1.0 - vsin(avg_horizon * RADPERDEG);
  const double slope_excess = (slp > avg_horizon) ? slp - avg_horizon : 0.0;
  double slope_scalar;
  if (2.0 * avg_horizon > 180.0) slope_scalar = 0.0;
  else {
    slope_scalar = 1.0 - (slope_excess / (180.0 - 2.0 * avg_horizon));
    if (slope_scalar < 0.0) slope_scalar = 0.0;
  }
  const double sky_prop = horizon_scalar * slope_scalar;
  // maximum daily transmittance factor; first guess Tdew = Tmin
  for (int i = 0; i < nd; i++) {
    const double b = B0 + B1 * vexp(-B2 * s.daily(S::D_smdtr, i, cell));
    double tf = 1.0 - 0.9 * vexp(-b * vpow(s.daily(S::D_dtr, i, cell), C_MT));
    if (s.daily(S::D_prec, i, cell) / 10. > d.SW_PREC_THRESH) tf *= RAIN_SCALAR;
    s.daily(S::D_tfmax, i, cell) = tf;
    const double tdew = s.daily(S::D_tmin, i, cell);
    s.daily(S::D_tdew, i, cell) = tdew;
    s.daily(S::D_pva, i, cell) = svp(tdew);
  }
  const double pa = P_STD * vpow(1.0 - (LR_STD * site_elev) / T_STD, G_STD / (LR_STD * (R / MA)));
  for (int i = 0; i < nd; i++) {
    const int yday = (int)s.daily(S::D_yday, i, cell) - 1;
    s.daily(S::D_sdayl, i, cell) = s.at(s.o_dayl() + yday, cell);
    s.daily(S::D_tdew_save, i, cell) = s.daily(S::D_tdew, i, cell);
  }
  disagg_srad_humidity_onetime(d, s, cell, nd, sky_prop, pa);
  double sum_pet = 0.0;
  for (int i = 0; i < nd; i++) sum_pet += s.daily(S::D_pet, i, cell);
  const double ann_pet = (sum_pet / (double)nd) * 365.25;
  const bool arid = (d.VP_ITER == VP_ITER_ANNUAL && ann_pet / ann_prcp >= 2.5);
  if (arid) {  // restore the first guess (mtclim_vic.c:1552-1557)
    for (int i = 0; i < nd; i++) {
      const double tdew = s.daily(S::D_tdew_save, i, cell);
      s.daily(S::D_tdew, i, cell) = tdew;
      s.daily(S::D_pva, i, cell) = svp(tdew);
    }
  }
  int max_iter;
  if (d.VP_ITER == VP_ITER_ALWAYS || arid || d.VP_ITER == VP_ITER_CONVERGE) max_iter = (d.VP_ITER == VP_ITER_CONVERGE) ? 100 : 2;
  else max_iter = 1;
  const double tol = 0.01;
  int iter = 1;
  double rmse = tol + 1;
  while (rmse > tol && iter < max_iter) {
    for (int i = 0; i < nd; i++) s.daily(S::D_tdew_save, i, cell) = s.daily(S::D_tdew, i, cell);
    disagg_srad_humidity_onetime(d, s, cell, nd, sky_prop, pa);
    rmse = 0;
    for (int i = 0; i < nd; i++) {
      const double e = s.daily(S::D_tdew, i, cell) - s.daily(S::D_tdew_save, i, cell);
      rmse += e * e;
    }
    rmse /= nd;
    rmse = vpow(rmse, 0.5);
    iter++;
  }
}

// ---- S3: hourly shortwave of one local day, hours of Tmax / Tmin, spline knots ------------------
VIC_HDI void disagg_day_radiation(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, int cell, int day) {
  typedef DisaggScratch S;
  const LocalTime lt = local_time(cp, d);
  if (day >= lt.Ndays_local) return;
  const int yday = (int)s.daily(S::D_yday, day, cell) - 1;
  const double tmp_rad = s.daily(S::D_srad, day, cell) * s.daily(S::D_sdayl, day, cell) / 3600.;
  for (int j = 0; j < 24; j++) {
    double r = 0;
    r += s.at(s.o_hourfrac() + (size_t)yday * 24 + j, cell);
    r *= tmp_rad;
    s.hourly(0, day * 24 + j, cell) = r;
  }
}

// set_max_min_hour for one day (calc_air_temperature.c:144-198); reads the last hour of the previous day.
// For day 0 the reference reads one element before its heap array (a malloc header, a tiny positive denormal
// when read as a double): treated as "> 0" here.
VIC_HDI void disagg_day_maxmin(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, int cell, int day) {
  typedef DisaggScratch S;
  const LocalTime lt = local_time(cp, d);
  const int nd = lt.Ndays_local;
  if (day >= nd) return;
  int risehour = INT_MIN, sethour = INT_MIN;
  for (int hour = 0; hour < 12; hour++) {
    const int idx = day * 24 + hour;
    const double prev = (idx == 0) ? 1e-300 : s.hourly(0, idx - 1, cell);
    if (s.hourly(0, idx, cell) > 0 && prev <= 0) risehour = hour;
  }
  for (int hour = 12; hour < 24; hour++) {
    const int idx = day * 24 + hour;
    if (s.hourly(0, idx, cell) <= 0 && s.hourly(0, idx - 1, cell) > 0) sethour = hour;
  }
  int tmaxhour, tminhour;
  if (risehour != INT_MIN && sethour != INT_MIN) {
    tmaxhour = (int)(0.67 * (sethour - risehour) + risehour);
    tminhour = risehour - 1;
  } else {
    tminhour = 2;
    tmaxhour = 14;
  }
  s.daily(S::D_tmaxhour, day, cell) = tmaxhour;
  s.daily(S::D_tminhour, day, cell) = tminhour;
  // spline knots (HourlyT, calc_air_temperature.c:101-120): two per day, plus one mirrored at each end
  const int hour0 = day * 24;
  const int j = 1 + 2 * day;
  const double tmin = s.daily(S::D_tmin, day, cell), tmax = s.daily(S::D_tmax, day, cell);
  if (tminhour < tmaxhour) {
    s.knot(0, j, cell) = tminhour + hour0; s.knot(1, j, cell) = tmin;
    s.knot(0, j + 1, cell) = tmaxhour + hour0; s.knot(1, j + 1, cell) = tmax;
  } else {
    s.knot(0, j, cell) = tmaxhour + hour0; s.knot(1, j, cell) = tmax;
    s.knot(0, j + 1, cell) = tminhour + hour0; s.knot(1, j + 1, cell) = tmin;
  }
  const int n = nd * 2 + 2;
  if (day == 0) {  // x[0] = x[2] - 24, y[0] = y[2]
    s.knot(0, 0, cell) = s.knot(0, 2, cell) - 24;
    s.knot(1, 0, cell) = s.knot(1, 2, cell);
  }
  if (day == nd - 1) {  // x[n-1] = x[n-3] + 24, y[n-1] = y[n-3]
    s.knot(0, n - 1, cell) = s.knot(0, n - 3, cell) + 24;
    s.knot(1, n - 1, cell) = s.knot(1, n - 3, cell);
  }
}

// ---- S4: Hermite coefficients of one knot interval (zero slopes at the knots) -------------------
VIC_HDI void disagg_knot_coeff(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, int cell, int i) {
  const LocalTime lt = local_time(cp, d);
  const int n = lt.Ndays_local * 2 + 2;
  if (i >= n - 1) return;
  const double dx = s.knot(0, i + 1, cell) - s.knot(0, i, cell);
  const double divdf1 = (s.knot(1, i + 1, cell) - s.knot(1, i, cell)) / dx;
  const double divdf3 = 0. + 0. - 2 * divdf1;
  s.knot(2, i, cell) = (divdf1 - 0. - divdf3) / dx;
  s.knot(3, i, cell) = divdf3 / (dx * dx);
}

// ---- S5: hourly air temperature and vapour pressure of one local day ---------------------------
VIC_HDI void disagg_day_hourly(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, int cell, int day) {
  typedef DisaggScratch S;
  const LocalTime lt = local_time(cp, d);
  const int nd = lt.Ndays_local;
  if (day >= nd) return;
  const int n = nd * 2 + 2;
  for (int hh = 0; hh < 24; hh++) {
    const int hour = day * 24 + hh;
    int klo = 0, khi = n - 1;
    while (khi - klo > 1) {
      const int k = (khi + klo) >> 1;
      if (s.knot(0, k, cell) > (double)hour) khi = k;
      else klo = k;
    }
    const double dx = (double)hour - s.knot(0, klo, cell);
    s.hourly(1, hour, cell) = s.knot(1, klo, cell) + dx * (0. + dx * (s.knot(2, klo, cell) + dx * s.knot(3, klo, cell)));
  }
  // vapour pressure: linear between the hours of Tmin of successive days (initialize_atmos.c:1103-1151)
  const double vp0 = s.daily(S::D_pva, day, cell);
  if (d.VP_INTERP) {
    const int tmh = (int)s.daily(S::D_tminhour, day, cell);
    const int tmh_m = (day > 0) ? (int)s.daily(S::D_tminhour, day - 1, cell) : 0;
    const int tmh_p = (day < nd - 1) ? (int)s.daily(S::D_tminhour, day + 1, cell) : 0;
    double delta_t_minus, delta_t_plus;
    if (day == 0 && nd == 1) { delta_t_minus = 24; delta_t_plus = 24; }
    else if (day == 0) { delta_t_minus = 24; delta_t_plus = tmh_p + 24 - tmh; }
    else if (day == nd - 1) { delta_t_minus = tmh + 24 - tmh_m; delta_t_plus = 24; }
    else { delta_t_minus = tmh + 24 - tmh_m; delta_t_plus = tmh_p + 24 - tmh; }
    for (int hour = 0; hour < 24; hour++) {
      double v;
      if (hour < tmh) {
        if (day > 0) {
          const double vm = s.daily(S::D_pva, day - 1, cell);
          v = vm + (vp0 - vm) * (hour + 24 - tmh_m) / delta_t_minus;
        } else v = vp0;
      } else {
        if (day < nd - 1) {
          const double vpn = s.daily(S::D_pva, day + 1, cell);
          v = vp0 + (vpn - vp0) * (hour - tmh) / delta_t_plus;
        } else v = vp0;
      }
      s.hourly(2, day * 24 + hour, cell) = v;
    }
  } else {
    for (int hour = 0; hour < 24; hour++) s.hourly(2, day * 24 + hour, cell) = vp0;
  }
}

VIC_HD double mt_longwave(double tskc, double air_temp, double vp, int LW_TYPE, int LW_CLOUD) {  // calc_longwave.c:8-73
  double emissivity_clear = 0, x;
  air_temp += KELVIN;
  vp /= 100;
  if (LW_TYPE == LW_TVA) emissivity_clear = 0.740 + 0.0049 * vp;
  else if (LW_TYPE == LW_ANDERSON) emissivity_clear = 0.68 + 0.036 * vpow(vp, 0.5);
  else if (LW_TYPE == LW_BRUTSAERT) { x = vp / air_temp; emissivity_clear = 1.24 * vpow(x, 0.14285714); }
  else if (LW_TYPE == LW_SATTERLUND) emissivity_clear = 1.08 * (1 - vexp(-1 * vpow(vp, (air_temp / 2016))));
  else if (LW_TYPE == LW_IDSO) emissivity_clear = 0.7 + 5.95e-5 * vp * vexp(1500 / air_temp);
  else if (LW_TYPE == LW_PRATA) { x = 46.5 * vp / air_temp; emissivity_clear = 1 - (1 + x) * vexp(-1 * vpow((1.2 + 3 * x), 0.5)); }
  double emissivity;
  if (LW_CLOUD == LW_CLOUD_DEARDORFF) {
    const double cloudfrac = 0.65 * tskc * tskc;
    emissivity = cloudfrac * 1.0 + (1 - cloudfrac) * emissivity_clear;
  } else {
    const double cloudfactor = 1.0 + 0.17 * tskc * tskc;
    emissivity = cloudfactor * emissivity_clear;
  }
  return emissivity * STEFAN_B * air_temp * air_temp * air_temp * air_temp / 1.0;  // LWAVE_COR 1
}

// ---- S6: one model record of one cell -> forcing record [FV][slot] (column-major over cells) --------
// daily: as in S2 (wind is column 3).  frec: this record's slab [f_stride][ncell].
VIC_HDI void disagg_record(const CellPar& cp, const DisaggOpts& d, const DisaggScratch& s, const double* daily, double* frec, int cell, int rec) {
  typedef DisaggScratch S;
  const LocalTime lt = local_time(cp, d);
  const size_t nc = s.ntotal;
  const int NF = d.NF, NR = d.NR, ns = d.f_nslot;
  const int stepspday = 24 / d.dt;
  const double elev = (double)(float)cp(CP_elevation);
  const double Rd = 287, Gg = 9.81;
#define FOUT(var, slot) frec[((size_t)(var) * ns + (slot)) * nc + cell]
  double sum_prec = 0, sum_wind = 0, sum_sw = 0, sum_t = 0, sum_vp = 0, sum_vpd = 0, sum_tskc = 0, sum_lw = 0;
  for (int i = 0; i < NF; i++) {
    const int hour = local_hour(d, lt, rec, i);
    const int dayidx = (int)((float)hour / 24.0);
    // daily values indexed by local day; prec is spread evenly over the sub-steps of its day
    int src = dayidx;
    if (lt.hour_offset_int > 0) src--;
    if (src < 0) src = 0;
    if (src >= d.Ndays) src = d.Ndays - 1;
    const double prec = daily[((size_t)src * 4 + 0) * nc + cell] / (float)(NF * stepspday);
    const double wind = daily[((size_t)src * 4 + 3) * nc + cell];
    FOUT(FV_prec, i) = prec;
    sum_prec += prec;
    FOUT(FV_wind, i) = wind;
    sum_wind += wind;
    double sw = 0, ta = 0, vp = 0;
    for (int idx = hour; idx < hour + d.SNOW_STEP; idx++) {
      sw += s.hourly(0, idx, cell);
      ta += s.hourly(1, idx, cell);
      vp += s.hourly(2, idx, cell);
    }
    sw /= d.SNOW_STEP;
    ta /= d.SNOW_STEP;
    vp /= d.SNOW_STEP;
    FOUT(FV_shortwave, i) = sw;
    sum_sw += sw;
    FOUT(FV_air_temp, i) = ta;
    sum_t += ta;
    FOUT(FV_vp, i) = vp;
    sum_vp += vp;
    const double tskc = s.daily(S::D_tskc, dayidx, cell);
    FOUT(FV_tskc, i) = tskc;
    sum_tskc += tskc;
  }
  if (NF > 1) {
    FOUT(FV_prec, NR) = sum_prec;
    FOUT(FV_wind, NR) = sum_wind / (float)NF;
    FOUT(FV_shortwave, NR) = sum_sw / (float)NF;
    FOUT(FV_air_temp, NR) = sum_t / (float)NF;
    FOUT(FV_vp, NR) = sum_vp / (float)NF;
    FOUT(FV_tskc, NR) = sum_tskc / (float)NF;
  }
  if (d.dt == 24 && NF > 1) {  // initialize_atmos.c:512-515: index j == NF after the loop, i.e. the step mean
    if (FOUT(FV_wind, NR) < d.MIN_WIND_SPEED) FOUT(FV_wind, NR) = d.MIN_WIND_SPEED;
  }
  // pressure and density from the hypsometric relation (PLAPSE) or constants
  for (int i = 0; i <= (NF > 1 ? NF : 0); i++) {
    const int slot = (i == NF) ? NR : i;
    const double ta = FOUT(FV_air_temp, slot);
    double p, rho;
    if (d.PLAPSE) {
      p = PS_PM * vexp(-elev * Gg / (Rd * (KELVIN + ta + 0.5 * elev * LAPSE_PM)));
      rho = p / (Rd * (KELVIN + ta));
    } else {
      p = 95500.;
      rho = 0.003486 * p / (275.0 + ta);
    }
    FOUT(FV_pressure, slot) = p;
    FOUT(FV_density, slot) = rho;
  }
  // vapour pressure deficit, longwave
  double sum2 = 0;
  for (int i = 0; i < NF; i++) {
    double vpd = svp(FOUT(FV_air_temp, i)) - FOUT(FV_vp, i);
    if (vpd < 0) {
      vpd = 0;
      FOUT(FV_vp, i) = svp(FOUT(FV_air_temp, i));
    }
    FOUT(FV_vpd, i) = vpd;
    sum_vpd += vpd;
    sum2 += FOUT(FV_vp, i);
    const double lw = mt_longwave(FOUT(FV_tskc, i), FOUT(FV_air_temp, i), FOUT(FV_vp, i), d.LW_TYPE, d.LW_CLOUD);
    FOUT(FV_longwave, i) = lw;
    sum_lw += lw;
  }
  if (d.VP_INTERP) {
    if (NF > 1) {
      FOUT(FV_vpd, NR) = sum_vpd / (float)NF;
      FOUT(FV_vp, NR) = sum2 / (float)NF;
    }
  } else FOUT(FV_vpd, NR) = (svp(FOUT(FV_air_temp, NR)) - FOUT(FV_vp, NR));
  if (NF > 1) FOUT(FV_longwave, NR) = sum_lw / (float)NF;
  // snow flag: can snow fall in any band during this step?
  if (!d.OUTPUT_FORCE) {
    double min_Tfactor = cp.band(CB_Tfactor, 0);
    for (int b = 1; b < d.Nbands; b++)
      if (cp.band(CB_Tfactor, b) < min_Tfactor) min_Tfactor = cp.band(CB_Tfactor, b);
    double any = 0;
    const double thr = (d.TEMP_TH_TYPE == VIC_412) ? cp(CP_MAX_SNOW_TEMP) : (cp(CP_MAX_SNOW_TEMP) + cp(CP_MIN_RAIN_TEMP) / 2);
    for (int i = 0; i < NF; i++) {
      const bool f = ((FOUT(FV_air_temp, i) + min_Tfactor) < thr && FOUT(FV_prec, i) > 0);
      FOUT(FV_snowflag, i) = f ? 1.0 : 0.0;
      if (f) any = 1.0;
    }
    FOUT(FV_snowflag, NR) = any;
  } else {
    for (int i = 0; i < ns; i++) FOUT(FV_snowflag, i) = 0.0;
  }
#undef FOUT
}

// options of the two ABI structs as the stage functions want them
inline DisaggOpts disagg_opts_from_abi(const vicgpu_options& a, const vicgpu_disagg_options& b, int f_nslot) {
  DisaggOpts d;
  d.dt = a.dt; d.SNOW_STEP = a.SNOW_STEP; d.NF = a.NF; d.NR = a.NR; d.nrecs = a.nrecs; d.TEMP_TH_TYPE = a.TEMP_TH_TYPE; d.Nbands = a.Nbands;
  d.MIN_WIND_SPEED = a.MIN_WIND_SPEED;
  d.starthour = b.starthour; d.startyear = b.startyear; d.startmonth = b.startmonth; d.startday = b.startday; d.Ndays = b.Ndays;
  d.PLAPSE = b.PLAPSE; d.MTCLIM_SWE_CORR = b.MTCLIM_SWE_CORR; d.VP_ITER = b.VP_ITER; d.VP_INTERP = b.VP_INTERP; d.LW_TYPE = b.LW_TYPE;
  d.LW_CLOUD = b.LW_CLOUD; d.OUTPUT_FORCE = b.OUTPUT_FORCE; d.SW_PREC_THRESH = b.SW_PREC_THRESH; d.f_nslot = f_nslot;
  return d;
}

}  // namespace vic
#endif

// vic_leaf.cuh -- closed-form leaf relations of the VIC land-surface step.
// Each function states the reference routine whose arithmetic (operation order included,
// because FP64 parity is judged at 1e-9) it reproduces.
#ifndef VIC_LEAF_CUH
#define VIC_LEAF_CUH
#include "vic_common.cuh"
#include "vic_types.cuh"

namespace vic {

// saturated vapour pressure [Pa], svp.c:7-23 (Handbook of Hydrology 4.2.2)
VIC_HD double svp(double temp) {
  double SVP = A_SVP * vexp((B_SVP * temp) / (C_SVP + temp));
  if (temp < 0) SVP *= 1.0 + .00972 * temp + .000042 * temp * temp;
  return SVP * 1000.;
}

// d(svp)/dT [Pa/K], svp.c:25-37
VIC_HD double svp_slope(double temp) {
  return (B_SVP * C_SVP) / ((C_SVP + temp) * (C_SVP + temp)) * svp(temp);
}

VIC_HD double linear_interp(double x, double lx, double ux, double ly, double uy) {  // modify_Ksat.c:7
  return (x - lx) / (ux - lx) * (uy - ly) + ly;
}

// vegetation height from displacement, calc_veg_params.c:26-37 (COEF_DRAG 0.2)
VIC_HD double calc_veg_height(double displacement, double lai) {
  double X = 0.2 * lai;
  return displacement / (1.1 * vlog(1 + vpow(X, 0.25)));
}

// canopy resistance, penman.c:44-90 (Wigmosta et al. 1994)
VIC_HD double calc_rc(double rs, double net_short, float RGL, double tair, double vpd, double lai,
                      double gsm_inv, bool ref_crop) {
  const double CLOSURE = 4000, RSMAX = 5000, VPDMINFACTOR = 0.1;
  double rc;
  if (rs == 0) {
    rc = 0;
  } else if (lai == 0) {
    rc = HUGE_RESIST;
  } else if (ref_crop) {
    rc = rs / (lai * 0.5);
  } else {
    double DAYfactor;
    if (rs > 0.) {
      double f = net_short / RGL;
      DAYfactor = (1. + f) / (f + rs / RSMAX);
    } else DAYfactor = 1.;
    double Tfactor = .08 * tair - 0.0016 * tair * tair;
    Tfactor = (Tfactor <= 0.0) ? 1e-10 : Tfactor;
    double vpdfactor = 1 - vpd / CLOSURE;
    vpdfactor = (vpdfactor < VPDMINFACTOR) ? VPDMINFACTOR : vpdfactor;
    rc = rs / (lai * gsm_inv * Tfactor * vpdfactor) * DAYfactor;
    rc = (rc > RSMAX) ? RSMAX : rc;
  }
  return rc;
}

// Penman-Monteith evaporation [mm/day], penman.c:96-144.  The terms that depend on air temperature and elevation only are
// split off (PenmanPre) so that a root finder that re-evaluates the evaporation with a new net radiation does not recompute
// them: the same operations in the same order, hence the same bits.
struct PenmanPre {
  double slope, lv, gamma, r_air;
};
VIC_HD PenmanPre penman_pre(double tair, double elevation) {
  PenmanPre p;
  p.slope = svp_slope(tair);
  double h = 287 / 9.81 * ((tair + 273.15) + 0.5 * (double)elevation * LAPSE_PM);
  double pz = PS_PM * vexp(-(double)elevation / h);
  p.lv = 2501000 - 2361 * tair;
  p.gamma = 1628.6 * pz / p.lv;
  p.r_air = 0.003486 * pz / (275 + tair);
  return p;
}
VIC_HD double penman_eval(const PenmanPre& p, double rad, double vpd, double ra, double rc, double rarc) {
  double evap = (p.slope * rad + p.r_air * CP_PM * vpd / ra) / (p.lv * (p.slope + p.gamma * (1 + div_zn(rc + rarc, ra)))) * SEC_PER_DAY;
  if (vpd >= 0.0 && evap < 0.0) evap = 0.0;
  return evap;
}
VIC_HD double penman(double tair, double elevation, double rad, double vpd, double ra, double rc, double rarc) {
  return penman_eval(penman_pre(tair, elevation), rad, vpd, ra, rc, rarc);
}

// Richardson-number stability multiplier, StabilityCorrection.c:44-81; lg = log((Z - d) / Z0)
VIC_HD double stability_correction_lg(double Z, double d, double TSurf, double Tair, double Wind, double lg) {
  double Correction = 1.0;
  const double RiCr = 0.2;
  if (TSurf != Tair) {
    double Ri = G_GRAV * (Tair - TSurf) * (Z - d) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * Wind * Wind);
    double RiLimit = (Tair + 273.15) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * (lg + 5));
    if (Ri > RiLimit) Ri = RiLimit;
    if (Ri > 0.0) Correction = (1 - Ri / RiCr) * (1 - Ri / RiCr);
    else {
      if (Ri < -0.5) Ri = -0.5;
      Correction = sqrt(1 - 16 * Ri);
    }
  }
  return Correction;
}
VIC_HD double stability_correction(double Z, double d, double TSurf, double Tair, double Wind, double Z0) {
  if (TSurf == Tair) return 1.0;
  return stability_correction_lg(Z, d, TSurf, Tair, Wind, vlog((Z - d) / Z0));
}
// log((Z - d) / Z0) does not change between the evaluations of one solve: computed at the first evaluation that needs it
struct StabLog {
  double lg;
  int ok;
  VIC_HD void reset() { lg = 0; ok = 0; }
  VIC_HD double correction(double Z, double d, double TSurf, double Tair, double Wind, double Z0) {
    if (TSurf == Tair) return 1.0;
    if (!ok) {
      lg = vlog((Z - d) / Z0);
      ok = 1;
    }
    return stability_correction_lg(Z, d, TSurf, Tair, Wind, lg);
  }
};

// rain / snow partition, calc_rainonly.c:12-103 (mu == 1 on this path)
VIC_HD double calc_rainonly(double air_temp, double prec, double MAX_SNOW_TEMP, double MIN_RAIN_TEMP, double mu, int TEMP_TH_TYPE) {
  const double MIN_PREC = 1.e-5;
  double rainonly = 0;
  if (TEMP_TH_TYPE == VIC_412) {
    if (air_temp < MAX_SNOW_TEMP && air_temp > MIN_RAIN_TEMP)
      rainonly = (air_temp - MIN_RAIN_TEMP) / (MAX_SNOW_TEMP - MIN_RAIN_TEMP) * prec;
    else if (air_temp >= MAX_SNOW_TEMP) rainonly = prec;
  } else {
    double rfrac;
    double TT = MIN_RAIN_TEMP, TR = MAX_SNOW_TEMP;
    double D = 1.4 * TR;
    double E1 = 5. * vpow((air_temp - TT) / D, 3.0);
    double E2 = 6.76 * vpow((air_temp - TT) / D, 2.0);
    double E3 = 3.19 * (air_temp - TT) / D;
    if (air_temp <= TT) rfrac = E1 + E2 + E3 + 0.5;
    else rfrac = E1 - E2 + E3 + 0.5;
    if (rfrac < 0.) rfrac = 0.;
    if (rfrac > 1.) rfrac = 1.;
    rainonly = rfrac * prec;
  }
  if (rainonly < MIN_PREC) rainonly = 0.;
  if ((prec - rainonly) < MIN_PREC) rainonly = prec;
  if (mu < 1) rainonly = prec;
  return rainonly;
}

// density of fresh snow, snow_utility.c:199-227
VIC_HD double new_snow_density(double air_temp, int SNOW_DENSITY) {
  double density_new;
  if (SNOW_DENSITY == DENS_SNTHRM) {
    density_new = 67.9 + 51.3 * vexp(air_temp / 2.6);
  } else {
    air_temp = air_temp * 9. / 5. + 32.;
    if (air_temp > 0) density_new = (double)NEW_SNOW_DENSITY + 1000. * (air_temp / 100.) * (air_temp / 100.);
    else density_new = (double)NEW_SNOW_DENSITY;
  }
  return density_new;
}

// pack density after compaction / ageing, snow_utility.c:9-197
VIC_HD double snow_density(const SnowPack& snow, double new_snow, double sswq, double Tgrnd, double Tair, double dt, int SNOW_DENSITY) {
  const double MAX_CHANGE = 0.9;
  double density_new, density, depth, swq, delta_depth, depth_new;
  (void)Tgrnd;
  if (new_snow > 0.) density_new = new_snow_density(Tair, SNOW_DENSITY);
  else density_new = 0.0;
  double Tavg = snow.surf_temp + KELVIN;
  if (SNOW_DENSITY == DENS_SNTHRM) {
    if (new_snow > 0.) {
      if (snow.depth > 0.0) density = snow.density;
      else density = density_new;
    } else density = snow.density;
    double dexpf = vexp(-SNDENS_C1 * (KELVIN - Tavg));
    double dm;
    if (new_snow > 0.0 && density_new > 0.0) dm = (SNDENS_DMLIMIT > 1.15 * density_new) ? SNDENS_DMLIMIT : 1.15 * density_new;
    else dm = SNDENS_DMLIMIT;
    double c3, c4;
    if (density <= dm) { c3 = 1.0; c4 = 1.0; }
    else { c3 = vexp(-0.046 * (density - dm)); c4 = 1.0; }
    if ((snow.surf_water + snow.pack_water) / snow.depth > 0.01) c4 = 2.0;
    double ddz1 = -SNDENS_C2 * c3 * c4 * dexpf;
    double f = SNDENS_F;
    swq = new_snow / 1000. + f * sswq;
    double ddz2;
    if (new_snow > 0.0) {
      double Ps = 0.5 * G_GRAV * RHO_W * swq;
      ddz2 = -Ps / SNDENS_ETA0 * vexp(-(-SNDENS_C5 * (Tavg - KELVIN) + SNDENS_C6 * density));
    } else ddz2 = 0.0;
    double CR = -ddz1 - ddz2;
    density = density * (1 + CR * dt * SECPHOUR);
  } else {
    depth = snow.depth;
    swq = sswq;
    if (new_snow > 0) {
      if (depth > 0.) {
        delta_depth = (((new_snow / 25.4) * (depth / 0.0254)) / (swq / 0.0254) * vpow((depth / 0.0254) / 10., 0.35)) * 0.0254;
        if (delta_depth > MAX_CHANGE * depth) delta_depth = MAX_CHANGE * depth;
        depth_new = new_snow / density_new;
        depth = depth - delta_depth + depth_new;
        swq += new_snow / 1000.;
        density = 1000. * swq / depth;
      } else {
        density = density_new;
        swq += new_snow / 1000.;
        depth = 1000. * swq / density;
      }
    } else density = 1000. * swq / snow.depth;
    if (depth > 0.) {
      double overburden = 0.5 * G_GRAV * RHO_W * swq;
      double viscosity = SNDENS_ETA0 * vexp(-SNDENS_C5 * (Tavg - KELVIN) + SNDENS_C6 * density);
      delta_depth = overburden / viscosity * depth * dt * SECPHOUR;
      if (delta_depth > MAX_CHANGE * depth) delta_depth = MAX_CHANGE * depth;
      depth -= delta_depth;
      density = 1000. * swq / depth;
    }
  }
  return density;
}

// snow surface albedo, snow_utility.c:229-307
VIC_HD double snow_albedo(double new_snow, double swq, double depth, double albedo, double cold_content, double dt,
                          int last_snow, bool MELTING, const CellPar& cp, int SNOW_ALBEDO) {
  if (new_snow > TraceSnow && cold_content < 0.0) albedo = cp(CP_NEW_SNOW_ALB);
  else if (swq > 0.0) {
    if (SNOW_ALBEDO == SUN1999) {
      if (depth > 0.025) albedo = 0.5 + (albedo - 0.5) * vexp(-0.01 * dt / 24);
      else if (cold_content < 0.0) albedo = albedo - 0.006 * dt / 24;
      else albedo = albedo - 0.071 * dt / 24;
      if (albedo < 0) albedo = 0;
    } else {
      if (cold_content < 0.0 && !MELTING)
        albedo = cp(CP_NEW_SNOW_ALB) * vpow(cp(CP_SNOW_ALB_ACCUM_A), vpow((double)last_snow * dt / 24., cp(CP_SNOW_ALB_ACCUM_B)));
      else
        albedo = cp(CP_NEW_SNOW_ALB) * vpow(cp(CP_SNOW_ALB_THAW_A), vpow((double)last_snow * dt / 24., cp(CP_SNOW_ALB_THAW_B)));
    }
  } else albedo = 0;
  return albedo;
}

// latent heat over snow, latent_heat_from_snow.c:7-68
VIC_HD void latent_heat_from_snow(double AirDens, double EactAir, double Lv, double Press, double Ra, double TMean, double Vpd,
                                  double* LatentHeat, double* LatentHeatSublimation, double* VaporMassFlux,
                                  double* BlowingMassFlux, double* SurfaceMassFlux) {
  double EsSnow = svp(TMean);
  *SurfaceMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra;
  if (Vpd == 0.0 && *SurfaceMassFlux < 0.0) *SurfaceMassFlux = 0.0;
  *VaporMassFlux = *SurfaceMassFlux + *BlowingMassFlux;
  if (TMean >= 0.0) {
    *LatentHeat = Lv * (*VaporMassFlux);
    *LatentHeatSublimation = 0;
  } else {
    double Ls = (677. - 0.07 * TMean) * JOULESPCAL * GRAMSPKG;
    *LatentHeatSublimation = Ls * (*VaporMassFlux);
    *LatentHeat = 0;
  }
}

// Johansen thermal conductivity, soil_conduction.c:7-105.  Everything that depends on the layer's soil constants only -- the dry and
// solid conductivities and the two powers of the unfrozen saturated conductivity -- is split off (SoilKPre): the library evaluates it
// once per (cell, layer) when the cells are set (derive_cell_constants, vic_engine.cuh) instead of at every node and layer of every
// step; the same operations in the same order, hence the same bits.
struct SoilKPre {
  double Kdry, porosity, Ks_pow, Ksat_unfrozen;  // Ks_pow = Ks^(1 - porosity); Ksat_unfrozen = Ks_pow * Kw^porosity
};
#define VIC_NKPRE 4
VIC_HD SoilKPre soil_k_pre(double soil_dens_min, double bulk_dens_min, double quartz, double soil_density, double bulk_density, double organic) {
  const double Kw = 0.57, Kdry_org = 0.05, Ks_org = 0.25;
  SoilKPre p;
  double Kdry_min = (0.135 * bulk_dens_min + 64.7) / (soil_dens_min - 0.947 * bulk_dens_min);
  p.Kdry = (1 - organic) * Kdry_min + organic * Kdry_org;
  p.porosity = 1.0 - bulk_density / soil_density;
  double Ks_min;
  if (quartz < .2) Ks_min = vpow(7.7, quartz) * vpow(3.0, 1.0 - quartz);
  else Ks_min = vpow(7.7, quartz) * vpow(2.2, 1.0 - quartz);
  double Ks = (1 - organic) * Ks_min + organic * Ks_org;
  p.Ks_pow = vpow(Ks, 1.0 - p.porosity);
  p.Ksat_unfrozen = p.Ks_pow * vpow(Kw, p.porosity);
  return p;
}
VIC_HD double soil_conductivity_pre(double moist, double Wu, const SoilKPre& p) {
  const double Ki = 2.2, Kw = 0.57;
  double Ke, Ksat, K;
  if (moist > 0.) {
    double Sr = moist / p.porosity;
    if (Wu == moist) {
      Ksat = p.Ksat_unfrozen;
      Ke = 0.7 * vlog10(Sr) + 1.0;
    } else {
      Ksat = p.Ks_pow * vpow(Ki, p.porosity - Wu) * vpow(Kw, Wu);
      Ke = Sr;
    }
    K = (Ksat - p.Kdry) * Ke + p.Kdry;
    if (K < p.Kdry) K = p.Kdry;
  } else K = p.Kdry;
  return K;
}
VIC_HD double soil_conductivity(double moist, double Wu, double soil_dens_min, double bulk_dens_min, double quartz,
                                double soil_density, double bulk_density, double organic) {
  return soil_conductivity_pre(moist, Wu, soil_k_pre(soil_dens_min, bulk_dens_min, quartz, soil_density, bulk_density, organic));
}

// volumetric heat capacity, soil_conduction.c:108-139
VIC_HD double volumetric_heat_capacity(double soil_fract, double water_fract, double ice_fract, double organic_fract) {
  double Cs = 2.0e6 * soil_fract * (1 - organic_fract);
  Cs += 2.7e6 * soil_fract * organic_fract;
  Cs += 4.2e6 * water_fract;
  Cs += 1.9e6 * ice_fract;
  Cs += 1.3e3 * (1. - (soil_fract + water_fract + ice_fract));
  return Cs;
}

// unfrozen water at T, soil_conduction.c:830-863
VIC_HD double maximum_unfrozen_water(double T, double max_moist, double bubble, double expt) {
  double unfrozen;
  if (T <= 0) {
    unfrozen = max_moist * vpow((-Lf * T) / 273.16 / (9.81 * bubble / 100.), -(2.0 / (expt - 3.0)));
    if (unfrozen > max_moist) unfrozen = max_moist;
    if (unfrozen < 0) unfrozen = 0;
  } else unfrozen = max_moist;
  return unfrozen;
}

// Liang et al. (1999) first-node temperature, estimate_T1.c:8-47
VIC_HD double estimate_T1(double Ts, double T1_old, double T2, double D1, double D2, double kappa1, double kappa2,
                          double Cs1, double Cs2, double dp, double delta_t) {
  (void)Cs1;
  double C1 = Cs2 * dp / D2 * (1. - vexp(-D2 / dp));
  double C2 = -(1. - vexp(D1 / dp)) * vexp(-D2 / dp);
  double C3 = kappa1 / D1 - kappa2 / D1 + kappa2 / D1 * vexp(-D1 / dp);
  double T1 = (kappa1 / 2. / D1 / D2 * (Ts) + C1 / delta_t * T1_old + (2. * C2 - 1. + vexp(-D1 / dp)) * kappa2 / 2. / D1 / D2 * T2) /
              (C1 / delta_t + kappa2 / D1 / D2 * C2 + C3 / 2. / D2);
  return T1;
}

// water-table position from the zwt-v-moisture curve, compute_zwt.c:7-40
VIC_HD double compute_zwt(const CellPar& cp, int curve, double moist) {
  double zwt = vnan();
  int i = VICGPU_NZWT - 1;
  while (i >= 1 && moist > cp.zwt(CZ_moist, curve, i)) i--;
  if (i == VICGPU_NZWT - 1) {
    if (moist < cp.zwt(CZ_moist, curve, i)) zwt = vnan();
    else if (moist == cp.zwt(CZ_moist, curve, i)) zwt = cp.zwt(CZ_zwt, curve, i);
  } else {
    zwt = cp.zwt(CZ_zwt, curve, i + 1) + (cp.zwt(CZ_zwt, curve, i) - cp.zwt(CZ_zwt, curve, i + 1)) *
                                             (moist - cp.zwt(CZ_moist, curve, i + 1)) /
                                             (cp.zwt(CZ_moist, curve, i) - cp.zwt(CZ_moist, curve, i + 1));
  }
  return zwt;
}

}  // namespace vic
#endif

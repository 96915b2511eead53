// vic_glacier.cuh -- land-surface step of a glacier HRU (PCIC glacier mass-balance mode):
//   ice-surface energy balance residual          GlacierEnergyBalance.c:15-93, latent_heat_from_glacier.c:8-52
//   bare-ice surface temperature / melt          glacier_melt.c:65-218, solve_glacier.c:5-108
//   snow on ice (incl. firn -> ice conversion)   solve_snow_glac.c:5-290, snow_melt_glac.c:14-406 (vic::snow_melt, GLAC mode)
//   sub-step loop, mass balance, linear-reservoir outflow, runoff   surface_fluxes_glac.c:6-613
//
// Reference behaviour that is undefined and is DEFINED here (and in the oracle build through the
// documented one-line patch of oracle/Makefile): surface_fluxes_glac.c:68 leaves `delta_coverage`
// uninitialised and reads it (:431-440) on every sub-step that takes the bare-ice branch, which never
// assigns it.  Here, and in the patched oracle, it starts at 0 on every call.
#ifndef VIC_GLACIER_CUH
#define VIC_GLACIER_CUH
#include "vic_surface.cuh"

namespace vic {

struct GlacierEB {
  double Dt, Ra, Z, Z0_snow, AirDens, EactAir, LongSnowIn, Lv, Press, Rain, NetShortUnder, Vpd, Wind, OldTSurf, IceDepth, Tair, TGrnd;
  RaUsed* Ra_used;
  double *AdvectedEnergy, *DeltaColdContent, *GroundFlux, *LatentHeat, *LatentHeatSub, *NetLongUnder, *SensibleHeat, *vapor_flux;
  StabLog stab;

  VIC_HDI double operator()(double TSurf) { return eval(TSurf); }
  // the residual proper, inlined into the one call site of glacier_melt's solve
  VIC_HD double eval(double TSurf) {
    const double TMean = (TSurf + TGrnd) / 2;
    const double OldTMean = (OldTSurf + TGrnd) / 2;
    const double Density = RHO_W;
    const double temp_IceDepth = IceDepth / 1000.;
    if (Wind > 0.0) Ra_used->surface = Ra / stab.correction(Z, 0., TSurf, Tair, Wind, Z0_snow);
    else Ra_used->surface = HUGE_RESIST;
    const double Tmp = TSurf + KELVIN;
    (*NetLongUnder) = LongSnowIn - STEFAN_B * Tmp * Tmp * Tmp * Tmp;
    const double NetRad = NetShortUnder + (*NetLongUnder);
    *SensibleHeat = AirDens * Cp * (Tair - TSurf) / Ra_used->surface;
    double VaporMassFlux = div_pos(*vapor_flux * Density, Dt);
    // latent_heat_from_glacier.c
    {
      const double EsSnow = svp(TSurf);
      VaporMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra_used->surface;
      if (Vpd == 0.0 && VaporMassFlux < 0.0) VaporMassFlux = 0.0;
      if (TSurf >= 0.0) {
        *LatentHeat = Lv * (VaporMassFlux);
        *LatentHeatSub = 0;
      } else {
        const double Ls = (677. - 0.07 * TSurf) * JOULESPCAL * GRAMSPKG;
        *LatentHeatSub = Ls * (VaporMassFlux);
        *LatentHeat = 0;
      }
    }
    *vapor_flux = VaporMassFlux * Dt / Density;
    if (TSurf == 0) *AdvectedEnergy = div_pos((CH_WATER * (Tair)*Rain), (Dt));
    else *AdvectedEnergy = 0.;
    *DeltaColdContent = div_pos(CH_ICE * temp_IceDepth * (TMean - OldTMean), (Dt));
    *GroundFlux = (GLAC_K_ICE + TSurf * (-0.0142)) * (TGrnd - TSurf) / temp_IceDepth;
    const double Fbal = NetRad + *SensibleHeat + *LatentHeat + *LatentHeatSub + *AdvectedEnergy;
    double RestTerm = Fbal - *DeltaColdContent + *GroundFlux;
    if (TSurf == 0.0 && RestTerm >= 0.) RestTerm = 0.;
    return RestTerm;
  }
};

// glacier_melt.c:65-218.  Returns 0 or ERROR_I; *melt in metres of water.
template <int NN>
VIC_HDI int glacier_melt(double Le, double NetShort, double Tgrnd, double Z0_snow, double aero_resist, RaUsed& aero_resist_used, double air_temp,
                         double delta_t, double density, double LongIn, double pressure, double rainfall, double vp, double vpd, double wind,
                         double z2, double* NetLong, double* OldTSurf, double* melt, EnergyBal<NN>& energy, Glacier& glacier, const CellPar& cp,
                         const Opts& o) {
  double advection = 0, deltaCC_glac = 0, latent_heat = 0, latent_heat_sub = 0, sensible_heat = 0, melt_energy = 0., grnd_flux = 0;
  double GlacMelt = 0, GlacCC = 0;  // the reference leaves both unset when the solve falls back (glacier_melt.c:164-168)
  const double RainFall = rainfall / 1000.;
  (*OldTSurf) = glacier.surf_temp;
  GlacierEB eb;
  eb.stab.reset();
  eb.Dt = delta_t; eb.Ra = aero_resist; eb.Z = z2; eb.Z0_snow = Z0_snow; eb.AirDens = density; eb.EactAir = vp; eb.LongSnowIn = LongIn; eb.Lv = Le;
  eb.Press = pressure; eb.Rain = RainFall; eb.NetShortUnder = NetShort; eb.Vpd = vpd; eb.Wind = wind; eb.OldTSurf = *OldTSurf;
  eb.IceDepth = cp(CP_GLAC_SURF_THICK); eb.Tair = air_temp; eb.TGrnd = Tgrnd; eb.Ra_used = &aero_resist_used;
  eb.AdvectedEnergy = &advection; eb.DeltaColdContent = &deltaCC_glac; eb.GroundFlux = &grnd_flux; eb.LatentHeat = &latent_heat;
  eb.LatentHeatSub = &latent_heat_sub; eb.NetLongUnder = NetLong; eb.SensibleHeat = &sensible_heat; eb.vapor_flux = &glacier.vapor_flux;
  // the balance at 0 C, the solve when that is not zero and the last evaluation (glacier_melt.c:134-176) through one residual call site
  struct Inlined {
    GlacierEB& f;
    VIC_HD double operator()(double x) { return f.eval(x); }
    VIC_HD void before_final() {}
  } call{eb};
  BrentFinal fin;
  fin.allow_fallback = o.TFALLBACK;
  fin.fallback_x = *OldTSurf;
  fin.nosolve_x = 0;
  fin.f_final = 0;
  fin.fell_back = 0;
  double Qnet = 0;
  const double T_guess = glacier.surf_temp;
  auto decide = [&](double f0) {
    Qnet = f0;
    fin.do_solve = (f0 != 0.0);
    fin.do_final = fin.do_solve;
    fin.lo = T_guess - SNOW_DT;
    fin.hi = T_guess + SNOW_DT;
  };
  const double T_solved = root_brent_ss_impl<true, true>(0., 0., call, &fin, true, 0.0, decide);
  if (Qnet == 0.0) {
    // surplus energy at a melting surface: all of it melts ice
    glacier.surf_temp = 0.;
    melt_energy = NetShort + (*NetLong) + sensible_heat + latent_heat + latent_heat_sub + advection - deltaCC_glac;
    GlacMelt = melt_energy / (Lf * RHO_W) * delta_t;
    GlacCC = 0.;
  } else {
    glacier.surf_temp = T_solved;
    if (fin.fell_back) {
      glacier.surf_temp_fbflag = 1;
      glacier.surf_temp_fbcount += 1;
    } else if (result_is_error(glacier.surf_temp)) return ERROR_I;
    {
      Qnet = fin.f_final;
      GlacMelt = 0.0;
      GlacCC = CH_ICE * glacier.surf_temp * cp(CP_GLAC_SURF_THICK) / 1000.;
    }
  }
  melt[0] = GlacMelt;
  glacier.cold_content = GlacCC;
  glacier.vapor_flux *= -1.;
  energy.advection = advection;
  energy.deltaCC_glac = deltaCC_glac;
  energy.glacier_melt_energy = melt_energy;
  energy.grnd_flux = grnd_flux;
  energy.latent = latent_heat;
  energy.latent_sub = latent_heat_sub;
  energy.sensible = sensible_heat;
  energy.error = Qnet;
  return 0;
}

// surface_fluxes_glac.c:6-613.  Returns 0 or ERROR_I.
template <int NN, bool ONE>
VIC_HDI int surface_fluxes_glac(double BareAlbedo, double ice0, double moist0, Hru<NN>& hru, const AeroState& as, const double* gauge_correction,
                                int band, const Ctx& cx, int veg_class, SurfaceFluxOut& out) {
  (void)ice0;
  (void)moist0;
  const Opts& o = *cx.o;
  const CellPar& cp = cx.cp;
  const Forcing& f = cx.f;
  const int NL = VICGPU_NLAYER;
  const double Tfactor = cp.band(CB_Tfactor, band), Pfactor = cp.band(CB_Pfactor, band);
  double coverage = hru.snow.coverage;
  double delta_coverage = 0;  // see the header of this file
  EnergyBal<NN> step_energy = hru.energy;
  SnowPack step_snow = hru.snow;
  Glacier step_glacier = hru.glac;
  // always sub-step at the snow step for snow + glaciers
  int hidx = 0;
  const int endhidx = hidx + o.NF;
  const int step_dt = o.SNOW_STEP;
  double st_AlbedoUnder = 0, st_AtmosLatent = 0, st_AtmosLatentSub = 0, st_AtmosSensible = 0, st_LongUnderIn = 0, st_LongUnderOut = 0,
         st_NetLongAtmos = 0, st_NetLongUnder = 0, st_NetShortAtmos = 0, st_NetShortUnder = 0, st_ShortUnderIn = 0, st_advected_sensible = 0,
         st_advection = 0, st_deltaCC = 0, st_grnd_flux = 0, st_latent = 0, st_latent_sub = 0, st_melt_energy = 0, st_refreeze_energy = 0,
         st_sensible = 0, st_snow_flux = 0, st_deltaCC_glac = 0, st_glacier_flux = 0, st_glacier_melt_energy = 0;
  double st_melt_glac = 0, st_vapor_flux_glac = 0, st_accum_glac = 0;
  double st_melt = 0, st_vapor_flux = 0, st_blowing_flux = 0, st_surface_flux = 0, st_ppt = 0;
  RaUsed st_aero_cond_used = {0, 0};
  double st_pot_evap[N_PET_TYPES];
  for (int p = 0; p < N_PET_TYPES; p++) st_pot_evap[p] = 0;
  double snow_inflow = 0;
  out.out_prec = out.out_rain = out.out_snow = 0;
  int N_steps = 0;
  double latent_heat_Le = 0;
  double LongUnderIn = 0, NetLongSnow = 0, NetShortSnow = 0, ShortUnderIn = 0, OldTSurf = 0;
  const double Tcanopy = 0.;
  const double VPDcanopy = 0.;
  do {
    const double Tair = f(FV_air_temp, hidx) + Tfactor;
    const double step_prec = f(FV_prec, hidx) / hru.mu * Pfactor;
    const double rainOnly = calc_rainonly(Tair, step_prec, cp(CP_MAX_SNOW_TEMP), cp(CP_MIN_RAIN_TEMP), hru.mu, o.TEMP_TH_TYPE);
    double snowfall = gauge_correction[1] * (step_prec - rainOnly) * cp(CP_PADJ_S);
    double rainfall = gauge_correction[0] * rainOnly * cp(CP_PADJ_R);
    const double step_out_prec = snowfall + rainfall, step_out_rain = rainfall, step_out_snow = snowfall;
    const double Tgrnd = GLAC_TEMP;
    // mass flux of blowing snow (surface_fluxes_glac.c:261-275); as in surface_fluxes, not in the three-node kernel
    step_snow.blowing_flux = 0.0;
    if constexpr (NN > 3) {
      if (o.BLOWING && step_snow.swq > 0.) {
        const double Ls = (677. - 0.07 * step_snow.surf_temp) * JOULESPCAL * GRAMSPKG;
        step_snow.blowing_flux = blow::calc_blowing_snow((double)step_dt, Tair, (int)step_snow.last_snow, step_snow.surf_water, as.wind_speed[SNOW_COVERED], Ls,
                                                         f(FV_density, hidx), f(FV_vp, hidx), as.roughness[SNOW_COVERED], step_snow.depth,
                                                         (float)cx.hp(HP_lag_one), (float)cx.hp(HP_sigma_slope), cx.hp(HP_isArtBare) != 0.0,
                                                         (float)cx.hp(HP_fetch), as.displacement[CANOPY_OVER], as.roughness[CANOPY_OVER], &step_snow.transport);
        if ((int)step_snow.blowing_flux == ERROR_I) return ERROR_I;
        step_snow.blowing_flux *= step_dt * SECPHOUR / RHO_W;
      }
    }
    const Surf4& temp_aero_resist = as.aero_resist[N_PET_TYPES];
    RaUsed aero_used;
    aero_used.surface = hru.cell.aero_surface;
    aero_used.overstory = hru.cell.aero_overstory;
    step_snow.canopy_vapor_flux = 0;
    step_snow.vapor_flux = 0;
    step_snow.surface_flux = 0;
    int UnderStory;
    double step_melt, step_melt_glac, step_melt_energy = 0., step_ppt = 0.;
    if (step_snow.swq > 0. || snowfall > 0.) {
      // ---- snow on the glacier: solve_snow_glac.c
      double melt = 0.;
      latent_heat_Le = (2.501e6 - 0.002361e6 * Tair);
      if (hru.mu != 1 && o.FULL_ENERGY) return ERROR_I;
      ShortUnderIn = f(FV_shortwave, hidx);
      LongUnderIn = f(FV_longwave, hidx);
      step_snow.snow = 1.0;
      const double old_coverage = step_snow.coverage;
      step_energy.NetLongOver = 0;
      step_energy.LongOverIn = 0;
      snow_inflow = rainfall + snowfall;
      const double old_swq = step_snow.swq;
      UnderStory = SNOW_COVERED;
      if (step_snow.swq > 0. && snowfall == 0.) {
        step_snow.last_snow += 1;
        step_snow.albedo = snow_albedo(snowfall, step_snow.swq, step_snow.depth, step_snow.albedo, step_snow.coldcontent, (double)step_dt,
                                       (int)step_snow.last_snow, step_snow.MELTING != 0.0, cp, o.SNOW_ALBEDO);
        hru.energy.AlbedoUnder = (coverage * step_snow.albedo + (1. - coverage) * BareAlbedo);
      } else {
        step_snow.last_snow = 0;
        step_snow.albedo = cp(CP_NEW_SNOW_ALB);
        hru.energy.AlbedoUnder = step_snow.albedo;
      }
      NetShortSnow = (1.0 - hru.energy.AlbedoUnder) * ShortUnderIn;
      SnowMeltOut m;
      m.NetLongSnow = NetLongSnow;
      int e = snow_melt(latent_heat_Le, NetShortSnow, /*Tair of the residual*/ Tair, Tgrnd, as.roughness[SNOW_COVERED], temp_aero_resist[UnderStory],
                        aero_used, Tair, (double)step_dt * SECPHOUR, f(FV_density, hidx), 0.0, LongUnderIn, f(FV_pressure, hidx), rainfall, snowfall,
                        f(FV_vp, hidx), f(FV_vpd, hidx), as.wind_speed[UnderStory], as.ref_height[UnderStory], false, step_snow, o, m, true,
                        &step_glacier.accumulation);
      if (e == ERROR_I) return ERROR_I;
      NetLongSnow = m.NetLongSnow;
      OldTSurf = m.OldTSurf;
      melt = m.melt;
      step_energy.error = m.Qnet;
      step_energy.advected_sensible = m.advected_sensible;
      step_energy.advection = m.advection;
      step_energy.deltaCC = m.deltaCC;
      step_energy.grnd_flux = m.grnd_flux;
      step_energy.latent = m.latent;
      step_energy.latent_sub = m.latent_sub;
      step_energy.refreeze_energy = m.refreeze_energy;
      step_energy.sensible = m.sensible;
      step_ppt += melt;
      step_energy.AlbedoUnder = hru.energy.AlbedoUnder;
      if (step_snow.swq > 0.) {
        if (is_valid(step_snow.surf_temp) && step_snow.surf_temp <= 0)
          step_snow.density = snow_density(step_snow, snowfall, old_swq, Tgrnd, Tair, (double)step_dt, o.SNOW_DENSITY);
        else if (step_snow.last_snow == 0) step_snow.density = new_snow_density(Tair, o.SNOW_DENSITY);
        step_snow.depth = 1000. * step_snow.swq / step_snow.density;
        const int diy = cx.dmy.day_in_year;
        const double lat = cp(CP_lat);
        if (step_snow.coldcontent >= 0 && ((lat >= 0 && (diy > 60 && diy < 273)) || (lat < 0 && (diy < 60 || diy > 273)))) step_snow.MELTING = 1.0;
        else if ((step_snow.MELTING != 0.0) && snowfall > TraceSnow) step_snow.MELTING = 0.0;
        if (step_snow.swq > 0) step_snow.coverage = 1.;
        else step_snow.coverage = 0.;
      } else step_snow.coverage = 0.;
      delta_coverage = old_coverage - step_snow.coverage;
      if (delta_coverage != 0) {
        if (old_coverage > step_snow.coverage) {
          coverage = (old_coverage);
          hru.energy.AlbedoUnder = (coverage - step_snow.coverage) / (1. - step_snow.coverage) * step_snow.albedo;
          hru.energy.AlbedoUnder += (1. - coverage) / (1. - step_snow.coverage) * BareAlbedo;
          step_melt_energy = (delta_coverage) * (step_energy.advection - step_energy.deltaCC + step_energy.latent + step_energy.latent_sub +
                                                 step_energy.sensible + step_energy.refreeze_energy + step_energy.advected_sensible);
        } else {
          coverage = step_snow.coverage;
          delta_coverage = 0;
        }
      } else if (old_coverage == 0 && step_snow.coverage == 0) {
        delta_coverage = 1.;
        coverage = 0.;
        step_melt_energy = (step_energy.advection - step_energy.deltaCC + step_energy.latent + step_energy.latent_sub + step_energy.sensible +
                            step_energy.refreeze_energy + step_energy.advected_sensible);
      }
      NetLongSnow *= (step_snow.coverage + delta_coverage);
      NetShortSnow *= (step_snow.coverage + delta_coverage);
      step_energy.latent *= (step_snow.coverage + delta_coverage);
      step_energy.latent_sub *= (step_snow.coverage + delta_coverage);
      step_energy.sensible *= (step_snow.coverage + delta_coverage);
      if (step_snow.swq == 0) {
        step_snow.density = 0.;
        step_snow.depth = 0.;
        step_snow.surf_water = 0;
        step_snow.pack_water = 0;
        step_snow.surf_temp = 0;
        step_snow.pack_temp = 0;
        step_snow.coverage = 0;
        step_snow.swq_slope = 0;
        step_snow.store_snow = 1.0;
        step_snow.MELTING = 0.0;
      }
      snowfall = 0;
      rainfall = 0;
      step_energy.melt_energy *= -1.;
      step_melt = melt;
      // back in surface_fluxes_glac.c:301-307
      step_melt_glac = 0.;
      step_glacier.vapor_flux = 0.;
      step_energy.glacier_flux = 0.;
      step_energy.deltaCC_glac = 0.;
      step_energy.glacier_melt_energy = 0.;
      step_energy.snow_flux = -step_energy.grnd_flux;
      step_energy.LongUnderOut = LongUnderIn - NetLongSnow;
    } else {
      // ---- bare ice: solve_glacier.c
      double melt = 0.;
      latent_heat_Le = (2.501e6 - 0.002361e6 * Tair);
      ShortUnderIn = f(FV_shortwave, hidx);
      LongUnderIn = f(FV_longwave, hidx);
      hru.energy.AlbedoUnder = BareAlbedo;
      NetShortSnow = (1.0 - hru.energy.AlbedoUnder) * ShortUnderIn;
      UnderStory = GLACIER_SURF;
      int e = glacier_melt<NN>(latent_heat_Le, NetShortSnow, Tgrnd, as.roughness[SNOW_COVERED], temp_aero_resist[UnderStory], aero_used, Tair,
                               (double)step_dt * SECPHOUR, f(FV_density, hidx), LongUnderIn, f(FV_pressure, hidx), rainfall, f(FV_vp, hidx),
                               f(FV_vpd, hidx), as.wind_speed[UnderStory], as.ref_height[UnderStory], &NetLongSnow, &OldTSurf, &melt, step_energy,
                               step_glacier, cp, o);
      if (e == ERROR_I) return ERROR_I;
      step_ppt = (melt + rainfall / 1000.);
      step_energy.AlbedoUnder = hru.energy.AlbedoUnder;
      rainfall = 0;
      step_melt_glac = melt;
      step_melt = 0.;
      // surface_fluxes_glac.c:324-328 writes these straight into the HRU's own snow record, which is
      // replaced by the working copy after the loop (:467) -- so they never survive; kept for the record.
      step_energy.deltaCC = 0.;
      step_energy.refreeze_energy = 0.;
      step_energy.snow_flux = 0.;
      step_energy.advected_sensible = 0.;
      step_energy.glacier_flux = -step_energy.grnd_flux;
      step_energy.LongUnderOut = LongUnderIn - NetLongSnow;
      step_glacier.accumulation = 0.;
    }
    cx.rendezvous(4, 1);
    step_energy.AtmosLatent = step_energy.latent;
    step_energy.AtmosLatentSub = step_energy.latent_sub;
    step_energy.AtmosSensible = step_energy.sensible;
    step_energy.NetLongAtmos = step_energy.NetLongUnder;
    step_energy.NetShortAtmos = step_energy.NetShortUnder;
    double stability_factor[2];
    if (aero_used.surface == HUGE_RESIST) stability_factor[0] = HUGE_RESIST;
    else stability_factor[0] = aero_used.surface / as.aero_resist[N_PET_TYPES][UnderStory];
    if (aero_used.overstory == aero_used.surface) stability_factor[1] = stability_factor[0];
    else {
      if (aero_used.overstory == HUGE_RESIST) stability_factor[1] = HUGE_RESIST;
      else stability_factor[1] = aero_used.overstory / as.aero_resist[N_PET_TYPES][CANOPY_OVER];
    }
    RaUsed step_aero[N_PET_TYPES];
    for (int p = 0; p < N_PET_TYPES; p++) {
      if (stability_factor[0] == HUGE_RESIST) step_aero[p].surface = HUGE_RESIST;
      else step_aero[p].surface = as.aero_resist[p][UnderStory] * stability_factor[0];
      if (stability_factor[1] == HUGE_RESIST) step_aero[p].overstory = HUGE_RESIST;
      else step_aero[p].overstory = as.aero_resist[p][CANOPY_OVER] * stability_factor[1];
    }
    double step_pot_evap[N_PET_TYPES];
    compute_pot_evap(cx.vl, o.NVegLibTypes, veg_class, cx.dmy.month - 1, o.dt, f(FV_shortwave, hidx), step_energy.NetLongAtmos, Tair, VPDcanopy,
                     cp(CP_elevation), step_aero, step_pot_evap);
    // ---- accumulate
    st_ppt += step_ppt;
    if (aero_used.surface > 0) st_aero_cond_used.surface += 1 / aero_used.surface;
    else st_aero_cond_used.surface += HUGE_RESIST;
    if (aero_used.overstory > 0) st_aero_cond_used.overstory += 1 / aero_used.overstory;
    else st_aero_cond_used.overstory += HUGE_RESIST;
    st_melt += step_melt;
    st_vapor_flux += step_snow.vapor_flux;
    st_surface_flux += step_snow.surface_flux;
    st_blowing_flux += step_snow.blowing_flux;
    out.out_prec += step_out_prec * hru.mu;
    out.out_rain += step_out_rain * hru.mu;
    out.out_snow += step_out_snow * hru.mu;
    st_AlbedoUnder += step_energy.AlbedoUnder;
    st_AtmosLatent += step_energy.AtmosLatent;
    st_AtmosLatentSub += step_energy.AtmosLatentSub;
    st_AtmosSensible += step_energy.AtmosSensible;
    st_LongUnderIn += LongUnderIn;
    st_LongUnderOut += step_energy.LongUnderOut;
    st_NetLongAtmos += NetLongSnow;
    st_NetLongUnder += NetLongSnow;
    st_NetShortAtmos += NetShortSnow;
    st_NetShortUnder += NetShortSnow;
    st_ShortUnderIn += ShortUnderIn;
    st_latent += step_energy.latent;
    st_latent_sub += step_energy.latent_sub;
    st_melt_energy += step_melt_energy;
    st_sensible += step_energy.sensible;
    st_grnd_flux += step_energy.grnd_flux;
    st_melt_glac += step_melt_glac;
    st_vapor_flux_glac += step_glacier.vapor_flux;
    st_accum_glac += step_glacier.accumulation;
    st_glacier_flux += step_energy.glacier_flux;
    st_deltaCC_glac += step_energy.deltaCC_glac;
    st_glacier_melt_energy += step_energy.glacier_melt_energy;
    const double cov = (step_snow.coverage + delta_coverage);
    st_advected_sensible += step_energy.advected_sensible * cov;
    st_advection += step_energy.advection * cov;
    st_deltaCC += step_energy.deltaCC * cov;
    st_snow_flux += step_energy.snow_flux * cov;
    st_refreeze_energy += step_energy.refreeze_energy * cov;
    for (int p = 0; p < N_PET_TYPES; p++) st_pot_evap[p] += step_pot_evap[p];
    N_steps++;
    hidx += 1;
  } while (!ONE && hidx < endhidx);

  const double N = ONE ? 1.0 : (double)N_steps;
  auto mean = [&](double x) { return ONE ? x : div_pos(x, N); };
  hru.glac = step_glacier;
  hru.glac.melt = st_melt_glac;
  hru.glac.vapor_flux = st_vapor_flux_glac;
  hru.glac.accumulation = st_accum_glac;
  const double wdew_in = hru.veg.Wdew;
  hru.snow = step_snow;
  hru.snow.vapor_flux = st_vapor_flux;
  hru.snow.blowing_flux = st_blowing_flux;
  hru.snow.surface_flux = st_surface_flux;
  hru.snow.canopy_vapor_flux = 0;
  out.Melt = st_melt + st_melt_glac;
  hru.snow.melt = st_melt;
  double ppt = st_ppt;
  // glacier mass balance [m w.e.]: precipitation - melt water leaving - sublimation
  hru.glac.mass_balance = out.out_prec / 1000. - ppt - hru.snow.vapor_flux - hru.glac.vapor_flux;
  hru.glac.ice_mass_balance = hru.glac.accumulation - hru.glac.melt - hru.glac.vapor_flux;
  hru.energy = step_energy;
  hru.energy.AlbedoOver = 0 / N;
  hru.energy.AlbedoUnder = mean(st_AlbedoUnder);
  hru.energy.AtmosLatent = mean(st_AtmosLatent);
  hru.energy.AtmosLatentSub = mean(st_AtmosLatentSub);
  hru.energy.AtmosSensible = mean(st_AtmosSensible);
  hru.energy.LongOverIn = 0 / N;
  hru.energy.LongUnderIn = mean(st_LongUnderIn);
  hru.energy.LongUnderOut = mean(st_LongUnderOut);
  hru.energy.NetLongAtmos = mean(st_NetLongAtmos);
  hru.energy.NetLongOver = 0 / N;
  hru.energy.NetLongUnder = mean(st_NetLongUnder);
  hru.energy.NetShortAtmos = mean(st_NetShortAtmos);
  hru.energy.NetShortGrnd = 0 / N;
  hru.energy.NetShortOver = 0 / N;
  hru.energy.NetShortUnder = mean(st_NetShortUnder);
  hru.energy.ShortOverIn = 0 / N;
  hru.energy.ShortUnderIn = mean(st_ShortUnderIn);
  hru.energy.advected_sensible = mean(st_advected_sensible);
  hru.energy.canopy_advection = 0 / N;
  hru.energy.canopy_latent = 0 / N;
  hru.energy.canopy_latent_sub = 0 / N;
  hru.energy.canopy_refreeze = 0 / N;
  hru.energy.canopy_sensible = 0 / N;
  hru.energy.deltaH = 0 / N;
  hru.energy.fusion = 0 / N;
  hru.energy.grnd_flux = mean(st_grnd_flux);
  hru.energy.latent = mean(st_latent);
  hru.energy.latent_sub = mean(st_latent_sub);
  hru.energy.melt_energy = mean(st_melt_energy);
  hru.energy.sensible = mean(st_sensible);
  hru.energy.glacier_flux = mean(st_glacier_flux);
  hru.energy.deltaCC_glac = mean(st_deltaCC_glac);
  hru.energy.glacier_melt_energy = mean(st_glacier_melt_energy);
  hru.energy.advection = mean(st_advection);
  hru.energy.deltaCC = mean(st_deltaCC);
  hru.energy.refreeze_energy = mean(st_refreeze_energy);
  hru.energy.snow_flux = mean(st_snow_flux);
  hru.energy.Tcanopy = Tcanopy;
  // the canopy stores of a glacier tile are never touched by the sub-steps
  hru.veg.throughfall = 0.;
  hru.veg.canopyevap = 0.;
  hru.veg.Wdew = wdew_in;
  for (int l = 0; l < NL; l++) hru.cell.layer[l].evap = 0.;
  if (st_aero_cond_used.surface > 0 && st_aero_cond_used.surface < HUGE_RESIST) hru.cell.aero_surface = 1 / (ONE ? st_aero_cond_used.surface : st_aero_cond_used.surface / N);
  else if (st_aero_cond_used.surface >= HUGE_RESIST) hru.cell.aero_surface = 0;
  else hru.cell.aero_surface = HUGE_RESIST;
  if (st_aero_cond_used.overstory > 0 && st_aero_cond_used.overstory < HUGE_RESIST) hru.cell.aero_overstory = 1 / (ONE ? st_aero_cond_used.overstory : st_aero_cond_used.overstory / N);
  else if (st_aero_cond_used.overstory >= HUGE_RESIST) hru.cell.aero_overstory = 0;
  else hru.cell.aero_overstory = HUGE_RESIST;
  for (int p = 0; p < N_PET_TYPES; p++) hru.cell.pot_evap[p] = ONE ? st_pot_evap[p] : st_pot_evap[p] / N;
  out.snow_inflow = snow_inflow;
  // glacier water storage: linear reservoir whose coefficient decays with snow depth on the ice
  hru.glac.inflow = ppt + 0.;
  ppt = hru.cell.excess_moist;
  hru.cell.excess_moist = 0.;
  hru.glac.outflow_coef = cp(CP_GLAC_KMIN) + cp(CP_GLAC_DK) * vexp(-cp(CP_GLAC_A) * hru.snow.swq);
  hru.glac.water_storage += hru.glac.inflow;
  hru.glac.outflow = hru.glac.outflow_coef * hru.glac.water_storage;
  hru.glac.water_storage -= hru.glac.outflow;
  hru.cell.inflow = ppt;
  cx.rendezvous(5, 1);
  int e = runoff<NN>(hru.cell, hru.energy, cp, ppt, o);
  hru.cell.runoff += (hru.glac.outflow * 1000.);
  return e;
}

}  // namespace vic
#endif

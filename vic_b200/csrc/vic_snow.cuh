// vic_snow.cuh -- snow on the ground and in the canopy for one HRU sub-step:
//   snow-pack surface energy balance residual   SnowPackEnergyBalance.c:85-197
//   two-layer pack mass / energy update          snow_melt.c:119-564
//   canopy energy balance residual               func_canopy_energy_bal.c:9-149
//   interception, unloading, drip                snow_intercept.c:81-582, massrelease.c:38-93
//   driver for one sub-step                      solve_snow.c:7-544
// Residuals are functors evaluated by vic::root_brent; they write their flux terms through
// references, and the values left by the LAST evaluation are the result (as in the reference).
#ifndef VIC_SNOW_CUH
#define VIC_SNOW_CUH
#include "vic_brent.cuh"
#include "vic_evap.cuh"

namespace vic {

// ---- snow pack surface energy balance ----------------------------------------------------
struct SnowPackEB {
  // inputs (by value, as captured by the reference's functor constructor)
  double Dt, Ra, Z, Z0_snow, AirDens, EactAir, LongSnowIn, Lv, Press, Rain, NetShortUnder, Vpd, Wind, OldTSurf;
  double SnowDepth, SnowDensity, SurfaceLiquidWater, SweSurfaceLayer, Tair, TGrnd;
  // outputs
  // (values, not pointers to the caller's variables: an evaluation then reads and writes the functor only; snow_melt binds
  // references to these members)
  RaUsed* Ra_used;
  SnowPack* sn;  // vapor_flux, blowing_flux, surface_flux
  double* NetLongUnder;
  double AdvectedEnergy, AdvectedSensibleHeat, DeltaColdContent, GroundFlux, LatentHeat, LatentHeatSub, RefreezeEnergy, SensibleHeat;
  StabLog stab;

  VIC_HDI double operator()(double TSurf) { return eval(TSurf); }
  // the residual proper, inlined into the one call site of snow_melt's solve
  VIC_HD double eval(double TSurf) {
    const double TMean = TSurf;
    const double Density = RHO_W;
    if (Wind > 0.0) Ra_used->surface = Ra / stab.correction(Z, 0., TMean, Tair, Wind, Z0_snow);
    else Ra_used->surface = HUGE_RESIST;
    const double Tmp = TMean + KELVIN;
    *NetLongUnder = LongSnowIn - STEFAN_B * Tmp * Tmp * Tmp * Tmp;
    const double NetRad = NetShortUnder + (*NetLongUnder);
    SensibleHeat = AirDens * Cp * (Tair - TMean) / Ra_used->surface;
    AdvectedSensibleHeat = 0;
    double VaporMassFlux = div_pos(sn->vapor_flux * Density, Dt);
    double BlowingMassFlux = div_pos(sn->blowing_flux * Density, Dt);
    double SurfaceMassFlux = div_pos(sn->surface_flux * Density, Dt);
    latent_heat_from_snow(AirDens, EactAir, Lv, Press, Ra_used->surface, TMean, Vpd, &LatentHeat, &LatentHeatSub, &VaporMassFlux,
                          &BlowingMassFlux, &SurfaceMassFlux);
    sn->vapor_flux = div_pos(VaporMassFlux * Dt, Density);
    sn->blowing_flux = div_pos(BlowingMassFlux * Dt, Density);
    sn->surface_flux = div_pos(SurfaceMassFlux * Dt, Density);
    if (TMean == 0) AdvectedEnergy = div_pos((CH_WATER * (Tair)*Rain), (Dt));
    else AdvectedEnergy = 0.;
    DeltaColdContent = div_pos(CH_ICE * SweSurfaceLayer * (TSurf - OldTSurf), (Dt));
    if (SnowDepth > 0.) GroundFlux = K_SNOW * SnowDensity * SnowDensity * (TGrnd - TMean) / SnowDepth / (Dt);
    else GroundFlux = 0;
    double RestTerm = NetRad + SensibleHeat + LatentHeat + LatentHeatSub + AdvectedEnergy + AdvectedSensibleHeat - DeltaColdContent + GroundFlux;
    RefreezeEnergy = div_pos((SurfaceLiquidWater * Lf * Density), (Dt));
    if (TSurf == 0.0 && RestTerm > -RefreezeEnergy) {
      RefreezeEnergy = -RestTerm;
      RestTerm = 0.0;
    } else RestTerm += RefreezeEnergy;
    return RestTerm;
  }
};

struct SnowMeltOut {
  double NetLongSnow, OldTSurf, melt, Qnet, advected_sensible, advection, deltaCC, grnd_flux, latent, latent_sub, refreeze_energy, sensible;
};

// snow_melt.c:119-564.  Returns 0, or ERROR_I when the surface solve fails and TFALLBACK is off.
VIC_HDI int snow_melt(double latent_heat_Le, double NetShortSnow, double Tcanopy, double Tgrnd, double Z0_snow, double aero_resist,
                      RaUsed& aero_resist_used, double air_temp, double delta_t, double density, double grnd_flux_in, double LongSnowIn,
                      double pressure, double rainfall, double snowfall, double vp, double vpd, double wind, double z2, bool UNSTABLE_SNOW,
                      SnowPack& snow, const Opts& o, SnowMeltOut& out, bool GLAC = false, double* firn_to_ice = nullptr) {
  double DeltaPackCC, DeltaPackSwq, SnowMelt = 0, RefrozenWater;
  SnowPackEB eb;
  double &advection = eb.AdvectedEnergy, &deltaCC = eb.DeltaColdContent, &latent_heat = eb.LatentHeat, &latent_heat_sub = eb.LatentHeatSub,
         &sensible_heat = eb.SensibleHeat, &advected_sensible_heat = eb.AdvectedSensibleHeat, &RefreezeEnergy = eb.RefreezeEnergy,
         &grnd_flux = eb.GroundFlux;
  advection = 0; deltaCC = 0; latent_heat = 0; latent_heat_sub = 0; sensible_heat = 0; advected_sensible_heat = 0; RefreezeEnergy = 0;
  grnd_flux = grnd_flux_in;
  double melt_energy = 0.;
  const double SnowFall = snowfall / 1000.;
  const double RainFall = rainfall / 1000.;
  const double InitialSwq = snow.swq;
  out.OldTSurf = snow.surf_temp;
  double Ice = snow.swq - snow.pack_water - snow.surf_water;
  double SurfaceSwq = (Ice > MAX_SURFACE_SWE) ? MAX_SURFACE_SWE : Ice;
  double PackSwq = Ice - SurfaceSwq;
  double SurfaceCC = CH_ICE * SurfaceSwq * snow.surf_temp;
  double PackCC = CH_ICE * PackSwq * snow.pack_temp;
  const double SnowFallCC = (air_temp > 0.0) ? 0.0 : CH_ICE * SnowFall * air_temp;
  if (SnowFall > (MAX_SURFACE_SWE - SurfaceSwq) && (MAX_SURFACE_SWE - SurfaceSwq) > SMALL) {
    DeltaPackSwq = SurfaceSwq + SnowFall - MAX_SURFACE_SWE;
    if (DeltaPackSwq > SurfaceSwq) DeltaPackCC = SurfaceCC + (SnowFall - MAX_SURFACE_SWE) / SnowFall * SnowFallCC;
    else DeltaPackCC = DeltaPackSwq / SurfaceSwq * SurfaceCC;
    SurfaceSwq = MAX_SURFACE_SWE;
    SurfaceCC += SnowFallCC - DeltaPackCC;
    PackSwq += DeltaPackSwq;
    PackCC += DeltaPackCC;
  } else {
    SurfaceSwq += SnowFall;
    SurfaceCC += SnowFallCC;
  }
  snow.surf_temp = (SurfaceSwq > 0.0) ? SurfaceCC / (CH_ICE * SurfaceSwq) : 0.0;
  if (!GLAC) {
    snow.pack_temp = (PackSwq > 0.0) ? PackCC / (CH_ICE * PackSwq) : 0.0;
  } else {
    // snow on glacier ice (snow_melt_glac.c:110-132): the part of the pack denser than the firn/ice
    // cut-off becomes glacier ice.  When the whole pack converts, pack_temp is 0/0 for a moment, as in
    // the reference (it is reassigned below before anything reads it).
    double FirnToIce = 0.;
    if (PackSwq > 0.0) {
      if (snow.density > SNOW_SURF_DENSITY) {
        const double zco = (CUTOFF_DENSITY - SNOW_SURF_DENSITY) * (snow.depth / 2) / (snow.density - SNOW_SURF_DENSITY);
        if (zco < snow.depth) {
          const double density_zsnow = SNOW_SURF_DENSITY + 2 * (snow.density - SNOW_SURF_DENSITY);
          FirnToIce = (density_zsnow + CUTOFF_DENSITY) / (2 * RHO_W) * (snow.depth - zco);
          if (FirnToIce >= PackSwq) {
            FirnToIce = PackSwq;
            PackSwq = 0.0;
            snow.pack_temp = 0.0;
            PackCC = 0.0;
          } else PackSwq -= FirnToIce;
        }
      }
      snow.pack_temp = PackCC / (CH_ICE * PackSwq);
    } else snow.pack_temp = 0.0;
    *firn_to_ice = FirnToIce;
  }
  Ice += SnowFall;
  snow.surf_water += RainFall;

  eb.stab.reset();
  eb.Dt = delta_t; eb.Ra = aero_resist; eb.Z = z2; eb.Z0_snow = Z0_snow; eb.AirDens = density; eb.EactAir = vp;
  eb.LongSnowIn = LongSnowIn; eb.Lv = latent_heat_Le; eb.Press = pressure; eb.Rain = RainFall; eb.NetShortUnder = NetShortSnow;
  eb.Vpd = vpd; eb.Wind = wind; eb.OldTSurf = out.OldTSurf; eb.SnowDepth = snow.depth; eb.SnowDensity = snow.density;
  eb.SurfaceLiquidWater = snow.surf_water; eb.SweSurfaceLayer = SurfaceSwq; eb.Tair = Tcanopy; eb.TGrnd = Tgrnd;
  eb.Ra_used = &aero_resist_used;
  eb.NetLongUnder = &out.NetLongSnow;
  eb.sn = &snow;

  // The balance at 0 C, the solve it may call for and the last evaluation at the accepted temperature (snow_melt.c:284-379) all go
  // through the one residual call site of root_brent_ss_impl, so the residual is inlined here.
  struct Inlined {
    SnowPackEB& f;
    VIC_HD double operator()(double x) { return f.eval(x); }
    VIC_HD void before_final() {}
  } call{eb};
  BrentFinal fin;
  fin.do_solve = false;
  fin.do_final = false;
  fin.final_needs_valid = true;
  fin.allow_fallback = o.TFALLBACK;
  fin.fallback_x = out.OldTSurf;
  fin.nosolve_x = 0;
  fin.f_final = 0;
  fin.fell_back = 0;
  double Qnet = 0;
  const bool thick = GLAC || SurfaceSwq > MIN_SWQ_EB_THRES;
  const double T_guess = snow.surf_temp;
  auto decide = [&](double f0) {
    Qnet = f0;
    if (!UNSTABLE_SNOW && f0 != 0.0 && thick) {
      fin.do_solve = true;
      fin.do_final = true;
      fin.lo = T_guess - SNOW_DT;
      fin.hi = T_guess + SNOW_DT;
    }
  };
  const double T_solved = root_brent_ss_impl<true, true>(0., 0., call, &fin, true, 0.0, decide);
  if (!UNSTABLE_SNOW) {
    if (Qnet == 0.0) {
      // pack is melting or isothermal at 0 C
      snow.surf_temp = 0.0;
      if (RefreezeEnergy >= 0.0) {
        RefrozenWater = RefreezeEnergy / (Lf * RHO_W) * delta_t;
        if (RefrozenWater > snow.surf_water) {
          RefrozenWater = snow.surf_water;
          RefreezeEnergy = RefrozenWater * Lf * RHO_W / (delta_t);
        }
        melt_energy += RefreezeEnergy;
        SurfaceSwq += RefrozenWater;
        Ice += RefrozenWater;
        snow.surf_water -= RefrozenWater;
        if (snow.surf_water < 0.0) snow.surf_water = 0.0;
        SnowMelt = 0.0;
      } else {
        SnowMelt = fabs(RefreezeEnergy) / (Lf * RHO_W) * delta_t;
        melt_energy += RefreezeEnergy;
      }
      if (snow.surf_water < -(snow.vapor_flux)) {
        snow.blowing_flux *= -(snow.surf_water / snow.vapor_flux);
        snow.vapor_flux = -(snow.surf_water);
        snow.surface_flux = -(snow.surf_water) - snow.blowing_flux;
        snow.surf_water = 0.0;
      } else snow.surf_water += snow.vapor_flux;
      if (SnowMelt < Ice) {
        if (SnowMelt <= PackSwq) {
          snow.surf_water += SnowMelt;
          PackSwq -= SnowMelt;
          Ice -= SnowMelt;
        } else {
          snow.surf_water += SnowMelt + snow.pack_water;
          snow.pack_water = 0.0;
          PackSwq = 0.0;
          Ice -= SnowMelt;
          SurfaceSwq = Ice;
        }
      } else {
        SnowMelt = Ice;
        snow.surf_water += Ice;
        SurfaceSwq = 0.0;
        snow.surf_temp = 0.0;
        PackSwq = 0.0;
        snow.pack_temp = 0.0;
        Ice = 0.0;
        melt_energy -= RefreezeEnergy;
        RefreezeEnergy = RefreezeEnergy / fabs(RefreezeEnergy) * SnowMelt * Lf * RHO_W / (delta_t);
        melt_energy += RefreezeEnergy;
      }
    } else {
      // pack surface below freezing: solve for its temperature
      if (thick) {
        snow.surf_temp = T_solved;
        if (fin.fell_back) {
          snow.surf_temp_fbflag = 1;
          snow.surf_temp_fbcount += 1;
        } else if (result_is_error(snow.surf_temp)) return ERROR_I;
      } else {
        snow.surf_temp = vnan();  // thin pack: solved together with the ground surface
      }
      if (is_valid(snow.surf_temp) && !result_is_error(snow.surf_temp)) {
        Qnet = fin.f_final;
        SnowMelt = 0.0;
        SurfaceSwq += snow.surf_water;
        Ice += snow.surf_water;
        snow.surf_water = 0.0;
        melt_energy += snow.surf_water * Lf * RHO_W / (delta_t);
        if (SurfaceSwq < -(snow.vapor_flux)) {
          snow.blowing_flux *= -(SurfaceSwq / snow.vapor_flux);
          snow.vapor_flux = -SurfaceSwq;
          snow.surface_flux = -SurfaceSwq - snow.blowing_flux;
          SurfaceSwq = 0.0;
          Ice = PackSwq;
        } else {
          SurfaceSwq += snow.vapor_flux;
          Ice += snow.vapor_flux;
        }
      }
    }
  } else {
    snow.surf_temp = vnan();
  }
  (void)SnowMelt;
  (void)melt_energy;
  // liquid water in the surface layer
  double MaxLiquidWater = LIQUID_WATER_CAPACITY * SurfaceSwq;
  double melt;
  if (snow.surf_water > MaxLiquidWater) {
    melt = snow.surf_water - MaxLiquidWater;
    snow.surf_water = MaxLiquidWater;
  } else melt = 0.0;
  // refreeze / drain in the pack layer
  snow.pack_water += melt;
  const double PackRefreezeEnergy = snow.pack_water * Lf * RHO_W;
  if (PackCC < -PackRefreezeEnergy) {
    PackSwq += snow.pack_water;
    Ice += snow.pack_water;
    snow.pack_water = 0.0;
    if (PackSwq > 0.0) {
      PackCC = PackSwq * CH_ICE * snow.pack_temp + PackRefreezeEnergy;
      snow.pack_temp = PackCC / (CH_ICE * PackSwq);
      if (snow.pack_temp > 0.) snow.pack_temp = 0.;
    } else snow.pack_temp = 0.0;
  } else {
    snow.pack_temp = 0.0;
    DeltaPackSwq = -PackCC / (Lf * RHO_W);
    snow.pack_water -= DeltaPackSwq;
    PackSwq += DeltaPackSwq;
    Ice += DeltaPackSwq;
  }
  MaxLiquidWater = LIQUID_WATER_CAPACITY * PackSwq;
  if (snow.pack_water > MaxLiquidWater) {
    melt = snow.pack_water - MaxLiquidWater;
    snow.pack_water = MaxLiquidWater;
  } else melt = 0.0;
  // re-partition ice between the two layers
  Ice = PackSwq + SurfaceSwq;
  if (Ice > MAX_SURFACE_SWE) {
    SurfaceCC = CH_ICE * snow.surf_temp * SurfaceSwq;
    PackCC = CH_ICE * snow.pack_temp * PackSwq;
    if (SurfaceSwq > MAX_SURFACE_SWE) {
      PackCC += SurfaceCC * (SurfaceSwq - MAX_SURFACE_SWE) / SurfaceSwq;
      SurfaceCC -= SurfaceCC * (SurfaceSwq - MAX_SURFACE_SWE) / SurfaceSwq;
      PackSwq += SurfaceSwq - MAX_SURFACE_SWE;
      SurfaceSwq -= SurfaceSwq - MAX_SURFACE_SWE;
    } else if (SurfaceSwq < MAX_SURFACE_SWE) {
      PackCC -= PackCC * (MAX_SURFACE_SWE - SurfaceSwq) / PackSwq;
      SurfaceCC += PackCC * (MAX_SURFACE_SWE - SurfaceSwq) / PackSwq;
      PackSwq -= MAX_SURFACE_SWE - SurfaceSwq;
      SurfaceSwq += MAX_SURFACE_SWE - SurfaceSwq;
    }
    snow.pack_temp = PackCC / (CH_ICE * PackSwq);
    snow.surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  } else {
    PackSwq = 0.0;
    PackCC = 0.0;
    snow.pack_temp = 0.0;
  }
  snow.swq = Ice + snow.pack_water + snow.surf_water;
  if (snow.swq == 0.0) {
    snow.surf_temp = 0.0;
    snow.pack_temp = 0.0;
  }
  const double MassBalanceError = (InitialSwq - snow.swq) + (RainFall + SnowFall) - melt + snow.vapor_flux;
  if (!GLAC) melt *= 1000.;  // snow_melt_glac.c:391 leaves the melt in metres
  snow.mass_error = MassBalanceError;
  snow.coldcontent = SurfaceCC;
  snow.vapor_flux *= -1.;
  out.melt = melt;
  out.advection = advection;
  out.deltaCC = deltaCC;
  out.grnd_flux = grnd_flux;
  out.latent = latent_heat;
  out.latent_sub = latent_heat_sub;
  out.sensible = sensible_heat;
  out.advected_sensible = advected_sensible_heat;
  out.refreeze_energy = RefreezeEnergy;
  out.Qnet = Qnet;
  return 0;
}

// ---- canopy (intercepted snow) energy balance --------------------------------------------
struct CanopyEB {
  // inputs
  double delta_t, elevation, AirDens, EactAir, Press, latent_heat_Le, Tcanopy, Vpd, IntRain, IntSnow, LongOverIn, LongUnderOut, NetShortOver;
  int AERO_RESIST_CANSNOW;
  double ra_free, ra_over, ws_over, zref_over, disp_over, rough_over;  // the entries of the aerodynamic tables the residual reads
  const VegNow* veg;
  const SoilET* soil;
  // in/out
  RaUsed* Ra_used;
  double* Rainfall;  // [m]
  double* Wdew;      // interception store, [m] between evaluations
  SoilLayer* layer;
  VegVar* vv;
  double *Evap, *AdvectedEnergy, *LatentHeat, *LatentHeatSub, *LongOverOut, *NetLongOver, *NetRadiation, *RefreezeEnergy, *SensibleHeat, *VaporMassFlux;
  StabLog stab;
  EvapMemo memo;

  VIC_HDI double operator()(double Tfoliage) { return eval(Tfoliage); }
  // the residual proper, inlined into the one call site of snow_intercept's solve
  VIC_HD double eval(double Tfoliage) {
    const double Tmp = Tfoliage + KELVIN;
    *LongOverOut = STEFAN_B * (Tmp * Tmp * Tmp * Tmp);
    *NetRadiation = NetShortOver + LongOverIn + LongUnderOut - 2 * (*LongOverOut);
    *NetLongOver = LongOverIn - (*LongOverOut);
    const int ar = AERO_RESIST_CANSNOW;
    if (IntSnow > 0) {
      Ra_used->surface = ra_free;
      Ra_used->overstory = ra_over;
      if (ar == AR_COMBO || ar == AR_406 || ar == AR_406_LS || ar == AR_406_FULL) Ra_used->overstory *= 10.;
      const double EsSnow = svp(Tfoliage);
      if (ar == AR_COMBO || ar == AR_410) {
        if (ws_over > 0.0)
          Ra_used->overstory /= stab.correction(zref_over, disp_over, Tfoliage, Tcanopy,
                                                     ws_over, rough_over);
        else Ra_used->overstory = HUGE_RESIST;
      }
      *VaporMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra_used->overstory / RHO_W;
      if (Vpd == 0.0 && *VaporMassFlux < 0.0) *VaporMassFlux = 0.0;
      const double Ls = (677. - 0.07 * Tfoliage) * JOULESPCAL * GRAMSPKG;
      *LatentHeatSub = Ls * *VaporMassFlux * RHO_W;
      *LatentHeat = 0;
      *Evap = 0;
      vv->throughfall = 0;
      if (ar == AR_406) Ra_used->overstory /= 10;
    } else {
      if (ar == AR_406_FULL || ar == AR_410 || ar == AR_COMBO) {
        Ra_used->surface = ra_free;
        Ra_used->overstory = ra_over;
      } else {
        Ra_used->surface = ra_free;
        Ra_used->overstory = ra_free;
      }
      *Wdew = IntRain * 1000.;
      const double prec = *Rainfall * 1000;
      // canopy_evap leaves the new store in vv->Wdew; in the reference vv->Wdew and *Wdew are the
      // same object (snow_intercept is handed &veg_var_wet->Wdew), so the write-back below is the
      // division the reference applies to that object.
      *Evap = canopy_evap(layer, *vv, false, *veg, *Wdew, delta_t, *NetRadiation, Vpd, NetShortOver, Tcanopy, Ra_used->overstory, elevation,
                          prec, *soil, &memo);
      *Wdew = div_pos(vv->Wdew, 1000.);
      vv->Wdew = *Wdew;
      *LatentHeat = latent_heat_Le * *Evap * RHO_W;
      *LatentHeatSub = 0;
    }
    *SensibleHeat = AirDens * Cp * (Tcanopy - Tfoliage) / Ra_used->overstory;
    *AdvectedEnergy = div_pos((4186.8 * Tcanopy * Rainfall[0]), (delta_t));
    double RestTerm = *SensibleHeat + *LatentHeat + *LatentHeatSub + *NetRadiation + *AdvectedEnergy;
    if (IntSnow > 0) {
      *RefreezeEnergy = div_pos((IntRain * Lf * RHO_W), (delta_t));
      if (Tfoliage == 0.0 && RestTerm > -(*RefreezeEnergy)) {
        *RefreezeEnergy = -RestTerm;
        RestTerm = 0.0;
      } else RestTerm += *RefreezeEnergy;
    } else *RefreezeEnergy = 0;
    return RestTerm;
  }
};

// massrelease.c:38-93 (tail recursion unrolled into a loop)
VIC_HDI void mass_release(double* InterceptedSnow, double* TempInterceptionStorage, double* ReleasedMass, double* Drip) {
  for (;;) {
    if (*InterceptedSnow > MIN_INTERCEPTION_STORAGE) {
      const double Threshold = 0.10 * *InterceptedSnow;
      const double MaxRelease = 0.17 * *InterceptedSnow;
      if ((*TempInterceptionStorage) >= Threshold) {
        *Drip += Threshold;
        *InterceptedSnow -= Threshold;
        *TempInterceptionStorage -= Threshold;
        double TempReleasedMass;
        if (*InterceptedSnow < MIN_INTERCEPTION_STORAGE) TempReleasedMass = 0.0;
        else TempReleasedMass = vmin((*InterceptedSnow - MIN_INTERCEPTION_STORAGE), MaxRelease);
        *ReleasedMass += TempReleasedMass;
        *InterceptedSnow -= TempReleasedMass;
        continue;
      } else {
        const double TempDrip = vmin(*TempInterceptionStorage, *InterceptedSnow);
        *Drip += TempDrip;
        *InterceptedSnow -= TempDrip;
      }
    } else {
      const double TempDrip = vmin(*TempInterceptionStorage, *InterceptedSnow);
      *Drip += TempDrip;
      *InterceptedSnow -= TempDrip;
      *TempInterceptionStorage = 0.0;
    }
    return;
  }
}

// snow_intercept.c:81-582 with F == 1.  energy: the sub-step's snow-side energy record;
// RainFall / SnowFall in mm in and out; LongOverOut is the canopy's downward longwave (becomes
// the understory's incoming longwave).
template <int NN, class EN>
VIC_HDI int snow_intercept(double Dt, double LAI, double latent_heat_Le, double LongOverIn, double LongUnderOut, double MaxInt,
                           double ShortOverIn, double Tcanopy, double bare_albedo, EN& energy, SnowPack& snow, VegVar& vv,
                           double* LongOverOut, const Surf4& Ra, RaUsed& Ra_used, double* RainFall, double* SnowFall,
                           const Surf4& wind_speed, const Surf4& displacement, const Surf4& ref_height, const Surf4& roughness,
                           const VegNow& veg, const SoilET& soil, SoilLayer* layer, double AirDens, double EactAir, double Press, double Vpd,
                           const CellPar& cp, const Opts& o) {
  const double F = 1.;
  double* IntRain = &vv.Wdew;
  double* IntSnow = &snow.snow_canopy;
  double* Tfoliage = &energy.Tfoliage;
  double* MeltEnergy = &energy.canopy_refreeze;
  double* VaporMassFlux = &snow.canopy_vapor_flux;
  double* TempIntStorage = &snow.tmp_int_storage;
  double Drip, ReleasedMass, Evap = 0, NetRadiation = 0, RefreezeEnergy = 0, Qnet;
  energy.Tfoliage_fbflag = 0;
  *RainFall /= 1000.;
  *SnowFall /= 1000.;
  *IntRain /= 1000.;
  MaxInt /= 1000.;
  const double IntRainOrg = *IntRain;
  *IntSnow /= F;
  *IntRain /= F;
  const double InitialSnowInt = *IntSnow;
  Drip = 0.0;
  ReleasedMass = 0.0;
  const double OldTfoliage = *Tfoliage;
  // maximum snow interception storage
  const double Imax1 = 4.0 * LAI_SNOW_MULTIPLIER * LAI;
  double MaxSnowInt;
  if ((*Tfoliage) < -1.0 && (*Tfoliage) > -3.0) MaxSnowInt = ((*Tfoliage) * 3.0 / 2.0) + (11.0 / 2.0);
  else if ((*Tfoliage) > -1.0) MaxSnowInt = 4.0;
  else MaxSnowInt = 1.0;
  MaxSnowInt *= LAI_SNOW_MULTIPLIER * LAI;
  double DeltaSnowInt = (1 - *IntSnow / MaxSnowInt) * *SnowFall;
  if (DeltaSnowInt + *IntSnow > MaxSnowInt) DeltaSnowInt = MaxSnowInt - *IntSnow;
  if (DeltaSnowInt < 0.0) DeltaSnowInt = 0.0;
  // snow blown off cold branches
  if ((*Tfoliage) < -3.0 && DeltaSnowInt > 0.0 && wind_speed[CANOPY_OVER] > 1.0) {
    double BlownSnow = (0.2 * wind_speed[CANOPY_OVER] - 0.2) * DeltaSnowInt;
    if (BlownSnow >= DeltaSnowInt) BlownSnow = DeltaSnowInt;
    DeltaSnowInt -= BlownSnow;
  }
  if (*IntSnow + DeltaSnowInt > Imax1) DeltaSnowInt = 0.0;
  double SnowThroughFall = (*SnowFall - DeltaSnowInt) * F + (*SnowFall) * (1 - F);
  if (*SnowFall == 0 && *IntSnow < MIN_SWQ_EB_THRES) {
    SnowThroughFall += *IntSnow;
    DeltaSnowInt -= *IntSnow;
  }
  *IntSnow += DeltaSnowInt;
  if (*IntSnow < SMALL) *IntSnow = 0.0;
  // rain interception
  double MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
  double RainThroughFall;
  if ((*IntRain + *RainFall) <= MaxWaterInt) {
    *IntRain += *RainFall;
    RainThroughFall = *RainFall * (1 - F);
  } else {
    RainThroughFall = (*IntRain + *RainFall - MaxWaterInt) * F + (*RainFall * (1 - F));
    *IntRain = MaxWaterInt;
  }
  if (*RainFall == 0 && *IntRain < MIN_SWQ_EB_THRES) {
    RainThroughFall += *IntRain;
    *IntRain = 0.0;
  }
  if (*IntRain + *IntSnow > Imax1) {
    const double Overload = (*IntSnow + *IntRain) - Imax1;
    const double IntRainFract = *IntRain / (*IntRain + *IntSnow);
    const double IntSnowFract = *IntSnow / (*IntRain + *IntSnow);
    *IntRain = *IntRain - Overload * IntRainFract;
    *IntSnow = *IntSnow - Overload * IntSnowFract;
    RainThroughFall = RainThroughFall + (Overload * IntRainFract) * F;
    SnowThroughFall = SnowThroughFall + (Overload * IntSnowFract) * F;
  }
  if (*IntRain + *IntSnow < SMALL) *Tfoliage = Tcanopy;

  // canopy energy balance
  CanopyEB eb;
  eb.stab.reset();
  eb.memo.reset();
  eb.delta_t = Dt; eb.elevation = cp(CP_elevation); eb.AirDens = AirDens; eb.EactAir = EactAir; eb.Press = Press;
  eb.latent_heat_Le = latent_heat_Le; eb.Tcanopy = Tcanopy; eb.Vpd = Vpd; eb.IntRain = IntRainOrg; eb.LongOverIn = LongOverIn;
  eb.LongUnderOut = LongUnderOut; eb.AERO_RESIST_CANSNOW = o.AERO_RESIST_CANSNOW;
  eb.ra_free = Ra[SNOW_FREE]; eb.ra_over = Ra[CANOPY_OVER]; eb.ws_over = wind_speed[CANOPY_OVER]; eb.zref_over = ref_height[CANOPY_OVER];
  eb.disp_over = displacement[CANOPY_OVER]; eb.rough_over = roughness[CANOPY_OVER];
  eb.veg = &veg; eb.soil = &soil; eb.Ra_used = &Ra_used; eb.Rainfall = RainFall; eb.Wdew = IntRain; eb.layer = layer; eb.vv = &vv;
  eb.Evap = &Evap; eb.AdvectedEnergy = &energy.canopy_advection; eb.LatentHeat = &energy.canopy_latent;
  eb.LatentHeatSub = &energy.canopy_latent_sub; eb.LongOverOut = LongOverOut; eb.NetLongOver = &energy.NetLongOver;
  eb.NetRadiation = &NetRadiation; eb.RefreezeEnergy = &RefreezeEnergy; eb.SensibleHeat = &energy.canopy_sensible;
  eb.VaporMassFlux = VaporMassFlux;

  // The balance at 0 C when there is snow in or falling on the canopy, the solve it may call for and the last evaluation at the
  // accepted temperature (snow_intercept.c:391-430) go through the one residual call site of root_brent_ss_impl.
  const bool snowy = (*IntSnow > 0 || *SnowFall > 0);
  energy.AlbedoOver = snowy ? cp(CP_NEW_SNOW_ALB) : bare_albedo;
  energy.NetShortOver = (1. - energy.AlbedoOver) * ShortOverIn;
  eb.IntSnow = *IntSnow;
  eb.NetShortOver = energy.NetShortOver;
  struct Inlined {
    CanopyEB& f;
    VIC_HD double operator()(double x) { return f.eval(x); }
    VIC_HD void before_final() {}
  } call{eb};
  BrentFinal fin;
  fin.allow_fallback = o.TFALLBACK;
  fin.fallback_x = OldTfoliage;
  fin.nosolve_x = 0;
  fin.f_final = 0;
  fin.fell_back = 0;
  const double T_guess = *Tfoliage;
  bool at_zero = false;
  Qnet = vnan();
  auto decide = [&](double f0) {
    Qnet = f0;
    if (f0 != 0) {
      fin.hi = 0;
      fin.lo = (T_guess <= 0.) ? T_guess - SNOW_DT : -SNOW_DT;
      fin.do_solve = is_valid(fin.lo);
    } else {
      at_zero = true;
      fin.do_solve = false;
    }
    fin.do_final = fin.do_solve;
  };
  if (!snowy) {
    fin.hi = T_guess + SNOW_DT;
    fin.lo = T_guess - SNOW_DT;
    fin.do_solve = is_valid(fin.hi) && is_valid(fin.lo);
    fin.do_final = fin.do_solve;
  }
  const double T_solved = root_brent_ss_impl<true, true>(0., 0., call, &fin, snowy, 0.0, decide);
  if (at_zero) *Tfoliage = 0.;
  if (fin.do_solve) {
    if (fin.fell_back) {
      *Tfoliage = OldTfoliage;
      energy.Tfoliage_fbflag = 1;
      energy.Tfoliage_fbcount += 1;
    } else if (result_is_error(T_solved)) return ERROR_I;
    else *Tfoliage = T_solved;
    Qnet = fin.f_final;
  }
  (void)Qnet;
  if (*IntSnow <= 0) RainThroughFall = vv.throughfall / 1000.;
  RefreezeEnergy *= Dt;
  MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
  *VaporMassFlux *= Dt;
  if (*Tfoliage == 0) {
    if (-(*VaporMassFlux) > *IntRain) {
      *VaporMassFlux = -(*IntRain);
      *IntRain = 0.;
    } else *IntRain += *VaporMassFlux;
    double PotSnowMelt;
    if (RefreezeEnergy < 0) {
      PotSnowMelt = vmin((-RefreezeEnergy / Lf / RHO_W), *IntSnow);
      *MeltEnergy -= (Lf * PotSnowMelt * RHO_W) / (Dt);
    } else {
      PotSnowMelt = 0;
      *MeltEnergy -= (Lf * PotSnowMelt * RHO_W) / (Dt);
    }
    if ((*IntRain + PotSnowMelt) <= MaxWaterInt) {
      *IntSnow -= PotSnowMelt;
      *IntRain += PotSnowMelt;
      PotSnowMelt = 0.0;
    } else {
      const double ExcessSnowMelt = PotSnowMelt + *IntRain - MaxWaterInt;
      *IntSnow -= MaxWaterInt - (*IntRain);
      *IntRain = MaxWaterInt;
      if (*IntSnow < 0.0) *IntSnow = 0.0;
      if (SnowThroughFall > 0.0 && InitialSnowInt <= MIN_INTERCEPTION_STORAGE) {
        Drip += ExcessSnowMelt;
        *IntSnow -= ExcessSnowMelt;
        if (*IntSnow < 0.0) *IntSnow = 0.0;
      } else *TempIntStorage += ExcessSnowMelt;
      mass_release(IntSnow, TempIntStorage, &ReleasedMass, &Drip);
    }
    MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
    if (*IntRain > MaxWaterInt) {
      Drip += *IntRain - MaxWaterInt;
      *IntRain = MaxWaterInt;
    }
  } else {
    *TempIntStorage = 0.0;
    if (-RefreezeEnergy > -(*IntRain) * Lf) {
      *IntSnow += fabs(RefreezeEnergy) / Lf;
      *IntRain -= fabs(RefreezeEnergy) / Lf;
      *MeltEnergy += (fabs(RefreezeEnergy) * RHO_W) / (Dt);
      RefreezeEnergy = 0.0;
    } else {
      *IntSnow += *IntRain;
      *MeltEnergy += (Lf * *IntRain * RHO_W) / (Dt);
      *IntRain = 0.0;
    }
    if (-(*VaporMassFlux) > *IntSnow) {
      *VaporMassFlux = -(*IntSnow);
      *IntSnow = 0.0;
    } else *IntSnow += *VaporMassFlux;
  }
  *IntSnow *= F;
  *IntRain *= F;
  *MeltEnergy *= F;
  *VaporMassFlux *= F;
  Drip *= F;
  ReleasedMass *= F;
  if (*IntSnow == 0 && *IntRain > MaxInt) {
    RainThroughFall += *IntRain - MaxInt;
    *IntRain = MaxInt;
  }
  *RainFall = RainThroughFall + Drip;
  *SnowFall = SnowThroughFall + ReleasedMass;
  *VaporMassFlux *= -1.;
  *RainFall *= 1000.;
  *SnowFall *= 1000.;
  *IntRain *= 1000.;
  *MeltEnergy = RefreezeEnergy / Dt;
  return 0;
}

}  // namespace vic
#endif

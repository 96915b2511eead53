// vic_lakeice.cuh -- the lake-ice / snow-on-ice surface solve of the reference's lake model:
//   IceEnergyBalance::calculate      IceEnergyBalance.c:60-175   the surface energy balance residual of the ice / snow pack
//   ice_melt                         ice_melt.c:30-585           mass and energy update of snow and lake ice over one sub-step
//   icerad                           lakes.eb.c:1092-1151        conductivity of the snow + ice slab, short wave absorbed in it
// restated operation by operation (bit-exact against the reference's own ice_melt(), oracle/icemeltcheck.cpp).  Served as a batch
// operator behind vicgpu_ice_melt (include/vicgpu.h): one thread per lake-ice column.  The lake itself -- solve_lake's water column,
// water_balance and the wetland rescaling of lakes.eb.c -- is not built, so LAKES TRUE stays rejected by vicgpu_create; ice_melt's
// BLOWING branch (ice_melt.c:239-258) is not served either (blowing_flux = 0, as with BLOWING FALSE).
#ifndef VIC_LAKEICE_CUH
#define VIC_LAKEICE_CUH
#include "vic_brent.cuh"
#include "vic_leaf.cuh"

namespace vic {

// LAKE.h:37-63
constexpr double LK_RHOSNOW = 250.;
constexpr double LK_CONDI = 2.3, LK_CONDS = 0.7;
constexpr double LK_lamisw = 1.5, LK_lamilw = 20, LK_lamssw = 6.0, LK_lamslw = 20;
constexpr double LK_a1 = 0.7, LK_a2 = 0.3;
// snow.h:34-68
constexpr double LK_LIQUID_WATER_CAPACITY = 0.035;
constexpr double LK_MAX_SURFACE_SWE = 0.125;
constexpr double LK_MIN_SWQ_EB_THRES = 0.0010;
constexpr double LK_SNOW_DT = 5.0;  // vicNl_def.h:298

// lakes.eb.c:1092-1151
VIC_HD void icerad(double sw, double hi, double hs, double* avgcond, double* SWnet, double* SW_under_ice) {
  *avgcond = (hs * LK_CONDI + hi * LK_CONDS) / (LK_CONDI * LK_CONDS);
  const double a = -1. * (1. - vexp(-LK_lamssw * hs)) / (LK_CONDS * LK_lamssw);
  const double b = -1. * vexp(-LK_lamssw * hs) * (1 - vexp(-LK_lamisw * hi)) / (LK_CONDI * LK_lamisw);
  const double c = -1. * (1. - vexp(-LK_lamslw * hs)) / (LK_CONDS * LK_lamslw);
  const double d = -1. * vexp(-LK_lamslw * hs) * (1 - vexp(-LK_lamilw * hi)) / (LK_CONDI * LK_lamilw);
  *SWnet = sw * LK_a1 * (a + b) + sw * LK_a2 * (c + d);
  *SW_under_ice = (LK_a1 * sw * (1 - vexp(-(LK_lamssw * hs + LK_lamisw * hi))) + LK_a2 * sw * (1 - vexp(-(LK_lamslw * hs + LK_lamilw * hi))));
}

// IceEnergyBalance.c:60-175.  The reference's functor writes its flux terms through pointers into ice_melt's locals; here they are
// members, and the values the LAST evaluation left are the result, as there.
struct IceEB {
  double Dt, Ra, Z, Z0, Wind, ShortRad, LongRadIn, AirDens, Lv, Tair, Press, Vpd, EactAir, Rain, SurfaceLiquidWater, Tfreeze, AvgCond, SWconducted;
  double Ra_used, RefreezeEnergy, vapor_flux, blowing_flux, surface_flux, AdvectedEnergy, qf, LatentHeat, LatentHeatSub, SensibleHeat, LongRadOut;
  StabLog stab;
  VIC_HDI double operator()(double TSurf) {
    const double TMean = TSurf;
    const double Density = RHO_W;
    if (Wind > 0.0) Ra_used = Ra / stab.correction(Z, 0., TMean, Tair, Wind, Z0);
    else Ra_used = HUGE_RESIST;
    LongRadOut = LongRadIn - STEFAN_B * (TMean + 273.15) * (TMean + 273.15) * (TMean + 273.15) * (TMean + 273.15);
    const double NetRad = ShortRad + LongRadOut;
    SensibleHeat = AirDens * CP_PM * (Tair - TMean) / Ra_used;
    double VaporMassFlux, SurfaceMassFlux;
    double BlowingMassFlux = blowing_flux * Density / (Dt * SECPHOUR);
    latent_heat_from_snow(AirDens, EactAir, Lv, Press, Ra, TMean, Vpd, &LatentHeat, &LatentHeatSub, &VaporMassFlux, &BlowingMassFlux, &SurfaceMassFlux);
    vapor_flux = VaporMassFlux * Dt * SECPHOUR / Density;
    surface_flux = SurfaceMassFlux * Dt * SECPHOUR / Density;
    AdvectedEnergy = (CH_WATER * Tair * Rain) / (Dt * SECPHOUR);
    const double qnull = (1 / AvgCond) * (Tfreeze - TMean + SWconducted);
    qf = qnull;
    double RestTerm = (NetRad + SensibleHeat + LatentHeat + LatentHeatSub + AdvectedEnergy + qnull);
    RefreezeEnergy = (SurfaceLiquidWater * Lf * Density) / (Dt * SECPHOUR);
    if (TSurf == 0.0 && RestTerm > -RefreezeEnergy) {
      RefreezeEnergy = -RestTerm;
      RestTerm = 0.0;
    } else {
      RestTerm += RefreezeEnergy;
    }
    return RestTerm;
  }
};

// members of snow_data_struct / lake_var_struct that ice_melt reads or writes (vicNl_def.h:1183-1232, 1285-1340)
struct IceSnow {
  double swq, surf_temp, pack_temp, pack_water, surf_water, vapor_flux, blowing_flux, surface_flux, surf_temp_fbflag, surf_temp_fbcount, coverage, mass_error,
      coldcontent;
};
struct IceLake {
  double ice_water_eq, areai, hice, volume;
};
struct IceMeltOut {
  double aero_resist_used, melt, advection, deltaCC, SnowFlux, latent, sensible, Qnet, refreeze_energy, LWnet;
};

// ice_melt.c:30-585.  Returns 0, or ERROR_I when the surface solve fails and TFALLBACK is off (outputs then undefined, as in the reference).
VIC_HDI int ice_melt(double z2, double aero_resist, double latent_heat_Le, IceSnow& snow, IceLake& lake, int delta_t, double Z0, double rainfall,
                     double snowfall, double wind, double Tcutoff, double air_temp, double net_short, double longwave, double density, double pressure,
                     double vpd, double vp, bool TFALLBACK, IceMeltOut& out) {
  double DeltaPackCC, DeltaPackSwq, MaxLiquidWater, Qnet, PackRefreezeEnergy, RefrozenWater, SnowFallCC, SurfaceCC, PackCC, SurfaceSwq, PackSwq, PackIce, SnowMelt, IceMelt;
  double avgcond, SWconducted, deltaCC;
  double melt_energy = 0.;
  const double SnowFall = snowfall / 1000.;
  const double RainFall = rainfall / 1000.;
  IceMelt = 0.0;
  RefrozenWater = 0.0;
  const double InitialSwq = snow.swq;
  const double OldTSurf = snow.surf_temp;
  double SnowIce = snow.swq - snow.pack_water - snow.surf_water;
  double LakeIce = lake.ice_water_eq / lake.areai;
  const double InitialIce = LakeIce;
  double Ice = SnowIce + LakeIce;
  if (Ice > LK_MAX_SURFACE_SWE) SurfaceSwq = LK_MAX_SURFACE_SWE;
  else SurfaceSwq = Ice;
  if (SurfaceSwq <= SnowIce) {
    PackSwq = SnowIce - SurfaceSwq;
    PackIce = LakeIce;
  } else {
    PackSwq = 0.;
    PackIce = Ice - SurfaceSwq;
  }
  SurfaceCC = CH_ICE * SurfaceSwq * snow.surf_temp;
  PackCC = CH_ICE * (PackSwq + PackIce) * snow.pack_temp;
  if (air_temp > 0.0) SnowFallCC = 0.0;
  else SnowFallCC = CH_ICE * SnowFall * air_temp;
  if (SnowFall > (LK_MAX_SURFACE_SWE - SurfaceSwq)) {
    DeltaPackSwq = SurfaceSwq + SnowFall - LK_MAX_SURFACE_SWE;
    if (DeltaPackSwq > SurfaceSwq) DeltaPackCC = SurfaceCC + (SnowFall - LK_MAX_SURFACE_SWE) / SnowFall * SnowFallCC;
    else DeltaPackCC = DeltaPackSwq / SurfaceSwq * SurfaceCC;
    SurfaceSwq = LK_MAX_SURFACE_SWE;
    SurfaceCC += SnowFallCC - DeltaPackCC;
    PackSwq += DeltaPackSwq;
    PackCC += DeltaPackCC;
  } else {
    SurfaceSwq += SnowFall;
    SurfaceCC += SnowFallCC;
    DeltaPackCC = 0;
  }
  if (SurfaceSwq > 0.0) snow.surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  else snow.surf_temp = 0.0;
  if (PackSwq + PackIce > 0.0) snow.pack_temp = PackCC / (CH_ICE * (PackSwq + PackIce));
  else snow.pack_temp = 0.0;
  SnowIce += SnowFall;
  Ice += SnowFall;
  snow.surf_water += RainFall;
  icerad(net_short, lake.hice, SnowIce * RHO_W / LK_RHOSNOW, &avgcond, &SWconducted, &deltaCC);
  snow.blowing_flux = 0.0;  // (BLOWING branch, ice_melt.c:239-258: not served)

  IceEB eb;
  eb.stab.reset();
  eb.Dt = (double)delta_t; eb.Ra = aero_resist; eb.Z = z2; eb.Z0 = Z0; eb.Wind = wind; eb.ShortRad = net_short; eb.LongRadIn = longwave;
  eb.AirDens = density; eb.Lv = latent_heat_Le; eb.Tair = air_temp; eb.Press = pressure * 1000.; eb.Vpd = vpd * 1000.; eb.EactAir = vp * 1000.;
  eb.Rain = RainFall; eb.SurfaceLiquidWater = snow.surf_water; eb.Tfreeze = Tcutoff; eb.AvgCond = avgcond; eb.SWconducted = SWconducted;
  eb.vapor_flux = snow.vapor_flux; eb.blowing_flux = snow.blowing_flux; eb.surface_flux = snow.surface_flux;
  double& RefreezeEnergy = eb.RefreezeEnergy;
  Qnet = eb(0.0);
  snow.vapor_flux = eb.vapor_flux;
  snow.surface_flux = eb.surface_flux;
  if (Qnet == 0.0) {
    snow.surf_temp = 0.0;
    if (RefreezeEnergy >= 0.0) {
      RefrozenWater = RefreezeEnergy / (Lf * RHO_W) * delta_t * SECPHOUR;
      if (RefrozenWater > snow.surf_water) {
        RefrozenWater = snow.surf_water;
        RefreezeEnergy = RefrozenWater * Lf * RHO_W / (delta_t * SECPHOUR);
      }
      melt_energy += RefreezeEnergy;
      SurfaceSwq += RefrozenWater;
      SnowIce += RefrozenWater;
      Ice += RefrozenWater;
      snow.surf_water -= RefrozenWater;
      if (snow.surf_water < 0.0) snow.surf_water = 0.0;
      SnowMelt = 0.0;
    } else {
      SnowMelt = fabs(RefreezeEnergy) / (Lf * RHO_W) * delta_t * SECPHOUR;
      melt_energy += RefreezeEnergy;
    }
    if (snow.surf_water < -(snow.vapor_flux)) {
      snow.blowing_flux *= -(snow.surf_water) / snow.vapor_flux;
      snow.vapor_flux = -(snow.surf_water);
      snow.surface_flux = -(snow.surf_water) - snow.blowing_flux;
      snow.surf_water = 0.0;
    } else {
      snow.surf_water += snow.vapor_flux;
    }
    if (SnowMelt < Ice) {
      if (SnowMelt <= PackSwq) {
        snow.surf_water += SnowMelt;
        PackSwq -= SnowMelt;
        Ice -= SnowMelt;
        SnowIce -= SnowMelt;
      } else if (SnowMelt <= SnowIce) {
        snow.surf_water += SnowMelt + snow.pack_water;
        snow.pack_water = 0.0;
        SurfaceSwq -= (SnowMelt - PackSwq);
        PackSwq = 0.0;
        SnowIce -= SnowMelt;
        Ice -= SnowMelt;
      } else {
        snow.surf_water += SnowIce + snow.pack_water;
        snow.pack_water = 0.0;
        PackSwq = 0.0;
        Ice -= SnowMelt;
        LakeIce -= SnowMelt - SnowIce;
        IceMelt = SnowMelt - SnowIce;
        if (SurfaceSwq > SnowMelt) {
          SurfaceSwq -= SnowMelt;
        } else {
          SurfaceSwq = 0.0;
          PackIce -= (SnowMelt - SurfaceSwq - PackSwq);
        }
        SnowIce = 0.0;
      }
    } else {
      snow.surf_water += SnowIce + snow.pack_water;
      snow.pack_water = 0.0;
      PackSwq = 0.0;
      SurfaceSwq = 0.0;
      SnowIce = 0.0;
      SnowMelt = Ice;
      IceMelt = LakeIce;
      LakeIce = 0.0;
      PackIce = 0.0;
      Ice = 0.0;
      snow.surf_temp = 0.0;
      snow.pack_temp = 0.0;
      melt_energy -= RefreezeEnergy;
      RefreezeEnergy = RefreezeEnergy / fabs(RefreezeEnergy) * SnowMelt * Lf * RHO_W / (delta_t);
      melt_energy += RefreezeEnergy;
    }
  } else {
    if (SurfaceSwq > LK_MIN_SWQ_EB_THRES) {
      snow.surf_temp = root_brent((double)(snow.surf_temp - LK_SNOW_DT), (double)(snow.surf_temp + LK_SNOW_DT), eb);
      if (snow.surf_temp <= -998) {
        if (TFALLBACK) {
          snow.surf_temp = OldTSurf;
          snow.surf_temp_fbflag = 1;
          snow.surf_temp_fbcount++;
        } else {
          return ERROR_I;
        }
      }
    } else {
      snow.surf_temp = NAN;  // INVALID
    }
    if (snow.surf_temp == snow.surf_temp && !(snow.surf_temp <= -998)) {
      Qnet = eb(snow.surf_temp);
      snow.vapor_flux = eb.vapor_flux;
      snow.surface_flux = eb.surface_flux;
      SnowMelt = 0.0;
      IceMelt = 0.0;
      SnowIce += snow.surf_water;
      Ice += snow.surf_water;
      melt_energy += snow.surf_water * Lf * RHO_W / (delta_t * SECPHOUR);
      RefrozenWater = snow.surf_water;
      snow.surf_water = 0.0;
      if (SurfaceSwq < -(snow.vapor_flux)) {
        if (SurfaceSwq > SnowIce) {
          snow.blowing_flux *= -(SurfaceSwq) / snow.vapor_flux;
          snow.vapor_flux = -SurfaceSwq;
          snow.surface_flux = -SurfaceSwq - snow.blowing_flux;
          LakeIce -= SurfaceSwq - SnowIce;
          Ice = PackIce;
          SnowIce = 0.0;
        } else {
          snow.blowing_flux *= -(SurfaceSwq) / snow.vapor_flux;
          snow.vapor_flux = -SurfaceSwq;
          snow.surface_flux = -SurfaceSwq - snow.blowing_flux;
          SurfaceSwq = 0.0;
          Ice = PackSwq + PackIce;
        }
      } else {
        SurfaceSwq += snow.vapor_flux;
        if (SnowIce > -(snow.vapor_flux)) SnowIce += snow.vapor_flux;
        else {
          LakeIce += (snow.vapor_flux + SnowIce);
          SnowIce = 0.;
        }
        Ice += snow.vapor_flux;
      }
    } else {
      snow.surf_temp = NAN;
    }
  }
  if (SnowIce > SurfaceSwq) MaxLiquidWater = LK_LIQUID_WATER_CAPACITY * SurfaceSwq;
  else MaxLiquidWater = LK_LIQUID_WATER_CAPACITY * SnowIce;
  double melt;
  if (snow.surf_water > MaxLiquidWater) {
    melt = snow.surf_water - MaxLiquidWater;
    snow.surf_water = MaxLiquidWater;
  } else melt = 0.0;
  snow.pack_water += melt;
  PackRefreezeEnergy = snow.pack_water * Lf * RHO_W;
  if (PackCC < -PackRefreezeEnergy) {
    PackSwq += snow.pack_water;
    Ice += snow.pack_water;
    SnowIce += snow.pack_water;
    snow.pack_water = 0.0;
    if (PackSwq + PackIce > 0.0) {
      PackCC = (PackSwq + PackIce) * CH_ICE * snow.pack_temp + PackRefreezeEnergy;
      snow.pack_temp = PackCC / (CH_ICE * (PackSwq + PackIce));
      if (snow.pack_temp > 0.) snow.pack_temp = 0.;
    } else snow.pack_temp = 0.0;
  } else {
    snow.pack_temp = 0.0;
    DeltaPackSwq = -PackCC / (Lf * RHO_W);
    snow.pack_water -= DeltaPackSwq;
    PackSwq += DeltaPackSwq;
    Ice += DeltaPackSwq;
    SnowIce += DeltaPackSwq;
  }
  MaxLiquidWater = LK_LIQUID_WATER_CAPACITY * PackSwq;
  if (snow.pack_water > MaxLiquidWater) {
    melt = snow.pack_water - MaxLiquidWater;
    snow.pack_water = MaxLiquidWater;
  } else melt = 0.0;
  Ice = PackIce + PackSwq + SurfaceSwq;
  if (Ice > LK_MAX_SURFACE_SWE) {
    SurfaceCC = CH_ICE * snow.surf_temp * SurfaceSwq;
    PackCC = CH_ICE * snow.pack_temp * (PackSwq + PackIce);
    if (SurfaceSwq > LK_MAX_SURFACE_SWE) {
      PackCC += SurfaceCC * (SurfaceSwq - LK_MAX_SURFACE_SWE) / SurfaceSwq;
      SurfaceCC -= SurfaceCC * (SurfaceSwq - LK_MAX_SURFACE_SWE) / SurfaceSwq;
      PackSwq += SurfaceSwq - LK_MAX_SURFACE_SWE;
      SurfaceSwq -= SurfaceSwq - LK_MAX_SURFACE_SWE;
    } else if (SurfaceSwq < LK_MAX_SURFACE_SWE) {
      PackCC -= PackCC * (LK_MAX_SURFACE_SWE - SurfaceSwq) / (PackSwq + PackIce);
      SurfaceCC += PackCC * (LK_MAX_SURFACE_SWE - SurfaceSwq) / (PackSwq + PackIce);
      PackSwq -= LK_MAX_SURFACE_SWE - SurfaceSwq;
      SurfaceSwq += LK_MAX_SURFACE_SWE - SurfaceSwq;
    }
    snow.pack_temp = PackCC / (CH_ICE * (PackSwq + PackIce));
    snow.surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  } else {
    PackSwq = 0.0;
    PackCC = 0.0;
    PackIce = 0.0;
    snow.pack_temp = 0.0;
  }
  snow.swq = SnowIce + snow.surf_water + snow.pack_water;
  lake.ice_water_eq = LakeIce * lake.areai;
  lake.volume -= (InitialIce - LakeIce - IceMelt) * lake.areai;
  if (lake.ice_water_eq <= 0.0) lake.ice_water_eq = 0.0;
  if (snow.swq > 0) snow.coverage = 1.;
  else snow.coverage = 0.;
  const double MassBalanceError = (InitialSwq - snow.swq) + (InitialIce - LakeIce) + (RainFall + SnowFall) - IceMelt - melt + snow.vapor_flux;
  melt *= 1000.;
  snow.mass_error = MassBalanceError;
  snow.coldcontent = SurfaceCC;
  snow.vapor_flux *= -1.;
  out.aero_resist_used = eb.Ra_used;
  out.melt = melt;
  out.LWnet = eb.LongRadOut;
  out.advection = eb.AdvectedEnergy;
  out.deltaCC = deltaCC;
  out.SnowFlux = eb.qf;
  out.latent = eb.LatentHeat + eb.LatentHeatSub;
  out.sensible = eb.SensibleHeat;
  out.refreeze_energy = RefreezeEnergy;
  out.Qnet = Qnet;
  (void)melt_energy; (void)RefrozenWater; (void)DeltaPackCC; (void)SnowMelt;
  return 0;
}

}  // namespace vic
#endif

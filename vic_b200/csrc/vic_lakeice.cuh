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

// The slab that lies on the lake water, in metres of water equivalent: snow over lake ice, cut into a surface layer that exchanges
// energy with the air (at most LK_MAX_SURFACE_SWE thick, snow first) and the pack below it.  ice_melt.c keeps these as a dozen loose
// locals; here they travel together through the phases of the sub-step.
struct IceSlab {
  double top;        // surface layer
  double pack_snow;  // snow below the surface layer
  double pack_ice;   // lake ice below the surface layer
  double snow;       // all frozen snow (surface layer + pack)
  double lake;       // all lake ice
  double all;        // snow + lake
  double top_cc;     // cold content of the surface layer [J/m2]
  double pack_cc;    // cold content of the pack

  // ice_melt.c:139-160: cut the slab at the surface-layer thickness
  VIC_HD void cut(double top_temp, double pack_temp) {
    all = snow + lake;
    top = all > LK_MAX_SURFACE_SWE ? LK_MAX_SURFACE_SWE : all;
    if (top <= snow) {
      pack_snow = snow - top;
      pack_ice = lake;
    } else {
      pack_snow = 0.;
      pack_ice = all - top;
    }
    top_cc = CH_ICE * top * top_temp;
    pack_cc = CH_ICE * (pack_snow + pack_ice) * pack_temp;
  }
  // ice_melt.c:162-205: new snow lands on the surface layer; what no longer fits moves down with its share of cold content
  VIC_HD void snowfall(double fall, double air_temp) {
    const double fall_cc = air_temp > 0.0 ? 0.0 : CH_ICE * fall * air_temp;
    if (fall > (LK_MAX_SURFACE_SWE - top)) {
      const double down = top + fall - LK_MAX_SURFACE_SWE;
      double down_cc;
      if (down > top) down_cc = top_cc + (fall - LK_MAX_SURFACE_SWE) / fall * fall_cc;
      else down_cc = down / top * top_cc;
      top = LK_MAX_SURFACE_SWE;
      top_cc += fall_cc - down_cc;
      pack_snow += down;
      pack_cc += down_cc;
    } else {
      top += fall;
      top_cc += fall_cc;
    }
    snow += fall;
    all += fall;
  }
  VIC_HD double top_temp() const { return top > 0.0 ? top_cc / (CH_ICE * top) : 0.0; }
  VIC_HD double pack_temp() const { return pack_snow + pack_ice > 0.0 ? pack_cc / (CH_ICE * (pack_snow + pack_ice)) : 0.0; }
};

// ice_melt.c:30-585.  Returns 0, or ERROR_I when the surface solve fails and TFALLBACK is off (outputs then undefined, as in the reference).
VIC_HDI int ice_melt(double z2, double aero_resist, double latent_heat_Le, IceSnow& sn, IceLake& lk, int delta_t, double Z0, double rainfall,
                     double snowfall, double wind, double Tcutoff, double air_temp, double net_short, double longwave, double density, double pressure,
                     double vpd, double vp, bool TFALLBACK, IceMeltOut& out) {
  const double fall = snowfall / 1000., rain = rainfall / 1000.;  // [m]
  const double swq0 = sn.swq, T_old = sn.surf_temp;
  IceSlab s;
  s.snow = sn.swq - sn.pack_water - sn.surf_water;
  s.lake = lk.ice_water_eq / lk.areai;
  const double lake0 = s.lake;
  s.cut(sn.surf_temp, sn.pack_temp);
  s.snowfall(fall, air_temp);
  sn.surf_temp = s.top_temp();
  sn.pack_temp = s.pack_temp();
  sn.surf_water += rain;

  double slab_cond, sw_conducted, sw_through;
  icerad(net_short, lk.hice, s.snow * RHO_W / LK_RHOSNOW, &slab_cond, &sw_conducted, &sw_through);
  sn.blowing_flux = 0.0;  // (BLOWING branch, ice_melt.c:239-258: not served)

  IceEB eb;
  eb.stab.reset();
  eb.Dt = (double)delta_t; eb.Ra = aero_resist; eb.Z = z2; eb.Z0 = Z0; eb.Wind = wind; eb.ShortRad = net_short; eb.LongRadIn = longwave;
  eb.AirDens = density; eb.Lv = latent_heat_Le; eb.Tair = air_temp; eb.Press = pressure * 1000.; eb.Vpd = vpd * 1000.; eb.EactAir = vp * 1000.;
  eb.Rain = rain; eb.SurfaceLiquidWater = sn.surf_water; eb.Tfreeze = Tcutoff; eb.AvgCond = slab_cond; eb.SWconducted = sw_conducted;
  eb.vapor_flux = sn.vapor_flux; eb.blowing_flux = sn.blowing_flux; eb.surface_flux = sn.surface_flux;
  double& refreeze = eb.RefreezeEnergy;  // [W/m2], left by the last evaluation of the balance
  double lake_melted = 0.0;
  const double step_s = delta_t * SECPHOUR;  // (int product, as in the reference)

  double Qnet = eb(0.0);  // the balance with the surface at the melting point
  sn.vapor_flux = eb.vapor_flux;
  sn.surface_flux = eb.surface_flux;
  if (Qnet == 0.0) {
    // ---- surface at 0 C: the residual went into refreezing (>= 0) or melting (< 0), ice_melt.c:292-392
    sn.surf_temp = 0.0;
    double melted;
    if (refreeze >= 0.0) {
      double refrozen = refreeze / (Lf * RHO_W) * delta_t * SECPHOUR;
      if (refrozen > sn.surf_water) {
        refrozen = sn.surf_water;
        refreeze = refrozen * Lf * RHO_W / step_s;
      }
      s.top += refrozen;
      s.snow += refrozen;
      s.all += refrozen;
      sn.surf_water -= refrozen;
      if (sn.surf_water < 0.0) sn.surf_water = 0.0;
      melted = 0.0;
    } else {
      melted = fabs(refreeze) / (Lf * RHO_W) * delta_t * SECPHOUR;
    }
    // vapour leaves the liquid water first; more than there is cannot leave
    if (sn.surf_water < -(sn.vapor_flux)) {
      sn.blowing_flux *= -(sn.surf_water) / sn.vapor_flux;
      sn.vapor_flux = -(sn.surf_water);
      sn.surface_flux = -(sn.surf_water) - sn.blowing_flux;
      sn.surf_water = 0.0;
    } else {
      sn.surf_water += sn.vapor_flux;
    }
    if (melted < s.all) {
      if (melted <= s.pack_snow) {  // (the reference takes it out of the pack's snow)
        sn.surf_water += melted;
        s.pack_snow -= melted;
        s.all -= melted;
        s.snow -= melted;
      } else if (melted <= s.snow) {  // all of the pack's snow and part of the surface layer
        sn.surf_water += melted + sn.pack_water;
        sn.pack_water = 0.0;
        s.top -= (melted - s.pack_snow);
        s.pack_snow = 0.0;
        s.snow -= melted;
        s.all -= melted;
      } else {  // all snow and some lake ice
        sn.surf_water += s.snow + sn.pack_water;
        sn.pack_water = 0.0;
        s.pack_snow = 0.0;
        s.all -= melted;
        s.lake -= melted - s.snow;
        lake_melted = melted - s.snow;
        if (s.top > melted) {
          s.top -= melted;
        } else {
          s.top = 0.0;
          s.pack_ice -= (melted - s.top - s.pack_snow);
        }
        s.snow = 0.0;
      }
    } else {  // everything melts
      sn.surf_water += s.snow + sn.pack_water;
      sn.pack_water = 0.0;
      s.pack_snow = 0.0;
      s.top = 0.0;
      s.snow = 0.0;
      melted = s.all;
      lake_melted = s.lake;
      s.lake = 0.0;
      s.pack_ice = 0.0;
      s.all = 0.0;
      sn.surf_temp = 0.0;
      sn.pack_temp = 0.0;
      refreeze = refreeze / fabs(refreeze) * melted * Lf * RHO_W / (delta_t);
    }
  } else {
    // ---- surface below 0 C: find its temperature (thick enough a layer only), ice_melt.c:394-500
    if (s.top > LK_MIN_SWQ_EB_THRES) {
      sn.surf_temp = root_brent((double)(sn.surf_temp - LK_SNOW_DT), (double)(sn.surf_temp + LK_SNOW_DT), eb);
      if (sn.surf_temp <= -998) {  // RootBrent::resultIsError
        if (!TFALLBACK) return ERROR_I;
        sn.surf_temp = T_old;
        sn.surf_temp_fbflag = 1;
        sn.surf_temp_fbcount++;
      }
    } else {
      sn.surf_temp = NAN;  // INVALID
    }
    if (sn.surf_temp == sn.surf_temp && !(sn.surf_temp <= -998)) {
      Qnet = eb(sn.surf_temp);
      sn.vapor_flux = eb.vapor_flux;
      sn.surface_flux = eb.surface_flux;
      // the liquid water of the surface layer freezes
      s.snow += sn.surf_water;
      s.all += sn.surf_water;
      sn.surf_water = 0.0;
      if (s.top < -(sn.vapor_flux)) {  // sublimation would take more than the surface layer holds
        sn.blowing_flux *= -(s.top) / sn.vapor_flux;
        sn.vapor_flux = -s.top;
        sn.surface_flux = -s.top - sn.blowing_flux;
        if (s.top > s.snow) {
          s.lake -= s.top - s.snow;
          s.all = s.pack_ice;
          s.snow = 0.0;
        } else {
          s.top = 0.0;
          s.all = s.pack_snow + s.pack_ice;
        }
      } else {
        s.top += sn.vapor_flux;
        if (s.snow > -(sn.vapor_flux)) s.snow += sn.vapor_flux;
        else {
          s.lake += (sn.vapor_flux + s.snow);
          s.snow = 0.;
        }
        s.all += sn.vapor_flux;
      }
    } else {
      sn.surf_temp = NAN;
    }
  }

  // ---- liquid water: what the surface layer cannot hold drains into the pack, refreezes there or leaves, ice_melt.c:502-560
  double hold = LK_LIQUID_WATER_CAPACITY * (s.snow > s.top ? s.top : s.snow);
  double drained = 0.0;
  if (sn.surf_water > hold) {
    drained = sn.surf_water - hold;
    sn.surf_water = hold;
  }
  sn.pack_water += drained;
  const double pack_refreeze = sn.pack_water * Lf * RHO_W;
  if (s.pack_cc < -pack_refreeze) {  // the pack is cold enough to freeze all of it
    s.pack_snow += sn.pack_water;
    s.all += sn.pack_water;
    s.snow += sn.pack_water;
    sn.pack_water = 0.0;
    if (s.pack_snow + s.pack_ice > 0.0) {
      s.pack_cc = (s.pack_snow + s.pack_ice) * CH_ICE * sn.pack_temp + pack_refreeze;
      sn.pack_temp = s.pack_cc / (CH_ICE * (s.pack_snow + s.pack_ice));
      if (sn.pack_temp > 0.) sn.pack_temp = 0.;
    } else sn.pack_temp = 0.0;
  } else {
    sn.pack_temp = 0.0;
    const double frozen = -s.pack_cc / (Lf * RHO_W);
    sn.pack_water -= frozen;
    s.pack_snow += frozen;
    s.all += frozen;
    s.snow += frozen;
  }
  hold = LK_LIQUID_WATER_CAPACITY * s.pack_snow;
  if (sn.pack_water > hold) {
    drained = sn.pack_water - hold;
    sn.pack_water = hold;
  } else drained = 0.0;

  // ---- bring the surface layer back to its thickness, ice_melt.c:562-600
  s.all = s.pack_ice + s.pack_snow + s.top;
  if (s.all > LK_MAX_SURFACE_SWE) {
    s.top_cc = CH_ICE * sn.surf_temp * s.top;
    s.pack_cc = CH_ICE * sn.pack_temp * (s.pack_snow + s.pack_ice);
    if (s.top > LK_MAX_SURFACE_SWE) {
      s.pack_cc += s.top_cc * (s.top - LK_MAX_SURFACE_SWE) / s.top;
      s.top_cc -= s.top_cc * (s.top - LK_MAX_SURFACE_SWE) / s.top;
      s.pack_snow += s.top - LK_MAX_SURFACE_SWE;
      s.top -= s.top - LK_MAX_SURFACE_SWE;
    } else if (s.top < LK_MAX_SURFACE_SWE) {
      s.pack_cc -= s.pack_cc * (LK_MAX_SURFACE_SWE - s.top) / (s.pack_snow + s.pack_ice);
      s.top_cc += s.pack_cc * (LK_MAX_SURFACE_SWE - s.top) / (s.pack_snow + s.pack_ice);
      s.pack_snow -= LK_MAX_SURFACE_SWE - s.top;
      s.top += LK_MAX_SURFACE_SWE - s.top;
    }
    sn.pack_temp = s.pack_cc / (CH_ICE * (s.pack_snow + s.pack_ice));
    sn.surf_temp = s.top_cc / (CH_ICE * s.top);
  } else {
    s.pack_snow = 0.0;
    s.pack_cc = 0.0;
    s.pack_ice = 0.0;
    sn.pack_temp = 0.0;
  }

  // ---- results, ice_melt.c:602-640
  sn.swq = s.snow + sn.surf_water + sn.pack_water;
  lk.ice_water_eq = s.lake * lk.areai;
  lk.volume -= (lake0 - s.lake - lake_melted) * lk.areai;
  if (lk.ice_water_eq <= 0.0) lk.ice_water_eq = 0.0;
  sn.coverage = sn.swq > 0 ? 1. : 0.;
  sn.mass_error = (swq0 - sn.swq) + (lake0 - s.lake) + (rain + fall) - lake_melted - drained + sn.vapor_flux;
  sn.coldcontent = s.top_cc;
  sn.vapor_flux *= -1.;
  out.aero_resist_used = eb.Ra_used;
  out.melt = drained * 1000.;
  out.LWnet = eb.LongRadOut;
  out.advection = eb.AdvectedEnergy;
  out.deltaCC = sw_through;
  out.SnowFlux = eb.qf;
  out.latent = eb.LatentHeat + eb.LatentHeatSub;
  out.sensible = eb.SensibleHeat;
  out.refreeze_energy = refreeze;
  out.Qnet = Qnet;
  return 0;
}

}  // namespace vic
#endif

// vic_surface.cuh -- land-surface step of a non-glacier HRU:
//   ground-surface energy balance residual      func_surf_energy_bal.c:9-403
//   surface temperature solve + bookkeeping     calc_surf_energy_bal.c:7-692
//   snow accumulation / ablation sub-step       solve_snow.c:7-544
//   sub-step loop, accumulation, runoff         surface_fluxes.c:17-956
//
// The reference's canopy-air / snow-flux iteration loops run exactly once (CLOSE_ENERGY is
// FALSE => MAX_ITER 0, surface_fluxes.c:8-13, 652-656), so the "iter_*" working copies it
// makes are identical to the "step_*" ones after every sub-step; this file keeps ONE working
// copy per quantity and saves the few pre-sub-step values the reference reads back from the
// older copy (snow depth, cold content, surface temperature).
#ifndef VIC_SURFACE_CUH
#define VIC_SURFACE_CUH
#include "vic_snow.cuh"
#include "vic_soil.cuh"
#include "vic_frozen.cuh"
#include "vic_blowing.cuh"

namespace vic {

// ---- ground surface energy balance --------------------------------------------------------
template <int NN>
struct SurfEB {
  const Opts* o;
  const CellPar* cp;
  const VegNow* veg;
  const SoilET* soil;
  // what the residual reads of *o, *cp, *veg and the aerodynamic tables, copied once per solve (prepare) so that an evaluation
  // does not chase pointers: each of these was two or three dependent thread-local / global loads per evaluation
  int QUICK_FLUX, GRND_FLUX_TYPE, FS_FROZEN, IMPLICIT;
  double elevation, b_infilt, depth0, resid_moist0, veg_LAI, ws_under, ra_under, zref_under, disp_under, rough_under;
  // scalars captured by value
  int VEG, UnderStory, overstory, INCLUDE_SNOW, NOFLUX, EXP_TRANS, SNOWING, Nnodes;
  double delta_t, Cs1, Cs2, D1, D2, T1_old, T2, Ts_old, bubble, dp, expt, ice0, kappa1, kappa2, max_moist, moist;
  double NetShortBare, NetShortGrnd, NetShortSnow, Tair, atmos_density, atmos_pressure, emissivity, LongBareIn, LongSnowIn;
  double surf_atten, vp, vpd, Wdew, rainfall, latent_heat_Le, Advection, OldTSurf, TPack, Tsnow_surf, kappa_snow, melt_energy;
  double snow_coverage, snow_density, snow_swq, snow_water;
  const Surf4 *displacement, *aero_resist, *ref_height, *roughness, *wind_speed;
  // in/out, BY VALUE: what an evaluation writes of the HRU's energy record, snow pack and the resistances in use.  The caller copies
  // them in before the solve and back after the last evaluation (calc_surf_energy_bal); written through pointers into the HRU they
  // were ~20 thread-local stores and reloads per evaluation, and every such store made the compiler reload whatever else it held.
  struct EnOut { double snow_flux, grnd_flux, deltaH, fusion, deltaCC, refreeze_energy, latent, latent_sub, sensible, error; } en;
  struct SnOut { double vapor_flux, blowing_flux, surface_flux; } sn;
  RaUsed aero_resist_used;
  double *Cs_node, *T_node, *Tnew_node, *Tnew_fbflag, *Tnew_fbcount, *ice_node, *kappa_node, *moist_node;
  SoilLayer* layer;
  VegVar* vv;
  int* FIRST_SOLN;
  double NetLongBare, NetLongSnow, T1;  // results the caller reads back after the solve
  // sub-expressions of the residual that do not depend on the trial temperature, evaluated once per solve (same operations,
  // same order as func_surf_energy_bal.c / estimate_T1.c evaluate them at every trial)
  double t1_k1, t1_b, t1_c, t1_den, gf_k1, gf_k2e, sc_lg;
  int sc_lg_ok;
  EvapMemo* memo;  // the caller's (the functor itself must not have its address taken, or its members stay in memory)
  double Tsnow_surf_final;  // what Tsnow_surf restarts from at the evaluation after the solve
  // QUICK_SOLVE (calc_surf_energy_bal.c:289-314, 400-475, 484-486): the search runs on the nodes above the thaw depth only, a second search
  // on the whole profile follows when the surface changes sign, and the evaluation at the accepted temperature always uses the whole profile
  int qs_pass, Nnodes_full, NOFLUX_full;  // qs_pass: 0 = option off, 1 = reduced-node search running, 2 = over
  VIC_HD void before_final() {
    Tsnow_surf = Tsnow_surf_final;
    if (qs_pass) {
      Nnodes = Nnodes_full;
      FIRST_SOLN[0] = 1;
    }
  }
  VIC_HD bool resolve(double Tsurf) {
    if (qs_pass != 1) return false;
    qs_pass = 2;
    if (Ts_old * Tsurf < 0) {
      Nnodes = Nnodes_full;
      NOFLUX = NOFLUX_full;
      FIRST_SOLN[0] = 1;
      Tsnow_surf = Tsnow_surf_final;  // (a fresh functor in the reference)
      return true;
    }
    return false;
  }

  VIC_HD void prepare() {
    memo->reset();
    sc_lg_ok = 0;
    sc_lg = 0;
    t1_k1 = t1_b = t1_c = t1_den = gf_k1 = gf_k2e = 0;
    QUICK_FLUX = o->QUICK_FLUX;
    IMPLICIT = o->IMPLICIT;
    GRND_FLUX_TYPE = o->GRND_FLUX_TYPE;
    FS_FROZEN = (((*cp)(CP_FS_ACTIVE) != 0.0) && o->FROZEN_SOIL) ? 1 : 0;
    elevation = (*cp)(CP_elevation);
    b_infilt = (*cp)(CP_b_infilt);
    depth0 = cp->layer(CL_depth, 0);
    resid_moist0 = cp->layer(CL_resid_moist, 0);
    veg_LAI = veg->LAI;
    ws_under = (*wind_speed)[UnderStory];
    ra_under = (*aero_resist)[UnderStory];
    zref_under = (*ref_height)[UnderStory];
    disp_under = (*displacement)[UnderStory];
    rough_under = (*roughness)[UnderStory];
    if (o->QUICK_FLUX) {
      // estimate_T1.c:8-47 with Ts factored out
      const double e_mD1 = vexp(-D1 / dp);
      const double C1 = Cs2 * dp / D2 * (1. - vexp(-D2 / dp));
      const double C2 = -(1. - vexp(D1 / dp)) * vexp(-D2 / dp);
      const double C3 = kappa1 / D1 - kappa2 / D1 + kappa2 / D1 * e_mD1;
      t1_k1 = kappa1 / 2. / D1 / D2;
      t1_b = C1 / delta_t * T1_old;
      t1_c = (2. * C2 - 1. + e_mD1) * kappa2 / 2. / D1 / D2 * T2;
      t1_den = (C1 / delta_t + kappa2 / D1 / D2 * C2 + C3 / 2. / D2);
      gf_k1 = kappa1 / D1;
      gf_k2e = kappa2 / D2 * (1. - e_mD1);
    }
  }

  VIC_HDI double operator()(double Ts) { return eval(Ts); }
  // the residual proper, inlined into the single call site of root_brent_ss_impl (calc_surf_energy_bal)
  VIC_HD double eval(double Ts) {
    const double TMean = Ts;
    const double Tmp = TMean + KELVIN;
    if (snow_coverage > 0 && !INCLUDE_SNOW) en.snow_flux = (kappa_snow * (Tsnow_surf - TMean));
    else if (INCLUDE_SNOW) {
      en.snow_flux = 0;
      Tsnow_surf = TMean;
    } else en.snow_flux = 0;
    const double cover = (snow_coverage + (1. - snow_coverage) * surf_atten);
    if (QUICK_FLUX) {
      T1 = (t1_k1 * (TMean) + t1_b + t1_c) / t1_den;
      if (GRND_FLUX_TYPE == GF_406) en.grnd_flux = cover * (gf_k1 * (T1 - TMean));
      else en.grnd_flux = cover * (gf_k1 * (T1 - TMean) + (gf_k2e * (T2 - T1))) / 2.;
    } else {
      T_node[0] = TMean;
      // IMPLICIT: Newton-Raphson on the whole profile first; the explicit sweeps are its fallback (func_surf_energy_bal.c:192-221)
      int Error = 0;
      // (not compiled into the three-node kernel, which exists for the QUICK_FLUX configurations: vic_node_width() sends an IMPLICIT
      // configuration with three nodes to the ten-node instantiation)
      if constexpr (NN > 3) {
        if (IMPLICIT) Error = solve_T_profile_implicit<NN>(Tnew_node, T_node, kappa_node, Cs_node, moist_node, delta_t, ice_node, dp, Nnodes, FIRST_SOLN,
                                                           NOFLUX, EXP_TRANS, *cp);
      }
      if (!IMPLICIT || Error == 1) {
        if (IMPLICIT) FIRST_SOLN[0] = 1;
        Error = solve_T_profile<NN>(Tnew_node, T_node, Tnew_fbflag, Tnew_fbcount, kappa_node, Cs_node, moist_node, delta_t, ice_node, dp,
                                    Nnodes, FIRST_SOLN, NOFLUX, EXP_TRANS, *cp, *o);
      }
      if (Error == ERROR_I) return ERROR_D;
      T1 = Tnew_node[1];
      if (GRND_FLUX_TYPE == GF_406) en.grnd_flux = cover * (kappa1 / D1 * (T1 - TMean));
      else en.grnd_flux = cover * (kappa1 / D1 * (T1 - TMean) + (kappa2 / D2 * (Tnew_node[2] - T1))) / 2.;
    }
    if (GRND_FLUX_TYPE == GF_FULL) en.deltaH = cover * (div_pos(Cs1 * ((Ts_old + T1_old) - (TMean + T1)) * D1, delta_t) / 2.);
    else en.deltaH = (div_pos(Cs1 * ((Ts_old + T1_old) - (TMean + T1)) * D1, delta_t) / 2.);
    if (FS_FROZEN) {
      double ice;
      if ((TMean + T1) / 2. < 0.) {
        ice = moist - maximum_unfrozen_water((TMean + T1) / 2., max_moist, bubble, expt);
        if (ice < 0.) ice = 0.;
      } else ice = 0.;
      if (GRND_FLUX_TYPE == GF_FULL) en.fusion = cover * (-ice_density * Lf * (ice0 - ice) * D1 / delta_t);
      else en.fusion = (-ice_density * Lf * (ice0 - ice) * D1 / delta_t);
    }
    if (INCLUDE_SNOW) {
      if (TMean > 0) en.deltaCC = div_pos(CH_ICE * (snow_swq - snow_water) * (0 - OldTSurf), delta_t);
      else en.deltaCC = div_pos(CH_ICE * (snow_swq - snow_water) * (TMean - OldTSurf), delta_t);
      en.refreeze_energy = div_pos((snow_water * Lf * snow_density), delta_t);
      en.deltaCC *= snow_coverage;
      en.refreeze_energy *= snow_coverage;
    }
    const double LongBareOut = STEFAN_B * Tmp * Tmp * Tmp * Tmp;
    if (INCLUDE_SNOW) NetLongSnow = (LongSnowIn - snow_coverage * LongBareOut);
    NetLongBare = (LongBareIn - (1. - snow_coverage) * LongBareOut);
    const double NetBareRad = (NetShortBare + NetLongBare + en.grnd_flux + en.deltaH + en.fusion);
    const double ws = ws_under;
    if (ws > 0.0) {
      // the displacement height is dropped under a snowing overstory (func_surf_energy_bal.c:280-296); which case applies is fixed for the solve
      const double Zr = zref_under, dr = (overstory && SNOWING) ? 0. : disp_under;
      if (!sc_lg_ok && TMean != Tair) {
        sc_lg = vlog((Zr - dr) / rough_under);
        sc_lg_ok = 1;
      }
      aero_resist_used.surface = ra_under / stability_correction_lg(Zr, dr, TMean, Tair, ws, sc_lg);
    } else aero_resist_used.surface = HUGE_RESIST;
    double Evap;
    if (VEG && !SNOWING && veg_LAI > 0) {
      Evap = canopy_evap(layer, *vv, true, *veg, Wdew, delta_t, NetBareRad, vpd, NetShortBare, Tair, aero_resist_used.overstory,
                         elevation, rainfall, *soil, memo);
    } else if (!SNOWING) {
      Evap = arno_evap(layer, NetBareRad, Tair, vpd, depth0, max_moist * depth0 * 1000., elevation,
                       b_infilt, aero_resist_used.surface, delta_t, resid_moist0, memo);
    } else Evap = 0.;
    en.latent = -RHO_W * latent_heat_Le * Evap;
    en.latent_sub = 0.;
    if (INCLUDE_SNOW) {
      double VaporMassFlux = div_pos(sn.vapor_flux * ice_density, delta_t);
      double BlowingMassFlux = div_pos(sn.blowing_flux * ice_density, delta_t);
      double SurfaceMassFlux = div_pos(sn.surface_flux * ice_density, delta_t);
      double tl, tls;
      latent_heat_from_snow(atmos_density, vp, latent_heat_Le, atmos_pressure, aero_resist_used.surface, TMean, vpd, &tl, &tls, &VaporMassFlux,
                            &BlowingMassFlux, &SurfaceMassFlux);
      en.latent += tl * snow_coverage;
      en.latent_sub = tls * snow_coverage;
      sn.vapor_flux = div_pos(VaporMassFlux * delta_t, ice_density);
      sn.blowing_flux = div_pos(BlowingMassFlux * delta_t, ice_density);
      sn.surface_flux = div_pos(SurfaceMassFlux * delta_t, ice_density);
    } else en.latent *= (1. - snow_coverage);
    if (snow_coverage < 1 || INCLUDE_SNOW) {
      en.sensible = atmos_density * Cp * (Tair - (TMean)) / aero_resist_used.surface;
      if (!INCLUDE_SNOW) en.sensible *= (1. - snow_coverage);
    } else en.sensible = 0.;
    double error = (NetBareRad + NetShortGrnd + NetShortSnow + emissivity * NetLongSnow) + en.sensible + (en.latent + en.latent_sub) +
                   en.snow_flux * snow_coverage + melt_energy + Advection - en.deltaCC;
    if (INCLUDE_SNOW) {
      if (Tsnow_surf == 0.0 && error > -(en.refreeze_energy)) {
        en.refreeze_energy = -error;
        error = 0.0;
      } else error += en.refreeze_energy;
    }
    en.error = error;
    return error;
  }
};

// calc_surf_energy_bal.c:7-692.  Returns Tsurf or ERROR_D.
template <int NN>
VIC_HDI double calc_surf_energy_bal(double latent_heat_Le, double LongUnderIn, double NetLongSnow, double NetShortGrnd, double NetShortSnow,
                                    double OldTSurf, double ShortUnderIn, double SnowAlbedo, double SnowLatent, double SnowLatentSub,
                                    double SnowSensible, double Tair, double VPDcanopy, double VPcanopy, double coldcontent,
                                    double delta_coverage, double dp, double ice0, double melt_energy, double moist, double snow_coverage,
                                    double snow_depth, double BareAlbedo, double surf_atten, const Surf4& aero_resist, RaUsed& aero_resist_used,
                                    const Surf4& displacement, double* melt, double* ppt, double rainfall, const Surf4& ref_height,
                                    const Surf4& roughness, const Surf4& wind_speed, int INCLUDE_SNOW, int UnderStory, int dt, int overstory,
                                    bool isArtificialBareSoil, double atmos_density, double atmos_pressure, EnergyBal<NN>& energy, SoilLayer* layer,
                                    SnowPack& snow, VegVar& vv, const VegNow& veg, const SoilET& soil, const CellPar& cp, const Opts& o) {
  (void)coldcontent;
  const int Nnodes = o.Nnode;
  int FIRST_SOLN[2] = {1, 1};
  double Tnew_node[NN + 1], Tnew_fbflag[NN], Tnew_fbcount[NN];  // (+1: the implicit scheme's cold-nose test reads one element past the unknowns)
  double Tsurf_fbflag = 0, Tsurf_fbcount = 0;
  for (int n = 0; n < NN; n++) { Tnew_fbflag[n] = 0; Tnew_fbcount[n] = 0; Tnew_node[n] = 0; }
  Tnew_node[NN] = 0;
  int VEG;
  if (!isArtificialBareSoil) VEG = (veg.LAI > 0.0) ? 1 : 0;
  else VEG = 0;
  const double T2 = energy.T[Nnodes - 1];
  const double Ts_old = energy.T[0];
  const double T1_old = energy.T[1];
  const double delta_t = (double)dt * 3600.;
  const double max_moist = cp.layer(CL_max_moist, 0) / (cp.layer(CL_depth, 0) * 1000.);
  double kappa_snow;
  if (snow.depth > 0.) kappa_snow = K_SNOW * (snow.density) * (snow.density) / snow_depth;
  else kappa_snow = 0;
  const double NetShortBare = (ShortUnderIn * (1. - (snow_coverage + delta_coverage)) * (1. - BareAlbedo) + ShortUnderIn * (delta_coverage) * (1. - SnowAlbedo));
  const double LongBareIn = (1. - snow_coverage) * LongUnderIn;
  double TmpNetLongSnow, TmpNetShortSnow, LongSnowIn;
  if (INCLUDE_SNOW || snow.swq == 0) {
    TmpNetLongSnow = NetLongSnow;
    TmpNetShortSnow = NetShortSnow;
    LongSnowIn = snow_coverage * LongUnderIn;
  } else {
    TmpNetShortSnow = 0.;
    TmpNetLongSnow = 0.;
    LongSnowIn = 0.;
  }
  double Tsurf;

  SurfEB<NN> eb;
  eb.o = &o; eb.cp = &cp; eb.veg = &veg; eb.soil = &soil;
  eb.VEG = VEG; eb.UnderStory = UnderStory; eb.overstory = overstory; eb.INCLUDE_SNOW = INCLUDE_SNOW; eb.SNOWING = (snow.snow != 0.0);
  eb.Nnodes = Nnodes; eb.NOFLUX = o.NOFLUX; eb.EXP_TRANS = o.EXP_TRANS;
  eb.delta_t = delta_t; eb.Cs1 = energy.Cs0; eb.Cs2 = energy.Cs1; eb.D1 = cp.node(CN_Zsum_node, 1) - cp.node(CN_Zsum_node, 0);
  eb.D2 = cp.node(CN_Zsum_node, 2) - cp.node(CN_Zsum_node, 1); eb.T1_old = T1_old; eb.T2 = T2; eb.Ts_old = Ts_old;
  eb.bubble = cp.layer(CL_bubble, 0); eb.dp = dp; eb.expt = cp.layer(CL_expt, 0); eb.ice0 = ice0; eb.kappa1 = energy.kappa0; eb.kappa2 = energy.kappa1;
  eb.max_moist = max_moist; eb.moist = moist; eb.NetShortBare = NetShortBare; eb.NetShortGrnd = NetShortGrnd; eb.NetShortSnow = TmpNetShortSnow;
  eb.Tair = Tair; eb.atmos_density = atmos_density; eb.atmos_pressure = atmos_pressure; eb.emissivity = 1.; eb.LongBareIn = LongBareIn;
  eb.LongSnowIn = LongSnowIn; eb.surf_atten = surf_atten; eb.vp = VPcanopy; eb.vpd = VPDcanopy; eb.Wdew = vv.Wdew; eb.rainfall = rainfall;
  eb.latent_heat_Le = latent_heat_Le; eb.Advection = energy.advection; eb.OldTSurf = OldTSurf; eb.TPack = snow.pack_temp; eb.Tsnow_surf = snow.surf_temp;
  eb.kappa_snow = kappa_snow; eb.melt_energy = melt_energy; eb.snow_coverage = snow_coverage; eb.snow_density = snow.density; eb.snow_swq = snow.swq;
  eb.snow_water = snow.surf_water;
  eb.displacement = &displacement; eb.aero_resist = &aero_resist; eb.ref_height = &ref_height; eb.roughness = &roughness; eb.wind_speed = &wind_speed;
  eb.aero_resist_used = aero_resist_used;
  eb.en.snow_flux = energy.snow_flux; eb.en.grnd_flux = energy.grnd_flux; eb.en.deltaH = energy.deltaH; eb.en.fusion = energy.fusion;
  eb.en.deltaCC = energy.deltaCC; eb.en.refreeze_energy = energy.refreeze_energy; eb.en.latent = energy.latent; eb.en.latent_sub = energy.latent_sub;
  eb.en.sensible = energy.sensible; eb.en.error = energy.error;
  eb.sn.vapor_flux = snow.vapor_flux; eb.sn.blowing_flux = snow.blowing_flux; eb.sn.surface_flux = snow.surface_flux;
  eb.Cs_node = energy.Cs_node; eb.T_node = energy.T; eb.Tnew_node = Tnew_node; eb.Tnew_fbflag = Tnew_fbflag; eb.Tnew_fbcount = Tnew_fbcount;
  eb.ice_node = energy.ice; eb.kappa_node = energy.kappa_node; eb.moist_node = energy.moist;
  eb.layer = layer; eb.vv = &vv; eb.FIRST_SOLN = FIRST_SOLN;
  eb.NetLongBare = 0; eb.NetLongSnow = TmpNetLongSnow; eb.T1 = 0;
  EvapMemo memo;
  eb.memo = &memo;
  eb.Tsnow_surf_final = snow.surf_temp;
  eb.prepare();

  // The solve (FULL_ENERGY) and the evaluation at the accepted temperature -- in the reference a fresh functor, i.e. Tsnow_surf restarts
  // from snow.surf_temp -- go through the one residual call site of root_brent_ss_impl, so the residual is inlined here and the
  // solve's constants live in registers.  QUICK_SOLVE's second search (SurfEB::resolve) goes through the same call site.
  double T_lower, T_upper;
  if (INCLUDE_SNOW) {
    T_lower = energy.T[0] - SURF_DT;
    T_upper = 0.;
  } else {
    T_lower = 0.5 * (energy.T[0] + Tair) - SURF_DT;
    T_upper = 0.5 * (energy.T[0] + Tair) + SURF_DT;
  }
  struct Inlined {
    SurfEB<NN>& f;
    VIC_HD double operator()(double x) { return f.eval(x); }
    VIC_HD void before_final() { f.before_final(); }
    VIC_HD bool resolve(double x) { return f.resolve(x); }
  } call{eb};
  eb.qs_pass = 0; eb.Nnodes_full = Nnodes; eb.NOFLUX_full = o.NOFLUX;
  if constexpr (NN > 3) {
    if (o.QUICK_SOLVE && !o.QUICK_FLUX) {
      if (o.FULL_ENERGY) {
        // iterate on the nodes down to four below the thaw front (or three, or all: calc_surf_energy_bal.c:289-300), open bottom, linear grid
        int tmpNnodes = 0;
        for (int nidx = Nnodes - 5; nidx >= 0; nidx--)
          if (energy.T[nidx] >= 0 && energy.T[nidx + 1] < 0) tmpNnodes = nidx + 1;
        if (tmpNnodes == 0) {
          if (energy.T[0] <= 0 && energy.T[1] >= 0) tmpNnodes = Nnodes;
          else tmpNnodes = 3;
        } else tmpNnodes += 4;
        eb.Nnodes = tmpNnodes;
        eb.NOFLUX = 0;
        eb.EXP_TRANS = 0;
      }
      eb.qs_pass = 1;
    }
  }
  BrentFinal fin;
  fin.do_solve = o.FULL_ENERGY != 0;
  fin.allow_fallback = o.TFALLBACK != 0;
  fin.fallback_x = Ts_old;
  fin.nosolve_x = Tair;
  fin.f_final = 0.;
  fin.fell_back = 0;
  Tsurf = root_brent_ss_impl<true, false, (NN > 3)>(T_lower, T_upper, call, &fin);
  // what the last evaluation left behind
  aero_resist_used = eb.aero_resist_used;
  energy.snow_flux = eb.en.snow_flux; energy.grnd_flux = eb.en.grnd_flux; energy.deltaH = eb.en.deltaH; energy.fusion = eb.en.fusion;
  energy.deltaCC = eb.en.deltaCC; energy.refreeze_energy = eb.en.refreeze_energy; energy.latent = eb.en.latent; energy.latent_sub = eb.en.latent_sub;
  energy.sensible = eb.en.sensible; energy.error = eb.en.error;
  snow.vapor_flux = eb.sn.vapor_flux; snow.blowing_flux = eb.sn.blowing_flux; snow.surface_flux = eb.sn.surface_flux;
  if (o.FULL_ENERGY && !fin.fell_back && result_is_error(Tsurf)) return ERROR_D;  // the solve failed and TFALLBACK is off
  if (fin.fell_back) {
    Tsurf_fbflag = 1;
    Tsurf_fbcount += fin.fell_back;
  }
  const double error = fin.f_final;
  if (error == ERROR_D) return ERROR_D;
  energy.error = error;
  if (o.QUICK_FLUX || !(o.FULL_ENERGY || (o.FROZEN_SOIL && (cp(CP_FS_ACTIVE) != 0.0)))) {
    Tnew_node[0] = Tsurf;
    Tnew_node[1] = eb.T1;
    Tnew_node[2] = T2;
  }
  if (calc_layer_average_thermal_props<NN>(energy, layer, cp, Nnodes, Tnew_node, o) == ERROR_I) return ERROR_D;
  if (!(snow.snow != 0.0) && !INCLUDE_SNOW) {
    if (!isArtificialBareSoil) {
      if (veg.LAI <= 0.0) {
        vv.throughfall = rainfall;
        ppt[0] = vv.throughfall;
      } else ppt[0] = vv.throughfall;
    } else ppt[0] = rainfall;
  }
  energy.NetShortGrnd = NetShortGrnd;
  if (INCLUDE_SNOW) {
    energy.NetLongUnder = eb.NetLongBare + eb.NetLongSnow;
    energy.NetShortUnder = NetShortBare + TmpNetShortSnow + NetShortGrnd;
  } else {
    energy.NetLongUnder = eb.NetLongBare + NetLongSnow;
    energy.NetShortUnder = NetShortBare + NetShortSnow + NetShortGrnd;
    energy.latent = (SnowLatent + energy.latent);
    energy.latent_sub = (SnowLatentSub + energy.latent_sub);
    energy.sensible = (SnowSensible + energy.sensible);
  }
  energy.LongUnderOut = LongUnderIn - energy.NetLongUnder;
  energy.AlbedoUnder = ((1. - (snow_coverage + delta_coverage)) * BareAlbedo + (snow_coverage + delta_coverage) * SnowAlbedo);
  energy.melt_energy = melt_energy;
  energy.Tsurf = (snow.coverage * snow.surf_temp + (1. - snow.coverage) * Tsurf);
  if (INCLUDE_SNOW) {
    // thin pack solved with the ground surface: update its mass here
    if (-(snow.vapor_flux) > snow.swq) {
      snow.blowing_flux *= -(snow.swq / snow.vapor_flux);
      snow.vapor_flux = -(snow.swq);
      snow.surface_flux = snow.vapor_flux - snow.blowing_flux;
    }
    snow.swq += snow.vapor_flux;
    snow.surf_water += snow.vapor_flux;
    snow.surf_water = (snow.surf_water < 0) ? 0. : snow.surf_water;
    if (energy.refreeze_energy >= 0.0) {
      double refrozen_water = energy.refreeze_energy / (Lf * RHO_W) * delta_t;
      if (refrozen_water > snow.surf_water) {
        refrozen_water = snow.surf_water;
        energy.refreeze_energy = refrozen_water * Lf * RHO_W / delta_t;
      }
      snow.surf_water -= refrozen_water;
      if (snow.surf_water < 0.0) snow.surf_water = 0.0;
      (*melt) = 0.0;
    } else {
      (*melt) = fabs(energy.refreeze_energy) / (Lf * RHO_W) * delta_t;
      snow.swq -= *melt;
      if (snow.swq < 0) {
        (*melt) += snow.swq;
        snow.swq = 0;
      }
    }
    if (snow.swq > 0) {
      snow.surf_temp = (Tsurf > 0) ? 0 : Tsurf;
      snow.coldcontent = CH_ICE * snow.surf_temp * snow.swq;
      snow.depth = 1000. * snow.swq / snow.density;
      if (snow.swq > 0) snow.coverage = 1.;
      else snow.coverage = 0.;
      if (is_invalid(snow.surf_temp) || snow.surf_temp > 0) energy.snow_flux = (energy.grnd_flux + energy.deltaH + energy.fusion);
    } else {
      snow.density = 0.;
      snow.depth = 0.;
      snow.surf_water = 0;
      snow.pack_water = 0;
      snow.surf_temp = 0;
      snow.pack_temp = 0;
      snow.coverage = 0;
    }
    snow.vapor_flux *= -1;
  }
  energy.Tsurf_fbflag = Tsurf_fbflag;
  energy.Tsurf_fbcount += Tsurf_fbcount;
  for (int n = 0; n < NN; n++) {
    if (n < Nnodes) {
      energy.T_fbflag[n] = Tnew_fbflag[n];
      energy.T_fbcount[n] += Tnew_fbcount[n];
    }
  }
  return Tsurf;
}

// everything the sub-step routines share about "the air above this HRU"
struct AeroState {
  Surf4 aero_resist[N_PET_TYPES + 1];
  Surf4 displacement, ref_height, roughness, wind_speed;
};

struct SolveSnowOut {
  double LongUnderIn, NetLongSnow, NetShortGrnd, NetShortSnow, ShortUnderIn, Torg_snow, coverage, delta_coverage, melt_energy;
  double out_prec, out_rain, out_snow, ppt, rainfall, snowfall;
};

// solve_snow.c:7-544.  Returns melt [mm] or ERROR_D.  AlbedoUnder is the HRU's own record
// (surface_fluxes.c:537 passes &energy->AlbedoUnder), energy the sub-step's snow-side record.
template <int NN, class EN>
VIC_HDI double solve_snow(bool overstory, double BareAlbedo, double LongUnderOut, double Tcanopy, double Tgrnd, double air_temp, double mu,
                          double prec, double snow_grnd_flux, double* AlbedoUnder, double* latent_heat_Le, const Surf4& aero_resist,
                          RaUsed& aero_resist_used, const AeroState& as, const double* gauge_correction, double* snow_inflow, double* surf_atten,
                          bool UNSTABLE_SNOW, int dt, int hidx, bool isArtificialBareSoil, int* UnderStory, const Ctx& cx, EN& energy,
                          SoilLayer* layer, SnowPack& snow, VegVar& vv, const VegNow& veg, const SoilET& soil, SolveSnowOut& r) {
  const Opts& o = *cx.o;
  const CellPar& cp = cx.cp;
  const Forcing& f = cx.f;
  double melt = 0.;
  r.ppt = 0.;
  r.melt_energy = 0.;
  const double rainonly = calc_rainonly(air_temp, prec, cp(CP_MAX_SNOW_TEMP), cp(CP_MIN_RAIN_TEMP), mu, o.TEMP_TH_TYPE);
  r.snowfall = gauge_correction[1] * (prec - rainonly) * cp(CP_PADJ_S);
  r.rainfall = gauge_correction[0] * rainonly * cp(CP_PADJ_R);
  r.out_prec = r.snowfall + r.rainfall;
  r.out_rain = r.rainfall;
  r.out_snow = r.snowfall;
  const double store_snowfall = r.snowfall;
  (*latent_heat_Le) = (2.501e6 - 0.002361e6 * air_temp);
  if ((snow.swq > 0 || r.snowfall > 0. || (snow.snow_canopy > 0. && overstory))) {
    if (mu != 1 && o.FULL_ENERGY) return ERROR_D;
  }
  if (*UnderStory == SURF_UNSET) {
    if (snow.swq > 0 || r.snowfall > 0) *UnderStory = SNOW_COVERED;
    else *UnderStory = SNOW_FREE;
  }
  r.ShortUnderIn = f(FV_shortwave, hidx);
  r.LongUnderIn = f(FV_longwave, hidx);
  if ((snow.swq > 0 || r.snowfall > 0. || (snow.snow_canopy > 0. && overstory)) && mu == 1) {
    snow.snow = 1.0;
    if (!overstory) (*surf_atten) = 1.;
    const double old_coverage = snow.coverage;
    if (!isArtificialBareSoil) {
      if (overstory) {
        r.ShortUnderIn *= (*surf_atten);
        const double ShortOverIn = (1. - (*surf_atten)) * f(FV_shortwave, hidx);
        int e = snow_intercept<NN, EN>((double)dt * SECPHOUR, veg.LAI, (*latent_heat_Le), f(FV_longwave, hidx), LongUnderOut, veg.Wdmax, ShortOverIn,
                                   Tcanopy, BareAlbedo, energy, snow, vv, &r.LongUnderIn, aero_resist, aero_resist_used, &r.rainfall, &r.snowfall,
                                   as.wind_speed, as.displacement, as.ref_height, as.roughness, veg, soil, layer, f(FV_density, hidx),
                                   f(FV_vp, hidx), f(FV_pressure, hidx), f(FV_vpd, hidx), cp, o);
        if (e == ERROR_I) return ERROR_D;
        vv.throughfall = r.rainfall + r.snowfall;
        energy.LongOverIn = f(FV_longwave, hidx);
      } else if (r.snowfall > 0. && vv.Wdew > 0.) {
        // snow falling on a short canopy: the interception store drops to the ground
        r.rainfall += vv.Wdew;
        vv.throughfall = r.rainfall + r.snowfall;
        vv.Wdew = 0.;
        energy.NetLongOver = 0;
        energy.LongOverIn = 0;
        energy.Tfoliage = air_temp;
        energy.Tfoliage_fbflag = 0;
      } else {
        vv.throughfall = r.rainfall + r.snowfall;
        energy.NetLongOver = 0;
        energy.LongOverIn = 0;
        energy.Tfoliage = air_temp;
        energy.Tfoliage_fbflag = 0;
      }
    } else {
      energy.NetLongOver = 0;
      energy.LongOverIn = 0;
    }
    if (snow.swq > 0.0 || r.snowfall > 0) {
      r.NetShortGrnd = 0.;
      (*snow_inflow) += r.rainfall + r.snowfall;
      const double old_swq = snow.swq;
      *UnderStory = SNOW_COVERED;
      if (snow.swq > 0 && store_snowfall == 0) {
        // age the albedo
        snow.last_snow += 1;
        snow.albedo = snow_albedo(r.snowfall, snow.swq, snow.depth, snow.albedo, snow.coldcontent, (double)dt, (int)snow.last_snow,
                                  snow.MELTING != 0.0, cp, o.SNOW_ALBEDO);
        (*AlbedoUnder) = (r.coverage * snow.albedo + (1. - r.coverage) * BareAlbedo);
      } else {
        snow.last_snow = 0;
        snow.albedo = cp(CP_NEW_SNOW_ALB);
        (*AlbedoUnder) = snow.albedo;
      }
      r.NetShortSnow = (1.0 - *AlbedoUnder) * r.ShortUnderIn;
      SnowMeltOut m;
      m.NetLongSnow = r.NetLongSnow;
      int e = snow_melt((*latent_heat_Le), r.NetShortSnow, Tcanopy, Tgrnd, as.roughness[SNOW_COVERED], aero_resist[*UnderStory], aero_resist_used,
                        air_temp, (double)dt * SECPHOUR, f(FV_density, hidx), snow_grnd_flux, r.LongUnderIn, f(FV_pressure, hidx), r.rainfall,
                        r.snowfall, f(FV_vp, hidx), f(FV_vpd, hidx), as.wind_speed[*UnderStory], as.ref_height[*UnderStory], UNSTABLE_SNOW, snow, o, m);
      if (e == ERROR_I) return ERROR_D;
      r.NetLongSnow = m.NetLongSnow;
      r.Torg_snow = m.OldTSurf;
      melt = m.melt;
      energy.error = m.Qnet;
      energy.advected_sensible = m.advected_sensible;
      energy.advection = m.advection;
      energy.deltaCC = m.deltaCC;
      energy.latent = m.latent;
      energy.latent_sub = m.latent_sub;
      energy.refreeze_energy = m.refreeze_energy;
      energy.sensible = m.sensible;
      r.ppt += melt;
      energy.AlbedoUnder = *AlbedoUnder;
      if (snow.swq > 0.) {
        if (is_valid(snow.surf_temp) && snow.surf_temp <= 0)
          snow.density = snow_density(snow, r.snowfall, old_swq, Tgrnd, air_temp, (double)dt, o.SNOW_DENSITY);
        else if (snow.last_snow == 0) snow.density = new_snow_density(air_temp, o.SNOW_DENSITY);
        snow.depth = 1000. * snow.swq / snow.density;
        const int diy = cx.dmy.day_in_year;
        const double lat = cp(CP_lat);
        if (snow.coldcontent >= 0 && ((lat >= 0 && (diy > 60 && diy < 273)) || (lat < 0 && (diy < 60 || diy > 273)))) snow.MELTING = 1.0;
        else if ((snow.MELTING != 0.0) && r.snowfall > TraceSnow) snow.MELTING = 0.0;
        if (snow.swq > 0) snow.coverage = 1.;
        else snow.coverage = 0.;
      } else snow.coverage = 0.;
      r.delta_coverage = old_coverage - snow.coverage;
      if (r.delta_coverage != 0) {
        if (old_coverage > snow.coverage) {
          // melt of the whole pack within the sub-step
          r.coverage = (old_coverage);
          (*AlbedoUnder) = (r.coverage - snow.coverage) / (1. - snow.coverage) * snow.albedo;
          (*AlbedoUnder) += (1. - r.coverage) / (1. - snow.coverage) * BareAlbedo;
          r.melt_energy = (r.delta_coverage) * (energy.advection - energy.deltaCC + energy.latent + energy.latent_sub + energy.sensible +
                                               energy.refreeze_energy + energy.advected_sensible);
        } else {
          r.coverage = snow.coverage;
          r.delta_coverage = 0;
        }
      } else if (old_coverage == 0 && snow.coverage == 0) {
        // snow fell and melted within the sub-step
        r.delta_coverage = 1.;
        r.coverage = 0.;
        r.melt_energy = (energy.advection - energy.deltaCC + energy.latent + energy.latent_sub + energy.sensible + energy.refreeze_energy +
                         energy.advected_sensible);
      }
      r.NetLongSnow *= (snow.coverage);
      r.NetShortSnow *= (snow.coverage);
      r.NetShortGrnd *= (snow.coverage);
      energy.latent *= (snow.coverage + r.delta_coverage);
      energy.latent_sub *= (snow.coverage + r.delta_coverage);
      energy.sensible *= (snow.coverage + r.delta_coverage);
      if (snow.swq == 0) {
        snow.density = 0.;
        snow.depth = 0.;
        snow.surf_water = 0;
        snow.pack_water = 0;
        snow.surf_temp = 0;
        snow.pack_temp = 0;
        snow.coverage = 0;
        snow.swq_slope = 0;
        snow.store_snow = 1.0;
        snow.MELTING = 0.0;
      }
      r.snowfall = 0;
      r.rainfall = 0;
    } else {
      // no pack on the ground (intercepted snow only)
      r.ppt += r.rainfall;
      energy.AlbedoOver = 0.;
      (*AlbedoUnder) = BareAlbedo;
      r.NetLongSnow = 0.;
      r.NetShortSnow = 0.;
      r.NetShortGrnd = 0.;
      r.delta_coverage = 0.;
      energy.latent = 0.;
      energy.latent_sub = 0.;
      energy.sensible = 0.;
      snow.last_snow = (double)INVALID_INT;
      snow.store_swq = 0;
      snow.store_coverage = 1;
      snow.MELTING = 0.0;
    }
  } else {
    // no snow anywhere
    *UnderStory = SNOW_FREE;
    snow.snow = 0.0;
    energy.AlbedoOver = 0.;
    (*AlbedoUnder) = BareAlbedo;
    energy.NetLongOver = 0.;
    energy.LongOverIn = 0.;
    energy.NetShortOver = 0.;
    energy.ShortOverIn = 0.;
    energy.latent = 0.;
    energy.latent_sub = 0.;
    energy.sensible = 0.;
    r.NetLongSnow = 0.;
    r.NetShortSnow = 0.;
    r.NetShortGrnd = 0.;
    r.delta_coverage = 0.;
    energy.Tfoliage = Tcanopy;
    snow.store_swq = 0;
    snow.store_coverage = 1;
    snow.MELTING = 0.0;
    snow.last_snow = (double)INVALID_INT;
    snow.albedo = cp(CP_NEW_SNOW_ALB);
  }
  energy.melt_energy *= -1.;
  return melt;
}

struct SurfaceFluxOut {
  double out_prec, out_rain, out_snow, Melt, snow_inflow;
};

// surface_fluxes.c:17-956.  Returns 0 or ERROR_I.
// ONE: the model step is a single sub-step (NF == 1: every hourly configuration).  The sub-step loop then runs exactly once, so its ~60
// accumulators are `0.0 + x` at the point of use instead of values kept alive (in thread-local memory) across the solves, and the
// division by the number of sub-steps (x / 1.0 == x for every x) is dropped.  Same operations on the same values: bit-identical.
template <int NN, bool ONE>
VIC_HDI int surface_fluxes(bool overstory, double BareAlbedo, double ice0, double moist0, Hru<NN>& hru, double surf_atten, const AeroState& as,
                           const double* gauge_correction, bool isArtificialBareSoil, int band, const Ctx& cx, const VegNow& veg,
                           const SoilET& soil, int veg_class, SurfaceFluxOut& out) {
  const Opts& o = *cx.o;
  const CellPar& cp = cx.cp;
  const Forcing& f = cx.f;
  EnergyBal<NN>& energy = hru.energy;
  SnowPack& snow = hru.snow;
  SoilCol& cell = hru.cell;
  const int NL = VICGPU_NLAYER;
  int INCLUDE_SNOW = 0;
  const bool UNSTABLE_SNOW = false;  // only set inside iterations that never run (MAX_ITER 0)
  const double dp = cp(CP_dp);
  const double Tfactor = cp.band(CB_Tfactor, band), Pfactor = cp.band(CB_Pfactor, band);

  energy.advection = 0;
  energy.deltaCC = 0;
  double snow_flux;
  if (snow.swq > 0) snow_flux = energy.snow_flux;
  else snow_flux = -(energy.grnd_flux + energy.deltaH + energy.fusion);
  energy.refreeze_energy = 0;
  double coverage = snow.coverage;
  // The reference works on copies (snow-side and ground-side energy records, snow pack, layers, canopy stores) and copies
  // them back at the end of the step.  Here the ground-side energy record, the snow pack and the layers ARE the HRU's own
  // records: nothing reads the HRU's originals once the sub-steps have started except the ground temperature handed to
  // solve_snow (surface_fluxes.c:433 reads energy->T[0], which the reference only updates after the last sub-step), saved below.
  // Only the snow-side energy record, which evolves separately, is a copy.  A failed step returns ERROR_I and the caller
  // drops the whole working set, so the in-place updates never reach the stored state.
  SnowSideEnergy snow_energy;
  snow_energy.take(energy);
  EnergyBal<NN>& soil_energy = energy;
  const double Tgrnd_step = energy.T[0];
  VegVar snow_vv = hru.veg, soil_vv = hru.veg;
  SnowPack& step_snow = snow;
  SoilLayer* step_layer = cell.layer;
  for (int l = 0; l < NL; l++) step_layer[l].evap = 0;
  soil_vv.canopyevap = 0;
  snow_vv.canopyevap = 0;
  soil_vv.throughfall = 0;
  snow_vv.throughfall = 0;

  int hidx, endhidx, step_dt;
  if (snow.swq > 0 || snow.snow_canopy > 0 || f(FV_snowflag, o.NR) != 0.0) {
    hidx = 0;
    endhidx = hidx + o.NF;
    step_dt = o.SNOW_STEP;
  } else {
    hidx = o.NR;
    endhidx = hidx + 1;
    step_dt = o.dt;
  }
  // accumulators over the sub-steps
  double st_AlbedoOver = 0, st_AlbedoUnder = 0, st_AtmosLatent = 0, st_AtmosLatentSub = 0, st_AtmosSensible = 0, st_LongOverIn = 0,
         st_LongUnderIn = 0, st_LongUnderOut = 0, st_NetLongAtmos = 0, st_NetLongOver = 0, st_NetLongUnder = 0, st_NetShortAtmos = 0,
         st_NetShortGrnd = 0, st_NetShortOver = 0, st_NetShortUnder = 0, st_ShortOverIn = 0, st_ShortUnderIn = 0, st_advected_sensible = 0,
         st_advection = 0, st_canopy_advection = 0, st_canopy_latent = 0, st_canopy_latent_sub = 0, st_canopy_sensible = 0,
         st_canopy_refreeze = 0, st_deltaCC = 0, st_deltaH = 0, st_fusion = 0, st_grnd_flux = 0, st_latent = 0, st_latent_sub = 0,
         st_melt_energy = 0, st_refreeze_energy = 0, st_sensible = 0, st_snow_flux = 0;
  double last_snow_coverage = snow.coverage;
  double st_canopy_vapor_flux = 0, st_melt = 0, st_vapor_flux = 0, st_surface_flux = 0, st_blowing_flux = 0;
  double st_throughfall = 0., st_canopyevap = 0., st_layerevap[NL], st_ppt = 0;
  for (int l = 0; l < NL; l++) st_layerevap[l] = 0.;
  double step_Wdew = hru.veg.Wdew;
  RaUsed st_aero_cond_used = {0, 0};
  double st_pot_evap[N_PET_TYPES];
  for (int p = 0; p < N_PET_TYPES; p++) st_pot_evap[p] = 0;
  double snow_inflow = 0;
  out.out_prec = out.out_rain = out.out_snow = 0;
  int N_steps = 0;
  double latent_heat_Le = 0, delta_coverage = 0;

  do {
    const double Tair = f(FV_air_temp, hidx) + Tfactor;
    const double step_prec = f(FV_prec, hidx) / hru.mu * Pfactor;
    const double Tgrnd = Tgrnd_step;
    const double Tcanopy = Tair;
    const double VPcanopy = f(FV_vp, hidx);
    const double VPDcanopy = f(FV_vpd, hidx);
    // mass flux of blowing snow (surface_fluxes.c:440-453); not compiled into the three-node kernel (vic_node_width sends BLOWING configurations
    // to the ten-node instantiation)
    step_snow.blowing_flux = 0.0;
    if constexpr (NN > 3) {
      if (!overstory && o.BLOWING && step_snow.swq > 0.) {
        const double Ls = (677. - 0.07 * step_snow.surf_temp) * JOULESPCAL * GRAMSPKG;
        step_snow.blowing_flux = blow::calc_blowing_snow((double)step_dt, Tair, (int)step_snow.last_snow, step_snow.surf_water, as.wind_speed[SNOW_COVERED], Ls,
                                                         f(FV_density, hidx), f(FV_vp, hidx), as.roughness[SNOW_COVERED], step_snow.depth,
                                                         (float)cx.hp(HP_lag_one), (float)cx.hp(HP_sigma_slope), isArtificialBareSoil, (float)cx.hp(HP_fetch),
                                                         as.displacement[CANOPY_OVER], as.roughness[CANOPY_OVER], &step_snow.transport);
        if ((int)step_snow.blowing_flux == ERROR_I) return ERROR_I;
        step_snow.blowing_flux *= step_dt * SECPHOUR / RHO_W;  // m per time step
      }
    }
    int UnderStory = SURF_UNSET;
    const double snow_grnd_flux = -snow_flux;

    // values of the pre-sub-step pack the reference reads from its older copy
    const double prev_depth = step_snow.depth, prev_coldcontent = step_snow.coldcontent, prev_surf_temp = step_snow.surf_temp;
    snow_vv.Wdew = step_Wdew;
    soil_vv.Wdew = step_Wdew;
    snow_vv.canopyevap = 0;
    soil_vv.canopyevap = 0;
    for (int l = 0; l < NL; l++) step_layer[l].evap = 0;
    const Surf4& iter_aero_resist = as.aero_resist[N_PET_TYPES];
    RaUsed aero_used;
    aero_used.surface = cell.aero_surface;
    aero_used.overstory = cell.aero_overstory;
    step_snow.canopy_vapor_flux = 0;
    step_snow.vapor_flux = 0;
    step_snow.surface_flux = 0;
    const double LongUnderOut = soil_energy.LongUnderOut;

    cx.rendezvous(0, 0);
    SolveSnowOut ss;
    ss.coverage = coverage;
    ss.delta_coverage = delta_coverage;
    ss.NetLongSnow = 0; ss.NetShortGrnd = 0; ss.NetShortSnow = 0; ss.Torg_snow = 0;
    double step_melt = solve_snow<NN, SnowSideEnergy>(overstory, BareAlbedo, LongUnderOut, Tcanopy, Tgrnd, Tair, hru.mu, step_prec, snow_grnd_flux,
                                      &energy.AlbedoUnder, &latent_heat_Le, iter_aero_resist, aero_used, as, gauge_correction, &snow_inflow,
                                      &surf_atten, UNSTABLE_SNOW, step_dt, hidx, isArtificialBareSoil, &UnderStory, cx, snow_energy, step_layer,
                                      step_snow, snow_vv, veg, soil, ss);
    if (step_melt == ERROR_D) return ERROR_I;
    coverage = ss.coverage;
    delta_coverage = ss.delta_coverage;
    double step_melt_energy = ss.melt_energy;
    double step_ppt = ss.ppt;

    if ((is_invalid(step_snow.surf_temp) || UNSTABLE_SNOW) && step_snow.swq > 0) {
      INCLUDE_SNOW = UnderStory + 1;
      soil_energy.advection = snow_energy.advection;
      step_snow.surf_temp = prev_surf_temp;
      step_melt_energy = 0;
    } else INCLUDE_SNOW = 0;

    cx.rendezvous(1, 0);
    double Tsurf = calc_surf_energy_bal<NN>(latent_heat_Le, ss.LongUnderIn, ss.NetLongSnow, ss.NetShortGrnd, ss.NetShortSnow, ss.Torg_snow,
                                            ss.ShortUnderIn, step_snow.albedo, snow_energy.latent, snow_energy.latent_sub, snow_energy.sensible,
                                            Tcanopy, VPDcanopy, VPcanopy, prev_coldcontent, delta_coverage, dp, ice0, step_melt_energy, moist0,
                                            step_snow.coverage, (prev_depth + step_snow.depth) / 2., BareAlbedo, surf_atten, iter_aero_resist,
                                            aero_used, as.displacement, &step_melt, &step_ppt, ss.rainfall, as.ref_height, as.roughness,
                                            as.wind_speed, INCLUDE_SNOW, UnderStory, step_dt, (int)overstory, isArtificialBareSoil,
                                            f(FV_density, hidx), f(FV_pressure, hidx), soil_energy, step_layer, step_snow, soil_vv, veg, soil, cp, o);
    if ((int)Tsurf == ERROR_I) return ERROR_I;
    if (INCLUDE_SNOW) step_ppt += step_melt;

    // no canopy-air closure: the atmosphere sees the understory fluxes
    soil_energy.AtmosLatent = soil_energy.latent;
    soil_energy.AtmosLatentSub = soil_energy.latent_sub;
    soil_energy.AtmosSensible = soil_energy.sensible;
    soil_energy.NetLongAtmos = soil_energy.NetLongUnder;
    soil_energy.NetShortAtmos = soil_energy.NetShortUnder;
    soil_energy.Tcanopy = Tcanopy;
    snow_energy.Tcanopy = Tcanopy;

    cx.rendezvous(2, 0);
    // potential evaporation with the stability-corrected resistances
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
    double iter_pot_evap[N_PET_TYPES];
    compute_pot_evap(cx.vl, o.NVegLibTypes, veg_class, cx.dmy.month - 1, o.dt, f(FV_shortwave, hidx), soil_energy.NetLongAtmos, Tair, VPDcanopy,
                     cp(CP_elevation), step_aero, iter_pot_evap);

    // ---- accumulate the sub-step
    if (!isArtificialBareSoil) {
      if (step_snow.snow != 0.0) {
        st_throughfall += snow_vv.throughfall;
        st_canopyevap += snow_vv.canopyevap;
        soil_vv.Wdew = snow_vv.Wdew;
      } else {
        st_throughfall += soil_vv.throughfall;
        st_canopyevap += soil_vv.canopyevap;
        snow_vv.Wdew = soil_vv.Wdew;
      }
      step_Wdew = soil_vv.Wdew;
    }
    for (int l = 0; l < NL; l++) st_layerevap[l] += step_layer[l].evap;
    st_ppt += step_ppt;
    if (aero_used.surface > 0) st_aero_cond_used.surface += 1 / aero_used.surface;
    else st_aero_cond_used.surface += HUGE_RESIST;
    if (aero_used.overstory > 0) st_aero_cond_used.overstory += 1 / aero_used.overstory;
    else st_aero_cond_used.overstory += HUGE_RESIST;
    if (!isArtificialBareSoil) st_canopy_vapor_flux += step_snow.canopy_vapor_flux;
    st_melt += step_melt;
    st_vapor_flux += step_snow.vapor_flux;
    st_surface_flux += step_snow.surface_flux;
    st_blowing_flux += step_snow.blowing_flux;
    out.out_prec += ss.out_prec * hru.mu;
    out.out_rain += ss.out_rain * hru.mu;
    out.out_snow += ss.out_snow * hru.mu;
    if (INCLUDE_SNOW) {
      snow_energy.advected_sensible = soil_energy.advected_sensible;
      snow_energy.advection = soil_energy.advection;
      snow_energy.deltaCC = soil_energy.deltaCC;
      snow_energy.latent = soil_energy.latent;
      snow_energy.latent_sub = soil_energy.latent_sub;
      snow_energy.refreeze_energy = soil_energy.refreeze_energy;
      snow_energy.sensible = soil_energy.sensible;
      snow_energy.snow_flux = soil_energy.snow_flux;
    }
    st_AlbedoOver += snow_energy.AlbedoOver;
    st_AlbedoUnder += soil_energy.AlbedoUnder;
    st_AtmosLatent += soil_energy.AtmosLatent;
    st_AtmosLatentSub += soil_energy.AtmosLatentSub;
    st_AtmosSensible += soil_energy.AtmosSensible;
    st_LongOverIn += snow_energy.LongOverIn;
    st_LongUnderIn += ss.LongUnderIn;
    st_LongUnderOut += soil_energy.LongUnderOut;
    st_NetLongAtmos += soil_energy.NetLongAtmos;
    st_NetLongOver += snow_energy.NetLongOver;
    st_NetLongUnder += soil_energy.NetLongUnder;
    st_NetShortAtmos += soil_energy.NetShortAtmos;
    st_NetShortGrnd += ss.NetShortGrnd;
    st_NetShortOver += snow_energy.NetShortOver;
    st_NetShortUnder += soil_energy.NetShortUnder;
    st_ShortOverIn += snow_energy.ShortOverIn;
    st_ShortUnderIn += soil_energy.ShortUnderIn;
    st_canopy_advection += snow_energy.canopy_advection;
    st_canopy_latent += snow_energy.canopy_latent;
    st_canopy_latent_sub += snow_energy.canopy_latent_sub;
    st_canopy_sensible += snow_energy.canopy_sensible;
    st_canopy_refreeze += snow_energy.canopy_refreeze;
    st_deltaH += soil_energy.deltaH;
    st_fusion += soil_energy.fusion;
    st_grnd_flux += soil_energy.grnd_flux;
    st_latent += soil_energy.latent;
    st_latent_sub += soil_energy.latent_sub;
    st_melt_energy += step_melt_energy;
    st_sensible += soil_energy.sensible;
    if (step_snow.swq == 0 && INCLUDE_SNOW) {
      // surface_fluxes.c:795: the cast of an array address is always true
      if (last_snow_coverage == 0) last_snow_coverage = 1;
      st_advected_sensible += snow_energy.advected_sensible * last_snow_coverage;
      st_advection += snow_energy.advection * last_snow_coverage;
      st_deltaCC += snow_energy.deltaCC * last_snow_coverage;
      st_snow_flux += soil_energy.snow_flux * last_snow_coverage;
      st_refreeze_energy += snow_energy.refreeze_energy * last_snow_coverage;
    } else if ((step_snow.snow != 0.0) || INCLUDE_SNOW) {
      const double cov = (step_snow.coverage + delta_coverage);
      st_advected_sensible += snow_energy.advected_sensible * cov;
      st_advection += snow_energy.advection * cov;
      st_deltaCC += snow_energy.deltaCC * cov;
      st_snow_flux += soil_energy.snow_flux * cov;
      st_refreeze_energy += snow_energy.refreeze_energy * cov;
    }
    for (int p = 0; p < N_PET_TYPES; p++) st_pot_evap[p] += iter_pot_evap[p];
    N_steps++;
    hidx += 1;
  } while (!ONE && hidx < endhidx);

  // ---- store the step's results
  const double N = ONE ? 1.0 : (double)N_steps;
  auto mean = [&](double x) { return ONE ? x : div_pos(x, N); };
  snow.vapor_flux = st_vapor_flux;
  snow.blowing_flux = st_blowing_flux;
  snow.surface_flux = st_surface_flux;
  snow.canopy_vapor_flux = st_canopy_vapor_flux;
  out.Melt = st_melt;
  snow.melt = st_melt;
  double ppt = st_ppt;
  energy.AlbedoOver = mean(st_AlbedoOver);
  energy.AlbedoUnder = mean(st_AlbedoUnder);
  energy.AtmosLatent = mean(st_AtmosLatent);
  energy.AtmosLatentSub = mean(st_AtmosLatentSub);
  energy.AtmosSensible = mean(st_AtmosSensible);
  energy.LongOverIn = mean(st_LongOverIn);
  energy.LongUnderIn = mean(st_LongUnderIn);
  energy.LongUnderOut = mean(st_LongUnderOut);
  energy.NetLongAtmos = mean(st_NetLongAtmos);
  energy.NetLongOver = mean(st_NetLongOver);
  energy.NetLongUnder = mean(st_NetLongUnder);
  energy.NetShortAtmos = mean(st_NetShortAtmos);
  energy.NetShortGrnd = mean(st_NetShortGrnd);
  energy.NetShortOver = mean(st_NetShortOver);
  energy.NetShortUnder = mean(st_NetShortUnder);
  energy.ShortOverIn = mean(st_ShortOverIn);
  energy.ShortUnderIn = mean(st_ShortUnderIn);
  energy.advected_sensible = mean(st_advected_sensible);
  energy.canopy_advection = mean(st_canopy_advection);
  energy.canopy_latent = mean(st_canopy_latent);
  energy.canopy_latent_sub = mean(st_canopy_latent_sub);
  energy.canopy_refreeze = mean(st_canopy_refreeze);
  energy.canopy_sensible = mean(st_canopy_sensible);
  energy.deltaH = mean(st_deltaH);
  energy.fusion = mean(st_fusion);
  energy.grnd_flux = mean(st_grnd_flux);
  energy.latent = mean(st_latent);
  energy.latent_sub = mean(st_latent_sub);
  energy.melt_energy = mean(st_melt_energy);
  energy.sensible = mean(st_sensible);
  if ((snow.snow != 0.0) || INCLUDE_SNOW) {
    energy.advection = mean(st_advection);
    energy.deltaCC = mean(st_deltaCC);
    energy.refreeze_energy = mean(st_refreeze_energy);
    energy.snow_flux = mean(st_snow_flux);
  }
  energy.Tfoliage = snow_energy.Tfoliage;
  energy.Tfoliage_fbflag = snow_energy.Tfoliage_fbflag;
  energy.Tfoliage_fbcount = snow_energy.Tfoliage_fbcount;
  if (!isArtificialBareSoil) {
    hru.veg.throughfall = st_throughfall;
    hru.veg.canopyevap = st_canopyevap;
    if (snow.snow != 0.0) hru.veg.Wdew = snow_vv.Wdew;
    else hru.veg.Wdew = soil_vv.Wdew;
  }
  for (int l = 0; l < NL; l++) cell.layer[l].evap = st_layerevap[l];
  if (st_aero_cond_used.surface > 0 && st_aero_cond_used.surface < HUGE_RESIST) cell.aero_surface = 1 / (ONE ? st_aero_cond_used.surface : st_aero_cond_used.surface / N);
  else if (st_aero_cond_used.surface >= HUGE_RESIST) cell.aero_surface = 0;
  else cell.aero_surface = HUGE_RESIST;
  if (st_aero_cond_used.overstory > 0 && st_aero_cond_used.overstory < HUGE_RESIST) cell.aero_overstory = 1 / (ONE ? st_aero_cond_used.overstory : st_aero_cond_used.overstory / N);
  else if (st_aero_cond_used.overstory >= HUGE_RESIST) cell.aero_overstory = 0;
  else cell.aero_overstory = HUGE_RESIST;
  for (int p = 0; p < N_PET_TYPES; p++) cell.pot_evap[p] = ONE ? st_pot_evap[p] : st_pot_evap[p] / N;
  out.snow_inflow = snow_inflow;

  // ---- soil column
  ppt += cell.excess_moist;
  cell.excess_moist = 0.;
  cell.inflow = ppt;
  return runoff<NN>(cell, energy, cp, ppt, o);
}

}  // namespace vic
#endif

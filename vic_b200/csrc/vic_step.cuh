// vic_step.cuh -- one model step of one HRU (vegetation tile x snow band): the body of the
// HRU loop of full_energy() (full_energy.c:216-496): radiation attenuation, top-layer thermal
// properties, the seven aerodynamic-resistance sets (six potential-evaporation land covers and
// the tile's own cover, in that order because each call overwrites the shared roughness /
// displacement / reference-height / wind records), the surface-flux step and the root-zone
// moisture / wetness diagnostics.
#ifndef VIC_STEP_CUH
#define VIC_STEP_CUH
#include "vic_aero.cuh"
#include "vic_surface.cuh"
#include "vic_glacier.cuh"

namespace vic {

struct HruPar {
  double Cv;
  float root[VICGPU_NLAYER];
  int vegIndex, band;
  bool isGlacier, isArtBare;
};

VIC_HD HruPar load_hrupar(const Col& hp) {
  HruPar p;
  p.Cv = hp(HP_Cv);
  p.root[0] = (float)hp(HP_root0);
  p.root[1] = (float)hp(HP_root1);
  p.root[2] = (float)hp(HP_root2);
  p.vegIndex = (int)hp(HP_vegIndex);
  p.band = (int)hp(HP_band);
  p.isGlacier = hp(HP_isGlacier) != 0.0;
  p.isArtBare = hp(HP_isArtBare) != 0.0;
  return p;
}

// ---- aerodynamic tables ---------------------------------------------------------------------------------------------------------
// The seven CalcAerodynamic() evaluations of full_energy.c:302-354 split into the part that depends only on the HRU's land cover in
// the current month (AeroGeom: logarithms and exponentials of roughness lengths and canopy heights, ~25 vlog/vexp/vpow per HRU-step
// when recomputed every record) and the scaling with the record's wind (aero_apply).  The device evaluates aero_geom once per HRU
// and month into a table (kernel k_hru_aero, vicgpu_api.cu) and hru_step loads it; without a table (host port) hru_step calls
// aero_geom itself -- the same operations in the same order either way.
#define VIC_AERO_NCOL 53
struct AeroGeom {
  double v[VIC_AERO_NCOL];
  VIC_HD double& R0(int p, int s) { return v[p * 4 + s]; }             // aero_resist[p][s] of a unit wind
  VIC_HD double R0(int p, int s) const { return v[p * 4 + s]; }
  VIC_HD double& wind_corr(int p) { return v[28 + p]; }                 // reference-height wind / forcing wind
  VIC_HD double wind_corr(int p) const { return v[28 + p]; }
  VIC_HD double* U0() { return v + 35; }                                // the shared tables as the last evaluation leaves them
  VIC_HD double* displacement() { return v + 39; }
  VIC_HD double* ref_height() { return v + 43; }
  VIC_HD double* roughness() { return v + 47; }
  VIC_HD const double* U0() const { return v + 35; }
  VIC_HD const double* displacement() const { return v + 39; }
  VIC_HD const double* ref_height() const { return v + 43; }
  VIC_HD const double* roughness() const { return v + 47; }
  VIC_HD double& valid() { return v[51]; }                              // bit p*4+s: wind_speed[s] is valid after evaluation p
  VIC_HD double valid() const { return v[51]; }
  VIC_HD double& err() { return v[52]; }                                // != 0: an evaluation returned ERROR
  VIC_HD double err() const { return v[52]; }
};

VIC_HDI void aero_geom(const VegLib& vl, const CellPar& cp, const Opts& o, bool isGlacier, int veg_class, int month0, const VegNow& veg, AeroGeom& g) {
  (void)isGlacier;
  for (int k = 0; k < VIC_AERO_NCOL; k++) g.v[k] = 0;
  Surf4 displacement, roughness, ref_height, wind_speed, resist, prev_resist;
  displacement.set_invalid(); roughness.set_invalid(); ref_height.set_invalid(); wind_speed.set_invalid(); prev_resist.set_invalid();
  const double wind_h = veg.wind_h;
  const double soil_rough = cp(CP_rough);
  double in_prev[7] = {-1, 0, 0, 0, 0, 0, 0};
  // terms of the loop below that do not depend on the land cover: the leaf-area factor of calc_veg_height()
  // (calc_veg_params.c:26-37) and the log-profile denominator of the wind correction
  const double height_den = 1.1 * vlog(1 + vpow(0.2 * veg.LAI, 0.25));
  const double wind_den = vlog((o.wind_h - 0.) / soil_rough);
  Surf4 snap_displacement, snap_ref_height, snap_roughness, snap_wind_speed;  // the tables as the last evaluation left them (set at p == 0)
  snap_displacement.set_invalid(); snap_ref_height.set_invalid(); snap_roughness.set_invalid(); snap_wind_speed.set_invalid();
  unsigned valid = 0, prev_valid = 0;
  #pragma unroll 1
  for (int p = 0; p < N_PET_TYPES + 1; p++) {
    const int pet_class = (p < N_PET_TYPES_NON_NAT) ? o.NVegLibTypes + p : veg_class;
    VegRow r = vl.row(pet_class);
    if (pet_class == o.GLACIER_ID) roughness[SNOW_FREE] = cp(CP_GLAC_ROUGH);
    else roughness[SNOW_FREE] = r.m(VM_roughness, month0);
    displacement[SNOW_FREE] = r.m(VM_displacement, month0);
    const bool overstory = r.overstory();
    if (p >= N_PET_TYPES_NON_NAT && roughness[SNOW_FREE] == 0) roughness[SNOW_FREE] = soil_rough;
    const double height = displacement[SNOW_FREE] / height_den;
    if (displacement[SNOW_FREE] < wind_h) ref_height[SNOW_FREE] = wind_h;
    else ref_height[SNOW_FREE] = displacement[SNOW_FREE] + wind_h + roughness[SNOW_FREE];
    // bring the forcing wind from its nominal height to the reference height (log profile over open ground)
    g.wind_corr(p) = vlog((ref_height[SNOW_FREE] - 0.) / soil_rough) / wind_den;
    wind_speed.set_invalid();
    resist.set_invalid();
    // calc_aerodynamic_geom is a pure function of the values just set (every shared entry is overwritten), so when a land cover
    // repeats the previous one (the two bare reference covers; the tile's own cover three times) the previous results are
    // reused bit for bit instead of being recomputed.
    const double in_now[7] = {(double)overstory, height, r.s(VL_trunk_ratio), r.s(VL_wind_atten), roughness[SNOW_FREE], displacement[SNOW_FREE],
                              ref_height[SNOW_FREE]};
    bool same = (p > 0);
    for (int k = 0; k < 7; k++) same = same && (in_now[k] == in_prev[k]);
    unsigned now_valid;
    if (same) {
      resist = prev_resist;
      now_valid = prev_valid;
      displacement = snap_displacement; ref_height = snap_ref_height; roughness = snap_roughness; wind_speed = snap_wind_speed;
    } else {
      int e = calc_aerodynamic_geom(overstory, height, r.s(VL_trunk_ratio), cp(CP_snow_rough), soil_rough, r.s(VL_wind_atten), resist, wind_speed,
                                    displacement, ref_height, roughness);
      if (e == ERROR_I) {
        g.err() = 1;
        return;
      }
      now_valid = 0;
      for (int sidx = 0; sidx < 4; sidx++)
        if (is_valid(wind_speed[sidx])) now_valid |= 1u << sidx;
      for (int k = 0; k < 7; k++) in_prev[k] = in_now[k];
      prev_resist = resist;
      prev_valid = now_valid;
      snap_displacement = displacement; snap_ref_height = ref_height; snap_roughness = roughness; snap_wind_speed = wind_speed;
    }
    for (int sidx = 0; sidx < 4; sidx++) g.R0(p, sidx) = resist[sidx];
    valid |= now_valid << (4 * p);
  }
  for (int sidx = 0; sidx < 4; sidx++) {
    g.U0()[sidx] = wind_speed[sidx];
    g.displacement()[sidx] = displacement[sidx];
    g.ref_height()[sidx] = ref_height[sidx];
    g.roughness()[sidx] = roughness[sidx];
  }
  g.valid() = (double)valid;
}

// my row of the per-month table [VIC_AERO_NCOL][nhru]
VIC_HD void load_aero_geom(const Col& c, AeroGeom& g) {
#pragma unroll
  for (int k = 0; k < VIC_AERO_NCOL; k++) g.v[k] = c(k);
}

// the record's wind scales the unit-wind tables: what the seven calc_aerodynamic() calls leave behind
VIC_HD void aero_apply(const AeroGeom& g, double wind_NR, AeroState& as) {
  const unsigned valid = (unsigned)g.valid();
  const double nanv = vnan();
#pragma unroll
  for (int p = 0; p < N_PET_TYPES + 1; p++) {
    const double tmp_wind = wind_NR * g.wind_corr(p);
    Surf4 ws;
#pragma unroll
    for (int sidx = 0; sidx < 4; sidx++) {
      as.aero_resist[p][sidx] = g.R0(p, sidx);
      const bool ok = (valid >> (4 * p + sidx)) & 1u;
      ws[sidx] = (p == N_PET_TYPES) ? g.U0()[sidx] : (ok ? 1.0 : nanv);
    }
    calc_aerodynamic_wind(tmp_wind, as.aero_resist[p], ws);
    if (p == N_PET_TYPES) as.wind_speed = ws;
  }
#pragma unroll
  for (int sidx = 0; sidx < 4; sidx++) {
    as.displacement[sidx] = g.displacement()[sidx];
    as.ref_height[sidx] = g.ref_height()[sidx];
    as.roughness[sidx] = g.roughness()[sidx];
  }
}

// per-HRU diagnostics of the step that are not part of the HRU record but feed the cell
// output (atmos->out_prec / out_rain / out_snow, full_energy.c:425-427)
struct HruStepDiag {
  double out_prec, out_rain, out_snow;
};

// Returns 0 or ERROR_I (the reference then invalidates the whole cell, vicNl.c:545-559).
template <int NN, bool ONE>
VIC_HDI int hru_step(Hru<NN>& hru, const HruPar& hp, const Ctx& cx, HruStepDiag& dg) {
  const Opts& o = *cx.o;
  const CellPar& cp = cx.cp;
  dg.out_prec = dg.out_rain = dg.out_snow = 0;
  if (!((hp.Cv > 0.0) || (hp.isGlacier && o.GLACIER_DYNAMICS && hp.Cv >= 0.0))) return 0;
  const int month0 = cx.dmy.month - 1;
  const double AreaFract = cp.band(CB_AreaFract, hp.band);
  // gauge undercatch (CORRPREC, full_energy.c:185-194, correct_precip.c:10-47: WMO equations for the shielded 8-inch gauge, wind
  // brought to the gauge height of 1 m over bare ground and over snow); [0] rain, [1] snow
  double gauge_correction[2] = {1, 1};
  if (o.CORRPREC && cx.f(FV_prec, o.NR) > 0) {
    const double wind = cx.f(FV_wind, o.NR), rough = cp(CP_rough), snow_rough = cp(CP_snow_rough);
    double gauge_wind = wind * (vlog((1.0 + rough) / rough) / vlog(o.wind_h / rough));
    gauge_correction[0] = 100. / vexp(4.606 - 0.041 * vpow(gauge_wind, 0.69));
    gauge_wind = wind * (vlog((1.0 + snow_rough) / snow_rough) / vlog(o.wind_h / snow_rough));
    gauge_correction[1] = 100. / vexp(4.606 - 0.036 * vpow(gauge_wind, 1.75));
  }
  if (AreaFract > 0) {
    hru.energy.shortwave = 0;
    hru.energy.longwave = 0.;
    hru.snow.vapor_flux = 0.;
    hru.snow.canopy_vapor_flux = 0.;
  }
  // the lanes that will run a surface-flux step register for the phase rendezvous of their path (vic_types.cuh PhaseSync)
  if ((AreaFract > 0) || (hp.isGlacier && o.GLACIER_DYNAMICS && AreaFract >= 0.0)) {
    if (hp.isGlacier) cx.join(1);
    else cx.join(0);
  }
  const int veg_class = hp.vegIndex;
  const VegNow veg = veg_now(cx.vl, veg_class, month0);
  double surf_atten = vexp(-veg.rad_atten * veg.LAI);
  double moist0 = 0, ice0 = 0;
  prepare_full_energy<NN>(hru, cp, AreaFract, o, &moist0, &ice0);
  const double bare_albedo = hp.isGlacier ? cp(CP_GLAC_ALBEDO) : veg.albedo;

  // aerodynamic resistances: 4 reference land covers, then the tile's own cover three times (natural vegetation, natural vegetation
  // without canopy resistance, current).  The wind-independent part comes from the per-month table when the kernel has one.
  AeroGeom ag;
#if defined(__CUDA_ARCH__) && !defined(VIC_NO_AERO_TABLE)
  load_aero_geom(cx.aero, ag);  // (no run-time alternative on the device: `ag` must not have its address taken, or it lives in local memory)
#else
  aero_geom(cx.vl, cp, o, hp.isGlacier, veg_class, month0, veg, ag);
#endif
  if (ag.err() != 0.0) return ERROR_I;
  AeroState as;
  aero_apply(ag, cx.f(FV_wind, o.NR), as);
  const bool overstory = veg.overstory;
  if (AreaFract > 0) {
    hru.cell.aero_surface = as.aero_resist[N_PET_TYPES][SNOW_FREE];
    hru.cell.aero_overstory = as.aero_resist[N_PET_TYPES][CANOPY_OVER];
  }
  if ((AreaFract > 0) || (hp.isGlacier && o.GLACIER_DYNAMICS && AreaFract >= 0.0)) {
    for (int p = 0; p < N_PET_TYPES; p++) hru.cell.pot_evap[p] = 0;
    SoilET soil;
    for (int l = 0; l < VICGPU_NLAYER; l++) {
      soil.Wcr[l] = cp.layer(CL_Wcr, l);
      soil.Wpwp[l] = cp.layer(CL_Wpwp, l);
      soil.root[l] = hp.root[l];
    }
    SurfaceFluxOut sf;
    int e;
    if (hp.isGlacier) e = surface_fluxes_glac<NN, ONE>(bare_albedo, ice0, moist0, hru, as, gauge_correction, hp.band, cx, veg_class, sf);
    else e = surface_fluxes<NN, ONE>(overstory, bare_albedo, ice0, moist0, hru, surf_atten, as, gauge_correction, hp.isArtBare, hp.band, cx, veg, soil, veg_class, sf);
    if (e == ERROR_I) return ERROR_I;
    dg.out_prec = sf.out_prec;
    dg.out_rain = sf.out_rain;
    dg.out_snow = sf.out_snow;
    // root-zone moisture and wetness
    hru.cell.rootmoist = 0;
    hru.cell.wetness = 0;
    for (int l = 0; l < VICGPU_NLAYER; l++) {
      if (hp.root[l] > 0) hru.cell.rootmoist += hru.cell.layer[l].moist;
      hru.cell.wetness += (hru.cell.layer[l].moist - cp.layer(CL_Wpwp, l)) / (cp.layer(CL_porosity, l) * cp.layer(CL_depth, l) * 1000 - cp.layer(CL_Wpwp, l));
    }
    hru.cell.wetness /= VICGPU_NLAYER;
  }
  return 0;
}

}  // namespace vic
#endif

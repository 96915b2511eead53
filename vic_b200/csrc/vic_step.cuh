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
  // CORRPREC is rejected at create time: no gauge correction
  const double gauge_correction[2] = {1, 1};
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
  const double wind_h = veg.wind_h;
  double surf_atten = vexp(-veg.rad_atten * veg.LAI);
  double moist0 = 0, ice0 = 0;
  prepare_full_energy<NN>(hru, cp, AreaFract, o, &moist0, &ice0);
  const double bare_albedo = hp.isGlacier ? cp(CP_GLAC_ALBEDO) : veg.albedo;

  // aerodynamic resistances: 4 reference land covers, then the tile's own cover three times
  // (natural vegetation, natural vegetation without canopy resistance, current)
  AeroState as;
  as.displacement.set_invalid();
  as.roughness.set_invalid();
  as.ref_height.set_invalid();
  double height = 0;
  bool overstory = false;
  const double soil_rough = cp(CP_rough);
  const double wind_NR = cx.f(FV_wind, o.NR);
  double in_prev[8] = {-1, 0, 0, 0, 0, 0, 0, 0};
  // terms of the loop below that do not depend on the land cover: the leaf-area factor of calc_veg_height()
  // (calc_veg_params.c:26-37) and the log-profile denominator of the wind correction
  const double height_den = 1.1 * vlog(1 + vpow(0.2 * veg.LAI, 0.25));
  const double wind_den = vlog((o.wind_h - 0.) / soil_rough);
  Surf4 snap_displacement, snap_ref_height, snap_roughness, snap_wind_speed;  // the tables as the last evaluation left them (set at p == 0)
  snap_displacement.set_invalid(); snap_ref_height.set_invalid(); snap_roughness.set_invalid(); snap_wind_speed.set_invalid();
  #pragma unroll 1
  for (int p = 0; p < N_PET_TYPES + 1; p++) {
    const int pet_class = (p < N_PET_TYPES_NON_NAT) ? o.NVegLibTypes + p : veg_class;
    VegRow r = cx.vl.row(pet_class);
    if (pet_class == o.GLACIER_ID) as.roughness[SNOW_FREE] = cp(CP_GLAC_ROUGH);
    else as.roughness[SNOW_FREE] = r.m(VM_roughness, month0);
    as.displacement[SNOW_FREE] = r.m(VM_displacement, month0);
    overstory = r.overstory();
    if (p >= N_PET_TYPES_NON_NAT && as.roughness[SNOW_FREE] == 0) as.roughness[SNOW_FREE] = soil_rough;
    height = as.displacement[SNOW_FREE] / height_den;
    if (as.displacement[SNOW_FREE] < wind_h) as.ref_height[SNOW_FREE] = wind_h;
    else as.ref_height[SNOW_FREE] = as.displacement[SNOW_FREE] + wind_h + as.roughness[SNOW_FREE];
    // bring the forcing wind from its nominal height to the reference height (log profile over open ground)
    const double wind_corr = vlog((as.ref_height[SNOW_FREE] - 0.) / soil_rough) / wind_den;
    as.wind_speed[SNOW_FREE] = wind_NR * wind_corr;
    as.wind_speed[CANOPY_OVER] = vnan();
    as.wind_speed[SNOW_COVERED] = vnan();
    as.wind_speed[GLACIER_SURF] = vnan();
    as.aero_resist[p].set_invalid();
    // calc_aerodynamic is a pure function of the values just set (every shared entry is overwritten), so when
    // a land cover repeats the previous one (the two bare reference covers; the tile's own cover three times)
    // the previous results are reused bit for bit instead of being recomputed.
    const double in_now[8] = {(double)overstory, height, r.s(VL_trunk_ratio), r.s(VL_wind_atten), as.roughness[SNOW_FREE],
                              as.displacement[SNOW_FREE], as.ref_height[SNOW_FREE], as.wind_speed[SNOW_FREE]};
    bool same = (p > 0);
    for (int k = 0; k < 8; k++) same = same && (in_now[k] == in_prev[k]);
    if (same) {
      as.aero_resist[p] = as.aero_resist[p - 1];
      as.displacement = snap_displacement; as.ref_height = snap_ref_height; as.roughness = snap_roughness; as.wind_speed = snap_wind_speed;
    } else {
      int e = calc_aerodynamic(overstory, height, r.s(VL_trunk_ratio), cp(CP_snow_rough), soil_rough, r.s(VL_wind_atten), as.aero_resist[p],
                               as.wind_speed, as.displacement, as.ref_height, as.roughness);
      if (e == ERROR_I) return ERROR_I;
      for (int k = 0; k < 8; k++) in_prev[k] = in_now[k];
      snap_displacement = as.displacement; snap_ref_height = as.ref_height; snap_roughness = as.roughness; snap_wind_speed = as.wind_speed;
    }
  }
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

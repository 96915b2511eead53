// vic_evap.cuh -- evapotranspiration of one HRU: interception-store evaporation and
// transpiration with soil-moisture stress (canopy_evap.c:46-442), bare-soil ARNO
// evaporation (arno_evap.c:61-228) and the six potential-evaporation diagnostics
// (compute_pot_evap.c:8-78).
//
// The wet/dry precipitation split (DIST_PRCP) is not on the hot path (vicgpu_create rejects
// it), so every routine handles the single "wet" fraction with mu == 1.
#ifndef VIC_EVAP_CUH
#define VIC_EVAP_CUH
#include "vic_leaf.cuh"

namespace vic {

// vegetation-library values of one class in the current month
struct VegNow {
  bool overstory;
  float RGL;
  double rarc, rmin, LAI, Wdmax, albedo, displacement, roughness, wind_h, rad_atten, wind_atten, trunk_ratio;
};

VIC_HD VegNow veg_now(const VegLib& vl, int cls, int month0) {
  VegRow r = vl.row(cls);
  VegNow v;
  v.overstory = r.overstory();
  v.RGL = (float)r.s(VL_RGL);
  v.rarc = r.s(VL_rarc);
  v.rmin = r.s(VL_rmin);
  v.wind_h = r.s(VL_wind_h);
  v.rad_atten = r.s(VL_rad_atten);
  v.wind_atten = r.s(VL_wind_atten);
  v.trunk_ratio = r.s(VL_trunk_ratio);
  v.LAI = r.m(VM_LAI, month0);
  v.Wdmax = r.m(VM_Wdmax, month0);
  v.albedo = r.m(VM_albedo, month0);
  v.displacement = r.m(VM_displacement, month0);
  v.roughness = r.m(VM_roughness, month0);
  return v;
}

// per-layer soil constants used by the ET routines
struct SoilET {
  double Wcr[VICGPU_NLAYER], Wpwp[VICGPU_NLAYER];
  float root[VICGPU_NLAYER];
};

// Memo of one surface-temperature solve.  The root finders call the evaporation routines a dozen times per solve with a new
// net radiation (and aerodynamic resistance) each time and everything else unchanged; the sub-results that do not depend on
// those two are pure functions of unchanged inputs, so they are computed at the first evaluation and reused: bit-identical to
// recomputing them.  A null memo means "compute everything" (the calls outside a solve).
struct EvapMemo {
  int pm_ok, rc_ok[VICGPU_NLAYER + 1], p23_ok[2], arno_ok, arno_kind;
  PenmanPre pm;                  // penman terms of (air temperature, elevation)
  double rc[VICGPU_NLAYER + 1];  // canopy resistance: [0] unstressed, [1 + i] layer i under moisture stress
  double p23[2];                 // (Wdew / Wdmax)^(2/3): [0] interception store in canopy_evap, [1] in transpiration
  double arno_beta;              // bare soil: evaporation = Epot * arno_beta (arno_kind 1) or Epot (arno_kind 0)
  VIC_HD void reset() {
    pm_ok = arno_ok = arno_kind = 0;
    p23_ok[0] = p23_ok[1] = 0;
    for (int i = 0; i <= VICGPU_NLAYER; i++) rc_ok[i] = 0;
  }
};
VIC_HD double penman_m(EvapMemo* m, double tair, double elevation, double rad, double vpd, double ra, double rc, double rarc) {
  if (!m) return penman(tair, elevation, rad, vpd, ra, rc, rarc);
  if (!m->pm_ok) {
    m->pm = penman_pre(tair, elevation);
    m->pm_ok = 1;
  }
  return penman_eval(m->pm, rad, vpd, ra, rc, rarc);
}
VIC_HD double calc_rc_m(EvapMemo* m, int site, double rs, double net_short, float RGL, double tair, double vpd, double lai, double gsm_inv) {
  if (!m) return calc_rc(rs, net_short, RGL, tair, vpd, lai, gsm_inv, false);
  if (!m->rc_ok[site]) {
    m->rc[site] = calc_rc(rs, net_short, RGL, tair, vpd, lai, gsm_inv, false);
    m->rc_ok[site] = 1;
  }
  return m->rc[site];
}
VIC_HD double pow23_m(EvapMemo* m, int site, double x) {
  if (!m) return vpow(x, (2.0 / 3.0));
  if (!m->p23_ok[site]) {
    m->p23[site] = vpow(x, (2.0 / 3.0));
    m->p23_ok[site] = 1;
  }
  return m->p23[site];
}

// canopy_evap.c:218-442
VIC_HD void transpiration(SoilLayer* layer, const VegNow& veg, double rad, double vpd, double net_short, double air_temp,
                           double ra, double f, double delta_t, double Wdew, double elevation, const SoilET& s,
                           double* layerevap, EvapMemo* memo = nullptr) {
  const int NL = VICGPU_NLAYER;
  double avail_moist[NL], ice[NL];
  for (int i = 0; i < NL; i++) ice[i] = layer[i].soil_ice;
  double moist1 = 0.0, Wcr1 = 0.0;
  for (int i = 0; i < NL - 1; i++) {
    if (s.root[i] > 0.) {
      avail_moist[i] = layer[i].moist - layer[i].soil_ice;
      moist1 += avail_moist[i];
      Wcr1 += s.Wcr[i];
    } else avail_moist[i] = 0.;
  }
  const double moist2 = layer[NL - 1].moist - layer[NL - 1].soil_ice;
  avail_moist[NL - 1] = moist2;
  const double wet_canopy = 1.0 - f * pow23_m(memo, 1, div_zn(Wdew, veg.Wdmax));
  // (1 - root) is evaluated in single precision by the reference (float operand)
  const double one_minus_rootN = (double)(1.0f - s.root[NL - 1]);
  if ((moist1 >= Wcr1 && moist2 >= s.Wcr[NL - 1] && Wcr1 > 0.) || (moist1 >= Wcr1 && one_minus_rootN >= 0.5) ||
      (moist2 >= s.Wcr[NL - 1] && s.root[NL - 1] >= 0.5)) {
    double rc = calc_rc_m(memo, 0, veg.rmin, net_short, veg.RGL, air_temp, vpd, veg.LAI, 1.0);
    double evap = div_pos(penman_m(memo, air_temp, elevation, rad, vpd, ra, rc, veg.rarc) * delta_t, SEC_PER_DAY) * wet_canopy;
    double root_sum = 1.0, spare_evap = 0.0;
    #pragma unroll
    for (int i = 0; i < NL; i++) {
      if (avail_moist[i] >= s.Wcr[i]) layerevap[i] = evap * (double)s.root[i];
      else {
        double gsm_inv;
        if (avail_moist[i] >= s.Wpwp[i]) gsm_inv = (avail_moist[i] - s.Wpwp[i]) / (s.Wcr[i] - s.Wpwp[i]);
        else gsm_inv = 0.0;
        layerevap[i] = evap * gsm_inv * (double)s.root[i];
        root_sum -= s.root[i];
        spare_evap = evap * (double)s.root[i] * (1.0 - gsm_inv);
      }
    }
    if (spare_evap > 0.0)
      #pragma unroll
      for (int i = 0; i < NL; i++)
        if (avail_moist[i] >= s.Wcr[i]) layerevap[i] += (double)s.root[i] * spare_evap / root_sum;
  } else {
    #pragma unroll
    for (int i = 0; i < NL; i++) {
      double gsm_inv;
      if (avail_moist[i] >= s.Wcr[i]) gsm_inv = 1.0;
      else if (avail_moist[i] >= s.Wpwp[i]) gsm_inv = (avail_moist[i] - s.Wpwp[i]) / (s.Wcr[i] - s.Wpwp[i]);
      else gsm_inv = 0.0;
      if (gsm_inv > 0.0) {
        double rc = calc_rc_m(memo, 1 + i, veg.rmin, net_short, veg.RGL, air_temp, vpd, veg.LAI, gsm_inv);
        layerevap[i] = div_pos(penman_m(memo, air_temp, elevation, rad, vpd, ra, rc, veg.rarc) * delta_t, SEC_PER_DAY) * (double)s.root[i] * wet_canopy;
      } else layerevap[i] = 0.0;
    }
  }
  #pragma unroll
  for (int i = 0; i < NL; i++) {
    if (ice[i] > 0) {
      if (ice[i] >= s.Wpwp[i]) {
        if (layerevap[i] > avail_moist[i]) layerevap[i] = avail_moist[i];
      } else {
        if (layerevap[i] > layer[i].moist - s.Wpwp[i]) layerevap[i] = layer[i].moist - s.Wpwp[i];
      }
    } else {
      if (layerevap[i] > layer[i].moist - s.Wpwp[i]) layerevap[i] = layer[i].moist - s.Wpwp[i];
    }
    if (layerevap[i] < 0.0) layerevap[i] = 0.0;
  }
}

// canopy_evap.c:46-212.  Wdew_in: interception store at the start of the sub-step [mm];
// ppt: rain reaching the canopy [mm]; returns total evaporation [m/s] and leaves
// canopyevap / throughfall / Wdew in veg and layer[].evap.
VIC_HD double canopy_evap(SoilLayer* layer, VegVar& vv, bool CALC_EVAP, const VegNow& veg, double Wdew_in, double delta_t,
                           double rad, double vpd, double net_short, double air_temp, double ra, double elevation, double ppt,
                           const SoilET& s, EvapMemo* memo = nullptr) {
  double layerevap[VICGPU_NLAYER];
  for (int i = 0; i < VICGPU_NLAYER; i++) layerevap[i] = 0;
  double throughfall = 0;
  double tmp_Wdew = Wdew_in;
  vv.Wdew = tmp_Wdew;
  if (tmp_Wdew > veg.Wdmax) {
    throughfall = tmp_Wdew - veg.Wdmax;
    tmp_Wdew = veg.Wdmax;
  }
  double rc = calc_rc(0.0, net_short, veg.RGL, air_temp, vpd, veg.LAI, 1.0, false);
  double canopyevap = div_pos(pow23_m(memo, 0, div_zn(tmp_Wdew, veg.Wdmax)) * penman_m(memo, air_temp, elevation, rad, vpd, ra, rc, veg.rarc) * delta_t, SEC_PER_DAY);
  double f;
  if (canopyevap > 0.0 && delta_t == SEC_PER_DAY) f = vmin(1.0, ((tmp_Wdew + ppt) / canopyevap));
  else if (canopyevap > 0.0) f = vmin(1.0, ((tmp_Wdew) / canopyevap));
  else f = 1.0;
  canopyevap *= f;
  tmp_Wdew += ppt - canopyevap;
  if (tmp_Wdew < 0.0) tmp_Wdew = 0.0;
  if (tmp_Wdew <= veg.Wdmax) throughfall += 0.0;
  else {
    throughfall += tmp_Wdew - veg.Wdmax;
    tmp_Wdew = veg.Wdmax;
  }
  if (CALC_EVAP) transpiration(layer, veg, rad, vpd, net_short, air_temp, ra, f, delta_t, vv.Wdew, elevation, s, layerevap, memo);
  vv.canopyevap = canopyevap;
  vv.throughfall = throughfall;
  vv.Wdew = tmp_Wdew;
  double tmp_Evap = canopyevap;
  for (int i = 0; i < VICGPU_NLAYER; i++) {
    layer[i].evap = layerevap[i];
    tmp_Evap += layerevap[i];
  }
  return 0 + div_pos(tmp_Evap * 1.0, (1000. * delta_t));
}

// arno_evap.c:61-228; returns evaporation [m/s] or ERROR_D.  Only Epot depends on the net radiation and the resistance: the
// infiltration-curve factor (three pow and a 30-term series) goes through the memo.
VIC_HD double arno_evap(SoilLayer* layer, double rad, double air_temp, double vpd, double depth1, double max_moist, double elevation,
                         double b_infilt, double ra, double delta_t, double moist_resid, EvapMemo* memo = nullptr) {
  double tmp, ratio, as, evap;
  double moist = layer[0].moist - layer[0].soil_ice;
  if (moist > max_moist) moist = max_moist;
  double Epot = div_pos(penman_m(memo, air_temp, elevation, rad, vpd, ra, 0.0, 0.0) * delta_t, SEC_PER_DAY);
  int kind;
  double beta_asp = 0;
  if (memo && memo->arno_ok) {
    kind = memo->arno_kind;
    beta_asp = memo->arno_beta;
  } else {
    double max_infil = (1.0 + b_infilt) * max_moist;
    if (b_infilt == -1.0) tmp = max_infil;
    else {
      ratio = 1.0 - (moist) / (max_moist);
      if (ratio > 1.0) return ERROR_D;
      else if (ratio < 0.0) return ERROR_D;
      else ratio = vpow(ratio, (1.0 / (b_infilt + 1.0)));
      tmp = max_infil * (1.0 - ratio);
    }
    if (tmp >= max_infil) kind = 0;
    else {
      kind = 1;
      ratio = tmp / max_infil;
      ratio = 1.0 - ratio;
      if (ratio > 1.0) return ERROR_D;
      else if (ratio < 0.0) return ERROR_D;
      else if (ratio != 0.0) ratio = vpow(ratio, b_infilt);
      as = 1 - ratio;
      ratio = vpow(ratio, (1.0 / b_infilt));
      // 30-term series; the running power reproduces the reference's repeated product
      // tmpsum = ratio * ratio * ... (left to right), so the partial products are identical
      double dummy = 1.0, tmpsum = 1.0;
      #pragma unroll 1
      for (int num_term = 1; num_term <= 30; num_term++) {
        tmpsum = (num_term == 1) ? ratio : tmpsum * ratio;
        dummy += b_infilt * tmpsum / (b_infilt + num_term);
      }
      beta_asp = as + (1.0 - as) * (1.0 - ratio) * dummy;
    }
    if (memo) {
      memo->arno_kind = kind;
      memo->arno_beta = beta_asp;
      memo->arno_ok = 1;
    }
  }
  if (kind == 0) evap = Epot;
  else evap = Epot * beta_asp;
  if (evap > 0.0) {
    if (moist > moist_resid * depth1 * 1000.) {
      if (evap > moist - moist_resid * depth1 * 1000.) evap = moist - moist_resid * depth1 * 1000.;
    } else evap = 0.0;
  }
  layer[0].evap = evap;
  return 0 + div_pos(div_pos(evap, 1000.), delta_t) * 1.0;
}

// compute_pot_evap.c:8-78.  aero[p] = {surface, overstory} resistances already corrected for
// stability (surface_fluxes.c:662-688).  net_short of PET type i is the one of type i-1
// (compute_pot_evap.c:68 vs :73); type 0 never reads it (rs == 0).
VIC_HDI void compute_pot_evap(const VegLib& vl, int NVegLibTypes, int veg_class, int month0, int dt, double shortwave,
                              double net_longwave, double tair, double vpd, double elevation, const RaUsed* aero, double* pot_evap) {
  double net_short = 0.0;
  const bool cur_over = vl.row(veg_class).overstory();
  const PenmanPre pm = penman_pre(tair, elevation);  // the same air temperature and elevation for all six land covers
  #pragma unroll 1
  for (int i = 0; i < N_PET_TYPES; i++) {
    const int cls = (i < N_PET_TYPES_NON_NAT) ? NVegLibTypes + i : veg_class;
    VegRow r = vl.row(cls);
    double rs = r.s(VL_rmin);
    if (i == PET_VEGNOCR) rs = 0;
    const double rarc = r.s(VL_rarc);
    const float RGL = (float)r.s(VL_RGL);
    const double lai = r.m(VM_LAI, month0);
    const double albedo = r.m(VM_albedo, month0);
    const bool ref_crop = (i == 2 || i == 3);
    double rc = calc_rc(rs, net_short, RGL, tair, vpd, lai, 1.0, ref_crop);
    double ra = (i < N_PET_TYPES_NON_NAT || !cur_over) ? aero[i].surface : aero[i].overstory;
    net_short = (1.0 - albedo) * shortwave;
    double net_rad = net_short + net_longwave;
    pot_evap[i] = penman_eval(pm, net_rad, vpd, ra, rc, rarc) * dt / 24.0;
  }
}

}  // namespace vic
#endif

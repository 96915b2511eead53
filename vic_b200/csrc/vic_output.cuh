// vic_output.cuh -- per-cell output of one model step: the area-weighted reduction of the HRU
// records into the 184 output variables, the derived variables, the water / energy balance
// checks and the temporal aggregation, as put_data() does it (put_data.c:7-760,
// collect_wb_terms :762-939, collect_eb_terms :941-1232,
// calc_water_energy_balance_errors.c:7-93).  HRUs are visited in hruList order so that every
// sum is formed in the reference's order.
//
// Lake / wetland terms are always zero here (LAKES is rejected at create time).
#ifndef VIC_OUTPUT_CUH
#define VIC_OUTPUT_CUH
#include "vic_types.cuh"

namespace vic {

// per-cell carry-over of put_data: save_data (vicNl_def.h save_data_struct) and CellBalanceErrors
enum CellCarry { CC_total_soil_moist = 0, CC_surfstor, CC_swe, CC_wdew, CC_water_last_storage, CC_water_cum_error, CC_water_max_error,
                 CC_energy_cum_error, CC_energy_max_error, CC_N };

// strided accessor: element k of a column-major table, for one row
struct RowRW {
  double* p;
  size_t n;
  VIC_HD double& operator[](int k) const { return p[(size_t)k * n]; }
};

// the row of one cell while it is built up on the device: a thread-local array
struct RowLocal {
  double* p;
  VIC_HD double& operator[](int k) const { return p[k]; }
};
// one HRU record, read in place from the tile-major state table
struct RecTile {
  const double* p;  // hrec + hr_off(h, hr_stride)
  VIC_HD double operator()(int k) const {
    return VIC_REC_LD(p + (size_t)k * VIC_HR_TILE);
  }
};

// what put_data carries from HRU to HRU of a cell
struct PutDataCtx {
  double TreeAdjustFactor[VICGPU_MAX_BANDS];
  double cv_baresoil, cv_veg, cv_overstory, cv_snow, cv_glacier;
};

#define OUT(v, e) out[L.out_off[VOUT_##v] + (e)]
#define HR(k) hr(k)
#define HP(k) hpar[(size_t)(k) * nhru + h]

// put_data() is restated in four parts (per cell / per HRU / per cell / aggregation) so that the driver decides where the row and
// the HRU records live while the cell is reduced (vic_engine.cuh cell_output).
//
// Part 1, once per cell: tree-line adjustment factors, the forcing variables and the Cv-weighted precipitation sums.  `out` has been
// zeroed.  [h0, h1) are the cell's HRUs in hruList order; slot (may be null = identity) maps an HRU to the table row it currently
// occupies (the device keeps HRUs binned by kind, not by cell).  rec < 0 reproduces the storage initialisation call
// put_data(rec = -nrecs) (vicNl.c:524-541).
template <class OutRow>
VIC_HD void put_data_begin(const Opts& o, const CellPar& cp, const VegLib& vl, const Forcing* f, const double* __restrict__ hpar,
                           const double* __restrict__ hdiag, size_t nhru, const int* __restrict__ slot, int h0, int h1, int rec, OutRow out, PutDataCtx& pc) {
  const vicgpu_layout& L = o.L;
  const int Nbands = o.Nbands;
  double bandCv[VICGPU_MAX_BANDS];
  for (int b = 0; b < VICGPU_MAX_BANDS; b++) bandCv[b] = 0;
  for (int hh = h0; hh < h1; hh++) {
    const int h = slot ? slot[hh] : hh;
    if (vl.row((int)HP(HP_vegIndex)).overstory()) bandCv[(int)HP(HP_band)] += HP(HP_Cv);
  }
  for (int b = 0; b < Nbands; b++) {
    if (cp.band(CB_AboveTreeLine, b) != 0.0) pc.TreeAdjustFactor[b] = 1. / (1. - bandCv[b]);
    else pc.TreeAdjustFactor[b] = 1.;
  }
  pc.cv_baresoil = pc.cv_veg = pc.cv_overstory = pc.cv_snow = pc.cv_glacier = 0;
  double out_prec = 0, out_rain = 0, out_snow = 0;
  if (rec >= 0) {
    // atmos->out_prec etc. (full_energy.c:425-427): Cv-weighted sums over the HRUs that were stepped
    for (int hh = h0; hh < h1; hh++) {
    const int h = slot ? slot[hh] : hh;
      out_prec += hdiag[(size_t)0 * nhru + h];
      out_rain += hdiag[(size_t)1 * nhru + h];
      out_snow += hdiag[(size_t)2 * nhru + h];
    }
    const int NR = o.NR;
    OUT(AIR_TEMP, 0) = (*f)(FV_air_temp, NR);
    OUT(DENSITY, 0) = (*f)(FV_density, NR);
    OUT(LONGWAVE, 0) = (*f)(FV_longwave, NR);
    OUT(PREC, 0) = out_prec;
    OUT(PRESSURE, 0) = (*f)(FV_pressure, NR) / 1000.;
    OUT(QAIR, 0) = EPS * (*f)(FV_vp, NR) / (*f)(FV_pressure, NR);
    OUT(RAINF, 0) = out_rain;
    OUT(REL_HUMID, 0) = 100. * (*f)(FV_vp, NR) / ((*f)(FV_vp, NR) + (*f)(FV_vpd, NR));
    OUT(LAKE_CHAN_IN, 0) = 0;
    OUT(SHORTWAVE, 0) = (*f)(FV_shortwave, NR);
    OUT(SNOWF, 0) = out_snow;
    OUT(TSKC, 0) = (*f)(FV_tskc, NR);
    OUT(VP, 0) = (*f)(FV_vp, NR) / 1000.;
    OUT(VPD, 0) = (*f)(FV_vpd, NR) / 1000.;
    OUT(WIND, 0) = (*f)(FV_wind, NR);
  }
}

// Part 2, once per HRU in hruList order: h is the HRU's table row, hr its record.
// G selects which groups of output variables are accumulated (bits): PD_WB the water-balance terms (collect_wb_terms), PD_EB the
// energy-balance terms of the cell (collect_eb_terms), PD_BAND the per-band terms.  Every output variable belongs to exactly one
// group (out_var_group below), so three threads can reduce the three groups of a cell side by side -- each variable is still summed
// by ONE thread over the HRUs in hruList order -- and PD_ALL in one thread is put_data as the reference runs it.
enum { PD_WB = 1, PD_EB = 2, PD_BAND = 4, PD_ALL = 7 };
template <int G, class Rec, class OutRow>
VIC_HD void put_data_hru(const Opts& o, const CellPar& cp, const VegLib& vl, Rec hr, const double* __restrict__ hpar, size_t nhru, int h, OutRow out,
                         PutDataCtx& pc) {
  const vicgpu_layout& L = o.L;
  const int NL = VICGPU_NLAYER;
  {
    const double Cv = HP(HP_Cv);
    const bool isArtBare = HP(HP_isArtBare) != 0.0, HasGlac = HP(HP_isGlacier) != 0.0;
    const bool HasVeg = !(isArtBare || HasGlac);
    if (!(Cv > 0)) return;
    const int band = (int)HP(HP_band);
    const bool overstory = vl.row((int)HP(HP_vegIndex)).overstory();
    const double ThisAreaFract = cp.band(CB_AreaFract, band);
    const double ThisTreeAdjust = pc.TreeAdjustFactor[band];
    const bool above = cp.band(CB_AboveTreeLine, band) != 0.0;
    if (!(ThisAreaFract > 0. && (isArtBare || (!above || (above && !overstory))))) return;
    if (G & PD_BAND) OUT(ELEV_BAND, band) = cp.band(CB_BandElev, band);
    const double mu = HR(HR_H_mu);
    const double swq = HR(HR_S_swq);
    if (G & PD_WB) {  // the area sums the derived variables of put_data_finish divide by (the thread that reduces PD_WB runs it)
      if (HasVeg) pc.cv_veg += Cv * mu * ThisTreeAdjust;
      else pc.cv_baresoil += Cv * mu * ThisTreeAdjust;
      if (overstory) pc.cv_overstory += Cv * mu * ThisTreeAdjust;
      if (swq > 0.0) pc.cv_snow += Cv * mu * ThisTreeAdjust;
      if (HasGlac) pc.cv_glacier += Cv * mu * ThisTreeAdjust;
    }
    // ---- water balance terms
    if (G & PD_WB) {
      const double AreaFactor = Cv * mu * ThisTreeAdjust * 1.0;
      double tmp_evap = 0.0;
      for (int l = 0; l < NL; l++) tmp_evap += HR(VICGPU_HR_LAYER(&L, HRL_evap, l));
      if (HasVeg) OUT(TRANSP_VEG, 0) += tmp_evap * AreaFactor;
      else OUT(EVAP_BARE, 0) += tmp_evap * AreaFactor;
      tmp_evap += HR(HR_S_vapor_flux) * 1000.;
      OUT(SUB_SNOW, 0) += HR(HR_S_vapor_flux) * 1000. * AreaFactor;
      OUT(SUB_SURFACE, 0) += HR(HR_S_surface_flux) * 1000. * AreaFactor;
      OUT(SUB_BLOWING, 0) += HR(HR_S_blowing_flux) * 1000. * AreaFactor;
      if (HasVeg) {
        tmp_evap += HR(HR_S_canopy_vapor_flux) * 1000.;
        OUT(SUB_CANOP, 0) += HR(HR_S_canopy_vapor_flux) * 1000. * AreaFactor;
        tmp_evap += HR(HR_V_canopyevap);
        OUT(EVAP_CANOP, 0) += HR(HR_V_canopyevap) * AreaFactor;
      }
      if (HasGlac) tmp_evap += HR(HR_G_vapor_flux) * 1000.;
      OUT(EVAP, 0) += tmp_evap * AreaFactor;
      OUT(PET_SATSOIL, 0) += HR(VICGPU_HR_PET(&L, 0)) * AreaFactor;
      OUT(PET_H2OSURF, 0) += HR(VICGPU_HR_PET(&L, 1)) * AreaFactor;
      OUT(PET_SHORT, 0) += HR(VICGPU_HR_PET(&L, 2)) * AreaFactor;
      OUT(PET_TALL, 0) += HR(VICGPU_HR_PET(&L, 3)) * AreaFactor;
      OUT(PET_NATVEG, 0) += HR(VICGPU_HR_PET(&L, 4)) * AreaFactor;
      OUT(PET_VEGNOCR, 0) += HR(VICGPU_HR_PET(&L, 5)) * AreaFactor;
      OUT(ASAT, 0) += HR(HR_C_asat) * AreaFactor;
      OUT(RUNOFF, 0) += HR(HR_C_runoff) * AreaFactor;
      OUT(BASEFLOW, 0) += HR(HR_C_baseflow) * AreaFactor;
      OUT(INFLOW, 0) += (HR(HR_C_inflow)) * AreaFactor;
      if (HasVeg) OUT(WDEW, 0) += HR(HR_V_Wdew) * AreaFactor;
      double tmp_cond1, tmp_cond2;
      if (HR(HR_C_aero_surface) > SMALL) tmp_cond1 = (1 / HR(HR_C_aero_surface)) * AreaFactor;
      else tmp_cond1 = HUGE_RESIST;
      OUT(AERO_COND1, 0) += tmp_cond1;
      if (overstory) {
        if (HR(HR_C_aero_overstory) > SMALL) tmp_cond2 = (1 / HR(HR_C_aero_overstory)) * AreaFactor;
        else tmp_cond2 = HUGE_RESIST;
      } else tmp_cond2 = HUGE_RESIST;
      OUT(AERO_COND2, 0) += tmp_cond2;
      if (overstory) OUT(AERO_COND, 0) += tmp_cond2;
      else OUT(AERO_COND, 0) += tmp_cond1;
      for (int l = 0; l < NL; l++) {
        double tmp_moist = HR(VICGPU_HR_LAYER(&L, HRL_moist, l));
        double tmp_ice = HR(VICGPU_HR_LAYER(&L, HRL_soil_ice, l));
        tmp_moist -= tmp_ice;
        if (o.MOISTFRACT) {
          tmp_moist /= cp.layer(CL_depth, l) * 1000.;
          tmp_ice /= cp.layer(CL_depth, l) * 1000.;
        }
        OUT(SOIL_LIQ, l) += tmp_moist * AreaFactor;
        OUT(SOIL_ICE, l) += tmp_ice * AreaFactor;
      }
      OUT(SOIL_WET, 0) += HR(HR_C_wetness) * AreaFactor;
      OUT(ROOTMOIST, 0) += HR(HR_C_rootmoist) * AreaFactor;
      OUT(ZWT, 0) += HR(HR_C_zwt) * AreaFactor;
      OUT(ZWT2, 0) += HR(HR_C_zwt2) * AreaFactor;
      OUT(ZWT3, 0) += HR(HR_C_zwt3) * AreaFactor;
      for (int l = 0; l < NL; l++) OUT(ZWTL, l) += HR(VICGPU_HR_LAYER(&L, HRL_zwt, l)) * AreaFactor;
      for (int l = 0; l < NL; l++) OUT(SOIL_TEMP, l) += HR(VICGPU_HR_LAYER(&L, HRL_T, l)) * AreaFactor;
      OUT(SWE, 0) += swq * AreaFactor * 1000.;
      OUT(SNOW_DEPTH, 0) += HR(HR_S_depth) * AreaFactor * 100.;
      if (swq > 0.0) {
        OUT(SALBEDO, 0) += HR(HR_S_albedo) * AreaFactor;
        OUT(SNOW_SURF_TEMP, 0) += HR(HR_S_surf_temp) * AreaFactor;
        OUT(SNOW_PACK_TEMP, 0) += HR(HR_S_pack_temp) * AreaFactor;
      }
      if (HasVeg) OUT(SNOW_CANOPY, 0) += (HR(HR_S_snow_canopy)) * AreaFactor * 1000.;
      OUT(SNOW_MELT, 0) += HR(HR_S_melt) * AreaFactor * 1000.;
      OUT(SNOW_COVER, 0) += HR(HR_S_coverage) * AreaFactor;
      if (HasGlac) {
        OUT(GLAC_WAT_STOR, 0) += HR(HR_G_water_storage) * AreaFactor * 1000.;
        OUT(GLAC_AREA, 0) += AreaFactor;
        OUT(GLAC_MBAL, 0) += HR(HR_G_mass_balance) * AreaFactor * 1000.;
        OUT(GLAC_IMBAL, 0) += HR(HR_G_ice_mass_balance) * AreaFactor * 1000.;
        OUT(GLAC_ACCUM, 0) += HR(HR_G_accumulation) * AreaFactor * 1000.;
        OUT(GLAC_MELT, 0) += HR(HR_G_melt) * AreaFactor * 1000.;
        OUT(GLAC_SUB, 0) += HR(HR_G_vapor_flux) * AreaFactor * 1000.;
        OUT(GLAC_INFLOW, 0) += HR(HR_G_inflow) * AreaFactor * 1000.;
        OUT(GLAC_OUTFLOW, 0) += HR(HR_G_outflow) * AreaFactor * 1000.;
        OUT(GLAC_OUTFLOW_COEF, 0) += HR(HR_G_outflow_coef) * AreaFactor;
      }
    }
    // ---- energy balance terms
    const bool snowing = HR(HR_S_snow) != 0.0;
    if (G & PD_EB) {
      const double AreaFactor = Cv * ThisTreeAdjust * 1.0;
      if (o.FROZEN_SOIL) {
        for (int i = 0; i < VICGPU_NFRONTS; i++) {
          const double fd = HR(VICGPU_HR_FRONT(&L, HRF_fdepth, i)), td = HR(VICGPU_HR_FRONT(&L, HRF_tdepth, i));
          if (is_valid(fd)) OUT(FDEPTH, i) += fd * AreaFactor * 100.;
          if (is_valid(td)) OUT(TDEPTH, i) += td * AreaFactor * 100.;
        }
      }
      double tmp_fract = 0;
      if (HR(VICGPU_HR_LAYER(&L, HRL_soil_ice, 0)) > 0) tmp_fract = 1.;
      OUT(SURF_FROST_FRAC, 0) += tmp_fract * AreaFactor;
      double rad_temp;
      if (overstory && snowing) rad_temp = HR(HR_E_Tcanopy) + KELVIN;
      else rad_temp = HR(HR_E_Tsurf) + KELVIN;
      const double surf_temp = HR(HR_E_Tsurf);
      // put_data.c:1025-1036 (the BARESOILT / VEGT labels are swapped in the reference)
      if (HasVeg) OUT(BARESOILT, 0) += (rad_temp - KELVIN) * AreaFactor;
      else {
        if (overstory && !snowing) OUT(VEGT, 0) += HR(HR_E_Tfoliage) * AreaFactor;
        else OUT(VEGT, 0) += (rad_temp - KELVIN) * AreaFactor;
      }
      OUT(SURF_TEMP, 0) += surf_temp * AreaFactor;
      for (int n = 0; n < o.Nnode; n++) OUT(SOIL_TNODE, n) += HR(VICGPU_HR_NODE(&L, HRN_T, n)) * AreaFactor;
      OUT(SURFT_FBFLAG, 0) += HR(HR_E_Tsurf_fbflag) * AreaFactor;
      for (int n = 0; n < o.Nnode; n++) OUT(SOILT_FBFLAG, n) += HR(VICGPU_HR_NODE(&L, HRN_T_fbflag, n)) * AreaFactor;
      OUT(SNOWT_FBFLAG, 0) += HR(HR_S_surf_temp_fbflag) * AreaFactor;
      OUT(TFOL_FBFLAG, 0) += HR(HR_E_Tfoliage_fbflag) * AreaFactor;
      OUT(TCAN_FBFLAG, 0) += HR(HR_E_Tcanopy_fbflag) * AreaFactor;
      OUT(GLAC_TSURF_FBFLAG, 0) += HR(HR_G_surf_temp_fbflag) * AreaFactor;
      OUT(NET_SHORT, 0) += HR(HR_E_NetShortAtmos) * AreaFactor;
      OUT(NET_LONG, 0) += HR(HR_E_NetLongAtmos) * AreaFactor;
      if (snowing && overstory) OUT(IN_LONG, 0) += HR(HR_E_LongOverIn) * AreaFactor;
      else OUT(IN_LONG, 0) += HR(HR_E_LongUnderIn) * AreaFactor;
      if (snowing && overstory) OUT(ALBEDO, 0) += HR(HR_E_AlbedoOver) * AreaFactor;
      else OUT(ALBEDO, 0) += HR(HR_E_AlbedoUnder) * AreaFactor;
      OUT(LATENT, 0) -= HR(HR_E_AtmosLatent) * AreaFactor;
      OUT(LATENT_SUB, 0) -= HR(HR_E_AtmosLatentSub) * AreaFactor;
      OUT(SENSIBLE, 0) -= HR(HR_E_AtmosSensible) * AreaFactor;
      OUT(GRND_FLUX, 0) -= HR(HR_E_grnd_flux) * AreaFactor;
      OUT(DELTAH, 0) -= HR(HR_E_deltaH) * AreaFactor;
      OUT(FUSION, 0) -= HR(HR_E_fusion) * AreaFactor;
      OUT(ENERGY_ERROR, 0) += HR(HR_E_error) * AreaFactor;
      OUT(RAD_TEMP, 0) += ((rad_temp) * (rad_temp) * (rad_temp) * (rad_temp)) * AreaFactor;
      OUT(DELTACC, 0) += HR(HR_E_deltaCC) * AreaFactor;
      if (snowing && overstory) OUT(ADVECTION, 0) += HR(HR_E_canopy_advection) * AreaFactor;
      OUT(ADVECTION, 0) += HR(HR_E_advection) * AreaFactor;
      OUT(SNOW_FLUX, 0) += HR(HR_E_snow_flux) * AreaFactor;
      if (snowing && overstory) OUT(RFRZ_ENERGY, 0) += HR(HR_E_canopy_refreeze) * AreaFactor;
      OUT(RFRZ_ENERGY, 0) += HR(HR_E_refreeze_energy) * AreaFactor;
      OUT(MELT_ENERGY, 0) += HR(HR_E_melt_energy) * AreaFactor;
      if (!overstory) OUT(ADV_SENS, 0) -= HR(HR_E_advected_sensible) * AreaFactor;
      if (HasGlac) {
        OUT(GLAC_SURF_TEMP, 0) += HR(HR_G_surf_temp) * AreaFactor;
        OUT(GLAC_DELTACC, 0) += HR(HR_E_deltaCC_glac) * AreaFactor;
        OUT(GLAC_FLUX, 0) += HR(HR_E_glacier_flux) * AreaFactor;
        OUT(GLAC_MELT_ENERGY, 0) += HR(HR_E_glacier_melt_energy) * AreaFactor;
      }
    }
    // ---- band-specific terms
    if (G & PD_BAND) {
      const double bandFactor = Cv * 1.0 / ThisAreaFract;
      OUT(AREA_BAND, band) += (Cv * 1.0);
      OUT(SWE_BAND, band) += swq * bandFactor * 1000.;
      OUT(SNOW_DEPTH_BAND, band) += HR(HR_S_depth) * bandFactor * 100.;
      if (HasVeg) OUT(SNOW_CANOPY_BAND, band) += (HR(HR_S_snow_canopy)) * bandFactor * 1000.;
      OUT(SNOW_MELT_BAND, band) += HR(HR_S_melt) * bandFactor;
      OUT(SNOW_COVER_BAND, band) += HR(HR_S_coverage) * bandFactor;
      OUT(DELTACC_BAND, band) += HR(HR_E_deltaCC) * bandFactor;
      OUT(ADVECTION_BAND, band) += HR(HR_E_advection) * bandFactor;
      OUT(SNOW_FLUX_BAND, band) += HR(HR_E_snow_flux) * bandFactor;
      OUT(RFRZ_ENERGY_BAND, band) += HR(HR_E_refreeze_energy) * bandFactor;
      OUT(MELT_ENERGY_BAND, band) += HR(HR_E_melt_energy) * bandFactor;
      OUT(ADV_SENS_BAND, band) -= HR(HR_E_advected_sensible) * bandFactor;
      OUT(SNOW_SURFT_BAND, band) += HR(HR_S_surf_temp) * bandFactor;
      OUT(SNOW_PACKT_BAND, band) += HR(HR_S_pack_temp) * bandFactor;
      OUT(LATENT_SUB_BAND, band) += HR(HR_E_latent_sub) * bandFactor;
      OUT(NET_SHORT_BAND, band) += HR(HR_E_NetShortAtmos) * bandFactor;
      OUT(NET_LONG_BAND, band) += HR(HR_E_NetLongAtmos) * bandFactor;
      if (snowing && overstory) OUT(ALBEDO_BAND, band) += HR(HR_E_AlbedoOver) * bandFactor;
      else OUT(ALBEDO_BAND, band) += HR(HR_E_AlbedoUnder) * bandFactor;
      OUT(LATENT_BAND, band) -= HR(HR_E_latent) * bandFactor;
      OUT(SENSIBLE_BAND, band) -= HR(HR_E_sensible) * bandFactor;
      OUT(GRND_FLUX_BAND, band) -= HR(HR_E_grnd_flux) * bandFactor;
      if (HasGlac) {
        OUT(GLAC_DELTACC_BAND, band) += HR(HR_E_deltaCC_glac);
        OUT(GLAC_FLUX_BAND, band) += HR(HR_E_glacier_flux);
        OUT(GLAC_WAT_STOR_BAND, band) += HR(HR_G_water_storage) * 1000.;
        OUT(GLAC_AREA_BAND, band) += Cv;
        OUT(GLAC_MBAL_BAND, band) += HR(HR_G_mass_balance) * 1000.;
        OUT(GLAC_IMBAL_BAND, band) += HR(HR_G_ice_mass_balance) * 1000.;
        OUT(GLAC_ACCUM_BAND, band) += HR(HR_G_accumulation) * 1000.;
        OUT(GLAC_MELT_BAND, band) += HR(HR_G_melt) * 1000.;
        OUT(GLAC_SUB_BAND, band) += HR(HR_G_vapor_flux) * 1000.;
        OUT(GLAC_INFLOW_BAND, band) += HR(HR_G_inflow) * 1000.;
        OUT(GLAC_OUTFLOW_BAND, band) += HR(HR_G_outflow) * 1000.;
      }
    }
  }
}

// the group put_data_hru accumulates an output variable in; variables no HRU term feeds (forcing, derived, lake terms) count as PD_WB
VIC_HD int out_var_group(int v) {
  switch (v) {
#define X(n) case VOUT_##n:
    X(FDEPTH) X(TDEPTH) X(SURF_FROST_FRAC) X(BARESOILT) X(VEGT) X(SURF_TEMP) X(SOIL_TNODE) X(SURFT_FBFLAG) X(SOILT_FBFLAG) X(SNOWT_FBFLAG)
    X(TFOL_FBFLAG) X(TCAN_FBFLAG) X(GLAC_TSURF_FBFLAG) X(NET_SHORT) X(NET_LONG) X(IN_LONG) X(ALBEDO) X(LATENT) X(LATENT_SUB) X(SENSIBLE)
    X(GRND_FLUX) X(DELTAH) X(FUSION) X(ENERGY_ERROR) X(RAD_TEMP) X(DELTACC) X(ADVECTION) X(SNOW_FLUX) X(RFRZ_ENERGY) X(MELT_ENERGY)
    X(ADV_SENS) X(GLAC_SURF_TEMP) X(GLAC_DELTACC) X(GLAC_FLUX) X(GLAC_MELT_ENERGY)
    return PD_EB;
    X(ELEV_BAND) X(AREA_BAND) X(SWE_BAND) X(SNOW_DEPTH_BAND) X(SNOW_CANOPY_BAND) X(SNOW_MELT_BAND) X(SNOW_COVER_BAND) X(DELTACC_BAND)
    X(ADVECTION_BAND) X(SNOW_FLUX_BAND) X(RFRZ_ENERGY_BAND) X(MELT_ENERGY_BAND) X(ADV_SENS_BAND) X(SNOW_SURFT_BAND) X(SNOW_PACKT_BAND)
    X(LATENT_SUB_BAND) X(NET_SHORT_BAND) X(NET_LONG_BAND) X(ALBEDO_BAND) X(LATENT_BAND) X(SENSIBLE_BAND) X(GRND_FLUX_BAND)
    X(GLAC_DELTACC_BAND) X(GLAC_FLUX_BAND) X(GLAC_WAT_STOR_BAND) X(GLAC_AREA_BAND) X(GLAC_MBAL_BAND) X(GLAC_IMBAL_BAND) X(GLAC_ACCUM_BAND)
    X(GLAC_MELT_BAND) X(GLAC_SUB_BAND) X(GLAC_INFLOW_BAND) X(GLAC_OUTFLOW_BAND)
    return PD_BAND;
#undef X
    default:
      return PD_WB;
  }
}

// Part 3, once per cell: derived variables, storage changes, the water / energy balance checks
template <class OutRow>
VIC_HD void put_data_finish(const Opts& o, const CellPar& cp, int rec, RowRW carry, OutRow out, const PutDataCtx& pc) {
  const vicgpu_layout& L = o.L;
  const int NL = VICGPU_NLAYER;
  // ---- derived variables
  if (pc.cv_baresoil > 0) OUT(BARESOILT, 0) /= pc.cv_baresoil;
  if (pc.cv_veg > 0) OUT(VEGT, 0) /= pc.cv_veg;
  if (pc.cv_overstory > 0) OUT(AERO_COND2, 0) /= pc.cv_overstory;
  if (pc.cv_snow > 0) {
    OUT(SALBEDO, 0) /= pc.cv_snow;
    OUT(SNOW_SURF_TEMP, 0) /= pc.cv_snow;
    OUT(SNOW_PACK_TEMP, 0) /= pc.cv_snow;
  }
  if (pc.cv_glacier > 0) OUT(GLAC_SURF_TEMP, 0) /= pc.cv_glacier;
  OUT(RAD_TEMP, 0) = vpow(OUT(RAD_TEMP, 0), 0.25);
  OUT(AERO_RESIST1, 0) = (OUT(AERO_COND1, 0) > SMALL) ? 1 / OUT(AERO_COND1, 0) : HUGE_RESIST;
  OUT(AERO_RESIST2, 0) = (OUT(AERO_COND2, 0) > SMALL) ? 1 / OUT(AERO_COND2, 0) : HUGE_RESIST;
  OUT(AERO_RESIST, 0) = (OUT(AERO_COND, 0) > SMALL) ? 1 / OUT(AERO_COND, 0) : HUGE_RESIST;
  OUT(DELSOILMOIST, 0) = 0;
  for (int l = 0; l < NL; l++) {
    OUT(SOIL_LIQ_TOT, 0) += OUT(SOIL_LIQ, l);
    OUT(SOIL_ICE_TOT, 0) += OUT(SOIL_ICE, l);
    OUT(SOIL_MOIST, l) = OUT(SOIL_LIQ, l) + OUT(SOIL_ICE, l);
    OUT(DELSOILMOIST, 0) += OUT(SOIL_MOIST, l);
    OUT(SMLIQFRAC, l) = OUT(SOIL_LIQ, l) / OUT(SOIL_MOIST, l);
    OUT(SMFROZFRAC, l) = 1 - OUT(SMLIQFRAC, l);
  }
  if (rec >= 0) {
    OUT(DELSOILMOIST, 0) -= carry[CC_total_soil_moist];
    OUT(DELSWE, 0) = OUT(SWE, 0) + OUT(SNOW_CANOPY, 0) - carry[CC_swe];
    OUT(DELINTERCEPT, 0) = OUT(WDEW, 0) - carry[CC_wdew];
    OUT(DELSURFSTOR, 0) = OUT(SURFSTOR, 0) - carry[CC_surfstor];
  }
  const int dt_sec = o.dt * SECPHOUR;
  OUT(REFREEZE, 0) = (OUT(RFRZ_ENERGY, 0) / Lf) * dt_sec;
  OUT(R_NET, 0) = OUT(NET_SHORT, 0) + OUT(NET_LONG, 0);
  double tsm = 0;
  for (int l = 0; l < NL; l++) tsm += OUT(SOIL_MOIST, l);
  carry[CC_total_soil_moist] = tsm;
  OUT(SOIL_MOIST_TOT, 0) = tsm;
  carry[CC_surfstor] = OUT(SURFSTOR, 0);
  carry[CC_swe] = OUT(SWE, 0) + OUT(SNOW_CANOPY, 0);
  carry[CC_wdew] = OUT(WDEW, 0);
  // ---- water balance check
  const double inflow = OUT(PREC, 0) + OUT(LAKE_CHAN_IN, 0);
  const double outflow = OUT(EVAP, 0) + OUT(RUNOFF, 0) + OUT(BASEFLOW, 0);
  const double glac_icebal = OUT(GLAC_IMBAL, 0);
  double storage = 0.;
  for (int l = 0; l < NL; l++) {
    if (o.MOISTFRACT) storage += (OUT(SOIL_LIQ, l) + OUT(SOIL_ICE, l)) * cp.layer(CL_depth, l) * 1000;
    else storage += OUT(SOIL_LIQ, l) + OUT(SOIL_ICE, l);
  }
  storage += OUT(SWE, 0) + OUT(SNOW_CANOPY, 0) + OUT(WDEW, 0) + OUT(SURFSTOR, 0) + OUT(GLAC_WAT_STOR, 0);
  if (rec < 0) {
    carry[CC_water_last_storage] = storage;
    carry[CC_water_cum_error] = 0.;
    carry[CC_water_max_error] = 0.;
    OUT(WATER_ERROR, 0) = 0.0;
  } else {
    const double error = inflow - outflow - (storage - carry[CC_water_last_storage]) - glac_icebal;
    carry[CC_water_cum_error] += error;
    if (fabs(error) > fabs(carry[CC_water_max_error]) && fabs(error) > 1e-5) carry[CC_water_max_error] = error;
    carry[CC_water_last_storage] = storage;
    OUT(WATER_ERROR, 0) = error;
  }
  // ---- energy balance check
  if (o.FULL_ENERGY) {
    if (rec < 0) {
      carry[CC_energy_cum_error] = 0;
      carry[CC_energy_max_error] = 0;
    } else {
      const double net_rad = OUT(NET_SHORT, 0) + OUT(NET_LONG, 0);
      const double latent = OUT(LATENT, 0) + OUT(LATENT_SUB, 0);
      const double sensible = OUT(SENSIBLE, 0) + OUT(ADV_SENS, 0);
      const double grnd_flux = OUT(GRND_FLUX, 0) + OUT(DELTAH, 0) + OUT(FUSION, 0);
      const double snow_fluxes = OUT(ADVECTION, 0) - OUT(DELTACC, 0) - OUT(SNOW_FLUX, 0) + OUT(RFRZ_ENERGY, 0);
      const double glac_fluxes = -OUT(GLAC_DELTACC, 0) - OUT(GLAC_MELT_ENERGY, 0);
      const double error = net_rad - latent - sensible - grnd_flux + snow_fluxes + glac_fluxes;
      carry[CC_energy_cum_error] += error;
      if (fabs(error) > fabs(carry[CC_energy_max_error]) && fabs(error) > 0.001) carry[CC_energy_max_error] = error;
    }
  }
}

// Part 4: temporal aggregation (put_data.c:664-680) of the variables v0, v0 + vstep, ... (independent of each other) ...
template <class OutRow, class AggRow>
VIC_HD void put_data_aggregate_vars(const Opts& o, const int* aggtype, OutRow out, AggRow agg, int v0, int vstep) {
  const vicgpu_layout& L = o.L;
  for (int v = v0; v < VICGPU_N_OUTVARS; v += vstep) {
    const int ne = L.out_nelem[v];
    const int off = L.out_off[v];
    const int at = aggtype[v];
    for (int i = 0; i < ne; i++) {
      if (at == VICGPU_AGG_END) agg[off + i] = out[off + i];
      else if (at == VICGPU_AGG_SUM) agg[off + i] += out[off + i];
      else if (at == VICGPU_AGG_AVG) agg[off + i] += out[off + i] / o.out_step_ratio;
    }
  }
}
// ... and what follows once all variables are aggregated: resistances from the aggregated conductances, ALMA unit conversions at an
// output step
template <class AggRow>
VIC_HD void put_data_aggregate_tail(const Opts& o, int step_count, AggRow agg) {
  const vicgpu_layout& L = o.L;
  const int NL = VICGPU_NLAYER;
  const int dt_sec = o.dt * SECPHOUR;
  agg[L.out_off[VOUT_AERO_RESIST]] = 1 / agg[L.out_off[VOUT_AERO_COND]];
  agg[L.out_off[VOUT_AERO_RESIST1]] = 1 / agg[L.out_off[VOUT_AERO_COND1]];
  agg[L.out_off[VOUT_AERO_RESIST2]] = 1 / agg[L.out_off[VOUT_AERO_COND2]];
  if (step_count == o.out_step_ratio && o.ALMA_OUTPUT) {
    // ALMA unit conversions of the aggregated values (put_data.c:694-755)
    const double out_dt_sec = (double)(o.out_step_ratio * dt_sec);
#define AGG(vn, e) agg[L.out_off[VOUT_##vn] + (e)]
    AGG(BASEFLOW, 0) /= out_dt_sec; AGG(EVAP, 0) /= out_dt_sec; AGG(EVAP_BARE, 0) /= out_dt_sec; AGG(EVAP_CANOP, 0) /= out_dt_sec;
    AGG(INFLOW, 0) /= out_dt_sec; AGG(PREC, 0) /= out_dt_sec; AGG(RAINF, 0) /= out_dt_sec; AGG(REFREEZE, 0) /= out_dt_sec;
    AGG(RUNOFF, 0) /= out_dt_sec; AGG(SNOW_MELT, 0) /= out_dt_sec; AGG(SNOWF, 0) /= out_dt_sec; AGG(SUB_BLOWING, 0) /= out_dt_sec;
    AGG(SUB_CANOP, 0) /= out_dt_sec; AGG(SUB_SNOW, 0) /= out_dt_sec; AGG(SUB_SNOW, 0) += AGG(SUB_CANOP, 0); AGG(SUB_SURFACE, 0) /= out_dt_sec;
    AGG(TRANSP_VEG, 0) /= out_dt_sec; AGG(BARESOILT, 0) += KELVIN; AGG(SNOW_PACK_TEMP, 0) += KELVIN; AGG(SNOW_SURF_TEMP, 0) += KELVIN;
    AGG(LAKE_ICE_TEMP, 0) += KELVIN; AGG(LAKE_SURF_TEMP, 0) += KELVIN;
    for (int l = 0; l < NL; l++) AGG(SOIL_TEMP, l) += KELVIN;
    for (int n = 0; n < o.Nnode; n++) { AGG(SOIL_TNODE, n) += KELVIN; AGG(SOIL_TNODE_WL, n) += KELVIN; }
    AGG(SURF_TEMP, 0) += KELVIN; AGG(VEGT, 0) += KELVIN; AGG(FDEPTH, 0) /= 100; AGG(TDEPTH, 0) /= 100;
    AGG(DELTACC, 0) *= out_dt_sec; AGG(DELTAH, 0) *= out_dt_sec; AGG(AIR_TEMP, 0) += KELVIN; AGG(PRESSURE, 0) *= 1000; AGG(VP, 0) *= 1000;
    AGG(VPD, 0) *= 1000;
#undef AGG
  }
}
#undef OUT
#undef HR
#undef HP

}  // namespace vic
#endif

// vicgpu_pack.h -- host-side glue between the reference's C++ structs and the flat
// records of the vic-b200 C-ABI (include/vicgpu.h).
//
// This is the ONLY file of the product that needs the reference's own headers: it is
// compiled inside the reference build -- and inside the oracle
// harness (oracle/ref_harness.cpp) -- with `-I<reference dir>`; it is never compiled
// into libvicgpu.so.  It replaces nothing in the reference: cell_info_struct,
// soil_con_struct, HRU and ProgramState keep their layout (vicNl_def.h:951-1100,
// :1375-1389, :1546-1581), so read_soilparam/read_vegparam/read_snowband,
// initialize_model_state, write_model_state and the NetCDF writers work unchanged.
#ifndef VICGPU_PACK_H
#define VICGPU_PACK_H

#include <limits.h>
#include <string.h>
#include <vector>
#include "vicNl.h"
#include "vicgpu.h"

// options consumed on the hot path: SURVEY.md section 8(b)
inline void vicgpu_pack_options(const ProgramState *state, vicgpu_options *o) {
  memset(o, 0, sizeof(*o));
  const option_struct &op = state->options;
  const global_param_struct &gp = state->global_param;
  o->abi_version = VICGPU_ABI_VERSION;
  o->Nlayer = op.Nlayer;
  o->Nnode = op.Nnode;
  o->Nbands = op.SNOW_BAND;
  o->dt = gp.dt;
  o->SNOW_STEP = op.SNOW_STEP;
  o->NR = state->NR;
  o->NF = state->NF;
  o->nrecs = gp.nrecs;
  o->out_step_ratio = state->out_step_ratio;
  o->FULL_ENERGY = op.FULL_ENERGY;
  o->FROZEN_SOIL = op.FROZEN_SOIL;
  o->QUICK_FLUX = op.QUICK_FLUX;
  o->QUICK_SOLVE = op.QUICK_SOLVE;
  o->IMPLICIT = op.IMPLICIT;
  o->EXP_TRANS = op.EXP_TRANS;
  o->NOFLUX = op.NOFLUX;
  o->GRND_FLUX_TYPE = op.GRND_FLUX_TYPE;
  o->AERO_RESIST_CANSNOW = op.AERO_RESIST_CANSNOW;
  o->SNOW_ALBEDO = op.SNOW_ALBEDO;
  o->SNOW_DENSITY = op.SNOW_DENSITY;
  o->TEMP_TH_TYPE = op.TEMP_TH_TYPE;
  o->TFALLBACK = op.TFALLBACK;
  o->BLOWING = op.BLOWING;
  o->DIST_PRCP = op.DIST_PRCP;
  o->CORRPREC = op.CORRPREC;
  o->LAKES = op.LAKES;
  o->COMPUTE_TREELINE = op.COMPUTE_TREELINE;
  o->GLACIER_ID = op.GLACIER_ID;
  o->GLACIER_DYNAMICS = op.GLACIER_DYNAMICS;
  o->MOISTFRACT = op.MOISTFRACT;
  o->ALMA_OUTPUT = op.ALMA_OUTPUT;
  o->NVegLibTypes = state->veg_lib ? state->veg_lib[0].NVegLibTypes : 0;
  o->glacierAccumStartYear = gp.glacierAccumStartYear;
  o->glacierAccumStartMonth = gp.glacierAccumStartMonth;
  o->glacierAccumStartDay = gp.glacierAccumStartDay;
  o->glacierAccumInterval = gp.glacierAccumInterval;
  o->wind_h = gp.wind_h;
  o->MIN_WIND_SPEED = op.MIN_WIND_SPEED;
}

// what initialize_atmos() reads besides the hot-path options (initialize_atmos.c:125-160, 905, 1103, 1274; mtclim_vic.c:1493, 1552-1571, 1698)
inline void vicgpu_pack_disagg_options(const ProgramState *state, const dmy_struct *dmy, vicgpu_disagg_options *d) {
  memset(d, 0, sizeof(*d));
  const option_struct &op = state->options;
  const global_param_struct &gp = state->global_param;
  d->abi_version = VICGPU_ABI_VERSION;
  d->starthour = gp.starthour; d->startyear = gp.startyear; d->startmonth = gp.startmonth; d->startday = gp.startday;
  const int tmp_starthour = 0, tmp_endhour = 24 - gp.dt;
  const int tmp_nrecs = gp.nrecs + gp.starthour - tmp_starthour + tmp_endhour - dmy[gp.nrecs - 1].hour;
  d->Ndays = (tmp_nrecs * gp.dt) / 24;
  d->PLAPSE = op.PLAPSE; d->MTCLIM_SWE_CORR = op.MTCLIM_SWE_CORR; d->VP_INTERP = op.VP_INTERP; d->OUTPUT_FORCE = op.OUTPUT_FORCE;
  d->VP_ITER = op.VP_ITER; d->LW_TYPE = op.LW_TYPE; d->LW_CLOUD = op.LW_CLOUD;
  d->SW_PREC_THRESH = (double)op.SW_PREC_THRESH;
}

// veg_lib_struct rows incl. the four reference PET classes (read_veglib.c:118-136)
inline void vicgpu_pack_veglib(const ProgramState *state, const vicgpu_layout *L, std::vector<double> &out) {
  const int nclass = state->veg_lib[0].NVegLibTypes + N_PET_TYPES_NON_NAT;
  out.assign((size_t)nclass * L->vl_stride, 0.0);
  for (int k = 0; k < nclass; k++) {
    const veg_lib_struct &v = state->veg_lib[k];
    double *r = &out[(size_t)k * L->vl_stride];
#define X(n, p) r[n] = (double)v.p;
    VICGPU_VEGLIB_SCALARS(X)
#undef X
    for (int i = 0; i < 12; i++) {
#define X(n, p) r[VICGPU_VL_MONTH(L, n, i)] = (double)v.p;
      VICGPU_VEGLIB_MONTHLY(X)
#undef X
    }
  }
}

inline void vicgpu_pack_cellpar(const soil_con_struct &soil_con, const vicgpu_layout *L, double *r) {
  const soil_con_struct &s = soil_con;
#define X(n, p) r[n] = (double)s.p;
  VICGPU_CPAR_SCALARS(X)
#undef X
  for (int i = 0; i < VICGPU_NLAYER; i++) {
#define X(n, p) r[VICGPU_CP_LAYER(L, n, i)] = (double)s.p;
    VICGPU_CPAR_LAYER(X)
#undef X
  }
  for (int i = 0; i < L->nnode; i++) {
#define X(n, p) r[VICGPU_CP_NODE(L, n, i)] = (double)s.p;
    VICGPU_CPAR_NODE(X)
#undef X
  }
  for (int i = 0; i < VICGPU_NZCURVE * VICGPU_NZWT; i++) {
#define X(n, p) r[VICGPU_CP_ZWT(L, n, i)] = (double)s.p;
    VICGPU_CPAR_ZWT(X)
#undef X
  }
  for (int i = 0; i < L->nbands; i++) {
#define X(n, p) r[VICGPU_CP_BAND(L, n, i)] = (double)s.p;
    VICGPU_CPAR_BAND(X)
#undef X
  }
}

inline void vicgpu_pack_hrupar(const HRU &hru, int cell, double *r) {
  r[HP_cell] = cell;
  r[HP_Cv] = hru.veg_con.Cv;
  // the artificial bare-soil HRU has no root zones; root[] is never read for it
  // (read_vegparam.c:313-340, calc_root_fraction.c), keep what the reference holds
  r[HP_root0] = hru.veg_con.root[0];
  r[HP_root1] = hru.veg_con.root[1];
  r[HP_root2] = hru.veg_con.root[2];
  r[HP_vegIndex] = hru.veg_con.vegIndex;
  r[HP_vegClass] = hru.veg_con.vegClass;
  r[HP_band] = hru.bandIndex;
  r[HP_isGlacier] = hru.isGlacier ? 1 : 0;
  r[HP_isArtBare] = hru.isArtificialBareSoil ? 1 : 0;
  r[HP_sigma_slope] = hru.veg_con.sigma_slope;
  r[HP_lag_one] = hru.veg_con.lag_one;
  r[HP_fetch] = hru.veg_con.fetch;
}

// HRU record <- HRU  (prognostic + diagnostic members; the list is the superset of
// processCellForStateFile(), write_model_state.c:107-371)
inline void vicgpu_pack_hrurec(const HRU &hru, const vicgpu_layout *L, double *r) {
#define X(n, p, c) r[HR_##n] = (double)hru.p;
  VICGPU_HRU_SCALARS(X)
#undef X
  for (int i = 0; i < VICGPU_NLAYER; i++) {
#define X(n, p, c) r[VICGPU_HR_LAYER(L, n, i)] = (double)hru.p;
    VICGPU_HRU_LAYER(X, HRL_)
#undef X
  }
  for (int i = 0; i < VICGPU_NFRONTS; i++) {
#define X(n, p, c) r[VICGPU_HR_FRONT(L, n, i)] = (double)hru.p;
    VICGPU_HRU_FRONT(X, HRF_)
#undef X
  }
  for (int i = 0; i < VICGPU_NPET; i++) r[VICGPU_HR_PET(L, i)] = hru.cell[0].pot_evap[i];
  for (int i = 0; i < L->nnode; i++) {
#define X(n, p, c) r[VICGPU_HR_NODE(L, n, i)] = (double)hru.p;
    VICGPU_HRU_NODE(X, HRN_)
#undef X
  }
}

template <typename T>
inline void vicgpu_assign_(T &dst, double v) { dst = (T)v; }
inline void vicgpu_assign_(bool &dst, double v) { dst = (v != 0.0); }

// HRU <- HRU record (so that put_data / write_model_state / glacier coupling keep working)
inline void vicgpu_unpack_hrurec(HRU &hru, const vicgpu_layout *L, const double *r) {
#define X(n, p, c) vicgpu_assign_(hru.p, r[HR_##n]);
  VICGPU_HRU_SCALARS(X)
#undef X
  for (int i = 0; i < VICGPU_NLAYER; i++) {
#define X(n, p, c) vicgpu_assign_(hru.p, r[VICGPU_HR_LAYER(L, n, i)]);
    VICGPU_HRU_LAYER(X, HRL_)
#undef X
  }
  for (int i = 0; i < VICGPU_NFRONTS; i++) {
#define X(n, p, c) vicgpu_assign_(hru.p, r[VICGPU_HR_FRONT(L, n, i)]);
    VICGPU_HRU_FRONT(X, HRF_)
#undef X
  }
  for (int i = 0; i < VICGPU_NPET; i++) hru.cell[0].pot_evap[i] = r[VICGPU_HR_PET(L, i)];
  for (int i = 0; i < L->nnode; i++) {
#define X(n, p, c) vicgpu_assign_(hru.p, r[VICGPU_HR_NODE(L, n, i)]);
    VICGPU_HRU_NODE(X, HRN_)
#undef X
  }
}

// one forcing record <- atmos_data_struct (alloc_atmos.c; arrays of NR+1 values)
inline void vicgpu_pack_forcing(const atmos_data_struct &a, const vicgpu_layout *L, double *r) {
  for (int s = 0; s < L->f_nslot; s++) {
#define X(n, p) r[VICGPU_F_IDX(L, n, s)] = (double)a.p[s];
    VICGPU_FORCING(X)
#undef X
  }
}

inline void vicgpu_unpack_forcing(atmos_data_struct &a, const vicgpu_layout *L, const double *r) {
  for (int s = 0; s < L->f_nslot; s++) {
#define X(n, p) vicgpu_assign_(a.p[s], r[VICGPU_F_IDX(L, n, s)]);
    VICGPU_FORCING(X)
#undef X
  }
}

// OutputData[N_OUTVAR_TYPES] -> flat per-cell output row
inline void vicgpu_pack_outdata(const OutputData *out_data, const vicgpu_layout *L, double *row, bool agg) {
  for (int v = 0; v < VICGPU_N_OUTVARS; v++)
    for (int e = 0; e < L->out_nelem[v]; e++)
      row[L->out_off[v] + e] = agg ? out_data[v].aggdata[e] : out_data[v].data[e];
}

inline void vicgpu_unpack_outdata(OutputData *out_data, const vicgpu_layout *L, const double *row, bool agg) {
  for (int v = 0; v < VICGPU_N_OUTVARS; v++)
    for (int e = 0; e < L->out_nelem[v]; e++)
      (agg ? out_data[v].aggdata[e] : out_data[v].data[e]) = row[L->out_off[v] + e];
}

#endif  // VICGPU_PACK_H

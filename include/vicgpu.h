/* vicgpu.h -- C-ABI of libvicgpu.so, the B200 (sm_100a) drop-in for the per-grid-cell
 * time-step loop of PCIC VIC 4.1.2.
 *
 * What it replaces in the reference (paths relative to /root/reference):
 *   - the body of the OpenMP cell loop in runModel()            vicNl.c:506-593
 *       put_data(rec = -nrecs) storage initialisation           vicNl.c:524-541
 *       dist_prec() -> full_energy() -> surface_fluxes[_glac]() vicNl.h:231-232, dist_prec.c:8-175
 *       accumulateGlacierMassBalance()                          vicNl.c:563
 *   - the output aggregation put_data() performs per cell       put_data.c:7-760
 *   - (forcing) initialize_atmos()/mtclim_wrapper()             vicNl.h:362, 390-392
 *
 * Conventions
 *   - plain C, pointers + sizes only; all arrays are caller-owned HOST memory unless a
 *     function name ends in _dev; the library copies in/out and never keeps host pointers.
 *   - every function returns 0 on success or a negative VICGPU_E* code; the reference's
 *     per-cell ERROR (-999, vicNl_def.h:146) is reported per cell through vicgpu_get_cell_status().
 *   - there is NO CPU fallback: without a CUDA device vicgpu_create() fails with VICGPU_ENODEV;
 *     option combinations that are not implemented on the device are rejected with VICGPU_EUNSUPPORTED.
 */
#ifndef VICGPU_H
#define VICGPU_H

#include <stddef.h>
#include "vicgpu_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VICGPU_ABI_VERSION 1

enum {
  VICGPU_OK = 0,
  VICGPU_EINVAL = -1,        /* bad argument */
  VICGPU_ENODEV = -2,        /* no usable CUDA device (never falls back to the CPU) */
  VICGPU_EUNSUPPORTED = -3,  /* option combination not implemented on the device */
  VICGPU_ECUDA = -4,         /* CUDA runtime error, see vicgpu_last_error() */
  VICGPU_ESTATE = -5         /* call order violated (e.g. step before set_state) */
};

/* Flat copy of the members of option_struct / global_param_struct that the hot path reads
 * (vicNl_def.h:653-780 option_struct, :871-914 global_param_struct; NR/NF :1561-1562). */
typedef struct vicgpu_options {
  int abi_version;        /* must be VICGPU_ABI_VERSION */
  int Nlayer;             /* options.Nlayer (must be 3) */
  int Nnode;              /* options.Nnode */
  int Nbands;             /* options.SNOW_BAND */
  int dt;                 /* global_param.dt [h] */
  int SNOW_STEP;          /* options.SNOW_STEP [h] */
  int NR, NF;             /* ProgramState::NR, NF */
  int nrecs;              /* global_param.nrecs */
  int out_step_ratio;     /* ProgramState::out_step_ratio */
  int FULL_ENERGY, FROZEN_SOIL, QUICK_FLUX, QUICK_SOLVE, IMPLICIT, EXP_TRANS, NOFLUX;
  int GRND_FLUX_TYPE;     /* GF_406=0, GF_410=1, GF_FULL=2 */
  int AERO_RESIST_CANSNOW;/* AR_406=0 .. AR_COMBO=4 */
  int SNOW_ALBEDO;        /* USACE=0, SUN1999=1 */
  int SNOW_DENSITY;       /* DENS_BRAS=0, DENS_SNTHRM=1 */
  int TEMP_TH_TYPE;       /* VIC_412=0, KIENZLE=1 */
  int TFALLBACK, BLOWING, DIST_PRCP, CORRPREC, LAKES, COMPUTE_TREELINE;
  int GLACIER_ID;         /* options.GLACIER_ID (veg-library INDEX compared in full_energy.c:313) */
  int GLACIER_DYNAMICS;
  int MOISTFRACT, ALMA_OUTPUT;
  int NVegLibTypes;       /* veg_lib[0].NVegLibTypes: rows [NVegLibTypes, +4) are the reference PET classes */
  int glacierAccumStartYear, glacierAccumStartMonth, glacierAccumStartDay, glacierAccumInterval; /* INT_MIN = unset */
  double wind_h;          /* global_param.wind_h */
  double MIN_WIND_SPEED;
} vicgpu_options;

/* Output variables: same order as enum OutputVariableIndices (vicNl_def.h:351-564, EXCESS_ICE FALSE).
 * X(name, element class, default aggregation)  element class: 1 scalar, L per layer, N per node,
 * B per band, F per front (MAX_FRONTS when FROZEN_SOIL else 1; output_list_utils.c:298-301). */
#define VICGPU_OUTVARS(X) \
  X(ASAT, 1, END) \
  X(LAKE_AREA_FRAC, 1, END) \
  X(LAKE_DEPTH, 1, END) \
  X(LAKE_ICE, 1, END) \
  X(LAKE_ICE_FRACT, 1, END) \
  X(LAKE_ICE_HEIGHT, 1, END) \
  X(LAKE_MOIST, 1, END) \
  X(LAKE_SURF_AREA, 1, END) \
  X(LAKE_SWE, 1, END) \
  X(LAKE_SWE_V, 1, END) \
  X(LAKE_VOLUME, 1, END) \
  X(ROOTMOIST, 1, END) \
  X(SMFROZFRAC, L, END) \
  X(SMLIQFRAC, L, END) \
  X(SNOW_CANOPY, 1, END) \
  X(SNOW_COVER, 1, END) \
  X(SNOW_DEPTH, 1, END) \
  X(SOIL_ICE, L, END) \
  X(SOIL_ICE_TOT, 1, END) \
  X(SOIL_LIQ, L, END) \
  X(SOIL_LIQ_TOT, 1, END) \
  X(SOIL_MOIST, L, END) \
  X(SOIL_MOIST_TOT, 1, END) \
  X(SOIL_WET, 1, END) \
  X(SURFSTOR, 1, END) \
  X(SURF_FROST_FRAC, 1, END) \
  X(SWE, 1, END) \
  X(WDEW, 1, END) \
  X(ZWT, 1, END) \
  X(ZWT2, 1, END) \
  X(ZWT3, 1, END) \
  X(ZWTL, L, END) \
  X(BASEFLOW, 1, SUM) \
  X(DELINTERCEPT, 1, SUM) \
  X(DELSOILMOIST, 1, SUM) \
  X(DELSURFSTOR, 1, SUM) \
  X(DELSWE, 1, SUM) \
  X(EVAP, 1, SUM) \
  X(EVAP_BARE, 1, SUM) \
  X(EVAP_CANOP, 1, SUM) \
  X(INFLOW, 1, SUM) \
  X(LAKE_BF_IN, 1, SUM) \
  X(LAKE_BF_IN_V, 1, SUM) \
  X(LAKE_BF_OUT, 1, SUM) \
  X(LAKE_BF_OUT_V, 1, SUM) \
  X(LAKE_CHAN_IN, 1, SUM) \
  X(LAKE_CHAN_IN_V, 1, SUM) \
  X(LAKE_CHAN_OUT, 1, SUM) \
  X(LAKE_CHAN_OUT_V, 1, SUM) \
  X(LAKE_DSTOR, 1, SUM) \
  X(LAKE_DSTOR_V, 1, SUM) \
  X(LAKE_DSWE, 1, SUM) \
  X(LAKE_DSWE_V, 1, SUM) \
  X(LAKE_EVAP, 1, SUM) \
  X(LAKE_EVAP_V, 1, SUM) \
  X(LAKE_PREC_V, 1, SUM) \
  X(LAKE_RCHRG, 1, SUM) \
  X(LAKE_RCHRG_V, 1, SUM) \
  X(LAKE_RO_IN, 1, SUM) \
  X(LAKE_RO_IN_V, 1, SUM) \
  X(LAKE_VAPFLX, 1, SUM) \
  X(LAKE_VAPFLX_V, 1, SUM) \
  X(PET_SATSOIL, 1, SUM) \
  X(PET_H2OSURF, 1, SUM) \
  X(PET_SHORT, 1, SUM) \
  X(PET_TALL, 1, SUM) \
  X(PET_NATVEG, 1, SUM) \
  X(PET_VEGNOCR, 1, SUM) \
  X(PREC, 1, SUM) \
  X(RAINF, 1, SUM) \
  X(REFREEZE, 1, SUM) \
  X(RUNOFF, 1, SUM) \
  X(SNOW_MELT, 1, SUM) \
  X(SNOWF, 1, SUM) \
  X(SUB_BLOWING, 1, SUM) \
  X(SUB_CANOP, 1, SUM) \
  X(SUB_SNOW, 1, SUM) \
  X(SUB_SURFACE, 1, SUM) \
  X(TRANSP_VEG, 1, SUM) \
  X(WATER_ERROR, 1, AVG) \
  X(ALBEDO, 1, AVG) \
  X(BARESOILT, 1, AVG) \
  X(FDEPTH, F, AVG) \
  X(LAKE_ICE_TEMP, 1, AVG) \
  X(LAKE_SURF_TEMP, 1, AVG) \
  X(RAD_TEMP, 1, AVG) \
  X(SALBEDO, 1, AVG) \
  X(SNOW_PACK_TEMP, 1, AVG) \
  X(SNOW_SURF_TEMP, 1, AVG) \
  X(SNOWT_FBFLAG, 1, SUM) \
  X(SOIL_TEMP, L, AVG) \
  X(SOIL_TNODE, N, AVG) \
  X(SOIL_TNODE_WL, N, AVG) \
  X(SOILT_FBFLAG, N, SUM) \
  X(SURF_TEMP, 1, AVG) \
  X(SURFT_FBFLAG, 1, SUM) \
  X(TCAN_FBFLAG, 1, SUM) \
  X(TDEPTH, F, AVG) \
  X(TFOL_FBFLAG, 1, SUM) \
  X(VEGT, 1, AVG) \
  X(ADV_SENS, 1, AVG) \
  X(ADVECTION, 1, AVG) \
  X(DELTACC, 1, AVG) \
  X(DELTAH, 1, AVG) \
  X(ENERGY_ERROR, 1, AVG) \
  X(FUSION, 1, AVG) \
  X(GRND_FLUX, 1, AVG) \
  X(IN_LONG, 1, AVG) \
  X(LATENT, 1, AVG) \
  X(LATENT_SUB, 1, AVG) \
  X(MELT_ENERGY, 1, AVG) \
  X(NET_LONG, 1, AVG) \
  X(NET_SHORT, 1, AVG) \
  X(R_NET, 1, AVG) \
  X(RFRZ_ENERGY, 1, AVG) \
  X(SENSIBLE, 1, AVG) \
  X(SNOW_FLUX, 1, AVG) \
  X(AERO_COND, 1, AVG) \
  X(AERO_COND1, 1, AVG) \
  X(AERO_COND2, 1, AVG) \
  X(AERO_RESIST, 1, AVG) \
  X(AERO_RESIST1, 1, AVG) \
  X(AERO_RESIST2, 1, AVG) \
  X(AIR_TEMP, 1, AVG) \
  X(DENSITY, 1, AVG) \
  X(LONGWAVE, 1, AVG) \
  X(PRESSURE, 1, AVG) \
  X(QAIR, 1, AVG) \
  X(REL_HUMID, 1, AVG) \
  X(SHORTWAVE, 1, AVG) \
  X(SURF_COND, 1, AVG) \
  X(TSKC, 1, AVG) \
  X(VP, 1, AVG) \
  X(VPD, 1, AVG) \
  X(WIND, 1, AVG) \
  X(ADV_SENS_BAND, B, AVG) \
  X(ADVECTION_BAND, B, AVG) \
  X(ALBEDO_BAND, B, AVG) \
  X(AREA_BAND, B, END) \
  X(DELTACC_BAND, B, SUM) \
  X(ELEV_BAND, B, END) \
  X(GRND_FLUX_BAND, B, AVG) \
  X(IN_LONG_BAND, B, AVG) \
  X(LATENT_BAND, B, AVG) \
  X(LATENT_SUB_BAND, B, AVG) \
  X(MELT_ENERGY_BAND, B, AVG) \
  X(NET_LONG_BAND, B, AVG) \
  X(NET_SHORT_BAND, B, AVG) \
  X(RFRZ_ENERGY_BAND, B, AVG) \
  X(SENSIBLE_BAND, B, AVG) \
  X(SNOW_CANOPY_BAND, B, END) \
  X(SNOW_COVER_BAND, B, END) \
  X(SNOW_DEPTH_BAND, B, END) \
  X(SNOW_FLUX_BAND, B, AVG) \
  X(SNOW_MELT_BAND, B, AVG) \
  X(SNOW_PACKT_BAND, B, AVG) \
  X(SNOW_SURFT_BAND, B, AVG) \
  X(SWE_BAND, B, END) \
  X(GLAC_WAT_STOR, 1, END) \
  X(GLAC_AREA, 1, END) \
  X(GLAC_MBAL, 1, SUM) \
  X(GLAC_IMBAL, 1, SUM) \
  X(GLAC_ACCUM, 1, SUM) \
  X(GLAC_MELT, 1, SUM) \
  X(GLAC_SUB, 1, SUM) \
  X(GLAC_INFLOW, 1, SUM) \
  X(GLAC_OUTFLOW, 1, SUM) \
  X(GLAC_SURF_TEMP, 1, END) \
  X(GLAC_TSURF_FBFLAG, 1, END) \
  X(GLAC_DELTACC, 1, AVG) \
  X(GLAC_FLUX, 1, AVG) \
  X(GLAC_MELT_ENERGY, 1, AVG) \
  X(GLAC_OUTFLOW_COEF, 1, END) \
  X(GLAC_DELTACC_BAND, B, AVG) \
  X(GLAC_FLUX_BAND, B, AVG) \
  X(GLAC_WAT_STOR_BAND, B, END) \
  X(GLAC_AREA_BAND, B, END) \
  X(GLAC_MBAL_BAND, B, SUM) \
  X(GLAC_IMBAL_BAND, B, SUM) \
  X(GLAC_ACCUM_BAND, B, SUM) \
  X(GLAC_MELT_BAND, B, SUM) \
  X(GLAC_SUB_BAND, B, SUM) \
  X(GLAC_INFLOW_BAND, B, SUM) \
  X(GLAC_OUTFLOW_BAND, B, SUM) \

enum vicgpu_outvar {
#define X(n, e, a) VOUT_##n,
  VICGPU_OUTVARS(X)
#undef X
  VICGPU_N_OUTVARS
};

enum { VICGPU_AGG_AVG = 0, VICGPU_AGG_BEG, VICGPU_AGG_END, VICGPU_AGG_MAX, VICGPU_AGG_MIN, VICGPU_AGG_SUM }; /* vicNl_def.h:590-597 */

/* Column layout of the flat records, as a function of (Nnode, Nbands). */
typedef struct vicgpu_layout {
  int nnode, nbands, nfront_out;
  /* HRU record */
  int hr_layer0, hr_front0, hr_pet0, hr_node0, hr_stride;
  /* cell parameter record */
  int cp_layer0, cp_node0, cp_zwt0, cp_band0, cp_stride;
  /* veg library row */
  int vl_month0, vl_stride;
  /* HRU parameter row */
  int hp_stride;
  /* forcing record: [var][slot] */
  int f_nslot, f_stride;
  /* per-cell output row: offsets of each variable's first element, and row length */
  int out_off[VICGPU_N_OUTVARS + 1];
  int out_nelem[VICGPU_N_OUTVARS];
} vicgpu_layout;

enum vicgpu_hru_scalar {
#define X(n, p, c) HR_##n,
  VICGPU_HRU_SCALARS(X)
#undef X
  HR_NSCALAR
};
enum vicgpu_hru_layer {
#define X(n, p, c) n,
  VICGPU_HRU_LAYER(X, HRL_)
#undef X
  HRL_N
};
enum vicgpu_hru_front {
#define X(n, p, c) n,
  VICGPU_HRU_FRONT(X, HRF_)
#undef X
  HRF_N
};
enum vicgpu_hru_node {
#define X(n, p, c) n,
  VICGPU_HRU_NODE(X, HRN_)
#undef X
  HRN_N
};
enum vicgpu_hpar {
#define X(n, p, c) n,
  VICGPU_HPAR_SCALARS(X)
#undef X
  HP_N
};
enum vicgpu_cpar_scalar {
#define X(n, p) n,
  VICGPU_CPAR_SCALARS(X)
#undef X
  CP_NSCALAR
};
enum vicgpu_cpar_layer {
#define X(n, p) n,
  VICGPU_CPAR_LAYER(X)
#undef X
  CL_N
};
enum vicgpu_cpar_node {
#define X(n, p) n,
  VICGPU_CPAR_NODE(X)
#undef X
  CN_N
};
enum vicgpu_cpar_zwt {
#define X(n, p) n,
  VICGPU_CPAR_ZWT(X)
#undef X
  CZ_N
};
enum vicgpu_cpar_band {
#define X(n, p) n,
  VICGPU_CPAR_BAND(X)
#undef X
  CB_N
};
enum vicgpu_veglib_scalar {
#define X(n, p) n,
  VICGPU_VEGLIB_SCALARS(X)
#undef X
  VL_NSCALAR
};
enum vicgpu_veglib_monthly {
#define X(n, p) n,
  VICGPU_VEGLIB_MONTHLY(X)
#undef X
  VM_N
};
enum vicgpu_forcing_var {
#define X(n, p) n,
  VICGPU_FORCING(X)
#undef X
  FV_N
};

#define VICGPU_NZCURVE (VICGPU_NLAYER + 2)

/* column helpers */
#define VICGPU_HR_LAYER(L, f, i) ((L)->hr_layer0 + (f) * VICGPU_NLAYER + (i))
#define VICGPU_HR_FRONT(L, f, i) ((L)->hr_front0 + (f) * VICGPU_NFRONTS + (i))
#define VICGPU_HR_PET(L, i) ((L)->hr_pet0 + (i))
#define VICGPU_HR_NODE(L, f, i) ((L)->hr_node0 + (f) * (L)->nnode + (i))
#define VICGPU_CP_LAYER(L, f, i) ((L)->cp_layer0 + (f) * VICGPU_NLAYER + (i))
#define VICGPU_CP_NODE(L, f, i) ((L)->cp_node0 + (f) * (L)->nnode + (i))
#define VICGPU_CP_ZWT(L, f, i) ((L)->cp_zwt0 + (f) * (VICGPU_NZCURVE * VICGPU_NZWT) + (i))
#define VICGPU_CP_BAND(L, f, i) ((L)->cp_band0 + (f) * (L)->nbands + (i))
#define VICGPU_VL_MONTH(L, f, m) ((L)->vl_month0 + (f) * 12 + (m))
#define VICGPU_F_IDX(L, var, slot) ((var) * (L)->f_nslot + (slot))

static inline void vicgpu_layout_init(vicgpu_layout *L, const vicgpu_options *o) {
  int v, off;
  L->nnode = o->Nnode;
  L->nbands = o->Nbands;
  L->nfront_out = o->FROZEN_SOIL ? VICGPU_NFRONTS : 1;
  L->hr_layer0 = HR_NSCALAR;
  L->hr_front0 = L->hr_layer0 + HRL_N * VICGPU_NLAYER;
  L->hr_pet0 = L->hr_front0 + HRF_N * VICGPU_NFRONTS;
  L->hr_node0 = L->hr_pet0 + VICGPU_NPET;
  L->hr_stride = L->hr_node0 + HRN_N * o->Nnode;
  L->cp_layer0 = CP_NSCALAR;
  L->cp_node0 = L->cp_layer0 + CL_N * VICGPU_NLAYER;
  L->cp_zwt0 = L->cp_node0 + CN_N * o->Nnode;
  L->cp_band0 = L->cp_zwt0 + CZ_N * VICGPU_NZCURVE * VICGPU_NZWT;
  L->cp_stride = L->cp_band0 + CB_N * o->Nbands;
  L->vl_month0 = VL_NSCALAR;
  L->vl_stride = VL_NSCALAR + VM_N * 12;
  L->hp_stride = HP_N;
  L->f_nslot = (o->NF > 1) ? o->NF + 1 : 1;
  L->f_stride = FV_N * L->f_nslot;
  off = 0;
  v = 0;
#define VG_E_1 1
#define VG_E_L VICGPU_NLAYER
#define VG_E_N (o->Nnode)
#define VG_E_B (o->Nbands)
#define VG_E_F (L->nfront_out)
#define X(n, e, a) L->out_off[v] = off; L->out_nelem[v] = VG_E_##e; off += VG_E_##e; v++;
  VICGPU_OUTVARS(X)
#undef X
#undef VG_E_1
#undef VG_E_L
#undef VG_E_N
#undef VG_E_B
#undef VG_E_F
  L->out_off[v] = off;
}

/* default aggregation type of each output variable (output_list_utils.c:353-470) */
static inline void vicgpu_default_aggtypes(int *aggtype /* [VICGPU_N_OUTVARS] */) {
  int v = 0;
#define X(n, e, a) aggtype[v++] = VICGPU_AGG_##a;
  VICGPU_OUTVARS(X)
#undef X
}

typedef struct vicgpu_handle vicgpu_handle;

/* ---- lifetime ------------------------------------------------------------------------ */
int vicgpu_abi_version(void);
const char *vicgpu_last_error(void);
/* device: CUDA ordinal.  Fails with VICGPU_ENODEV if no device is present. */
int vicgpu_create(vicgpu_handle **h, const vicgpu_options *opt, int device);
int vicgpu_destroy(vicgpu_handle *h);
int vicgpu_get_layout(const vicgpu_handle *h, vicgpu_layout *L);

/* ---- static inputs (stand in for ProgramState::veg_lib and cell_info_struct::soil_con / hruList) */
/* veglib: [nclass][L.vl_stride], nclass = NVegLibTypes + 4 (read_veglib.c:118-136) */
int vicgpu_set_veglib(vicgpu_handle *h, int nclass, const double *veglib);
/* cellpar [ncell][L.cp_stride]; hrupar [nhru][HP_N], HRUs grouped by cell in hruList order,
 * hrupar[HP_cell] ascending (read_vegparam.c:117-340 builds that order) */
int vicgpu_set_cells(vicgpu_handle *h, int ncell, const double *cellpar, int nhru, const double *hrupar);
/* aggregation type per output variable (OutputData::aggtype); NULL = reference defaults */
int vicgpu_set_output_spec(vicgpu_handle *h, const int *aggtype);

/* ---- model state (stands in for initialize_model_state() / read_initial_model_state() results) */
/* hrurec [nhru][L.hr_stride] */
int vicgpu_set_state(vicgpu_handle *h, const double *hrurec);
int vicgpu_get_state(vicgpu_handle *h, double *hrurec);

/* ---- forcing: hourly/sub-daily records as produced by initialize_atmos() ---------------- */
/* forcing [nrec][ncell][L.f_stride]  (var-major inside a record: [FV_*][slot]); copies H2D into a
 * device-resident forcing window that starts at record rec0.  The device keeps the TWO most recently
 * set windows, and the call returns as soon as the copy is queued: upload block b + 1, then step over
 * block b, and the transfer overlaps the kernels.  `forcing` must stay untouched until the next call
 * into the library has returned (with pageable memory the driver has already staged it on return). */
int vicgpu_set_forcing(vicgpu_handle *h, int rec0, int nrec, const double *forcing);

/* ---- time stepping --------------------------------------------------------------------- */
/* Advance every valid cell over records [rec0, rec0+nrec).  dmy: [nrec+1][5] ints
 * {day, day_in_year, hour, month, year} (dmy_struct, vicNl_def.h:1083-1089); entry nrec is the
 * date of the record after the block (accumulateGlacierMassBalance.c:52 reads dmy[rec+1]).
 * On the first call (rec0 == 0) the storage terms are initialised exactly as
 * put_data(rec = -nrecs) does (vicNl.c:524-541).
 * out_data : NULL or [nrec][ncell][L.out_off[N]]   per-step OutputData::data (all 184 variables)
 * out_agg  : NULL or [nout][ncell][L.out_off[N]]   OutputData::aggdata at every completed output
 *            interval inside the block, nout = number of records with step_count == out_step_ratio */
int vicgpu_step(vicgpu_handle *h, int rec0, int nrec, const int *dmy, double *out_data, double *out_agg);
/* The same with the outputs narrowed to float32 on the device, as the reference's NetCDF writer stores them
 * (WriteOutputNetCDF.c:279, 351, 412: plain (float) conversions): half the device-to-host bytes.
 * Either way the rows travel through two staging buffers on a copy stream, so the transfer of record r
 * overlaps the kernels of record r + 1. */
int vicgpu_step_f32(vicgpu_handle *h, int rec0, int nrec, const int *dmy, float *out_data, float *out_agg);

/* mark cells invalid before the run (cells whose initialisation failed on the host, vicNl.c:420-427);
 * status [ncell]: 0 = valid, -999 = skip */
int vicgpu_set_cell_status(vicgpu_handle *h, const int *status);
/* per-cell status: 0 = valid, -999 = the reference would have returned ERROR from dist_prec for
 * that cell (vicNl.c:545-559); such cells are skipped for the rest of the run. */
int vicgpu_get_cell_status(vicgpu_handle *h, int *status /* [ncell] */);
/* cumulative balance errors per cell: [ncell][5] = water_last_storage, water_cum_error,
 * water_max_error, energy_cum_error, energy_max_error (CellBalanceErrors, vicNl_def.h:1452-1462) */
int vicgpu_get_balance_errors(vicgpu_handle *h, double *err);
/* device time (ms) spent in kernels during the last vicgpu_step call, and number of kernel launches */
int vicgpu_get_last_step_timing(vicgpu_handle *h, double *kernel_ms, long long *launches);

/* ---- forcing disaggregation (stands in for initialize_atmos() / mtclim_wrapper(), vicNl.h:362, 390-392) ------------ */
/* members of global_param_struct / option_struct that only initialize_atmos reads */
typedef struct vicgpu_disagg_options {
  int abi_version;        /* VICGPU_ABI_VERSION */
  int starthour, startyear, startmonth, startday;   /* global_param.start* */
  int Ndays;              /* number of daily forcing records (initialize_atmos.c:146-149) */
  int PLAPSE, MTCLIM_SWE_CORR, VP_INTERP, OUTPUT_FORCE;
  int VP_ITER;            /* VP_ITER_NEVER=0, ALWAYS, ANNUAL, CONVERGE */
  int LW_TYPE;            /* LW_TVA=0, ANDERSON, BRUTSAERT, SATTERLUND, IDSO, PRATA */
  int LW_CLOUD;           /* LW_CLOUD_BRAS=0, LW_CLOUD_DEARDORFF */
  int reserved;
  double SW_PREC_THRESH;  /* options.SW_PREC_THRESH (float member) */
} vicgpu_disagg_options;

/* daily [ncell][Ndays][4] = PREC [mm/day], TMAX, TMIN [C], WIND [m/s] (FORCE_DT 24; no other forcing variable supplied).
 * Fills the device-resident forcing window with records [0, nrecs) -- what initialize_atmos() would have put into
 * cell->atmos[] for every cell -- and, when forcing_out is not NULL, copies it back as [nrecs][ncell][L.f_stride]
 * (the OUTPUT_FORCE use).  Needs vicgpu_set_cells (latitude, longitude, time zone, elevation, slope, aspect, horizons,
 * annual precipitation, band temperature factors, rain/snow thresholds are cell parameters). */
int vicgpu_disagg(vicgpu_handle *h, const vicgpu_disagg_options *dopt, const double *daily, double *forcing_out);
/* The same from TIME-MAJOR daily input, daily_tm [Ndays][4][ncell]: the layout a (time, lat, lon) forcing file yields when it is
 * read the way it is stored (vicgpu_nc_read_slab below), and the layout the device kernels use, so the upload needs no transpose.
 * Same forcing, bit for bit. */
int vicgpu_disagg_tm(vicgpu_handle *h, const vicgpu_disagg_options *dopt, const double *daily_tm, double *forcing_out);

/* ---- NetCDF forcing ingestion (stands in for read_atmos_data()'s NetCDF branch, read_atmos_data.c:109-338) ---------------
 * Host-side and device-free.  The reference pulls one cell's time series per call out of the (time, lat, lon) variables with a
 * strided nc_get_varm_*: Ncell x Nvar passes over the file.  These entry points read each (lat, lon) grid once, in file order,
 * and gather the modelled cells: out [nt][nvar][ncell].  Cell lookup (first exact match of the float/double "lat" / "lon"
 * coordinate), the (time, lat, lon) requirement and the value conversions (NC_SHORT / inverse_scale_factor or * scale_factor with
 * the attribute as float, NC_FLOAT and NC_DOUBLE as they are, anything else an error) are the reference's.  Containers: NetCDF
 * classic CDF-1 and CDF-2 (parsed by vic_b200/host/vicgpu_ncslab.h); a NetCDF-4/HDF5 file is refused with VICGPU_EUNSUPPORTED. */
typedef struct vicgpu_ncfile vicgpu_ncfile;
int vicgpu_nc_open(vicgpu_ncfile **nc, const char *path);
int vicgpu_nc_close(vicgpu_ncfile *nc);
/* lengths of the "time", "lat" and "lon" coordinate variables' dimensions */
int vicgpu_nc_dims(vicgpu_ncfile *nc, long long *ntime, long long *nlat, long long *nlon);
/* varnames [nvar]: the file's variable names (the reference maps FORCE_TYPE names through ProgramState::forcing_mapping);
 * lat, lng [ncell]: (double)soil_con.lat / .lng of the modelled cells; time steps [t0, t0 + nt) (t0 = the reference's skip_recs) */
int vicgpu_nc_read_slab(vicgpu_ncfile *nc, int nvar, const char *const *varnames, long long t0, long long nt, int ncell,
                        const double *lat, const double *lng, double *out);

/* ---- NetCDF output, one record per output step (stands in for WriteOutputNetCDF::initializeFile / write_data_all_cells,
 * WriteOutputNetCDF.c:163-299, 386-452) ----------------------------------------------------------------------------------
 * Host-side and device-free.  Same dimensions (lat, lon, bnds, time, depth), coordinate variables, per-variable attributes, fill
 * value and cell placement as the reference's file; `time` is the record dimension so that a step's grids of ALL variables are one
 * contiguous record, filled from the float32 rows vicgpu_step_f32 returns and written with one write.  Container: NetCDF classic
 * with 64-bit offsets (vic_b200/host/vicgpu_ncwrite.h); no compression. */
typedef struct vicgpu_ncout vicgpu_ncout;
typedef struct vicgpu_ncout_var {
  const char *name;          /* NetCDF variable name (VariableMetaData::name, variable_mapping.c) */
  int nelem;                 /* OutputData::nelem; > 1: (time, depth, lat, lon), elements beyond nelem stay at the fill value */
  const char *long_name, *units, *standard_name, *cell_methods, *internal_vic_name, *category;
} vicgpu_ncout_var;
typedef struct vicgpu_ncout_spec {
  int nlat, nlon, depth;     /* gridNumLatDivisions, gridNumLonDivisions, MAX_BANDS */
  double lat0, dlat, lon0, dlon;   /* gridStartLat, gridStepLat, gridStartLon, gridStepLon */
  const char *time_units;    /* "hours since Y-M-D H:00" / "days since Y-M-D" (WriteOutputNetCDF.c:221-229) */
  double time_step;          /* out_dt when out_dt < 24, else 1 (:240-243) */
  int nvar;
  const vicgpu_ncout_var *vars;
  int ntext;                 /* global text attributes (title, institution, source, history, frequency, Conventions ...) */
  const char *const *text_keys, *const *text_values;
  int nint;                  /* global integer attributes (model_start_year ... model_end_day) */
  const char *const *int_keys;
  const int *int_values;
  int ncell;                 /* modelled cells, in the order of the rows handed to write_step */
  const int *lat_index, *lon_index;   /* latitudeToIndex / longitudeToIndex of every cell */
} vicgpu_ncout_spec;
int vicgpu_ncout_create(vicgpu_ncout **w, const char *path, const vicgpu_ncout_spec *spec);
/* rows [ncell][row_stride] float32 (vicgpu_step_f32's out_agg of one output step); col_of_var [nvar]: first column of each variable */
int vicgpu_ncout_write_step(vicgpu_ncout *w, const float *rows, long long row_stride, const int *col_of_var);
int vicgpu_ncout_close(vicgpu_ncout *w);

/* ---- lake-ice surface solve as a batch operator ------------------------------------------------------
 * ice_melt() (ice_melt.c:30-585) with its residual IceEnergyBalance::calculate (IceEnergyBalance.c:60-175) and icerad()
 * (lakes.eb.c:1092-1151) for n independent lake-ice columns, one device thread each; bit-identical to the reference's ice_melt().
 * The lake model around it (solve_lake, water_balance: lakes.eb.c) is not built -- vicgpu_create still rejects LAKES TRUE -- so
 * this entry point serves a host-side lake loop that batches its cells' ice solves, and pins the residual the survey lists.
 * in  [n][VICGPU_ICE_NIN]  arguments and the members of snow_data_struct / lake_var_struct the function reads
 * out [n][VICGPU_ICE_NOUT] return code (0 or -999), the save_* results, and the members it writes
 * Arguments of the reference's signature that its body never reads (displacement, surf_atten, fracprv) are not carried;
 * options.BLOWING is taken as FALSE (ice_melt.c:239-258 not served); tfallback = options.TFALLBACK. */
#define VICGPU_ICE_IN(X) \
  X(z2) X(aero_resist) X(latent_heat_Le) X(Z0) X(rainfall) X(snowfall) X(wind) X(Tcutoff) X(air_temp) X(net_short) X(longwave) \
  X(density) X(pressure) X(vpd) X(vp) X(swq) X(surf_temp) X(pack_temp) X(pack_water) X(surf_water) X(vapor_flux) X(surface_flux) \
  X(surf_temp_fbflag) X(surf_temp_fbcount) X(ice_water_eq) X(areai) X(hice) X(volume)
#define VICGPU_ICE_OUT(X) \
  X(rc) X(aero_resist_used) X(melt) X(advection) X(deltaCC) X(SnowFlux) X(latent) X(sensible) X(Qnet) X(refreeze_energy) X(LWnet) \
  X(swq) X(surf_temp) X(pack_temp) X(pack_water) X(surf_water) X(vapor_flux) X(blowing_flux) X(surface_flux) X(surf_temp_fbflag) \
  X(surf_temp_fbcount) X(coverage) X(mass_error) X(coldcontent) X(ice_water_eq) X(volume)
#define VICGPU_ICE_ENUM_IN(n) ICEIN_##n,
#define VICGPU_ICE_ENUM_OUT(n) ICEOUT_##n,
enum { VICGPU_ICE_IN(VICGPU_ICE_ENUM_IN) VICGPU_ICE_NIN };
enum { VICGPU_ICE_OUT(VICGPU_ICE_ENUM_OUT) VICGPU_ICE_NOUT };
int vicgpu_ice_melt(int device, int n, int delta_t, int tfallback, const double *in, double *out);

/* accumulateGlacierMassBalance()'s per-cell result (vicNl.c:563, cell_info_struct::gmbEquation, written to the state file by
 * write_model_state.c:153-156): gmb[ncell][4] = b0, b1, b2, fitError of the quadratic fitted to (band elevation, cumulative mass
 * balance of the cell's glacier HRUs) at the end of the last completed accumulation interval; 0, 0, 0, -1 before the first. */
int vicgpu_get_glacier_fit(vicgpu_handle *h, double *gmb);

/* measurement aid: the device's FP64 fused-multiply-add throughput [TFLOP/s, 2 flops per FMA] from a register-resident DFMA loop
 * timed with CUDA events (best of 5) -- the denominator of the FP64 roofline fraction bench.py reports. */
int vicgpu_measure_fp64_peak(int device, double *tflops);

/* measurement aid for the kernel-structure decision: device time [us] of one pass that does what every
 * additional kernel boundary inside the step would add -- load the HRU record, load and store `nframe` doubles per HRU of values
 * live across the boundary, store the record -- over the current domain, averaged over `reps` launches.  Does not change the state. */
int vicgpu_measure_phase_tax(vicgpu_handle *h, int nframe, int reps, double *us_per_pass);

/* measurement aid: with profiling on, every launch of the per-HRU step kernel inside vicgpu_step is bracketed
 * by CUDA events on the library's stream; get_kernel_profile returns the summed duration and the launch count
 * since profiling was switched on. */
int vicgpu_set_profiling(vicgpu_handle *h, int on);
int vicgpu_get_kernel_profile(vicgpu_handle *h, double *hru_step_ms_total, long long *hru_step_launches);
/* measurement aid: per-warp start and end times (globaltimer, ns relative to the earliest start) of the LAST launch of the
 * per-HRU step kernel made while profiling was on; times[2*w] / times[2*w+1] for warp w of the launch, kind[w] the binning kind
 * of the warp's first row (vegetation class, + 1e6 bare soil, + 2e6 glacier).  Returns the number of warps (<= max_warps
 * are written).  Shows the load balance between warps and thread blocks.  The timers are only compiled into the launch when
 * the environment variable VICGPU_WARPTIME=1 was set at vicgpu_create (they cost ~15 % of the kernel). */
int vicgpu_get_warp_times(vicgpu_handle *h, double *times, double *kind, int max_warps);

#ifdef __cplusplus
}
#endif
#endif /* VICGPU_H */

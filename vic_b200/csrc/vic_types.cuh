// vic_types.cuh -- per-thread working set of one HRU (veg tile x snow band) and read-only
// views of the structure-of-arrays parameter tables in HBM.
//
// Device memory layout (all FP64, column-major = "structure of arrays"):
//   cellpar [L.cp_stride][ncell]   column k of cell c      at cellpar[k*ncell + c]
//   hrupar  [HP_N][nhru]
//   hrurec  [L.hr_stride][nhru]    the prognostic + diagnostic HRU record
//   veglib  [nclass][L.vl_stride]  (tiny, row-major, read through the read-only cache)
//   forcing [nrec][FV_N*nslot][ncell]
// so that consecutive threads (consecutive HRUs; the HRUs of one cell are adjacent) touch
// consecutive addresses.  The host-compiled port uses the same layout.
#ifndef VIC_TYPES_CUH
#define VIC_TYPES_CUH
#include "vic_common.cuh"

namespace vic {

// options as the kernels want them (subset of vicgpu_options, resolved once on the host)
struct Opts {
  int Nnode, Nbands, dt, SNOW_STEP, NR, NF, nrecs, out_step_ratio;
  int FULL_ENERGY, FROZEN_SOIL, QUICK_FLUX, QUICK_SOLVE, IMPLICIT, EXP_TRANS, NOFLUX;
  int GRND_FLUX_TYPE, AERO_RESIST_CANSNOW, SNOW_ALBEDO, SNOW_DENSITY, TEMP_TH_TYPE, TFALLBACK;
  int GLACIER_ID, GLACIER_DYNAMICS, MOISTFRACT, ALMA_OUTPUT, NVegLibTypes;
  int gaYear, gaMonth, gaDay, gaInterval;
  double wind_h;
  vicgpu_layout L;
};

// strided column view: element k of "my" row
struct Col {
  const double* p;  // already offset to my cell / my HRU
  size_t n;         // column stride (ncell or nhru)
  VIC_HD double operator()(int k) const {
#if defined(__CUDA_ARCH__)
    return __ldg(p + (size_t)k * n);
#else
    return p[(size_t)k * n];
#endif
  }
};

// one row of the vegetation library
struct VegRow {
  const double* r;
  const vicgpu_layout* L;
  VIC_HD double s(int k) const {
#if defined(__CUDA_ARCH__)
    return __ldg(r + k);
#else
    return r[k];
#endif
  }
  VIC_HD double m(int f, int month0) const { return s(VICGPU_VL_MONTH(L, f, month0)); }
  VIC_HD bool overstory() const { return s(VL_overstory) != 0.0; }
};

struct VegLib {
  const double* base;
  const vicgpu_layout* L;
  VIC_HD VegRow row(int cls) const { return VegRow{base + (size_t)cls * L->vl_stride, L}; }
};

// cell parameters with named access
struct CellPar {
  Col c;
  const vicgpu_layout* L;
  VIC_HD double operator()(int k) const { return c(k); }
  VIC_HD double layer(int f, int i) const { return c(VICGPU_CP_LAYER(L, f, i)); }
  VIC_HD double node(int f, int i) const { return c(VICGPU_CP_NODE(L, f, i)); }
  VIC_HD double zwt(int f, int curve, int pt) const { return c(VICGPU_CP_ZWT(L, f, curve * VICGPU_NZWT + pt)); }
  VIC_HD double band(int f, int b) const { return c(VICGPU_CP_BAND(L, f, b)); }
};

// one forcing record of my cell: value of variable var at sub-step slot
struct Forcing {
  Col c;
  int nslot;
  VIC_HD double operator()(int var, int slot) const { return c(var * nslot + slot); }
};

struct Dmy {
  int day, day_in_year, hour, month, year;
};

// ---- HRU working set -------------------------------------------------------------------
// Members are generated from the column tables (include/vicgpu_fields.h) so that the
// record <-> struct transfer below cannot drift from the C-ABI.
template <int NN>
struct EnergyBal {
#define X(n, p, c) double n;
  VICGPU_HRU_ENERGY(X, )
#undef X
  double fdepth[VICGPU_NFRONTS], tdepth[VICGPU_NFRONTS];
  // thermal nodes
  double Cs_node[NN], ice[NN], kappa_node[NN], moist[NN], T[NN], T_fbflag[NN], T_fbcount[NN];
};

struct SnowPack {
#define X(n, p, c) double n;
  VICGPU_HRU_SNOW(X, )
#undef X
};

struct SoilLayer {
#define X(n, p, c) double n;
  VICGPU_HRU_LAYER(X, )
#undef X
};

struct SoilCol {
#define X(n, p, c) double n;
  VICGPU_HRU_CELL(X, )
#undef X
  SoilLayer layer[VICGPU_NLAYER];
  double pot_evap[VICGPU_NPET];
};

struct VegVar {
#define X(n, p, c) double n;
  VICGPU_HRU_VEG(X, )
#undef X
};

struct Glacier {
#define X(n, p, c) double n;
  VICGPU_HRU_GLAC(X, )
#undef X
};

template <int NN>
struct Hru {
  EnergyBal<NN> energy;
  SnowPack snow;
  SoilCol cell;
  VegVar veg;
  Glacier glac;
  double mu;
};

// HRU record (column-major in memory, stride n) -> working set
template <int NN>
VIC_HDI void load_hru(Hru<NN>& h, const double* rec, size_t n, const vicgpu_layout* L) {
#define LD(k) rec[(size_t)(k) * n]
#define X(nm, p, c) h.energy.nm = LD(HR_E_##nm);
  VICGPU_HRU_ENERGY(X, )
#undef X
#define X(nm, p, c) h.snow.nm = LD(HR_S_##nm);
  VICGPU_HRU_SNOW(X, )
#undef X
#define X(nm, p, c) h.cell.nm = LD(HR_C_##nm);
  VICGPU_HRU_CELL(X, )
#undef X
#define X(nm, p, c) h.veg.nm = LD(HR_V_##nm);
  VICGPU_HRU_VEG(X, )
#undef X
#define X(nm, p, c) h.glac.nm = LD(HR_G_##nm);
  VICGPU_HRU_GLAC(X, )
#undef X
  h.mu = LD(HR_H_mu);
  for (int i = 0; i < VICGPU_NLAYER; i++) {
#define X(nm, p, c) h.cell.layer[i].nm = LD(VICGPU_HR_LAYER(L, HRL_##nm, i));
    VICGPU_HRU_LAYER(X, )
#undef X
  }
  for (int i = 0; i < VICGPU_NFRONTS; i++) {
    h.energy.fdepth[i] = LD(VICGPU_HR_FRONT(L, HRF_fdepth, i));
    h.energy.tdepth[i] = LD(VICGPU_HR_FRONT(L, HRF_tdepth, i));
  }
  for (int i = 0; i < VICGPU_NPET; i++) h.cell.pot_evap[i] = LD(VICGPU_HR_PET(L, i));
  for (int i = 0; i < NN; i++) {
    if (i < L->nnode) {
      h.energy.Cs_node[i] = LD(VICGPU_HR_NODE(L, HRN_Cs, i));
      h.energy.ice[i] = LD(VICGPU_HR_NODE(L, HRN_ice, i));
      h.energy.kappa_node[i] = LD(VICGPU_HR_NODE(L, HRN_kappa, i));
      h.energy.moist[i] = LD(VICGPU_HR_NODE(L, HRN_moist, i));
      h.energy.T[i] = LD(VICGPU_HR_NODE(L, HRN_T, i));
      h.energy.T_fbflag[i] = LD(VICGPU_HR_NODE(L, HRN_T_fbflag, i));
      h.energy.T_fbcount[i] = LD(VICGPU_HR_NODE(L, HRN_T_fbcount, i));
    }
  }
#undef LD
}

template <int NN>
VIC_HDI void store_hru(const Hru<NN>& h, double* rec, size_t n, const vicgpu_layout* L) {
#define ST(k) rec[(size_t)(k) * n]
#define X(nm, p, c) ST(HR_E_##nm) = h.energy.nm;
  VICGPU_HRU_ENERGY(X, )
#undef X
#define X(nm, p, c) ST(HR_S_##nm) = h.snow.nm;
  VICGPU_HRU_SNOW(X, )
#undef X
#define X(nm, p, c) ST(HR_C_##nm) = h.cell.nm;
  VICGPU_HRU_CELL(X, )
#undef X
#define X(nm, p, c) ST(HR_V_##nm) = h.veg.nm;
  VICGPU_HRU_VEG(X, )
#undef X
#define X(nm, p, c) ST(HR_G_##nm) = h.glac.nm;
  VICGPU_HRU_GLAC(X, )
#undef X
  ST(HR_H_mu) = h.mu;
  for (int i = 0; i < VICGPU_NLAYER; i++) {
#define X(nm, p, c) ST(VICGPU_HR_LAYER(L, HRL_##nm, i)) = h.cell.layer[i].nm;
    VICGPU_HRU_LAYER(X, )
#undef X
  }
  for (int i = 0; i < VICGPU_NFRONTS; i++) {
    ST(VICGPU_HR_FRONT(L, HRF_fdepth, i)) = h.energy.fdepth[i];
    ST(VICGPU_HR_FRONT(L, HRF_tdepth, i)) = h.energy.tdepth[i];
  }
  for (int i = 0; i < VICGPU_NPET; i++) ST(VICGPU_HR_PET(L, i)) = h.cell.pot_evap[i];
  for (int i = 0; i < NN; i++) {
    if (i < L->nnode) {
      ST(VICGPU_HR_NODE(L, HRN_Cs, i)) = h.energy.Cs_node[i];
      ST(VICGPU_HR_NODE(L, HRN_ice, i)) = h.energy.ice[i];
      ST(VICGPU_HR_NODE(L, HRN_kappa, i)) = h.energy.kappa_node[i];
      ST(VICGPU_HR_NODE(L, HRN_moist, i)) = h.energy.moist[i];
      ST(VICGPU_HR_NODE(L, HRN_T, i)) = h.energy.T[i];
      ST(VICGPU_HR_NODE(L, HRN_T_fbflag, i)) = h.energy.T_fbflag[i];
      ST(VICGPU_HR_NODE(L, HRN_T_fbcount, i)) = h.energy.T_fbcount[i];
    }
  }
#undef ST
}

// everything a physics routine needs to know about "where am I"
struct Ctx {
  const Opts* o;
  CellPar cp;
  VegLib vl;
  Col hp;        // my HRU parameter row
  Forcing f;     // my cell's forcing record of the current model step
  Dmy dmy;
  int rec;
};

}  // namespace vic
#endif

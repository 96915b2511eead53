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
  int GLACIER_ID, GLACIER_DYNAMICS, MOISTFRACT, ALMA_OUTPUT, NVegLibTypes, CORRPREC, BLOWING;
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
  Col d;  // constants derived from the cell parameters once per domain (derive_cell_constants); d.p == nullptr: not available
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
  VIC_HD double operator()(int var, int slot) const {
#if defined(VIC_FORCING_SMEM) && defined(__CUDA_ARCH__)
    return c.p[(size_t)(var * nslot + slot) * c.n];  // may point into shared memory (vicgpu_step.inc): a generic load, not ld.global.nc
#else
    return c(var * nslot + slot);
#endif
  }
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

// The members of the sub-step's snow-side energy record that solve_snow / snow_intercept / surface_fluxes read or write.  The
// reference works on a full copy of the HRU's energy record (surface_fluxes.c:385-389 snow_energy); everything else in that copy is
// never touched, so only these 25 values are carried (200 B of thread-local memory instead of 616 B, and 25 copies instead of 77).
#define VIC_SNOW_SIDE_ENERGY(X)                                                                                                      \
  X(AlbedoOver) X(AlbedoUnder) X(LongOverIn) X(NetLongOver) X(NetShortOver) X(ShortOverIn) X(Tcanopy) X(Tfoliage) X(Tfoliage_fbcount)  \
  X(Tfoliage_fbflag) X(advected_sensible) X(advection) X(canopy_advection) X(canopy_latent) X(canopy_latent_sub) X(canopy_refreeze)    \
  X(canopy_sensible) X(deltaCC) X(error) X(latent) X(latent_sub) X(melt_energy) X(refreeze_energy) X(sensible) X(snow_flux)
struct SnowSideEnergy {
#define X(n) double n;
  VIC_SNOW_SIDE_ENERGY(X)
#undef X
  template <int NN>
  VIC_HD void take(const EnergyBal<NN>& e) {
#define X(n) n = e.n;
    VIC_SNOW_SIDE_ENERGY(X)
#undef X
  }
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
// ---- layout of the HRU state tables ---------------------------------------------------------------------------------------------
// Tile-major: the records of 32 consecutive rows (one warp) form one contiguous tile [hr_stride][32], tiles follow each other.
// Column k of row h lives at base[hr_off(h, hr_stride) + k * VIC_HR_TILE]: the 32 lanes of a warp still read 256 contiguous bytes
// per column, and a warp's whole record (181 columns = 46 KB) is ONE contiguous block of memory instead of 181 pieces 0.4 MB apart
// (DRAM pages, TLB reach, prefetch).  Tables hold round32(nhru) rows.
#define VIC_HR_TILE 32
VIC_HD size_t hr_off(int h, int ncol) { return (size_t)(h >> 5) * (size_t)ncol * VIC_HR_TILE + (size_t)(h & 31); }
VIC_HD size_t hr_rows(int nhru) { return ((size_t)nhru + 31) / 32 * 32; }

// ---- record <-> working set ------------------------------------------------------------------------------------------
// The scalar columns of the HRU record (include/vicgpu_fields.h) and the members of EnergyBal / SnowPack / SoilCol / VegVar /
// Glacier are generated from the same X-macro tables in the same order, and every member is a double: a group of columns is
// a run of consecutive doubles of the working set.  The transfer is written as chunked loops -- a chunk of independent loads
// first, then the chunk's stores -- so that the memory round trips of a chunk overlap (a member-by-member copy compiles to
// load, store, load, store ... and pays one round trip per column).
#include <stddef.h>
#ifndef VIC_XFER_CHUNK
#define VIC_XFER_CHUNK 32  // (16: 527 -> 520 us per launch at 10,000 cells, 5,313 -> 5,218 us at 125,000: half as many dependent round trips per record)
#endif
// The HRU records stream through the step once per record (read from one state half, written to the other) while the step's own
// thread-local working set is re-read thousands of times: with -DVIC_STATE_CS the record transfers carry the evict-first hint
// (ld.global.cs / st.global.cs) so that they do not push the stack lines out of L2.
#if defined(VIC_STATE_CS) && defined(__CUDA_ARCH__)
#define VIC_REC_LD(ptr) __ldcs(ptr)
#define VIC_REC_ST(ptr, v) __stcs((ptr), (v))
#elif defined(__CUDA_ARCH__)
#define VIC_REC_LD(ptr) __ldg(ptr)
#define VIC_REC_ST(ptr, v) (*(ptr) = (v))
#else
#define VIC_REC_LD(ptr) (*(ptr))
#define VIC_REC_ST(ptr, v) (*(ptr) = (v))
#endif

// columns col0 .. col0+count-1 of my record row  ->  dst[0 .. count)
VIC_HD void cols_to_local(const double* __restrict__ rec, size_t n, int col0, int count, double* __restrict__ dst) {
  #pragma unroll 1
  for (int k = 0; k < count; k += VIC_XFER_CHUNK) {
    double r[VIC_XFER_CHUNK];
#pragma unroll
    for (int j = 0; j < VIC_XFER_CHUNK; j++) {
      if (k + j < count) {
        r[j] = VIC_REC_LD(rec + (size_t)(col0 + k + j) * n);
      }
    }
#pragma unroll
    for (int j = 0; j < VIC_XFER_CHUNK; j++)
      if (k + j < count) dst[k + j] = r[j];
  }
}
VIC_HD void local_to_cols(const double* __restrict__ src, double* __restrict__ rec, size_t n, int col0, int count) {
  #pragma unroll 1
  for (int k = 0; k < count; k += VIC_XFER_CHUNK) {
    double r[VIC_XFER_CHUNK];
#pragma unroll
    for (int j = 0; j < VIC_XFER_CHUNK; j++)
      if (k + j < count) r[j] = src[k + j];
#pragma unroll
    for (int j = 0; j < VIC_XFER_CHUNK; j++)
      if (k + j < count) VIC_REC_ST(rec + (size_t)(col0 + k + j) * n, r[j]);
  }
}

template <int NN>
struct HruGroups {
  static constexpr int nE = (int)(offsetof(EnergyBal<NN>, fdepth) / sizeof(double));
  static constexpr int nS = (int)(sizeof(SnowPack) / sizeof(double));
  static constexpr int nC = (int)(offsetof(SoilCol, layer) / sizeof(double));
  static constexpr int nV = (int)(sizeof(VegVar) / sizeof(double));
  static constexpr int nG = (int)(sizeof(Glacier) / sizeof(double));
  static constexpr int nLayerF = (int)(sizeof(SoilLayer) / sizeof(double));
  static_assert(nE + nS + nC + nV + nG == (int)HR_H_mu && (int)HR_H_mu + 1 == (int)HR_NSCALAR, "HRU scalar columns and working-set members out of step");
  static_assert(nLayerF == (int)HRL_N, "soil-layer columns and SoilLayer members out of step");
  static_assert(offsetof(EnergyBal<NN>, tdepth) == offsetof(EnergyBal<NN>, fdepth) + VICGPU_NFRONTS * sizeof(double), "fronts");
  static_assert(offsetof(EnergyBal<NN>, T_fbcount) == offsetof(EnergyBal<NN>, Cs_node) + 6 * NN * sizeof(double) && (int)HRN_N == 7, "nodes");
};

// HRU record (column-major in memory, stride n) -> working set
template <int NN>
VIC_HDI void load_hru(Hru<NN>& h, const double* __restrict__ rec, size_t n, const vicgpu_layout* Lg) {
  typedef HruGroups<NN> G;
  const int hr_layer0 = Lg->hr_layer0, hr_front0 = Lg->hr_front0, hr_pet0 = Lg->hr_pet0, hr_node0 = Lg->hr_node0, nnode = Lg->nnode;
  cols_to_local(rec, n, 0, G::nE, reinterpret_cast<double*>(&h.energy));
  cols_to_local(rec, n, G::nE, G::nS, reinterpret_cast<double*>(&h.snow));
  cols_to_local(rec, n, G::nE + G::nS, G::nC, reinterpret_cast<double*>(&h.cell));
  cols_to_local(rec, n, G::nE + G::nS + G::nC, G::nV, reinterpret_cast<double*>(&h.veg));
  cols_to_local(rec, n, G::nE + G::nS + G::nC + G::nV, G::nG, reinterpret_cast<double*>(&h.glac));
  h.mu = VIC_REC_LD(rec + (size_t)HR_H_mu * n);
  // layers: record [field][layer], working set layer[i].field
  {
    double r[HRL_N * VICGPU_NLAYER];
    cols_to_local(rec, n, hr_layer0, HRL_N * VICGPU_NLAYER, r);
    for (int i = 0; i < VICGPU_NLAYER; i++) {
      double* li = reinterpret_cast<double*>(&h.cell.layer[i]);
      for (int f = 0; f < HRL_N; f++) li[f] = r[f * VICGPU_NLAYER + i];
    }
  }
  cols_to_local(rec, n, hr_front0, 2 * VICGPU_NFRONTS, h.energy.fdepth);  // fdepth[], tdepth[] are adjacent
  cols_to_local(rec, n, hr_pet0, VICGPU_NPET, h.cell.pot_evap);
  // nodes: record [field][nnode], working set field[NN]
  if (nnode == NN) cols_to_local(rec, n, hr_node0, HRN_N * NN, h.energy.Cs_node);  // both sides contiguous
  else
    for (int f = 0; f < HRN_N; f++) cols_to_local(rec, n, hr_node0 + f * nnode, nnode < NN ? nnode : NN, h.energy.Cs_node + f * NN);
}

template <int NN>
VIC_HDI void store_hru(const Hru<NN>& h, double* __restrict__ rec, size_t n, const vicgpu_layout* Lg) {
  typedef HruGroups<NN> G;
  const int hr_layer0 = Lg->hr_layer0, hr_front0 = Lg->hr_front0, hr_pet0 = Lg->hr_pet0, hr_node0 = Lg->hr_node0, nnode = Lg->nnode;
  local_to_cols(reinterpret_cast<const double*>(&h.energy), rec, n, 0, G::nE);
  local_to_cols(reinterpret_cast<const double*>(&h.snow), rec, n, G::nE, G::nS);
  local_to_cols(reinterpret_cast<const double*>(&h.cell), rec, n, G::nE + G::nS, G::nC);
  local_to_cols(reinterpret_cast<const double*>(&h.veg), rec, n, G::nE + G::nS + G::nC, G::nV);
  local_to_cols(reinterpret_cast<const double*>(&h.glac), rec, n, G::nE + G::nS + G::nC + G::nV, G::nG);
  VIC_REC_ST(rec + (size_t)HR_H_mu * n, h.mu);
  {
    double r[HRL_N * VICGPU_NLAYER];
    for (int i = 0; i < VICGPU_NLAYER; i++) {
      const double* li = reinterpret_cast<const double*>(&h.cell.layer[i]);
      for (int f = 0; f < HRL_N; f++) r[f * VICGPU_NLAYER + i] = li[f];
    }
    local_to_cols(r, rec, n, hr_layer0, HRL_N * VICGPU_NLAYER);
  }
  local_to_cols(h.energy.fdepth, rec, n, hr_front0, 2 * VICGPU_NFRONTS);
  local_to_cols(h.cell.pot_evap, rec, n, hr_pet0, VICGPU_NPET);
  if (nnode == NN) local_to_cols(h.energy.Cs_node, rec, n, hr_node0, HRN_N * NN);
  else
    for (int f = 0; f < HRN_N; f++) local_to_cols(h.energy.Cs_node + f * NN, rec, n, hr_node0 + f * nnode, nnode < NN ? nnode : NN);
}

// everything a physics routine needs to know about "where am I"
// Rendezvous of the warps of a thread block at the phase boundaries of the step (device only; a hint, never a requirement).  The
// kernel stalls on instruction fetch as soon as the warps of an SM spread over different phases of its 0.6 MB of code (DESIGN.md
// section 6): at a phase boundary a warp waits until the block's other warps that take the same path have arrived, and the block
// then runs the next phase's code together -- measured 821 -> 665 us per launch.  Warps register for a path (group 0: surface_fluxes,
// group 1: surface_fluxes_glac) when they enter the step; the wait is bounded by `limit` clock cycles so that a warp that leaves its
// path early (an ERROR return) cannot hold the others for long.
#define VIC_NPHASE 8
struct PhaseSync {
  unsigned* count;  // [VIC_NPHASE + 2] in shared memory, zeroed at kernel start: arrivals per phase, then the two groups' sizes; null: off
  long long limit;  // clock cycles a warp is prepared to wait
  unsigned mask;    // bit p: rendezvous at phase boundary p (VICGPU_SYNCMASK, A/B knob; all by default)
};

struct Ctx {
  const Opts* o;
  CellPar cp;
  VegLib vl;
  Col hp;        // my HRU parameter row
  Forcing f;     // my cell's forcing record of the current model step
  Col aero;      // my row of the per-month aerodynamic table (vic_step.cuh AeroGeom); p == nullptr: none, hru_step computes it
  Dmy dmy;
  int rec;
  PhaseSync ps;
  // the calling lanes (one path of a possibly divergent warp) join group `group`
  VIC_HD void join(int group) const {
#if defined(__CUDA_ARCH__)
    if (!ps.count) return;
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&ps.count[VIC_NPHASE + group], 1u);
    __syncwarp(m);
#else
    (void)group;
#endif
  }
  VIC_HD void rendezvous(int phase, int group) const {
#if defined(__CUDA_ARCH__)
    if (!ps.count || !((ps.mask >> phase) & 1u)) return;
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) {
      atomicAdd(&ps.count[phase], 1u);
      const long long t0 = clock64();
      while (*(volatile unsigned*)&ps.count[phase] < *(volatile unsigned*)&ps.count[VIC_NPHASE + group] && clock64() - t0 < ps.limit) __nanosleep(100);
    }
    __syncwarp(m);
#else
    (void)phase;
    (void)group;
#endif
  }
};

}  // namespace vic
#endif

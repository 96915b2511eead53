// vic_engine.cuh -- the per-HRU and per-cell work items of one model record, written once and
// called from the CUDA kernels (vicgpu.cu: one thread per HRU / per cell) and from the
// host-compiled port used only by the tests (oracle/vicport.cpp: a loop).
//
// A record of the reference's time loop (vicNl.c:506-610) is, per cell:
//   [rec 0 only] put_data(rec = -nrecs)    -> cell_output(rec = -1)
//   dist_prec: full_energy + put_data       -> hru_work for every HRU, then cell_output(rec)
//   accumulateGlacierMassBalance            -> folded into hru_work (it only touches the HRU's
//                                              own glacier.cum_mass_balance, which put_data never reads)
#ifndef VIC_ENGINE_CUH
#define VIC_ENGINE_CUH
#include <algorithm>
#include <vector>
#include "vic_step.cuh"
#include "vic_output.cuh"
#include "vic_gmb.cuh"

namespace vic {

// device-resident tables (column-major, see vic_types.cuh)
struct Tables {
  int ncell, nhru, nclass;
  const double* veglib;   // [nclass][vl_stride]
  const double* cellpar;  // [cp_stride][ncell]
  const double* cellder;  // [VIC_NCELLDER][ncell] constants derived from cellpar (derive_cell_constants), or null
  const double* hrupar;   // [HP_N][nhru]
  const double* hrurec;   // tiles [hr_stride][32] (vic_types.cuh hr_off): state at the start of the record (read by hru_work)
  double* hrurec_out;     // [hr_stride][nhru]  state at the end of the record (written by hru_work, read by cell_output);
                          //                    the CUDA library ping-pongs two buffers so that cell_output of record r runs
                          //                    beside hru_work of record r+1; the host port passes hrurec_out == hrurec
  double* hdiag_out;      // [3][nhru]  Cv-weighted out_prec / out_rain / out_snow of the step (same ping-pong)
  const int* cell_h0;     // [ncell+1] first HRU of each cell (caller's HRU numbering: hruList order grouped by cell)
  const int* slot_of_hru; // [nhru] row of the HRU tables an HRU occupies (null = identity); threads of hru_work run over rows
  int* status;            // [ncell] 0 or -999
  int* fail_rec;          // [ncell] first record at which the cell is invalid (INT_MAX: never; -1: from the start)
  double* carry;          // [CC_N][ncell]
  double* out;            // [nout][ncell]  OutputData::data of the current record
  double* agg;            // [nout][ncell]  OutputData::aggdata
  const int* aggtype;     // [N_OUTVARS]
  int* cost;              // [nhru] or null: cost estimate of each row's last step (vic_frozen.cuh vic_count_work), read by the row binning only
  const double* aero;     // [VIC_AERO_NCOL][nhru] per-month aerodynamic geometry of each row (vic_step.cuh AeroGeom, kernel k_hru_aero), or null
  double* gmb_cum;        // [nhru] glacier.cum_mass_balance of an HRU at the end of the last accumulation interval, before its reset
  double* gmb;            // [4][ncell] b0, b1, b2, fitError of the cell's mass-balance curve (GraphingEquation), or null
};

// what accumulateGlacierMassBalance does at this record (decided on the host from the calendar,
// accumulateGlacierMassBalance.c:15-66; integer bookkeeping, bit-exact)
struct GlacAccum {
  int enabled, reset_first, accumulate, reset_after;
};

// an HRU that is not stepped (its cell is invalid) keeps its state: copy the record to the output buffer when there are two
VIC_HDI void carry_hru_record(const Tables& t, int h, int hr_stride) {
  if (t.hrurec_out == t.hrurec) return;
  const size_t off = hr_off(h, hr_stride);
  for (int k = 0; k < hr_stride; k++) t.hrurec_out[off + (size_t)k * VIC_HR_TILE] = t.hrurec[off + (size_t)k * VIC_HR_TILE];
}

// h: row of the HRU tables
template <int NN, bool ONE>
VIC_HDI void hru_work(const Opts* o, const Tables& t, const double* forcing_rec /* [f_stride][ncell] */, int h, Dmy dmy, int rec, GlacAccum ga,
                      PhaseSync ps = PhaseSync{nullptr, 0, 0}, const double* fstage = nullptr, int fstage_n = 0) {
  const size_t nh = (size_t)t.nhru;
  Col hpc{t.hrupar + h, nh};
  const int cell = (int)hpc(HP_cell);
  double* dg = t.hdiag_out + h;
  if (t.status[cell] != 0) {
    dg[0] = dg[nh] = dg[2 * nh] = 0;
    carry_hru_record(t, h, o->L.hr_stride);
    return;
  }
  Ctx cx;
  cx.o = o;
  cx.cp = CellPar{Col{t.cellpar + cell, (size_t)t.ncell}, &o->L, Col{t.cellder ? t.cellder + cell : nullptr, (size_t)t.ncell}};
  cx.vl = VegLib{t.veglib, &o->L};
  cx.hp = hpc;
  // fstage: this thread's column of the forcing record staged in shared memory by the kernel (vicgpu_step.inc), stride fstage_n
  cx.f = fstage ? Forcing{Col{fstage, (size_t)fstage_n}, o->L.f_nslot} : Forcing{Col{forcing_rec + cell, (size_t)t.ncell}, o->L.f_nslot};
  cx.aero = Col{t.aero ? t.aero + h : nullptr, nh};
  cx.dmy = dmy;
  cx.rec = rec;
  cx.ps = ps;
  const HruPar hp = load_hrupar(hpc);
  Hru<NN> hru;
  load_hru<NN>(hru, t.hrurec + hr_off(h, o->L.hr_stride), VIC_HR_TILE, &o->L);
#if defined(VIC_FORCING_SMEM) && defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");  // the staged forcing column is in shared memory (issued before the record load)
#endif
  HruStepDiag d;
  int e = hru_step<NN, ONE>(hru, hp, cx, d);
  if (e == ERROR_I) {
    t.status[cell] = ERROR_I;  // benign race: every writer stores the same value
#if defined(__CUDA_ARCH__)
    atomicMin(&t.fail_rec[cell], rec);
#else
    if (rec < t.fail_rec[cell]) t.fail_rec[cell] = rec;
#endif
    dg[0] = dg[nh] = dg[2 * nh] = 0;
    carry_hru_record(t, h, o->L.hr_stride);
    return;
  }
  if (ga.enabled && hp.isGlacier) {
    if (ga.reset_first) hru.glac.cum_mass_balance = 0;
    if (ga.accumulate && is_valid(hru.glac.mass_balance)) hru.glac.cum_mass_balance += hru.glac.mass_balance;
    if (ga.reset_after) {
      if (t.gmb_cum) t.gmb_cum[h] = hru.glac.cum_mass_balance;  // what GlacierMassBalanceResult reads (accumulateGlacierMassBalance.c:59-64)
      hru.glac.cum_mass_balance = 0;
    }
  }
  store_hru<NN>(hru, t.hrurec_out + hr_off(h, o->L.hr_stride), VIC_HR_TILE, &o->L);
  dg[0] = d.out_prec * hp.Cv;
  dg[nh] = d.out_rain * hp.Cv;
  dg[2 * nh] = d.out_snow * hp.Cv;
#if defined(VIC_WORK_BUFFER) && defined(__CUDA_ARCH__)
  if (t.cost && vic_work_buf) {  // the step's cost estimate (vic_frozen.cuh vic_count_work), for the row binning only
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    t.cost[h] = vic_work_buf[g];
    vic_work_buf[g] = 0;
  }
#endif
}

// put_data of one cell (vic_output.cuh), host port: one call does everything, the row is built in the output table itself.
// roles != 0 emulates the device's three-thread reduction (below) on the host: three passes with separate rows, combined by variable
// ownership -- the same arithmetic, and the check that out_var_group() agrees with what put_data_hru writes (oracle/vicport --roles).
VIC_HDI void cell_output(const Opts& o, const Tables& t, const double* forcing_rec, int cell, int rec, int step_count, int roles = 0) {
  if (rec >= 0 && t.fail_rec[cell] <= rec) return;  // the reference stops touching an invalid cell (vicNl.c:521): its row keeps the last values
  const size_t nc = (size_t)t.ncell, nh = (size_t)t.nhru;
  const int nout = o.L.out_off[VICGPU_N_OUTVARS], hs = o.L.hr_stride;
  CellPar cp{Col{t.cellpar + cell, nc}, &o.L};
  VegLib vl{t.veglib, &o.L};
  Forcing f{Col{forcing_rec ? forcing_rec + cell : nullptr, nc}, o.L.f_nslot};
  RowRW out{t.out + cell, nc}, agg{t.agg + cell, nc};
  PutDataCtx pc;
  const int h0 = t.cell_h0[cell], h1 = t.cell_h0[cell + 1];
  if (!roles) {
    for (int k = 0; k < nout; k++) out[k] = 0;
    put_data_begin(o, cp, vl, &f, t.hrupar, t.hdiag_out, nh, t.slot_of_hru, h0, h1, rec, out, pc);
    for (int hh = h0; hh < h1; hh++) {
      const int h = t.slot_of_hru ? t.slot_of_hru[hh] : hh;
      put_data_hru<PD_ALL>(o, cp, vl, RecTile{t.hrurec_out + hr_off(h, hs)}, t.hrupar, nh, h, out, pc);
    }
  } else {
#if !defined(__CUDA_ARCH__)
    std::vector<double> rows((size_t)3 * nout, 0.0);
    PutDataCtx pcs[3];
    for (int g = 0; g < 3; g++) {
      RowLocal row{rows.data() + (size_t)g * nout};
      put_data_begin(o, cp, vl, &f, t.hrupar, t.hdiag_out, nh, t.slot_of_hru, h0, h1, rec, row, pcs[g]);
      for (int hh = h0; hh < h1; hh++) {
        const int h = t.slot_of_hru ? t.slot_of_hru[hh] : hh;
        const RecTile hr{t.hrurec_out + hr_off(h, hs)};
        if (g == 0) put_data_hru<PD_WB>(o, cp, vl, hr, t.hrupar, nh, h, row, pcs[g]);
        else if (g == 1) put_data_hru<PD_EB>(o, cp, vl, hr, t.hrupar, nh, h, row, pcs[g]);
        else put_data_hru<PD_BAND>(o, cp, vl, hr, t.hrupar, nh, h, row, pcs[g]);
      }
    }
    for (int v = 0; v < VICGPU_N_OUTVARS; v++) {
      const int grp = out_var_group(v), g = grp == PD_WB ? 0 : grp == PD_EB ? 1 : 2;
      for (int i = 0; i < o.L.out_nelem[v]; i++) out[o.L.out_off[v] + i] = rows[(size_t)g * nout + o.L.out_off[v] + i];
    }
    pc = pcs[0];
#endif
  }
  put_data_finish(o, cp, rec, RowRW{t.carry + cell, nc}, out, pc);
  if (rec < 0) return;
  put_data_aggregate_vars(o, t.aggtype, out, agg, 0, 1);
  put_data_aggregate_tail(o, step_count, agg);
}

#if defined(__CUDACC__)
// Device: THREE threads per cell, in three warps of a 96-thread block that share 32 cells.  Every output variable is a sum over the
// cell's HRUs in hruList order -- a sequential chain by definition -- but the variables are independent of each other: warp 0
// reduces the water-balance terms, warp 1 the energy-balance terms, warp 2 the per-band terms (put_data_hru<PD_*>; each variable is
// still summed by ONE thread, in the reference's order).  Each thread builds its part of the row in a thread-local array (a
// read-modify-write chain through global memory per statement serialises on the L2 latency), writes the variables it owns to the
// output table, and after a block barrier the first thread of the cell runs the derived variables and balance checks on the whole
// row; the temporal aggregation is dealt over the three threads again.  The kernel runs as a programmatic dependent of the next
// record's step (vicgpu_api.cu) on the SMs the step grid leaves idle, so what counts is its duration with few warps: one thread per
// cell took 375 us alone and bounded the record at 590 us (profiles/r02_summary.md).
#define VIC_OUT_LOCAL_MAX 512  // doubles of thread-local row (the three-node layout has 365 columns, ten nodes 400)
__device__ __forceinline__ void cell_output_role(const Opts& o, const Tables& t, const double* forcing_rec, int cell, int rec, int step_count, int role, bool live) {
  const size_t nc = (size_t)t.ncell, nh = (size_t)t.nhru;
  const int nout = o.L.out_off[VICGPU_N_OUTVARS], hs = o.L.hr_stride;
  const int c = live ? cell : 0;
  CellPar cp{Col{t.cellpar + c, nc}, &o.L};
  VegLib vl{t.veglib, &o.L};
  Forcing f{Col{forcing_rec ? forcing_rec + c : nullptr, nc}, o.L.f_nslot};
  RowRW gout{t.out + c, nc}, agg{t.agg + c, nc};
  PutDataCtx pc;
  if (live) {
    double row[VIC_OUT_LOCAL_MAX];
    RowLocal out{row};
    for (int k = 0; k < nout; k++) row[k] = 0;
    const int h0 = t.cell_h0[c], h1 = t.cell_h0[c + 1];
    put_data_begin(o, cp, vl, &f, t.hrupar, t.hdiag_out, nh, t.slot_of_hru, h0, h1, rec, out, pc);
    for (int hh = h0; hh < h1; hh++) {
      const int h = t.slot_of_hru ? t.slot_of_hru[hh] : hh;
      const RecTile hr{t.hrurec_out + hr_off(h, hs)};
      if (role == 0) put_data_hru<PD_WB>(o, cp, vl, hr, t.hrupar, nh, h, out, pc);
      else if (role == 1) put_data_hru<PD_EB>(o, cp, vl, hr, t.hrupar, nh, h, out, pc);
      else put_data_hru<PD_BAND>(o, cp, vl, hr, t.hrupar, nh, h, out, pc);
    }
    const int mine = role == 0 ? PD_WB : role == 1 ? PD_EB : PD_BAND;
    for (int v = 0; v < VICGPU_N_OUTVARS; v++)
      if (out_var_group(v) == mine)
        for (int i = 0; i < o.L.out_nelem[v]; i++) gout[o.L.out_off[v] + i] = row[o.L.out_off[v] + i];
  }
  __syncthreads();  // the whole row of every cell of the block is in the output table
  if (live && role == 0) put_data_finish(o, cp, rec, RowRW{t.carry + c, nc}, gout, pc);
  if (rec < 0) return;  // (uniform)
  __syncthreads();
  if (live) put_data_aggregate_vars(o, t.aggtype, gout, agg, role, 3);
  __syncthreads();
  if (live && role == 0) put_data_aggregate_tail(o, step_count, agg);
}
#endif

// the cell's mass-balance curve at the end of an accumulation interval (GlacierMassBalanceResult.c:35-72): one point per band
// elevation that holds glacier HRUs, cumulative balances of HRUs at the same elevation added up, in hruList order
VIC_HDI void cell_gmb(const Opts* o, const Tables& t, int cell) {
  if (t.status[cell] != 0) return;  // the reference leaves an invalid cell alone (vicNl.c:521)
  const size_t nc = (size_t)t.ncell, nh = (size_t)t.nhru;
  CellPar cp{Col{t.cellpar + cell, nc}, &o->L};
  double x[VICGPU_MAX_BANDS], y[VICGPU_MAX_BANDS];
  int n = 0;
  for (int hh = t.cell_h0[cell]; hh < t.cell_h0[cell + 1]; hh++) {
    const int h = t.slot_of_hru ? t.slot_of_hru[hh] : hh;
    if (t.hrupar[(size_t)HP_isGlacier * nh + h] == 0.0) continue;
    const double cum = t.gmb_cum[h];
    if (!is_valid(cum)) continue;
    const double elev = cp.band(CB_BandElev, (int)t.hrupar[(size_t)HP_band * nh + h]);
    bool found = false;
    for (int k = 0; k < n; k++)
      if (x[k] == elev) {
        y[k] += cum;
        found = true;
      }
    if (!found && n < VICGPU_MAX_BANDS) {
      x[n] = elev;
      y[n] = cum;
      n++;
    }
  }
  // points at elevation 0 carry no data (GlacierMassBalanceResult.c:57-65: lastElevation stays 0)
  int m = 0;
  for (int k = 0; k < n; k++)
    if (!(x[k] == 0 && x[k] <= 0)) {
      x[m] = x[k];
      y[m] = y[k];
      m++;
    }
  double eq[4] = {0, 0, 0, -1};  // a GraphingEquation as constructed: no fit yet
  if (m > 0) gmb_fit(m, x, y, eq);
  for (int k = 0; k < 4; k++) t.gmb[(size_t)k * nc + cell] = eq[k];
}

// options as the kernels want them
inline int opts_from_abi(const vicgpu_options& a, Opts& o, const char** why) {
  *why = "";
  if (a.abi_version != VICGPU_ABI_VERSION) { *why = "abi_version mismatch"; return VICGPU_EINVAL; }
  if (a.Nlayer != VICGPU_NLAYER) { *why = "Nlayer must be 3"; return VICGPU_EUNSUPPORTED; }
  if (a.Nnode < 3 || a.Nnode > VICGPU_MAX_NODES) { *why = "Nnode out of range"; return VICGPU_EUNSUPPORTED; }
  if (a.Nbands < 1 || a.Nbands > VICGPU_MAX_BANDS) { *why = "Nbands out of range"; return VICGPU_EUNSUPPORTED; }
  if (a.DIST_PRCP) { *why = "DIST_PRCP is not implemented on the device"; return VICGPU_EUNSUPPORTED; }
  if (a.LAKES) { *why = "LAKES is not implemented on the device"; return VICGPU_EUNSUPPORTED; }
  // QUICK_SOLVE + IMPLICIT: fda_heat_eqn reads kappa_new[n + 1], which no evaluation of the same solve writes; with QUICK_SOLVE the number
  // of unknowns n changes between the searches of a step, so that entry holds what an earlier solve with more unknowns -- of whichever
  // HRU or cell the thread handled before -- left in the (static) array: the reference's answer depends on its OpenMP schedule.
  if (a.QUICK_SOLVE && a.IMPLICIT && !a.QUICK_FLUX) { *why = "QUICK_SOLVE with IMPLICIT has no defined answer in the reference"; return VICGPU_EUNSUPPORTED; }
  // GLACIER_DYNAMICS: glacier HRUs of zero area are stepped like the others (full_energy.c:220, 389; vic_step.cuh hru_step); the coupling
  // that changes their area between runs is host-side
  // COMPUTE_TREELINE: the host's initialize_atmos() decides which bands lie above the treeline (compute_treeline.c) and hands the flags
  // over in cellpar (CB_AboveTreeLine); the device side of the option is put_data's treatment of those bands (vic_output.cuh).
  // vicgpu_disagg refuses it: the decision needs the July temperatures of the whole forcing record.
  if (a.dt < 1 || a.SNOW_STEP < 1 || a.NF < 1 || a.out_step_ratio < 1) { *why = "bad time-step options"; return VICGPU_EINVAL; }
  o.Nnode = a.Nnode; o.Nbands = a.Nbands; o.dt = a.dt; o.SNOW_STEP = a.SNOW_STEP; o.NR = a.NR; o.NF = a.NF; o.nrecs = a.nrecs;
  o.out_step_ratio = a.out_step_ratio; o.FULL_ENERGY = a.FULL_ENERGY; o.FROZEN_SOIL = a.FROZEN_SOIL; o.QUICK_FLUX = a.QUICK_FLUX;
  o.QUICK_SOLVE = a.QUICK_SOLVE; o.IMPLICIT = a.IMPLICIT; o.EXP_TRANS = a.EXP_TRANS; o.NOFLUX = a.NOFLUX; o.GRND_FLUX_TYPE = a.GRND_FLUX_TYPE;
  o.AERO_RESIST_CANSNOW = a.AERO_RESIST_CANSNOW; o.SNOW_ALBEDO = a.SNOW_ALBEDO; o.SNOW_DENSITY = a.SNOW_DENSITY; o.TEMP_TH_TYPE = a.TEMP_TH_TYPE;
  o.TFALLBACK = a.TFALLBACK; o.GLACIER_ID = a.GLACIER_ID; o.GLACIER_DYNAMICS = a.GLACIER_DYNAMICS; o.MOISTFRACT = a.MOISTFRACT;
  o.ALMA_OUTPUT = a.ALMA_OUTPUT; o.NVegLibTypes = a.NVegLibTypes; o.CORRPREC = a.CORRPREC; o.BLOWING = a.BLOWING; o.gaYear = a.glacierAccumStartYear; o.gaMonth = a.glacierAccumStartMonth;
  o.gaDay = a.glacierAccumStartDay; o.gaInterval = a.glacierAccumInterval; o.wind_h = a.wind_h;
  vicgpu_layout_init(&o.L, &a);
  return VICGPU_OK;
}

// which instantiation (thermal-node array width) steps a configuration: 3, 10 or 32
inline int vic_node_width(const Opts& o) {
  if (o.Nnode <= 3 && !((o.IMPLICIT || o.QUICK_SOLVE) && !o.QUICK_FLUX) && !o.BLOWING) return 3;
  return o.Nnode <= 10 ? 10 : 32;
}

// Binning (host): the rows of the HRU tables are ordered by kind -- glacier / artificial bare soil / vegetation class --
// and by cell within a kind, so that the threads of a warp (consecutive rows) take the same branches of the step
// (surface_fluxes vs surface_fluxes_glac, overstory canopy balance, transpiration vs bare-soil evaporation) and the warps
// of a block run through the same code.  hrupar_rm: the caller's row-major HRU parameter records.
inline void bin_hrus(const double* hrupar_rm, int nhru, std::vector<int>& hru_of_slot, std::vector<int>& slot_of_hru) {
  std::vector<long long> key((size_t)nhru);
  for (int k = 0; k < nhru; k++) {
    const double* p = hrupar_rm + (size_t)k * HP_N;
    long long kind = (long long)p[HP_vegIndex];
    if (p[HP_isArtBare] != 0.0) kind += 1000000;
    if (p[HP_isGlacier] != 0.0) kind += 2000000;
    key[k] = kind;
  }
  hru_of_slot.resize((size_t)nhru);
  for (int k = 0; k < nhru; k++) hru_of_slot[k] = k;
  std::stable_sort(hru_of_slot.begin(), hru_of_slot.end(), [&](int a, int b) { return key[a] < key[b]; });
  slot_of_hru.resize((size_t)nhru);
  for (int s = 0; s < nhru; s++) slot_of_hru[hru_of_slot[s]] = s;
}

// accumulateGlacierMassBalance.c:13-66 calendar logic; `started` persists between records
inline GlacAccum glacier_accum_flags(const Opts& o, const int* dmy_rec, const int* dmy_next, int rec, bool* started) {
  GlacAccum g = {0, 0, 0, 0};
  if (o.gaYear == INT_MIN || o.gaMonth == INT_MIN || o.gaDay == INT_MIN || o.gaInterval == INT_MIN) return g;
  if (rec == o.nrecs) return g;
  g.enabled = 1;
  if (rec == 0) g.reset_first = 1;
  if (dmy_rec[4] == o.gaYear && dmy_rec[3] == o.gaMonth && dmy_rec[0] == o.gaDay) *started = true;
  if (*started) g.accumulate = 1;
  const int ny = dmy_next[4], nm = dmy_next[3], nd = dmy_next[0], nhour = dmy_next[2];
  int dy = ny - o.gaYear;
  if (dy < 0) dy = -dy;
  if ((ny > o.gaYear) && (dy % o.gaInterval == 0) && nm == o.gaMonth && nd == o.gaDay && (((o.dt <= 12) && (nhour == 0)) || (o.dt == 24)))
    g.reset_after = 1;
  return g;
}

}  // namespace vic
#endif

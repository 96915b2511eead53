// oracle/vicport.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, loaded by or shipped with the product).
//
// Host build (g++) of the SAME physics headers the CUDA library compiles (vic_b200/csrc/*.cuh),
// driven by a plain loop over records, HRUs and cells.  It exists so that the restated algorithm
// can be checked against the reference build (oracle/_ref/vic_ref_harness) in a container without
// a GPU: tests/ run it on a case file written by the harness and compare with the reference's
// answers stored in the same file.  libvicgpu.so has no path into this code.
//
// Usage: vicport <case.bin> <result.bin> [--nrec N]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "casefile.h"
#include "vic_engine.cuh"

using namespace vic;

static void die(const char* m) { fprintf(stderr, "vicport: %s\n", m); exit(2); }

template <int NN>
static void run_record(const Opts* o, Tables& t, const double* frec, Dmy d, int rec, GlacAccum ga) {
  if (o->NF == 1)
    for (int h = 0; h < t.nhru; h++) hru_work<NN, true>(o, t, frec, h, d, rec, ga);
  else
    for (int h = 0; h < t.nhru; h++) hru_work<NN, false>(o, t, frec, h, d, rec, ga);
}

// row-major record r -> column-major row (slot ? slot[r] : r)
static void to_colmajor(const std::vector<double>& rm, int nrow, int ncol, std::vector<double>& cm, const int* slot = nullptr) {
  cm.resize(rm.size());
  for (int r = 0; r < nrow; r++)
    for (int c = 0; c < ncol; c++) cm[(size_t)c * nrow + (slot ? slot[r] : r)] = rm[(size_t)r * ncol + c];
}
static void to_rowmajor(const double* cm, int nrow, int ncol, double* rm, const int* slot = nullptr) {
  for (int r = 0; r < nrow; r++)
    for (int c = 0; c < ncol; c++) rm[(size_t)r * ncol + c] = cm[(size_t)c * nrow + (slot ? slot[r] : r)];
}

int main(int argc, char** argv) {
  if (argc < 3) die("usage: vicport case.bin result.bin [--nrec N] [--binned]");
  int nrec_limit = -1;
  bool binned = false;  // keep the HRU tables in the device's binned row order (bin_hrus) instead of the caller's
  int roles = 0;        // reduce the cell outputs as the device does, three variable groups side by side (vic_engine.cuh cell_output)
  for (int i = 3; i < argc; i++) {
    if (!strcmp(argv[i], "--nrec") && i + 1 < argc) nrec_limit = atoi(argv[++i]);
    if (!strcmp(argv[i], "--binned")) binned = true;
    if (!strcmp(argv[i], "--roles")) roles = 1;
  }
  std::map<std::string, CaseArray> cs;
  if (!case_read(argv[1], cs)) die("cannot read case file");
  vicgpu_options abi;
  if (cs["options_raw"].i32.size() * 4 != sizeof(abi)) die("options size mismatch");
  memcpy(&abi, cs["options_raw"].i32.data(), sizeof(abi));
  Opts o;
  const char* why;
  if (opts_from_abi(abi, o, &why) != VICGPU_OK) { fprintf(stderr, "vicport: unsupported options: %s\n", why); return 3; }
  const vicgpu_layout& L = o.L;
  const int ncell = cs["meta"].i32[0], nhru = cs["meta"].i32[1];
  int nrec = cs["meta"].i32[2];
  if (nrec_limit > 0 && nrec_limit < nrec) nrec = nrec_limit;
  const int nout = L.out_off[VICGPU_N_OUTVARS];

  std::vector<double> cellpar, hrupar, hrurec;
  to_colmajor(cs["cellpar"].f64, ncell, L.cp_stride, cellpar);
  std::vector<int> hru_of_slot, slot_of_hru;
  if (binned) bin_hrus(cs["hrupar"].f64.data(), nhru, hru_of_slot, slot_of_hru);
  const int* slot = binned ? slot_of_hru.data() : nullptr;
  to_colmajor(cs["hrupar"].f64, nhru, HP_N, hrupar, slot);
  // the HRU state is tile-major (vic_types.cuh hr_off), like on the device
  hrurec.assign(hr_rows(nhru) * L.hr_stride, 0.0);
  for (int r = 0; r < nhru; r++)
    for (int c = 0; c < L.hr_stride; c++) hrurec[hr_off(slot ? slot[r] : r, L.hr_stride) + (size_t)c * VIC_HR_TILE] = cs["hrurec0"].f64[(size_t)r * L.hr_stride + c];
  std::vector<double> hdiag((size_t)3 * nhru, 0.0), carry((size_t)CC_N * ncell, 0.0), out((size_t)nout * ncell, 0.0), agg((size_t)nout * ncell, 0.0);
  std::vector<int> cell_h0(ncell + 1, 0), status(ncell, 0), fail_rec(ncell, INT_MAX);
  for (int h = 0; h < nhru; h++) cell_h0[(int)cs["hrupar"].f64[(size_t)h * HP_N + HP_cell] + 1]++;
  for (int c = 0; c < ncell; c++) cell_h0[c + 1] += cell_h0[c];
  if (cs.count("valid0"))
    for (int c = 0; c < ncell; c++) {
      status[c] = cs["valid0"].i32[c] ? 0 : ERROR_I;
      if (status[c] != 0) fail_rec[c] = -1;
    }

  // constants derived from the cell parameters, as the library computes them in vicgpu_set_cells
  std::vector<double> cellder((size_t)VIC_NCELLDER * ncell);
  for (int c = 0; c < ncell; c++) derive_cell_constants(CellPar{Col{cellpar.data() + c, (size_t)ncell}, &o.L}, cellder.data() + c, (size_t)ncell);
  std::vector<double> gmb_cum((size_t)nhru, 0.0), gmb((size_t)4 * ncell, 0.0);
  for (int c = 0; c < ncell; c++) gmb[(size_t)3 * ncell + c] = -1;  // GraphingEquation(): fitError -1
  Tables t;
  t.ncell = ncell; t.nhru = nhru; t.nclass = (int)cs["veglib"].dims[0];
  t.veglib = cs["veglib"].f64.data(); t.cellpar = cellpar.data(); t.cellder = cellder.data(); t.hrupar = hrupar.data(); t.hrurec = hrurec.data(); t.hrurec_out = hrurec.data(); t.hdiag_out = hdiag.data();
  t.cost = nullptr;
  t.aero = nullptr;  // the host port evaluates the aerodynamic geometry every record (vic_step.cuh aero_geom)
  t.cell_h0 = cell_h0.data(); t.status = status.data(); t.gmb_cum = gmb_cum.data(); t.gmb = gmb.data(); t.fail_rec = fail_rec.data(); t.carry = carry.data(); t.out = out.data(); t.agg = agg.data();
  t.aggtype = cs["aggtype"].i32.data();
  t.slot_of_hru = slot;

  const std::vector<double>& forcing = cs["forcing"].f64;  // [nrec][ncell][f_stride]
  const std::vector<int32_t>& dmy = cs["dmy"].i32;
  std::vector<int32_t> dump_recs;
  if (cs.count("dump_recs")) dump_recs = cs["dump_recs"].i32;
  std::vector<double> out_all((size_t)nrec * ncell * nout), agg_all, hru_all, frec((size_t)L.f_stride * ncell);
  std::vector<int32_t> agg_recs;
  int step_count = 0;
  bool started = false;
  size_t nd = 0;
  for (int rec = 0; rec < nrec; rec++) {
    step_count++;
    // forcing record -> [f_stride][ncell]
    for (int c = 0; c < ncell; c++)
      for (int k = 0; k < L.f_stride; k++) frec[(size_t)k * ncell + c] = forcing[((size_t)rec * ncell + c) * L.f_stride + k];
    Dmy d = {dmy[rec * 5 + 0], dmy[rec * 5 + 1], dmy[rec * 5 + 2], dmy[rec * 5 + 3], dmy[rec * 5 + 4]};
    if (rec == 0)
      for (int c = 0; c < ncell; c++) cell_output(o, t, nullptr, c, -1, step_count, roles);
    GlacAccum ga = glacier_accum_flags(o, &dmy[rec * 5], &dmy[(rec + 1) * 5], rec, &started);
    if (vic_node_width(o) == 3) run_record<3>(&o, t, frec.data(), d, rec, ga);
    else if (vic_node_width(o) == 10) run_record<10>(&o, t, frec.data(), d, rec, ga);
    else run_record<VICGPU_MAX_NODES>(&o, t, frec.data(), d, rec, ga);
    if (ga.enabled && ga.reset_after)
      for (int c = 0; c < ncell; c++) cell_gmb(&o, t, c);
    for (int c = 0; c < ncell; c++) cell_output(o, t, frec.data(), c, rec, step_count, roles);
    to_rowmajor(out.data(), ncell, nout, &out_all[(size_t)rec * ncell * nout]);
    while (nd < dump_recs.size() && dump_recs[nd] < rec) nd++;
    if (nd < dump_recs.size() && dump_recs[nd] == rec) {
      size_t base = hru_all.size();
      hru_all.resize(base + (size_t)nhru * L.hr_stride);
      for (int r = 0; r < nhru; r++)
        for (int c = 0; c < L.hr_stride; c++) hru_all[base + (size_t)r * L.hr_stride + c] = hrurec[hr_off(slot ? slot[r] : r, L.hr_stride) + (size_t)c * VIC_HR_TILE];
      nd++;
    }
    if (step_count == o.out_step_ratio) {
      size_t base = agg_all.size();
      agg_all.resize(base + (size_t)ncell * nout);
      to_rowmajor(agg.data(), ncell, nout, &agg_all[base]);
      agg_recs.push_back(rec);
      std::fill(agg.begin(), agg.end(), 0.0);
      step_count = 0;
    }
  }
  CaseWriter cw(argv[2]);
  if (!cw.ok()) die("cannot open result file");
  int64_t d3[3] = {nrec, ncell, nout};
  cw.f64("out", out_all.data(), 3, d3);
  int64_t d4[3] = {(int64_t)(hru_all.size() / ((size_t)nhru * L.hr_stride)), nhru, L.hr_stride};
  cw.f64("hrurec", hru_all.data(), 3, d4);
  int64_t d5[3] = {(int64_t)agg_recs.size(), ncell, nout};
  cw.f64("agg", agg_all.data(), 3, d5);
  std::vector<double> be((size_t)ncell * 5);
  for (int c = 0; c < ncell; c++)
    for (int k = 0; k < 5; k++) be[(size_t)c * 5 + k] = carry[(size_t)(CC_water_last_storage + k) * ncell + c];
  int64_t d6[2] = {ncell, 5};
  cw.f64("balance", be.data(), 2, d6);
  int64_t d7[1] = {ncell};
  std::vector<int32_t> st(status.begin(), status.end());
  cw.i32("status", st.data(), 1, d7);
  {
    std::vector<double> g((size_t)ncell * 4);
    for (int c = 0; c < ncell; c++)
      for (int k = 0; k < 4; k++) g[(size_t)c * 4 + k] = gmb[(size_t)k * ncell + c];
    int64_t d8[2] = {ncell, 4};
    cw.f64("gmb", g.data(), 2, d8);
  }
  return 0;
}

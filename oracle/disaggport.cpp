// oracle/disaggport.cpp -- TEST INFRASTRUCTURE ONLY.
// Host (g++) build of vic_b200/csrc/vic_disagg.cuh: runs the forcing-disaggregation stages as plain loops over
// cells and work items, so the restatement can be compared with the forcing the reference's initialize_atmos()
// produced (stored in every case file written by oracle/_ref/vic_ref_harness) without a GPU.
//
// Usage: disaggport <case.bin (options_raw, disagg_raw, meta, cellpar, daily)> <result.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "casefile.h"
#include "vic_engine.cuh"
#include "vic_disagg.cuh"

using namespace vic;

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: disaggport case.bin result.bin\n"); return 2; }
  std::map<std::string, CaseArray> cs;
  if (!case_read(argv[1], cs)) { fprintf(stderr, "cannot read case\n"); return 2; }
  vicgpu_options abi;
  vicgpu_disagg_options dabi;
  memcpy(&abi, cs["options_raw"].i32.data(), sizeof(abi));
  if (cs["disagg_raw"].i32.size() * 4 != sizeof(dabi)) { fprintf(stderr, "disagg options size mismatch\n"); return 2; }
  memcpy(&dabi, cs["disagg_raw"].i32.data(), sizeof(dabi));
  Opts o;
  const char* why;
  vicgpu_options abi2 = abi;
  abi2.LAKES = abi2.DIST_PRCP = abi2.BLOWING = abi2.CORRPREC = 0;
  if (opts_from_abi(abi2, o, &why) != VICGPU_OK) { fprintf(stderr, "options: %s\n", why); return 3; }
  const vicgpu_layout& L = o.L;
  const int ncell = (int)cs["cellpar"].dims[0];
  const DisaggOpts d = disagg_opts_from_abi(abi, dabi, L.f_nslot);
  // tables: column-major
  std::vector<double> cellpar((size_t)ncell * L.cp_stride), daily((size_t)ncell * d.Ndays * 4);
  for (int c = 0; c < ncell; c++)
    for (int k = 0; k < L.cp_stride; k++) cellpar[(size_t)k * ncell + c] = cs["cellpar"].f64[(size_t)c * L.cp_stride + k];
  const std::vector<double>& din = cs["daily"].f64;  // [ncell][Ndays][4]
  for (int c = 0; c < ncell; c++)
    for (int k = 0; k < d.Ndays * 4; k++) daily[(size_t)k * ncell + c] = din[(size_t)c * d.Ndays * 4 + k];
  DisaggScratch s;
  s.ncell = ncell; s.ntotal = ncell; s.cell0 = 0; s.Ndl = d.Ndays + 1;
  std::vector<double> scratch(s.per_cell() * (size_t)ncell, 0.0);
  s.base = scratch.data();
  std::vector<double> forcing((size_t)d.nrecs * L.f_stride * ncell, 0.0);  // [nrec][f_stride][ncell]
  for (int c = 0; c < ncell; c++) {
    CellPar cp{Col{cellpar.data() + c, (size_t)ncell}, &o.L};
    for (int i = 0; i < 365; i++) disagg_solar(cp, d, s, c, i);
    disagg_daily(cp, d, s, daily.data(), c);
    for (int day = 0; day < s.Ndl; day++) disagg_day_radiation(cp, d, s, c, day);
    for (int day = 0; day < s.Ndl; day++) disagg_day_maxmin(cp, d, s, c, day);
    for (int i = 0; i < 2 * s.Ndl + 2; i++) disagg_knot_coeff(cp, d, s, c, i);
    for (int day = 0; day < s.Ndl; day++) disagg_day_hourly(cp, d, s, c, day);
    for (int rec = 0; rec < d.nrecs; rec++) disagg_record(cp, d, s, daily.data(), forcing.data() + (size_t)rec * L.f_stride * ncell, c, rec);
  }
  // -> [nrec][ncell][f_stride]
  std::vector<double> out((size_t)d.nrecs * ncell * L.f_stride);
  for (int r = 0; r < d.nrecs; r++)
    for (int c = 0; c < ncell; c++)
      for (int k = 0; k < L.f_stride; k++) out[((size_t)r * ncell + c) * L.f_stride + k] = forcing[((size_t)r * L.f_stride + k) * ncell + c];
  CaseWriter cw(argv[2]);
  int64_t d3[3] = {d.nrecs, ncell, L.f_stride};
  cw.f64("forcing", out.data(), 3, d3);
  return 0;
}

// oracle/readercheck.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The indexed parameter lookups of the drop-in (vic_b200/host/vicgpu_fastread.h) against the reference's own scans, on the files of
// one global parameter file: every cell is read both ways -- read_vegparam() (read_vegparam.c:53) and read_snowband()
// (read_snowband.c:8) as the reference compiles them, and the same sources behind the index -- and the HRU lists, the band tables
// and ProgramState::initGrid()'s eight results (get_global_param.c:61-109) are compared bit for bit.  Prints the time both took.
// Usage: readercheck -g global.txt [--skip-stock-above N]   (exit 0: identical; N: the quadratic stock scans only on the first N cells)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "vicNl.h"
#include "WriteOutputNetCDF.h"
#include "vicgpu_pack.h"
#include "vicgpu_fastread.h"

void readSoilData(std::vector<cell_info_struct>& cell_data_structs, filep_struct filep, filenames_struct filenames, dmy_struct* dmy, ProgramState& state);  // vicNl.c:237

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Bands {
  std::vector<double> area, tf, pf;
  std::vector<float> elev;
  float elevation;
};
static soil_con_struct with_own_bands(const soil_con_struct& s, int nb) {
  soil_con_struct c = s;
  c.AreaFract = (double*)malloc(nb * sizeof(double)); memcpy(c.AreaFract, s.AreaFract, nb * sizeof(double));
  c.BandElev = (float*)malloc(nb * sizeof(float)); memcpy(c.BandElev, s.BandElev, nb * sizeof(float));
  c.Tfactor = (double*)malloc(nb * sizeof(double)); memcpy(c.Tfactor, s.Tfactor, nb * sizeof(double));
  c.Pfactor = (double*)malloc(nb * sizeof(double)); memcpy(c.Pfactor, s.Pfactor, nb * sizeof(double));
  return c;
}
static Bands take_bands(soil_con_struct& c, int nb) {
  Bands b{std::vector<double>(c.AreaFract, c.AreaFract + nb), std::vector<double>(c.Tfactor, c.Tfactor + nb), std::vector<double>(c.Pfactor, c.Pfactor + nb),
          std::vector<float>(c.BandElev, c.BandElev + nb), c.elevation};
  free(c.AreaFract); free(c.BandElev); free(c.Tfactor); free(c.Pfactor);
  return b;
}
static bool same(const Bands& a, const Bands& b) {
  auto eq = [](const void* x, const void* y, size_t n) { return memcmp(x, y, n) == 0; };
  return eq(a.area.data(), b.area.data(), a.area.size() * 8) && eq(a.tf.data(), b.tf.data(), a.tf.size() * 8) && eq(a.pf.data(), b.pf.data(), a.pf.size() * 8) &&
         eq(a.elev.data(), b.elev.data(), a.elev.size() * 4) && eq(&a.elevation, &b.elevation, 4);
}

int main(int argc, char** argv) {
  const char* global_file = NULL;
  long stock_limit = -1;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-g" && i + 1 < argc) global_file = argv[++i];
    else if (a == "--skip-stock-above" && i + 1 < argc) stock_limit = atol(argv[++i]);
    else { fprintf(stderr, "readercheck: bad argument\n"); return 2; }
  }
  if (!global_file) { fprintf(stderr, "readercheck: need -g <global file>\n"); return 2; }
  if (!freopen("/dev/null", "w", stderr)) {}

  // main(), vicNl.c:36-163
  ProgramState state;
  state.initialize_global();
  filenames_struct filenames;
  strcpy(filenames.global, global_file);
  state.build_forcing_variable_mapping();
  state.build_output_variable_mapping();
  state.init_global_param(&filenames, filenames.global);
  filep_struct filep = get_files(&filenames, &state);
  state.veg_lib = read_veglib(filep.veglib, &state.num_veg_types, state.options.LAI_SRC);
  dmy_struct* dmy = make_dmy(&state.global_param, &state);
  std::vector<cell_info_struct> cells;
  readSoilData(cells, filep, filenames, dmy, state);
  const size_t ncell = cells.size();
  const size_t nstock = stock_limit >= 0 && (size_t)stock_limit < ncell ? (size_t)stock_limit : ncell;
  const int nb = state.options.SNOW_BAND;
  int bad = 0;

  // ---- initGrid
  double t0 = now();
  global_param_struct g_fast = state.global_param;
  vicgpu_fastread::init_grid(g_fast, cells);
  const double t_grid_fast = now() - t0;
  double t_grid_stock = -1;
  if (nstock == ncell) {
    t0 = now();
    state.initGrid(cells);
    t_grid_stock = now() - t0;
    const global_param_struct& a = state.global_param;
    const double va[8] = {a.gridStartLat, a.gridStartLon, a.gridEndLat, a.gridEndLon, a.gridStepLat, a.gridStepLon, a.gridNumLatDivisions, a.gridNumLonDivisions};
    const double vb[8] = {g_fast.gridStartLat, g_fast.gridStartLon, g_fast.gridEndLat, g_fast.gridEndLon, g_fast.gridStepLat, g_fast.gridStepLon,
                          g_fast.gridNumLatDivisions, g_fast.gridNumLonDivisions};
    if (memcmp(va, vb, sizeof(va)) != 0) { printf("initGrid differs\n"); bad++; }
  }

  // ---- read_vegparam: the index first (all cells), then the reference's scan
  std::vector<std::vector<double>> hp_fast(ncell);
  std::vector<double> cvsum_fast(ncell);
  size_t nhru = 0;
  t0 = now();
  for (size_t c = 0; c < ncell; c++) {
    cell_info_struct cell = cells[c];
    vicgpu_fastread::read_vegparam_indexed(filep.vegparam, cell, &state);
    hp_fast[c].resize(cell.prcp.hruList.size() * (size_t)HP_N);
    for (size_t j = 0; j < cell.prcp.hruList.size(); j++) vicgpu_pack_hrupar(cell.prcp.hruList[j], (int)c, &hp_fast[c][j * HP_N]);
    cvsum_fast[c] = cell.Cv_sum;
    nhru += cell.prcp.hruList.size();
  }
  const double t_veg_fast = now() - t0;
  t0 = now();
  for (size_t c = 0; c < nstock; c++) {
    cell_info_struct cell = cells[c];
    read_vegparam(filep.vegparam, cell, &state);
    std::vector<double> hp(cell.prcp.hruList.size() * (size_t)HP_N);
    for (size_t j = 0; j < cell.prcp.hruList.size(); j++) vicgpu_pack_hrupar(cell.prcp.hruList[j], (int)c, &hp[j * HP_N]);
    if (hp.size() != hp_fast[c].size() || memcmp(hp.data(), hp_fast[c].data(), hp.size() * 8) != 0 || memcmp(&cell.Cv_sum, &cvsum_fast[c], 8) != 0) {
      if (bad < 5) printf("read_vegparam differs for cell %d\n", cells[c].soil_con.gridcel);
      bad++;
    }
  }
  const double t_veg_stock = now() - t0;

  // ---- read_snowband
  double t_band_fast = 0, t_band_stock = 0;
  if (nb > 1) {
    std::vector<Bands> bf(ncell);
    t0 = now();
    for (size_t c = 0; c < ncell; c++) {
      soil_con_struct s = with_own_bands(cells[c].soil_con, nb);
      vicgpu_fastread::read_snowband_indexed(filep.snowband, &s, nb);
      bf[c] = take_bands(s, nb);
    }
    t_band_fast = now() - t0;
    t0 = now();
    for (size_t c = 0; c < nstock; c++) {
      soil_con_struct s = with_own_bands(cells[c].soil_con, nb);
      read_snowband(filep.snowband, &s, nb);
      if (!same(take_bands(s, nb), bf[c])) {
        if (bad < 5) printf("read_snowband differs for cell %d\n", cells[c].soil_con.gridcel);
        bad++;
      }
    }
    t_band_stock = now() - t0;
  }
  printf("ncell %zu nhru %zu compared %zu bands %d\n", ncell, nhru, nstock, nb);
  printf("read_vegparam  indexed %.3f s (all cells)   reference scan %.3f s (%zu cells)\n", t_veg_fast, t_veg_stock, nstock);
  printf("read_snowband  indexed %.3f s (all cells)   reference scan %.3f s (%zu cells)\n", t_band_fast, t_band_stock, nstock);
  printf("initGrid       sorted  %.4f s               all pairs      %.3f s\n", t_grid_fast, t_grid_stock);
  printf(bad ? "DIFFERENT (%d)\n" : "identical\n", bad);
  return bad ? 1 : 0;
}

// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference physics (objects compiled from /root/reference by
// oracle/Makefile) through the same sequence runModel() uses (vicNl.c:390-654):
//   initializeCell (vicNl.c:295-385)  ->  per record, per cell:
//     put_data(rec=-nrecs) at rec 0 (vicNl.c:524-541), dist_prec (vicNl.c:543),
//     accumulateGlacierMassBalance (vicNl.c:563), aggdata reset at output steps (vicNl.c:596-609)
// and writes a "case file" (oracle/casefile.h) holding
//   (1) the flat C-ABI inputs produced by the product's host packer (vic_b200/host/vicgpu_pack.h):
//       options, veglib, cellpar, hrupar, hrurec0 (initial state), dmy, forcing
//   (2) the reference's answers: hrurec_ref at selected records, out_ref (OutputData::data of every
//       variable at every record), agg_ref (aggdata at output steps), balance errors, cell status.
// It is also the CPU baseline of bench.py (--time-only: OpenMP cell loop exactly as vicNl.c:514-517).
//
// Usage: vic_ref_harness -g global.txt [-o case.bin] [--nrec N] [--dump-every K] [--threads T]
//                        [--forcing-bin forcing.bin] [--time-only] [--time-from REC] [--no-run] [--verbose]
//                        [--agg-only]   keep out_ref only at output steps (agg_ref): large domains over long runs
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "vicNl.h"
#include "WriteOutputNetCDF.h"
#include "vicgpu_pack.h"
#include "casefile.h"
#if PARALLEL_AVAILABLE
#include <omp.h>
#endif

void readSoilData(std::vector<cell_info_struct> &cell_data_structs, filep_struct filep, filenames_struct filenames,
                  dmy_struct *dmy, ProgramState &state);  // vicNl.c:237

static void die(const char *msg) { fprintf(stderr, "vic_ref_harness: %s\n", msg); exit(2); }

// mirrors initializeCell(), vicNl.c:295-385, with the forcing source made switchable
static int init_cell(cell_info_struct &cell, filep_struct filep, dmy_struct *dmy, filenames_struct filenames,
                     const ProgramState *state, const double *forcing_bin, int cellidx, int ncell, const vicgpu_layout *L) {
  const int Ndist = state->options.DIST_PRCP ? 2 : 1;
  if (!state->options.OUTPUT_FORCE) {
    if (!forcing_bin) make_in_files(&filep, &filenames, &cell.soil_con, state);
    calc_root_fractions(cell.prcp.hruList, &cell.soil_con, state);
    read_snowband(filep.snowband, &cell.soil_con, state->options.SNOW_BAND);
  } else {
    make_in_files(&filep, &filenames, &cell.soil_con, state);
  }
  cell.atmos = alloc_atmos(state->global_param.nrecs, state->NR);
  if (forcing_bin) {
    for (int rec = 0; rec < state->global_param.nrecs; rec++)
      vicgpu_unpack_forcing(cell.atmos[rec], L, forcing_bin + ((size_t)rec * ncell + cellidx) * L->f_stride);
  } else {
    initialize_atmos(cell.atmos, dmy, filep.forcing, filep.forcing_ncid, &cell.soil_con, state);
    if (filep.forcing[0]) fclose(filep.forcing[0]);
    if (filep.forcing[1]) fclose(filep.forcing[1]);
  }
  cell.writeDebug.initialize(cell.prcp.hruList.size(), state);
  if (!state->options.OUTPUT_FORCE) {
    int ErrorFlag = initialize_model_state(&cell, dmy[0], filep, Ndist, filenames.init_state, state);
    if (ErrorFlag == ERROR) return ERROR;
  }
  return 0;
}

int main(int argc, char **argv) {
  const char *global_file = NULL, *out_path = NULL, *forcing_path = NULL;
  int nrec_limit = -1, dump_every = 0, threads = 1, time_from = 0;
  bool time_only = false, no_run = false, verbose = false, agg_only = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-g" && i + 1 < argc) global_file = argv[++i];
    else if (a == "-o" && i + 1 < argc) out_path = argv[++i];
    else if (a == "--nrec" && i + 1 < argc) nrec_limit = atoi(argv[++i]);
    else if (a == "--dump-every" && i + 1 < argc) dump_every = atoi(argv[++i]);
    else if (a == "--threads" && i + 1 < argc) threads = atoi(argv[++i]);
    else if (a == "--forcing-bin" && i + 1 < argc) forcing_path = argv[++i];
    else if (a == "--time-only") time_only = true;
    else if (a == "--time-from" && i + 1 < argc) time_from = atoi(argv[++i]);
    else if (a == "--no-run") no_run = true;
    else if (a == "--verbose") verbose = true;
    else if (a == "--agg-only") agg_only = true;
    else die("bad argument");
  }
  if (!global_file) die("need -g <global file>");
  if (!verbose) { if (!freopen("/dev/null", "w", stderr)) {} }
  FILE *log = stdout;

  // ---- exactly main(), vicNl.c:36-212, minus NetCDF output initialisation
  ProgramState state;
  state.initialize_global();
  filenames_struct filenames;
  strcpy(filenames.global, global_file);
  state.build_forcing_variable_mapping();
  state.build_output_variable_mapping();
  state.init_global_param(&filenames, filenames.global);
  OutputData *out_data_list = create_output_list(&state);
  out_data_file_struct *out_data_files = set_output_defaults(out_data_list, &state);
  parse_output_info(filenames.global, out_data_files, out_data_list, &state);
  filep_struct filep = get_files(&filenames, &state);
  if (!state.options.OUTPUT_FORCE) state.veg_lib = read_veglib(filep.veglib, &state.num_veg_types, state.options.LAI_SRC);
  dmy_struct *dmy = make_dmy(&state.global_param, &state);
  state.dt_sec = state.global_param.dt * SECPHOUR;
  state.out_dt_sec = state.global_param.out_dt * SECPHOUR;
  state.out_step_ratio = (int)(state.out_dt_sec / state.dt_sec);
  if (nrec_limit > 0 && nrec_limit < state.global_param.nrecs && forcing_path) state.global_param.nrecs = nrec_limit;
  std::vector<cell_info_struct> cells;
  readSoilData(cells, filep, filenames, dmy, state);
  const int ncell = (int)cells.size();
  if (!state.options.OUTPUT_FORCE)
    for (int c = 0; c < ncell; c++) {
      int numHRUs = read_vegparam(filep.vegparam, cells[c], &state);
      if (numHRUs > state.max_num_HRUs) state.update_max_num_HRUs(numHRUs);
    }
#if PARALLEL_AVAILABLE
  omp_set_num_threads(threads);
  omp_set_dynamic(0);
#endif
  state.global_param.num_threads = threads;

  vicgpu_options opt;
  vicgpu_pack_options(&state, &opt);
  vicgpu_layout L;
  vicgpu_layout_init(&L, &opt);
  const int nrecs_all = state.global_param.nrecs;
  const int nrec = (nrec_limit > 0 && nrec_limit < nrecs_all) ? nrec_limit : nrecs_all;

  // optional binary forcing (same layout vicgpu_set_forcing takes)
  std::vector<double> forcing_in;
  if (forcing_path) {
    std::map<std::string, CaseArray> fc;
    if (!case_read(forcing_path, fc) || !fc.count("forcing")) die("cannot read --forcing-bin");
    forcing_in.swap(fc["forcing"].f64);
    if (forcing_in.size() < (size_t)nrecs_all * ncell * L.f_stride) die("--forcing-bin too small");
  }

  auto t0 = std::chrono::steady_clock::now();
  for (int c = 0; c < ncell; c++) {
    if (init_cell(cells[c], filep, dmy, filenames, &state, forcing_path ? forcing_in.data() : NULL, c, ncell, &L) == ERROR)
      cells[c].isValid = FALSE;
  }
  auto t1 = std::chrono::steady_clock::now();
  fprintf(log, "init_seconds %.3f\n", std::chrono::duration<double>(t1 - t0).count());

  // ---- pack the flat C-ABI inputs with the product's host packer
  int nhru = 0;
  for (int c = 0; c < ncell; c++) nhru += (int)cells[c].prcp.hruList.size();
  const int nout = L.out_off[VICGPU_N_OUTVARS];
  fprintf(log, "ncell %d nhru %d nrec %d nout %d hr_stride %d cp_stride %d f_stride %d\n", ncell, nhru, nrec, nout, L.hr_stride, L.cp_stride, L.f_stride);

  CaseWriter *cw = NULL;
  if (out_path && !time_only) {
    cw = new CaseWriter(out_path);
    if (!cw->ok()) die("cannot open output case file");
    {
      int64_t d[1] = {(int64_t)(sizeof(opt) / sizeof(int32_t))};
      cw->i32("options_raw", (const int32_t *)&opt, 1, d);
      vicgpu_disagg_options dop;
      vicgpu_pack_disagg_options(&state, dmy, &dop);
      int64_t dd1[1] = {(int64_t)(sizeof(dop) / sizeof(int32_t))};
      cw->i32("disagg_raw", (const int32_t *)&dop, 1, dd1);
      int32_t meta[8] = {ncell, nhru, nrec, nout, L.hr_stride, L.cp_stride, L.f_stride, state.options.OUTPUT_FORCE};
      int64_t dm[1] = {8};
      cw->i32("meta", meta, 1, dm);
    }
    if (state.options.OUTPUT_FORCE) {  // the disaggregator mode has cells (soil parameters, bands) but no vegetation and no state
      std::vector<double> cellpar((size_t)ncell * L.cp_stride);
      for (int c = 0; c < ncell; c++) vicgpu_pack_cellpar(cells[c].soil_con, &L, &cellpar[(size_t)c * L.cp_stride]);
      int64_t dc[2] = {ncell, L.cp_stride};
      cw->f64("cellpar", cellpar.data(), 2, dc);
    }
    if (!state.options.OUTPUT_FORCE) {
      std::vector<double> veglib;
      vicgpu_pack_veglib(&state, &L, veglib);
      int64_t dv[2] = {(int64_t)(veglib.size() / L.vl_stride), L.vl_stride};
      cw->f64("veglib", veglib.data(), 2, dv);
      std::vector<double> cellpar((size_t)ncell * L.cp_stride), hrupar((size_t)nhru * HP_N), hrurec((size_t)nhru * L.hr_stride);
      int h = 0;
      for (int c = 0; c < ncell; c++) {
        vicgpu_pack_cellpar(cells[c].soil_con, &L, &cellpar[(size_t)c * L.cp_stride]);
        for (size_t k = 0; k < cells[c].prcp.hruList.size(); k++, h++) {
          vicgpu_pack_hrupar(cells[c].prcp.hruList[k], c, &hrupar[(size_t)h * HP_N]);
          vicgpu_pack_hrurec(cells[c].prcp.hruList[k], &L, &hrurec[(size_t)h * L.hr_stride]);
        }
      }
      int64_t dc[2] = {ncell, L.cp_stride}, dh[2] = {nhru, HP_N}, dr[2] = {nhru, L.hr_stride};
      cw->f64("cellpar", cellpar.data(), 2, dc);
      cw->f64("hrupar", hrupar.data(), 2, dh);
      cw->f64("hrurec0", hrurec.data(), 2, dr);
      std::vector<int32_t> agg(VICGPU_N_OUTVARS);
      for (int v = 0; v < VICGPU_N_OUTVARS; v++) agg[v] = out_data_list[v].aggtype;
      int64_t da[1] = {VICGPU_N_OUTVARS};
      cw->i32("aggtype", agg.data(), 1, da);
      std::vector<int32_t> valid(ncell);
      for (int c = 0; c < ncell; c++) valid[c] = cells[c].isValid ? 1 : 0;
      int64_t dvd[1] = {ncell};
      cw->i32("valid0", valid.data(), 1, dvd);
    }
    std::vector<int32_t> d5((size_t)(nrec + 1) * 5);
    for (int r = 0; r <= nrec; r++) {
      // make_dmy allocates and fills nrecs+1 entries (make_dmy.c:105-127)
      dmy_struct d = dmy[r];
      d5[(size_t)r * 5 + 0] = d.day; d5[(size_t)r * 5 + 1] = d.day_in_year; d5[(size_t)r * 5 + 2] = d.hour;
      d5[(size_t)r * 5 + 3] = d.month; d5[(size_t)r * 5 + 4] = d.year;
    }
    int64_t dd[2] = {nrec + 1, 5};
    cw->i32("dmy", d5.data(), 2, dd);
    // forcing [nrec][ncell][f_stride]
    int64_t df[3] = {nrec, ncell, L.f_stride};
    cw->header("forcing", 0, 3, df);
    std::vector<double> frow((size_t)ncell * L.f_stride);
    for (int r = 0; r < nrec; r++) {
      for (int c = 0; c < ncell; c++) vicgpu_pack_forcing(cells[c].atmos[r], &L, &frow[(size_t)c * L.f_stride]);
      cw->raw(frow.data(), frow.size() * 8);
    }
  }
  if (no_run || state.options.OUTPUT_FORCE) { delete cw; fprintf(log, "done (no run)\n"); return 0; }

  // ---- the time loop (vicNl.c:506-610)
  std::vector<OutputData *> current_output_data;
  for (int c = 0; c < ncell; c++) copy_output_data(current_output_data, out_data_list, &state);

  std::vector<int> dump_recs;
  if (dump_every > 0) for (int r = 0; r < nrec; r++) if ((r + 1) % dump_every == 0 || r == nrec - 1 || r == 0) dump_recs.push_back(r);
  std::vector<double> hru_dump, out_dump, agg_dump;
  std::vector<int32_t> agg_recs;
  if (cw && !agg_only) out_dump.resize((size_t)nrec * ncell * nout);

  auto t2 = std::chrono::steady_clock::now();
  size_t nd = 0;
  for (int rec = 0; rec < nrec; rec++) {
    if (rec == time_from) t2 = std::chrono::steady_clock::now();  // warm-up records are not timed
    state.step_count++;
#if PARALLEL_AVAILABLE
#pragma omp parallel for
#endif
    for (int c = 0; c < ncell; c++) {
      if (cells[c].isValid == FALSE) continue;
      if (rec == 0) {
        int e = put_data(&cells[c], cells[c].outputFormat, current_output_data[c], &dmy[0], -state.global_param.nrecs, &state);
        if (e == ERROR) { cells[c].isValid = FALSE; continue; }
      }
      int e = dist_prec(&cells[c], dmy, &filep, cells[c].outputFormat, current_output_data[c], rec, FALSE, &state);
      if (e == ERROR) cells[c].isValid = FALSE;
      if (cells[c].isValid)
        accumulateGlacierMassBalance(&(cells[c].gmbEquation), dmy, rec, &(cells[c].prcp), &(cells[c].soil_con), &state);
    }
    if (cw) {
      for (int c = 0; c < ncell && !agg_only; c++)
        vicgpu_pack_outdata(current_output_data[c], &L, &out_dump[((size_t)rec * ncell + c) * nout], false);
      if (nd < dump_recs.size() && dump_recs[nd] == rec) {
        size_t base = hru_dump.size();
        hru_dump.resize(base + (size_t)nhru * L.hr_stride);
        int h = 0;
        for (int c = 0; c < ncell; c++)
          for (size_t k = 0; k < cells[c].prcp.hruList.size(); k++, h++)
            vicgpu_pack_hrurec(cells[c].prcp.hruList[k], &L, &hru_dump[base + (size_t)h * L.hr_stride]);
        nd++;
      }
    }
    if (state.step_count == state.out_step_ratio) {
      if (cw) {
        size_t base = agg_dump.size();
        agg_dump.resize(base + (size_t)ncell * nout);
        for (int c = 0; c < ncell; c++) vicgpu_pack_outdata(current_output_data[c], &L, &agg_dump[base + (size_t)c * nout], true);
        agg_recs.push_back(rec);
      }
      for (int v = 0; v < N_OUTVAR_TYPES; v++)
        for (int c = 0; c < ncell; c++)
          for (int e = 0; e < out_data_list[v].nelem; e++) current_output_data[c][v].aggdata[e] = 0;
      state.step_count = 0;
    }
  }
  auto t3 = std::chrono::steady_clock::now();
  double secs = std::chrono::duration<double>(t3 - t2).count();
  int nvalid = 0;
  for (int c = 0; c < ncell; c++) nvalid += cells[c].isValid ? 1 : 0;
  fprintf(log, "run_seconds %.6f threads %d cell_steps %lld cell_steps_per_s %.1f valid_cells %d\n", secs, threads,
          (long long)ncell * (nrec - time_from), (double)ncell * (nrec - time_from) / secs, nvalid);

  if (cw) {
    std::vector<int32_t> dr(dump_recs.begin(), dump_recs.end());
    int64_t d1[1] = {(int64_t)dr.size()};
    cw->i32("dump_recs", dr.data(), 1, d1);
    int64_t d3[3] = {(int64_t)dr.size(), nhru, L.hr_stride};
    cw->f64("hrurec_ref", hru_dump.data(), 3, d3);
    int64_t d4[3] = {agg_only ? 0 : nrec, ncell, nout};
    cw->f64("out_ref", out_dump.data(), 3, d4);
    int64_t d5[1] = {(int64_t)agg_recs.size()};
    cw->i32("agg_recs", agg_recs.data(), 1, d5);
    int64_t d6[3] = {(int64_t)agg_recs.size(), ncell, nout};
    cw->f64("agg_ref", agg_dump.data(), 3, d6);
    std::vector<double> be((size_t)ncell * 5);
    std::vector<int32_t> st(ncell);
    for (int c = 0; c < ncell; c++) {
      be[c * 5 + 0] = cells[c].cellErrors.water_last_storage;
      be[c * 5 + 1] = cells[c].cellErrors.water_cum_error;
      be[c * 5 + 2] = cells[c].cellErrors.water_max_error;
      be[c * 5 + 3] = cells[c].cellErrors.energy_cum_error;
      be[c * 5 + 4] = cells[c].cellErrors.energy_max_error;
      st[c] = cells[c].isValid ? 0 : ERROR;
    }
    int64_t d7[2] = {ncell, 5};
    cw->f64("balance_ref", be.data(), 2, d7);
    int64_t d8[1] = {ncell};
    cw->i32("status_ref", st.data(), 1, d8);
    // the cell's glacier mass-balance curve as left by accumulateGlacierMassBalance() (what write_model_state.c:153-156 stores)
    std::vector<double> gm((size_t)ncell * 4);
    for (int c = 0; c < ncell; c++) {
      gm[c * 4 + 0] = cells[c].gmbEquation.b0;
      gm[c * 4 + 1] = cells[c].gmbEquation.b1;
      gm[c * 4 + 2] = cells[c].gmbEquation.b2;
      gm[c * 4 + 3] = cells[c].gmbEquation.fitError;
    }
    int64_t d9[2] = {ncell, 4};
    cw->f64("gmb_ref", gm.data(), 2, d9);
    delete cw;
  }
  return 0;
}

// vicNl_gpu.cpp -- the drop-in: runModel() of the reference's vicNl (vicNl.c:390-654) with the per-cell CPU dispatch
// (vicNl.c:506-610: put_data(rec = -nrecs), dist_prec, accumulateGlacierMassBalance for every cell and record) replaced by batched
// calls into libvicgpu.so.  Everything else stays the reference's own code, reached through its own headers: main() and the global
// parameter file (vicNl.c:36-212), the soil / vegetation / snow-band readers, initializeCell() with the forcing readers,
// initialize_atmos() and initialize_model_state() (vicNl.c:295-385), the output writer and write_model_state().
//
// Build (oracle/Makefile target _ref/vicNl_gpu): the reference's objects + this file + -lvicgpu.  vicNl.c is compiled as it is; its
// own runModel is made a weak symbol (objcopy --weaken-symbol) so that the definition below is the one main() calls.
//
// OUTPUT_FORCE runs (the meteorological disaggregator, vicNl.c:445-490) go through vicgpu_disagg: run_output_force() below.  There is no CPU
// fallback: without a CUDA device, or with an option the device code does not implement, the run stops with the library's message.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "vicNl.h"
#include "WriteOutputNetCDF.h"
#include "vicgpu_pack.h"

int initializeCell(cell_info_struct& cell, filep_struct filep, dmy_struct* dmy, filenames_struct filenames, const ProgramState* state);  // vicNl.c:295

static void gpu_check(int rc, const char* what) {
  if (rc == VICGPU_OK) return;
  char msg[MAXSTRING];
  snprintf(msg, sizeof(msg), "libvicgpu: %s failed (%d): %s", what, rc, vicgpu_last_error());
  vicerror(msg);
}

// ---- OUTPUT_FORCE TRUE: the meteorological disaggregator (vicNl.c:420-490) ---------------------------------------------------
// The reference, per cell: initializeCell() -> initialize_atmos() (read the daily forcing, MTCLIM, hourly arrays), then
// write_forcing_file() + write_data_one_cell() per record.  Here the cells' daily PREC / TMAX / TMIN / WIND are read with the
// reference's own read_forcing_data(), a chunk of cells is disaggregated on the device by vicgpu_disagg, and the records go through
// the reference's own write_forcing_file() / write_data_one_cell() in the reference's order (cell by cell, record by record).
// Served: what vicgpu_disagg takes -- daily PREC, TMAX, TMIN and WIND from one forcing file per cell (FORCE_DT 24), no ALMA_INPUT.
static void run_output_force(std::vector<cell_info_struct>& cells, filep_struct filep, filenames_struct filenames,
                             out_data_file_struct* out_data_files_template, OutputData* out_data_list, dmy_struct* dmy, ProgramState* state) {
  const int ncell = (int)cells.size(), nrecs = state->global_param.nrecs;
  for (int t = 0; t < N_FORCING_TYPES; t++) {
    const bool wanted = t == PREC || t == TMAX || t == TMIN || t == WIND;
    const int sup = state->param_set.TYPE[t].SUPPLIED;
    if (wanted != (sup != 0) || (wanted && (sup != 1 || state->param_set.FORCE_DT[0] != 24)))
      vicerror("vicNl_gpu: OUTPUT_FORCE on the device takes daily PREC, TMAX, TMIN and WIND from the first forcing file (FORCE_DT 24) and nothing else");
  }
  if (state->options.ALMA_INPUT) vicerror("vicNl_gpu: OUTPUT_FORCE with ALMA_INPUT is not served");
  vicgpu_options opt;
  vicgpu_pack_options(state, &opt);
  vicgpu_disagg_options dopt;
  vicgpu_pack_disagg_options(state, dmy, &dopt);
  vicgpu_handle* h = NULL;
  const char* dev = getenv("VICGPU_DEVICE");
  gpu_check(vicgpu_create(&h, &opt, dev ? atoi(dev) : 0), "vicgpu_create");
  vicgpu_layout L;
  gpu_check(vicgpu_get_layout(h, &L), "vicgpu_get_layout");
  {  // the disaggregator mode reads no vegetation (vicNl.c:145-151): placeholder library rows for set_cells' range checks
    std::vector<double> veglib((size_t)(opt.NVegLibTypes + 4) * L.vl_stride, 0.0);
    gpu_check(vicgpu_set_veglib(h, opt.NVegLibTypes + 4, veglib.data()), "vicgpu_set_veglib");
  }
  // cells per device pass: at most ~512 MB of hourly forcing on the host at a time
  const size_t per_cell = (size_t)nrecs * L.f_stride * sizeof(double);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ncell, ((size_t)512 << 20) / per_cell));
  std::vector<OutputData*> current_output_data;
  for (int i = 0; i < state->global_param.disagg_write_chunk_size; i++) copy_output_data(current_output_data, out_data_list, state);
  std::vector<double> daily, forcing, cellpar, hrupar;
  fprintf(stderr, "Disaggregating forcings on the GPU...\n");
  auto t_start = std::chrono::system_clock::now();
  for (int c0 = 0; c0 < ncell; c0 += chunk) {
    const int n = std::min(chunk, ncell - c0);
    daily.assign((size_t)n * dopt.Ndays * 4, 0.0);
    cellpar.assign((size_t)n * L.cp_stride, 0.0);
    hrupar.assign((size_t)n * HP_N, 0.0);
    for (int c = 0; c < n; c++) {
      cell_info_struct& cell = cells[c0 + c];
      make_in_files(&filep, &filenames, &cell.soil_con, state);  // vicNl.c:335-337
      double** fd = read_forcing_data(filep.forcing, filep.forcing_ncid, state->global_param, &cell.soil_con, state);  // initialize_atmos.c:247
      const int types[4] = {PREC, TMAX, TMIN, WIND};
      for (int d = 0; d < dopt.Ndays; d++)
        for (int k = 0; k < 4; k++) daily[((size_t)c * dopt.Ndays + d) * 4 + k] = fd[types[k]][d];
      for (int t = 0; t < N_FORCING_TYPES; t++) free(fd[t]);
      free(fd);
      if (filep.forcing[0]) fclose(filep.forcing[0]);
      if (filep.forcing[1]) fclose(filep.forcing[1]);
      vicgpu_pack_cellpar(cell.soil_con, &L, &cellpar[(size_t)c * L.cp_stride]);
      hrupar[(size_t)c * HP_N + HP_cell] = c;  // one placeholder HRU per cell (no vegetation in this mode); it is never stepped
      hrupar[(size_t)c * HP_N + HP_Cv] = 1.0;
    }
    gpu_check(vicgpu_set_cells(h, n, cellpar.data(), n, hrupar.data()), "vicgpu_set_cells");
    forcing.resize((size_t)nrecs * n * L.f_stride);
    gpu_check(vicgpu_disagg(h, &dopt, daily.data(), forcing.data()), "vicgpu_disagg");
    for (int c = 0; c < n; c++) {
      cell_info_struct& cell = cells[c0 + c];
      cell.atmos = alloc_atmos(nrecs, state->NR);
      for (int rec = 0; rec < nrecs; rec++) vicgpu_unpack_forcing(cell.atmos[rec], &L, &forcing[((size_t)rec * n + c) * L.f_stride]);
      copy_data_file_format(out_data_files_template, cell.outputFormat->dataFiles, state);
      make_out_files(&filep, &filenames, &cell.soil_con, cell.outputFormat, state);
      // vicNl.c:462-480, unchanged
      int chunk_step_count = 0, chunk_start_rec = 0;
      for (int rec = 0; rec < nrecs; rec++) {
        write_forcing_file(&cell, rec, cell.outputFormat, current_output_data[chunk_step_count], state, dmy);
        chunk_step_count++;
        if (rec >= nrecs - 1) {
          cell.outputFormat->write_data_one_cell(current_output_data, out_data_files_template, chunk_start_rec, nrecs - chunk_start_rec, state);
        } else if (chunk_step_count >= state->global_param.disagg_write_chunk_size) {
          cell.outputFormat->write_data_one_cell(current_output_data, out_data_files_template, chunk_start_rec, state->global_param.disagg_write_chunk_size, state);
          chunk_step_count = 0;
          chunk_start_rec = rec + 1;
        }
      }
      free_atmos(nrecs, &cell.atmos);
      delete cell.outputFormat;
    }
  }
  gpu_check(vicgpu_destroy(h), "vicgpu_destroy");
  std::chrono::duration<double> elapsed = std::chrono::system_clock::now() - t_start;
  fprintf(stderr, "\nVIC forcing disaggregation done. Execution time (GPU): %.3f seconds\n", elapsed.count());
  for (int c = 0; c < ncell; c++) {  // vicNl.c:640-652 (no vegetation, no atmos left)
    free(cells[c].soil_con.AreaFract);
    free(cells[c].soil_con.BandElev);
    free(cells[c].soil_con.Tfactor);
    free(cells[c].soil_con.Pfactor);
    free(cells[c].soil_con.AboveTreeLine);
  }
}

// ---- VICGPU_NC_OUTPUT=<file>: the output through the library's own NetCDF writer ------------------------------------------------
// What WriteOutputNetCDF::initializeFile() and write_data_all_cells() (WriteOutputNetCDF.c:163-299, 386-452) put into the file, taken
// from the same sources -- the output files' variable lists, ProgramState::output_mapping, the grid of initGrid() and the modelled-cell
// mask of initCellMask() -- but one record per output step, filled from the float32 rows vicgpu_step_f32 returns (vicgpu_ncwrite.h).
static vicgpu_ncout* open_nc_output(const char* path, out_data_file_struct* files, OutputData* out_data_list, const vicgpu_layout& L, int ncell,
                                    const ProgramState* state, std::vector<int>& col_of_var) {
  const global_param_struct& gp = state->global_param;
  const int nlat = (int)gp.gridNumLatDivisions, nlon = (int)gp.gridNumLonDivisions;
  // cell k sits where the k-th set bit of the mask is (WriteOutputNetCDF.c:419-430)
  std::vector<int> lat_index, lon_index;
  for (int g = 0; g < nlat * nlon; g++)
    if (state->modeled_cell_mask[g]) {
      lat_index.push_back(g / nlon);
      lon_index.push_back(g % nlon);
    }
  if ((int)lat_index.size() != ncell) vicerror("vicNl_gpu: the modelled-cell mask does not hold one grid point per cell");
  std::vector<std::string> keep;  // owns the strings the spec points to
  keep.reserve(8 * N_OUTVAR_TYPES + 64);
  auto own = [&](const std::string& x) { keep.push_back(x); return keep.back().c_str(); };
  std::vector<vicgpu_ncout_var> vars;
  col_of_var.clear();
  for (int f = 0; f < state->options.Noutfiles; f++)
    for (int v = 0; v < files[f].nvars; v++) {
      const int id = files[f].varid[v];
      const std::string varname = out_data_list[id].varname;
      if (state->output_mapping.find(varname) == state->output_mapping.end()) vicerror("vicNl_gpu: output variable missing from output_mapping");
      const VariableMetaData& m = state->output_mapping.at(varname);
      vicgpu_ncout_var a;
      a.name = own(m.name); a.nelem = out_data_list[id].nelem; a.long_name = own(m.longName); a.units = own(m.units); a.standard_name = own(m.standardName);
      a.cell_methods = own(m.cellMethods); a.internal_vic_name = own(varname); a.category = own(files[f].prefix);
      vars.push_back(a);
      col_of_var.push_back(L.out_off[id]);
    }
  std::string units = gp.out_dt < 24 ? "hours since " : "days since ";  // WriteOutputNetCDF.c:221-229
  units += std::to_string(gp.startyear) + "-" + std::to_string(gp.startmonth) + "-" + std::to_string(gp.startday);
  if (gp.out_dt < 24) units += " " + std::to_string(gp.starthour) + ":00";
  std::vector<const char*> tk, tv, ik;
  std::vector<int> iv;
  tk.push_back("title"); tv.push_back("VIC model run output.");
  for (const auto& a : gp.netCDFGlobalAttributes) { tk.push_back(own(a.first)); tv.push_back(own(a.second)); }
  tk.push_back("source"); tv.push_back("VIC (time loop on libvicgpu)");
  tk.push_back("frequency"); tv.push_back(own(gp.out_dt < 24 ? std::to_string(gp.out_dt) + " hour" : std::string("day")));
  tk.push_back("Conventions"); tv.push_back("CF-1.6");
  const char* inames[] = {"model_start_year", "model_start_month", "model_start_day", "model_start_hour", "model_end_year", "model_end_month", "model_end_day"};
  const int ivals[] = {gp.startyear, gp.startmonth, gp.startday, gp.starthour, gp.endyear, gp.endmonth, gp.endday};
  for (int k = 0; k < 7; k++) { ik.push_back(inames[k]); iv.push_back(ivals[k]); }
  vicgpu_ncout_spec sp;
  memset(&sp, 0, sizeof(sp));
  sp.nlat = nlat; sp.nlon = nlon; sp.depth = MAX_BANDS;
  sp.lat0 = gp.gridStartLat; sp.dlat = gp.gridStepLat; sp.lon0 = gp.gridStartLon; sp.dlon = gp.gridStepLon;
  sp.time_units = units.c_str();
  sp.time_step = gp.out_dt < 24 ? gp.out_dt : 1;
  sp.nvar = (int)vars.size(); sp.vars = vars.data();
  sp.ntext = (int)tk.size(); sp.text_keys = tk.data(); sp.text_values = tv.data();
  sp.nint = (int)ik.size(); sp.int_keys = ik.data(); sp.int_values = iv.data();
  sp.ncell = ncell; sp.lat_index = lat_index.data(); sp.lon_index = lon_index.data();
  vicgpu_ncout* w = NULL;
  gpu_check(vicgpu_ncout_create(&w, path, &sp), "vicgpu_ncout_create");
  return w;
}

void runModel(std::vector<cell_info_struct>& cell_data_structs, filep_struct filep, filenames_struct filenames,
              out_data_file_struct* out_data_files_template, OutputData* out_data_list, dmy_struct* dmy, ProgramState* state) {
  if (state->options.OUTPUT_FORCE) {
    run_output_force(cell_data_structs, filep, filenames, out_data_files_template, out_data_list, dmy, state);
    return;
  }
  std::vector<OutputData*> current_output_data;
  WriteOutputNetCDF* outputwriter = new WriteOutputNetCDF(state);
  outputwriter->openFile();
  const int ncell = (int)cell_data_structs.size();
  const int nrecs = state->global_param.nrecs;

  // ---- per-cell initialisation, as vicNl.c:420-445
  for (int c = 0; c < ncell; c++) {
    if (initializeCell(cell_data_structs[c], filep, dmy, filenames, state) == ERROR) cell_data_structs[c].isValid = FALSE;
    copy_data_file_format(out_data_files_template, cell_data_structs[c].outputFormat->dataFiles, state);
    make_out_files(&filep, &filenames, &cell_data_structs[c].soil_con, cell_data_structs[c].outputFormat, state);
    copy_output_data(current_output_data, out_data_list, state);
  }
  fprintf(stderr, "Running Model on the GPU...\n");
  auto t_start = std::chrono::system_clock::now();

  // ---- hand the domain to the device
  vicgpu_options opt;
  vicgpu_pack_options(state, &opt);
  vicgpu_handle* h = NULL;
  const char* dev = getenv("VICGPU_DEVICE");
  gpu_check(vicgpu_create(&h, &opt, dev ? atoi(dev) : 0), "vicgpu_create");
  vicgpu_layout L;
  gpu_check(vicgpu_get_layout(h, &L), "vicgpu_get_layout");
  const int nout = L.out_off[VICGPU_N_OUTVARS];
  {
    std::vector<double> veglib;
    vicgpu_pack_veglib(state, &L, veglib);
    gpu_check(vicgpu_set_veglib(h, (int)(veglib.size() / L.vl_stride), veglib.data()), "vicgpu_set_veglib");
  }
  int nhru = 0;
  for (int c = 0; c < ncell; c++) nhru += (int)cell_data_structs[c].prcp.hruList.size();
  std::vector<double> hrurec((size_t)nhru * L.hr_stride);
  std::vector<int> status(ncell);
  {
    std::vector<double> cellpar((size_t)ncell * L.cp_stride), hrupar((size_t)nhru * HP_N);
    int k = 0;
    for (int c = 0; c < ncell; c++) {
      vicgpu_pack_cellpar(cell_data_structs[c].soil_con, &L, &cellpar[(size_t)c * L.cp_stride]);
      status[c] = cell_data_structs[c].isValid ? 0 : ERROR;
      for (size_t j = 0; j < cell_data_structs[c].prcp.hruList.size(); j++, k++) {
        vicgpu_pack_hrupar(cell_data_structs[c].prcp.hruList[j], c, &hrupar[(size_t)k * HP_N]);
        vicgpu_pack_hrurec(cell_data_structs[c].prcp.hruList[j], &L, &hrurec[(size_t)k * L.hr_stride]);
      }
    }
    gpu_check(vicgpu_set_cells(h, ncell, cellpar.data(), nhru, hrupar.data()), "vicgpu_set_cells");
  }
  gpu_check(vicgpu_set_cell_status(h, status.data()), "vicgpu_set_cell_status");
  {
    std::vector<int> aggtype(N_OUTVAR_TYPES);
    for (int v = 0; v < N_OUTVAR_TYPES; v++) aggtype[v] = out_data_list[v].aggtype;
    gpu_check(vicgpu_set_output_spec(h, aggtype.data()), "vicgpu_set_output_spec");
  }
  gpu_check(vicgpu_set_state(h, hrurec.data()), "vicgpu_set_state");
  // VICGPU_NC_OUTPUT=<file>: float32 rows straight into the library's NetCDF writer instead of OutputData + the reference's writer
  const char* nc_path = getenv("VICGPU_NC_OUTPUT");
  vicgpu_ncout* ncout = NULL;
  std::vector<int> col_of_var;
  std::vector<float> agg32;
  if (nc_path && *nc_path) ncout = open_nc_output(nc_path, out_data_files_template, out_data_list, L, ncell, state, col_of_var);

  // the record after which the state file is written (vicNl.c:569-577)
  int state_rec = -1;
  if (state->options.SAVE_STATE == TRUE)
    for (int rec = 0; rec < nrecs; rec++)
      if (dmy[rec].year == state->global_param.stateyear && dmy[rec].month == state->global_param.statemonth && dmy[rec].day == state->global_param.stateday &&
          (rec + 1 == nrecs || dmy[rec + 1].day != state->global_param.stateday))
        state_rec = rec;

  // ---- time loop: one call per block of records instead of nrecs x ncell calls of dist_prec()
  // a block holds at most ~256 MB of packed forcing and never runs past the state-file record
  const size_t per_rec = (size_t)ncell * L.f_stride;
  const int Bmax = (int)std::max<size_t>(1, std::min<size_t>((size_t)nrecs, ((size_t)256 << 20) / (per_rec * sizeof(double))));
  std::vector<double> forcing((size_t)Bmax * per_rec), agg;
  std::vector<int> dmy5((size_t)(Bmax + 1) * 5), out_recs;
  for (int rec0 = 0; rec0 < nrecs;) {
    int n = std::min(Bmax, nrecs - rec0);
    if (state_rec >= rec0 && state_rec < rec0 + n) n = state_rec - rec0 + 1;
    for (int r = 0; r < n; r++)
      for (int c = 0; c < ncell; c++) vicgpu_pack_forcing(cell_data_structs[c].atmos[rec0 + r], &L, &forcing[((size_t)r * ncell + c) * L.f_stride]);
    for (int r = 0; r <= n; r++) {  // make_dmy() fills nrecs + 1 entries (make_dmy.c:105-127)
      const dmy_struct& d = dmy[rec0 + r];
      int* p = &dmy5[(size_t)r * 5];
      p[0] = d.day; p[1] = d.day_in_year; p[2] = d.hour; p[3] = d.month; p[4] = d.year;
    }
    // the output steps inside the block: state->step_count runs exactly as in vicNl.c:512, 596-609
    out_recs.clear();
    for (int r = 0; r < n; r++) {
      state->step_count++;
      if (state->step_count == state->out_step_ratio) {
        out_recs.push_back(rec0 + r);
        state->step_count = 0;
      }
    }
    gpu_check(vicgpu_set_forcing(h, rec0, n, forcing.data()), "vicgpu_set_forcing");
    if (ncout) {
      agg32.resize(std::max<size_t>(1, out_recs.size()) * (size_t)ncell * nout);
      gpu_check(vicgpu_step_f32(h, rec0, n, dmy5.data(), NULL, out_recs.empty() ? NULL : agg32.data()), "vicgpu_step_f32");
    } else {
      agg.resize(std::max<size_t>(1, out_recs.size()) * (size_t)ncell * nout);
      gpu_check(vicgpu_step(h, rec0, n, dmy5.data(), NULL, out_recs.empty() ? NULL : agg.data()), "vicgpu_step");
    }
    // cells the step invalidated (dist_prec would have returned ERROR, vicNl.c:545-559)
    gpu_check(vicgpu_get_cell_status(h, status.data()), "vicgpu_get_cell_status");
    for (int c = 0; c < ncell; c++)
      if (status[c] != 0 && cell_data_structs[c].isValid) {
        cell_data_structs[c].isValid = FALSE;
        if (state->options.CONTINUEONERROR == TRUE) {
          fprintf(stderr, "Error processing cell %d (method dist_prec) in records %d-%d.  Cell has been marked as invalid and will be skipped for remainder of model run.  "
                          "An incomplete output file has been generated, check your inputs before re-running the simulation.\n",
                  cell_data_structs[c].soil_con.gridcel, rec0, rec0 + n - 1);
        } else {
          sprintf(cell_data_structs[c].ErrStr, "Error processing cell %d (method dist_prec) in records %d-%d so the simulation has ended. Check your inputs before re-running the simulation.\n",
                  cell_data_structs[c].soil_con.gridcel, rec0, rec0 + n - 1);
          vicerror(cell_data_structs[c].ErrStr);
        }
      }
    // the unchanged output writer consumes OutputData::aggdata (vicNl.c:596-598)
    for (size_t s = 0; s < out_recs.size(); s++) {
      if (out_recs[s] < state->global_param.skipyear) continue;
      if (ncout) {
        gpu_check(vicgpu_ncout_write_step(ncout, &agg32[s * (size_t)ncell * nout], nout, col_of_var.data()), "vicgpu_ncout_write_step");
        continue;
      }
      for (int c = 0; c < ncell; c++) vicgpu_unpack_outdata(current_output_data[c], &L, &agg[(s * ncell + c) * nout], true);
      outputwriter->write_data_all_cells(current_output_data, out_data_files_template, out_recs[s] / state->out_step_ratio, state);
    }
    // the state file: pull the HRU records (and the glacier mass-balance curves) back into the reference's structs
    if (state_rec == rec0 + n - 1) {
      gpu_check(vicgpu_get_state(h, hrurec.data()), "vicgpu_get_state");
      std::vector<double> gmb((size_t)ncell * 4);
      gpu_check(vicgpu_get_glacier_fit(h, gmb.data()), "vicgpu_get_glacier_fit");
      int k = 0;
      for (int c = 0; c < ncell; c++) {
        for (size_t j = 0; j < cell_data_structs[c].prcp.hruList.size(); j++, k++)
          vicgpu_unpack_hrurec(cell_data_structs[c].prcp.hruList[j], &L, &hrurec[(size_t)k * L.hr_stride]);
        cell_data_structs[c].gmbEquation.b0 = gmb[(size_t)c * 4 + 0];
        cell_data_structs[c].gmbEquation.b1 = gmb[(size_t)c * 4 + 1];
        cell_data_structs[c].gmbEquation.b2 = gmb[(size_t)c * 4 + 2];
        cell_data_structs[c].gmbEquation.fitError = gmb[(size_t)c * 4 + 3];
        if (cell_data_structs[c].isValid) write_model_state(&cell_data_structs[c], filenames.statefile, state);
      }
    }
    rec0 += n;
  }
  if (ncout) gpu_check(vicgpu_ncout_close(ncout), "vicgpu_ncout_close");
  gpu_check(vicgpu_destroy(h), "vicgpu_destroy");
  std::chrono::duration<double> elapsed = std::chrono::system_clock::now() - t_start;
  fprintf(stderr, "\nVIC model run done. Model execution time (GPU): %.3f seconds\n", elapsed.count());

  // ---- clean-up, as vicNl.c:632-652
  if (state->param_set.FORCE_FORMAT[0] == NETCDF) close_files(&filep, &filenames, state->options.COMPRESS, state);
  for (int c = 0; c < ncell; c++) {
    cell_data_structs[c].writeDebug.cleanup(cell_data_structs[c].prcp.hruList.size(), state);
    free_atmos(state->global_param.nrecs, &cell_data_structs[c].atmos);
    delete cell_data_structs[c].outputFormat;
    free_vegcon(cell_data_structs[c]);
    free(cell_data_structs[c].soil_con.AreaFract);
    free(cell_data_structs[c].soil_con.BandElev);
    free(cell_data_structs[c].soil_con.Tfactor);
    free(cell_data_structs[c].soil_con.Pfactor);
    free(cell_data_structs[c].soil_con.AboveTreeLine);
  }
}

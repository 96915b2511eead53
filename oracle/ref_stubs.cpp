// oracle/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The reference's NetCDF writers (WriteOutputNetCDF.c, StateIONetCDF.c) need
// netcdf-cxx4, which this image does not have.  The oracle build leaves those two
// files out and supplies stand-ins of the two classes here so that the rest of the
// reference links.  Nothing on the dist_prec -> full_energy -> surface_fluxes path
// calls into them.  The output writer's stand-in appends OutputData::aggdata of every
// cell and variable as raw doubles to <RESULT_DIR>/<netCDF output name>.f64 at every
// output step (the real writer narrows to float32, WriteOutputNetCDF.c:279, useless
// for bit-level parity): that file is what tests/test_dropin.py compares between the
// stock vicNl and the GPU drop-in vicNl_gpu (and, for OUTPUT_FORCE runs, <...>.force.f64 written by write_data_one_cell).
#include <cstdio>
#include <stdexcept>
#include <string>
#include "vicNl.h"
#include "WriteOutputNetCDF.h"
#include "StateIONetCDF.h"

WriteOutputNetCDF::WriteOutputNetCDF(const ProgramState *state) : WriteOutputFormat(state), netCDF(NULL), timeIndexDivisor(1) {}
WriteOutputNetCDF::~WriteOutputNetCDF() {}
const char *WriteOutputNetCDF::getDescriptionOfOutputType() { return "oracle-stub"; }
void WriteOutputNetCDF::initializeFile(const ProgramState *, const OutputData *) {}
void WriteOutputNetCDF::openFile() {}
void WriteOutputNetCDF::compressFiles() {}
// disaggregator mode (OUTPUT_FORCE TRUE, vicNl.c:462-480): the chunk's aggdata of every variable, record by record, appended to
// <RESULT_DIR>/<netCDF output name>.force.f64 (cells follow each other in the order the run processes them)
void WriteOutputNetCDF::write_data_one_cell(std::vector<OutputData *> &chunk, out_data_file_struct *, const int, const int num_recs, const ProgramState *state) {
  static bool first = true;
  const std::string path = std::string(state->options.NETCDF_FULL_FILE_PATH) + ".force.f64";
  FILE *f = fopen(path.c_str(), first ? "wb" : "ab");
  first = false;
  if (!f) throw std::runtime_error("cannot open " + path);
  for (int r = 0; r < num_recs; r++)
    for (int v = 0; v < N_OUTVAR_TYPES; v++) fwrite(chunk[r][v].aggdata, sizeof(double), chunk[r][v].nelem, f);
  fclose(f);
}
void WriteOutputNetCDF::write_data_all_cells(std::vector<OutputData *> &all, out_data_file_struct *, const int output_rec, const ProgramState *state) {
  const std::string path = std::string(state->options.NETCDF_FULL_FILE_PATH) + ".f64";
  FILE *f = fopen(path.c_str(), output_rec == 0 ? "wb" : "ab");
  if (!f) throw std::runtime_error("cannot open " + path);
  for (size_t c = 0; c < all.size(); c++)
    for (int v = 0; v < N_OUTVAR_TYPES; v++) fwrite(all[c][v].aggdata, sizeof(double), all[c][v].nelem, f);
  fclose(f);
}
void WriteOutputNetCDF::write_header(OutputData *, const dmy_struct *, const ProgramState *) {}
int WriteOutputNetCDF::getLengthOfTimeDimension(const ProgramState *) { return 0; }
int WriteOutputNetCDF::getTimeIndex(const dmy_struct *, const int, const ProgramState *) { return 0; }

static void no_netcdf() { throw std::runtime_error("NetCDF state files are not available in the oracle build"); }
StateIONetCDF::StateIONetCDF(std::string filename, IOType ioType, const ProgramState *state) : StateIO(filename, ioType, state) { no_netcdf(); }
StateIONetCDF::~StateIONetCDF() {}
void StateIONetCDF::initializeOutput() {}
int StateIONetCDF::write(const int *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::write(const double *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::write(const float *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::write(const bool *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::write(const char *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::read(int *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::read(double *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::read(float *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::read(bool *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
int StateIONetCDF::read(char *, int, const StateVariables::StateMetaDataVariableIndices) { return -1; }
StateHeader StateIONetCDF::readHeader() { no_netcdf(); return StateHeader(0, 0, 0, 0, 0); }
void StateIONetCDF::notifyDimensionUpdate(StateVariables::StateVariableDimensionId, int) {}
void StateIONetCDF::initializeDimensionIndices() {}
int StateIONetCDF::getCurrentDimensionIndex(StateVariables::StateVariableDimensionId) { return 0; }
int StateIONetCDF::seekToCell(int, int *, int *) { return -1; }
void StateIONetCDF::flush() {}
void StateIONetCDF::rewindFile() {}

// vicgpu_internal.h -- the library handle and small helpers shared by the translation units of libvicgpu.so
#ifndef VICGPU_INTERNAL_H
#define VICGPU_INTERNAL_H
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "vicgpu_kernels.h"

extern thread_local std::string vicgpu_err;
inline int vicgpu_fail(int code, const std::string& msg) {
  vicgpu_err = msg;
  return code;
}
#define CK(call)                                                                                                \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return vicgpu_fail(VICGPU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct vicgpu_handle {
  int device = 0;
  vicgpu_options abi;
  vic::Opts o;
  vic::Opts* d_o = nullptr;
  vic::Tables t;
  int nout = 0;
  // HRU state and step diagnostics are double-buffered: record r reads buffer `cur` and writes `cur ^ 1` (vicgpu_step)
  double *d_hrurec2[2] = {nullptr, nullptr}, *d_hdiag2[2] = {nullptr, nullptr};
  int cur = 0;
  int* d_fail_rec = nullptr;
  cudaStream_t stream_out = nullptr;  // cell output (put_data) of record r runs here, beside the HRU step of record r + 1
  cudaEvent_t ev_step = nullptr, ev_out[2] = {nullptr, nullptr};
  bool overlap = true;
  double *d_veglib = nullptr, *d_cellpar = nullptr, *d_hrupar = nullptr, *d_carry = nullptr,
         *d_out = nullptr, *d_agg = nullptr, *d_stage = nullptr, *d_forcing = nullptr, *d_fstage = nullptr;
  size_t stage_elems = 0, forcing_cap = 0, fstage_cap = 0;
  int *d_cell_h0 = nullptr, *d_status = nullptr, *d_aggtype = nullptr, *d_slot_of_hru = nullptr, *d_hru_of_slot = nullptr;
  bool binned = true;
  // dynamic re-binning (vicgpu_api.cu rebin_rows)
  int rebin_interval = 24, recs_since_bin = 1 << 30;
  double* d_hrupar_alt = nullptr;
  unsigned long long* d_keys[2] = {nullptr, nullptr};
  int* d_oldslot[2] = {nullptr, nullptr};
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int hru_block = VICGPU_HRU_BLOCK;
  int frec0 = 0, fnrec = 0;
  bool have_cells = false, have_state = false, glac_started = false;
  int step_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0;
  long long last_launches = 0;
  // optional per-launch timing of the per-HRU step kernel (vicgpu_set_profiling)
  bool profiling = false, warp_timing = false;
  unsigned long long* d_warp_ns = nullptr;  // [2 * nwarps] start / end of every warp of the last profiled step launch
  std::vector<cudaEvent_t> pev;
  double prof_hru_ms = 0;
  long long prof_hru_launches = 0;
};


// in: [batch][rows][cols] row-major  ->  out: [batch][cols][rows]   (vicgpu_api.cu)
int vicgpu_transpose(vicgpu_handle* h, const double* d_in, double* d_out, int rows, int cols, int batch, cudaStream_t st = nullptr);
int vicgpu_ensure_forcing(vicgpu_handle* h, size_t elems);

#endif

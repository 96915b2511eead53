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

// one half of the double-buffered state (vicgpu_api.cu "state halves")
struct StateHalf {
  double* in = nullptr;      // [hr_stride][nhru] input state of a record block, in this half's row order
  double* snap = nullptr;    // tiles [hr_stride][32]: state after the record
  double* hdiag = nullptr;   // [3][nhru] Cv-weighted out_prec / out_rain / out_snow of the record
  int ord = 0;               // which RowOrder its rows are in
};
// a device-resident window of forcing records [rec0, rec0 + nrec): [nrec][f_stride][ncell]
struct ForcingWindow {
  double* d = nullptr;
  size_t cap = 0;              // doubles allocated
  int rec0 = 0, nrec = 0;
  cudaEvent_t ready = nullptr;  // recorded on the copy stream when the window's upload is complete
};
// a row order of the HRU tables: the parameter table in that order and the maps row <-> caller's HRU index (null: identity)
struct RowOrder {
  double* hrupar = nullptr;  // [HP_N][nhru]
  int *slot_of_hru = nullptr, *hru_of_slot = nullptr;
};

struct vicgpu_handle {
  int device = 0;
  vicgpu_options abi;
  vic::Opts o;
  vic::Opts* d_o = nullptr;
  vic::Tables t;
  int nout = 0;
  StateHalf half[2];
  RowOrder order[2];
  int rebin_every = 24, recs_since_rebin = 1 << 30;  // records between re-sorts of the rows
  int cur_half = 0;               // the half whose row order d_state_cur is in
  double* d_state_cur = nullptr;  // current state: half.in after set_state, else the last snapshot of the last block
  bool binned = true, rebin = true, pdl = true, bin_fine = false, pdl_wait = true, even = false;
  int* d_fail_rec = nullptr;
  int* d_cost = nullptr;  // [nhru] cost estimate of each row's last step (Tables::cost)
  int* d_work = nullptr;  // per-thread cost counters of the step kernel (vic_frozen.cuh vic_count_work)
  int bin_cost = 0, bin_cost_env = -1;  // VICGPU_BINCOST
  bool rebin_env = false;
  cudaStream_t stream = nullptr, stream_copy = nullptr;  // kernels / host <-> device copies
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // output staging: two buffers handed between the kernel stream (transpose) and the copy stream (D2H)
  double* d_ostage[2] = {nullptr, nullptr};
  cudaEvent_t ev_stage_full[2] = {nullptr, nullptr}, ev_stage_free[2] = {nullptr, nullptr};
  int stage_idx = 0;
  ForcingWindow fwin[2];
  int fwin_next = 0;
  double *d_gmb_cum = nullptr, *d_gmb = nullptr;  // glacier mass-balance fit (vic_engine.cuh cell_gmb)
  double *d_veglib = nullptr, *d_cellpar = nullptr, *d_cellder = nullptr, *d_carry = nullptr, *d_out = nullptr, *d_agg = nullptr, *d_stage = nullptr,
         *d_fstage = nullptr;
  size_t stage_elems = 0, fstage_cap = 0;
  int *d_cell_h0 = nullptr, *d_status = nullptr, *d_aggtype = nullptr;
  int hru_block = VICGPU_HRU_BLOCK;
  bool hru_block_fixed = false;  // set through VICGPU_BLOCK
  int sm_count = 148;
  int l2_persist_mb = 0;  // VICGPU_L2PERSIST
  bool aero_cache = true;   // VICGPU_AEROCACHE
  double* d_aero = nullptr;  // [VIC_AERO_NCOL][nhru] aerodynamic geometry of every row for month aero_month (k_hru_aero), in the current row order
  int aero_month = -1;
  // Balanced blocks of a single-wave domain (vicgpu_api.cu balance_blocks): the step kernel of a domain that fits the machine at once
  // ends when its slowest block does, and the land-cover kinds differ by 30 % in duration; kinds that run long get fewer warps per block
  bool balance = true;                 // VICGPU_BALANCE
  int bal_reserve = 28;
  bool bal_active = false;             // this domain is balanced (binned by kind first, one wave)
  bool bal_measured = false;           // d_warp_ns holds the warp times of a launch with the current layout
  bool bal_measure_next = false;       // time the warps of the next launch
  std::vector<int> warp_kind;          // [nwarp] dense index of the kind of each warp's first row (static: the kind is the first sort key)
  std::vector<int> kind_n;             // [nkind] warps per block of each kind in the current layout
  std::vector<int> h_block_w0;         // [nb + 1] first warp of each block
  int* d_block_w0 = nullptr;
  int nb_balanced = 0;
  long long sync_limit = 0;  // PhaseSync::limit
  // re-binning scratch (vicgpu_api.cu rebin_rows)
  unsigned long long* d_keys[2] = {nullptr, nullptr};
  int* d_oldslot[2] = {nullptr, nullptr};
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  bool have_cells = false, have_state = false, glac_started = false;
  int step_count = 0;
  double last_ms = 0;
  long long last_launches = 0;
  // optional timing of the step kernel launches (vicgpu_set_profiling)
  bool profiling = false, warp_timing = false;
  unsigned long long* d_warp_ns = nullptr;  // [2 * nwarps] start / end of every warp of the last profiled step launch
  std::vector<cudaEvent_t> pev;
  double prof_hru_ms = 0;
  long long prof_hru_launches = 0;
};


// in: [batch][rows][cols] row-major  ->  out: [batch][cols][rows]   (vicgpu_api.cu)
int vicgpu_transpose(vicgpu_handle* h, const double* d_in, double* d_out, int rows, int cols, int batch, cudaStream_t st = nullptr);
int vicgpu_ensure_window(vicgpu_handle* h, ForcingWindow& w, size_t elems);

#endif

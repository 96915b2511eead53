// vicgpu.cu -- libvicgpu.so: CUDA kernels (sm_100a, FP64) and the C-ABI declared in include/vicgpu.h.
//
// Device data layout: every table is column-major ("structure of arrays") so that consecutive
// threads (= consecutive HRUs, the HRUs of one cell being adjacent) read consecutive addresses:
//   cellpar [cp_stride][ncell]   hrupar [HP_N][nhru]   hrurec [hr_stride][nhru]
//   forcing [nrec][f_stride][ncell]   out / agg [nout][ncell]
// The C-ABI speaks row-major records (what a host packer naturally produces); the conversion is
// a tiled transpose ON THE DEVICE after a straight host->device copy, never a host loop.
//
// There is no CPU path in this file: without a CUDA device vicgpu_create fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include "vicgpu_internal.h"

using namespace vic;

thread_local std::string vicgpu_err;
#define fail vicgpu_fail

// ---- kernels ---------------------------------------------------------------------------------
// in: [batch][rows][cols] row-major  ->  out: [batch][cols][rows]
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
  __shared__ double tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[boff + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// The same conversions when the device row of record r is not r (HRU tables are kept binned by kind, vic_engine.cuh bin_hrus):
// host record r (row-major) <-> device row row_of[r] (column-major).  One thread per element, coalesced on the device side.
__global__ void k_scatter_rows(const double* __restrict__ in /* [rows][cols] */, double* __restrict__ out /* [cols][rows] */, int rows, int cols,
                               const int* __restrict__ rec_of_row) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int c = (int)(i / rows), r = (int)(i % rows);
  out[i] = in[(size_t)rec_of_row[r] * cols + c];
}
__global__ void k_gather_rows(const double* __restrict__ in /* [cols][rows] */, double* __restrict__ out /* [rows][cols] */, int rows, int cols,
                              const int* __restrict__ row_of_rec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  out[i] = in[(size_t)c * rows + row_of_rec[r]];
}

// the same two conversions for the tile-major HRU state (vic_types.cuh hr_off)
__global__ void k_scatter_state(const double* __restrict__ in /* [rows][cols] */, double* __restrict__ out /* tiles */, int rows, int cols,
                                const int* __restrict__ rec_of_row /* may be null */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int c = (int)(i / rows), r = (int)(i % rows);
  out[hr_off(r, cols) + (size_t)c * VIC_HR_TILE] = in[(size_t)(rec_of_row ? rec_of_row[r] : r) * cols + c];
}
__global__ void k_gather_state(const double* __restrict__ in /* tiles */, double* __restrict__ out /* [rows][cols] */, int rows, int cols,
                               const int* __restrict__ row_of_rec /* may be null */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  out[i] = in[hr_off(row_of_rec ? row_of_rec[r] : r, cols) + (size_t)c * VIC_HR_TILE];
}
// tile-major state: out row s = in row src[s]
__global__ void k_permute_state(const double* __restrict__ in, double* __restrict__ out, int rows, int cols, const int* __restrict__ src) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int c = (int)(i / rows), r = (int)(i % rows);
  out[hr_off(r, cols) + (size_t)c * VIC_HR_TILE] = in[hr_off(src[r], cols) + (size_t)c * VIC_HR_TILE];
}

// ---- dynamic binning: rows of the HRU tables re-ordered by (kind, snow on the ground or in the canopy, cell) ----------------
// The step's control flow differs most between glacier / bare / vegetated HRUs (static: bin_hrus) and between HRUs with and
// without snow (dynamic: solve_snow's pack and canopy balances, sub-stepping, evaporation switched off under snow).  Every
// `rebin_interval` records the rows are re-sorted on the device so that the 32 HRUs of a warp share both.
__global__ void k_bin_keys(const double* __restrict__ hrupar, const double* __restrict__ hrurec, int nhru, int hr_stride, const int* __restrict__ hru_of_slot,
                           unsigned long long* __restrict__ keys, int* __restrict__ old_slot, int fine, const int* __restrict__ cost, int cost_mode) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nhru) return;
  const size_t n = (size_t)nhru;
  unsigned long long kind = (unsigned long long)(long long)hrupar[(size_t)HP_vegIndex * n + s] & 0xfffffull;
  if (hrupar[(size_t)HP_isArtBare * n + s] != 0.0) kind |= 1ull << 20;
  if (hrupar[(size_t)HP_isGlacier * n + s] != 0.0) kind |= 1ull << 21;
  const size_t so = hr_off(s, hr_stride);
  const double swq = hrurec[so + (size_t)HR_S_swq * VIC_HR_TILE], canopy = hrurec[so + (size_t)HR_S_snow_canopy * VIC_HR_TILE];
  unsigned long long regime = (swq > 0.0 || canopy > 0.0) ? 1 : 0;
  if (fine) {
    // VICGPU_BINFINE=1: the pack's regime decides which branches of snow_melt / solve_snow run -- no pack on the ground, a pack whose
    // surface is at the melting point (the balance at 0 C usually closes without a solve) or a cold pack (Brent solve) -- and snow in
    // the canopy decides whether snow_intercept's canopy balance runs
    const double ts = hrurec[so + (size_t)HR_S_surf_temp * VIC_HR_TILE];
    const unsigned long long pack = !(swq > 0.0) ? 0 : ((ts < 0.0) ? 1 : 2);
    regime = (pack << 1) | (canopy > 0.0 ? 1 : 0);
  }
  // VICGPU_BINCOST: the cost of the row's last step (residual evaluations + frozen-node solves, HruStepDiag::work) in powers of two.
  // A warp costs what its slowest lane costs, so rows of similar cost share warps: mode 1 sorts by cost inside a (kind, regime) bin,
  // mode 2 sorts by cost FIRST (the frozen-soil configuration, where an HRU with frozen nodes costs 10-100 times one without and
  // the code path is the same for all of them).
  unsigned long long bucket = 0;
  if (cost_mode && cost) {
    int c = cost[s];
    while (c > 0 && bucket < 15) { c >>= 1; bucket++; }
  }
  // low 32 bits: the HRU's own index, i.e. cell order within a bin (and a total order: the sort is deterministic)
  if (cost_mode == 2) keys[s] = (bucket << 58) | (kind << 35) | (regime << 32) | (unsigned long long)(unsigned)hru_of_slot[s];
  else keys[s] = (kind << 39) | (regime << 36) | (bucket << 32) | (unsigned long long)(unsigned)hru_of_slot[s];
  old_slot[s] = s;
}
// out[c][s] = in[c][src[s]]
__global__ void k_permute_rows(const double* __restrict__ in, double* __restrict__ out, int rows, int cols, const int* __restrict__ src) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int c = (int)(i / rows), r = (int)(i % rows);
  out[i] = in[(size_t)c * rows + src[r]];
}
__global__ void k_slot_maps(const unsigned long long* __restrict__ keys, int nhru, int* __restrict__ hru_of_slot, int* __restrict__ slot_of_hru) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nhru) return;
  const int hru = (int)(unsigned)(keys[s] & 0xffffffffull);
  hru_of_slot[s] = hru;
  slot_of_hru[hru] = s;
}

// FP64 FMA throughput probe: 8 independent chains per thread, nothing but DFMA in the loop
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;  // keeps the chains alive
}

// constants derived from the cell parameters, once per vicgpu_set_cells (vic_soil.cuh derive_cell_constants)
__global__ void k_derive_cells(const Opts* __restrict__ o, const double* __restrict__ cellpar, double* __restrict__ cellder, int ncell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  derive_cell_constants(CellPar{Col{cellpar + c, (size_t)ncell}, &o->L}, cellder + c, (size_t)ncell);
}

// the cells' glacier mass-balance curves at the end of an accumulation interval (vic_engine.cuh cell_gmb)
// The wind-independent part of the seven CalcAerodynamic() evaluations of every row (vic_step.cuh AeroGeom) for the month `month0`:
// run when the month or the row order changes, read by every step of the month.
__global__ void __launch_bounds__(128) k_hru_aero(const Opts* __restrict__ o, Tables t, double* __restrict__ aero, int month0) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= t.nhru) return;
  const size_t nh = (size_t)t.nhru;
  Col hpc{t.hrupar + h, nh};
  const int cell = (int)hpc(HP_cell);
  const CellPar cp{Col{t.cellpar + cell, (size_t)t.ncell}, &o->L, Col{nullptr, 0}};
  const VegLib vl{t.veglib, &o->L};
  const int veg_class = (int)hpc(HP_vegIndex);
  const VegNow veg = veg_now(vl, veg_class, month0);
  AeroGeom g;
  aero_geom(vl, cp, *o, hpc(HP_isGlacier) != 0.0, veg_class, month0, veg, g);
  for (int k = 0; k < VIC_AERO_NCOL; k++) aero[(size_t)k * nh + h] = g.v[k];
}

__global__ void k_cell_gmb(const Opts* __restrict__ o, Tables t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= t.ncell) return;
  cell_gmb(o, t, c);
}

// three threads per cell in three warps of a 96-thread block (vic_engine.cuh cell_output_role); the options travel as a kernel
// parameter: the output offsets L.out_off[VOUT_x] are then operands in the constant bank instead of a dependent global load each
__global__ void __launch_bounds__(96) k_cell_output(const __grid_constant__ Opts o, Tables t, const double* __restrict__ forcing_rec, int rec, int step_count,
                                                    int wait_primary) {
  const int role = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool live = c < t.ncell && !(rec >= 0 && t.fail_rec[c] <= rec);
  cell_output_role(o, t, forcing_rec, c, rec, step_count, role, live);
  // Launched as a programmatic dependent of step(rec + 1) (vicgpu_step), this grid starts while that step is still running.  It
  // reads only what step(rec) left behind -- hru_work of step(rec + 1) writes the OTHER state half, and its atomicMin on fail_rec
  // cannot change the `fail_rec <= rec` test made here -- so the work above needs no ordering with it.  The wait below is what makes
  // COMPLETION of this grid imply completion of the step grid it rode on: step(rec + 2), the transposes and copies that follow are
  // ordinary launches ordered after this grid only (PTX ISA griddepcontrol.wait).  A no-op for an ordinary launch.
  if (wait_primary) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// [rows][cols] -> [cols][rows] with the float32 narrowing WriteOutputNetCDF applies to every value it writes (WriteOutputNetCDF.c:279,
// 351, 412): a plain (float) conversion, round to nearest
__global__ void k_transpose_f32(const double* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = (float)tile[threadIdx.x][i];
  }
}

int vicgpu_transpose(vicgpu_handle* h, const double* d_in, double* d_out, int rows, int cols, int batch, cudaStream_t st) {
  if (!st) st = h->stream;
  // the tile rows ride on gridDim.y: 65,535 x 32 = 2,097,120 rows (cells of a shard, or columns of a table) per launch
  if ((rows + 31) / 32 > 65535 || batch > 65535) return vicgpu_fail(VICGPU_EUNSUPPORTED, "more than 2,097,120 cells per device: shard the domain (vic_b200/shard.py)");
  dim3 b(32, 8), g((cols + 31) / 32, (rows + 31) / 32, batch);
  k_transpose<<<g, b, 0, st>>>(d_in, d_out, rows, cols);
  h->last_launches++;
  CK(cudaGetLastError());
  return VICGPU_OK;
}

int vicgpu_ensure_window(vicgpu_handle* h, ForcingWindow& w, size_t need) {
  if (need > w.cap) {
    // nothing queued may still read or write the old buffer
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->stream_copy));
    cudaFree(w.d);
    w.d = nullptr;
    w.cap = 0;
    CK(cudaMalloc(&w.d, need * sizeof(double)));
    w.cap = need;
  }
  return VICGPU_OK;
}

static int ensure_stage(vicgpu_handle* h, size_t elems) {
  if (elems <= h->stage_elems) return VICGPU_OK;
  if (h->d_stage) cudaFree(h->d_stage);
  h->d_stage = nullptr;
  h->stage_elems = 0;
  CK(cudaMalloc(&h->d_stage, elems * sizeof(double)));
  h->stage_elems = elems;
  return VICGPU_OK;
}

// host row-major [rows][cols] -> device column-major [cols][rows]; rec_of_row (device, may be null): record held by each device row
static int upload_transposed(vicgpu_handle* h, const double* host, double* d_dst, int rows, int cols, const int* rec_of_row = nullptr) {
  int rc = ensure_stage(h, (size_t)rows * cols);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_stage, host, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (rec_of_row) {
    const size_t n = (size_t)rows * cols;
    k_scatter_rows<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_stage, d_dst, rows, cols, rec_of_row);
    h->last_launches++;
    CK(cudaGetLastError());
  } else {
    rc = vicgpu_transpose(h, h->d_stage, d_dst, rows, cols, 1);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  return VICGPU_OK;
}
// device column-major [cols][rows] -> host row-major [rows][cols]; row_of_rec (device, may be null): device row of each record
static int download_transposed(vicgpu_handle* h, const double* d_src, double* host, int rows, int cols, bool sync, const int* row_of_rec = nullptr) {
  int rc = ensure_stage(h, (size_t)rows * cols);
  if (rc) return rc;
  if (row_of_rec) {
    const size_t n = (size_t)rows * cols;
    k_gather_rows<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_src, h->d_stage, rows, cols, row_of_rec);
    h->last_launches++;
    CK(cudaGetLastError());
  } else {
    rc = vicgpu_transpose(h, d_src, h->d_stage, cols, rows, 1);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(host, h->d_stage, (size_t)rows * cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (sync) CK(cudaStreamSynchronize(h->stream));
  return VICGPU_OK;
}

// ---- state halves ------------------------------------------------------------------------------------------------------------
// The step kernel of record r reads the state in one half and writes the state after the record into the other half's snapshot, so
// that the cell output of record r (which reads that snapshot) can run beside the step of record r + 1.  A half also owns the row
// order its snapshot is in (the rows are re-sorted by kind and snow state every `rebin_every` records) together with the HRU
// parameter table in that order.
static void free_half(StateHalf& s) {
  cudaFree(s.in); cudaFree(s.snap); cudaFree(s.hdiag);
  s.in = s.snap = s.hdiag = nullptr;
  s.ord = 0;
}
static void free_order(RowOrder& r) {
  cudaFree(r.hrupar); cudaFree(r.slot_of_hru); cudaFree(r.hru_of_slot);
  r.hrupar = nullptr;
  r.slot_of_hru = r.hru_of_slot = nullptr;
}

// Re-sort the rows by (kind, snow, HRU): state `src_state` in row order S  ->  buffer dst_state in the new row order D (parameter
// table and maps).  Runs on h->stream; the caller has made sure nothing still reads order D or dst_state.
static int rebin_rows(vicgpu_handle* h, const RowOrder& S, RowOrder& D, const double* src_state, double* dst_state) {
  const int nhru = h->t.nhru;
  const vicgpu_layout& L = h->o.L;
  const int T = 256, G = (nhru + T - 1) / T;
  k_bin_keys<<<G, T, 0, h->stream>>>(S.hrupar, src_state, nhru, L.hr_stride, S.hru_of_slot, h->d_keys[0], h->d_oldslot[0], h->bin_fine ? 1 : 0, h->d_cost, h->bin_cost);
  size_t need = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, need, h->d_keys[0], h->d_keys[1], h->d_oldslot[0], h->d_oldslot[1], nhru, 0, 64, h->stream));
  if (need > h->sort_tmp_bytes) {
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_sort_tmp);
    h->d_sort_tmp = nullptr;
    h->sort_tmp_bytes = 0;
    CK(cudaMalloc(&h->d_sort_tmp, need));
    h->sort_tmp_bytes = need;
  }
  CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp, need, h->d_keys[0], h->d_keys[1], h->d_oldslot[0], h->d_oldslot[1], nhru, 0, 64, h->stream));
  const size_t ns = (size_t)nhru * L.hr_stride, np = (size_t)nhru * HP_N;
  k_permute_state<<<(unsigned)((ns + T - 1) / T), T, 0, h->stream>>>(src_state, dst_state, nhru, L.hr_stride, h->d_oldslot[1]);
  k_permute_rows<<<(unsigned)((np + T - 1) / T), T, 0, h->stream>>>(S.hrupar, D.hrupar, nhru, HP_N, h->d_oldslot[1]);
  k_slot_maps<<<G, T, 0, h->stream>>>(h->d_keys[1], nhru, D.hru_of_slot, D.slot_of_hru);
  h->last_launches += 5 + 2;  // + the radix sort's own passes (counted as two)
  CK(cudaGetLastError());
  return VICGPU_OK;
}

// ---- balanced blocks (single-wave domains) -------------------------------------------------------------------------------------
// The warps of kind k run in blocks of kind_n[k] warps (16 = the register file's worth, the initial layout).  upload_block_layout
// turns kind_n into the table the kernel reads; balance_blocks updates kind_n from the measured warp durations of a launch with the
// current layout: a warp gets ~10 us faster per co-resident warp it loses (profiles/r02_summary.md, domain-size sweep), so every kind
// is moved towards the common duration T that makes the blocks fit the SMs the cell-output grid does not need.
static int upload_block_layout(vicgpu_handle* h) {
  const int nw = (int)h->warp_kind.size();
  h->h_block_w0.clear();
  for (int w = 0; w < nw;) {
    h->h_block_w0.push_back(w);
    w += std::max(1, std::min(h->kind_n[h->warp_kind[w]], VICGPU_HRU_BLOCK_MAX / 32));
  }
  h->h_block_w0.push_back(nw);
  h->nb_balanced = (int)h->h_block_w0.size() - 1;
  // (at most one block per warp: the table is allocated once per domain, in vicgpu_set_cells; the copy is ordered behind the launches
  // that still read the old table and h_block_w0 stays untouched until the next revision, a day of records later)
  if (!h->d_block_w0) CK(cudaMalloc(&h->d_block_w0, ((size_t)nw + 1) * sizeof(int)));
  CK(cudaMemcpyAsync(h->d_block_w0, h->h_block_w0.data(), h->h_block_w0.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  h->bal_measured = false;
  h->bal_measure_next = true;
  return VICGPU_OK;
}

static int balance_blocks(vicgpu_handle* h) {
  const int nw = (int)h->warp_kind.size(), nk = (int)h->kind_n.size();
  CK(cudaStreamSynchronize(h->stream));
  std::vector<unsigned long long> ns((size_t)2 * nw);
  CK(cudaMemcpy(ns.data(), h->d_warp_ns, ns.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<double> sum((size_t)nk, 0.0);
  std::vector<int> cnt((size_t)nk, 0), wk((size_t)nk, 0);
  for (int w = 0; w < nw; w++) {
    wk[h->warp_kind[w]]++;
    if (ns[2 * w + 1] > ns[2 * w] && ns[2 * w] != 0) {
      sum[h->warp_kind[w]] += (double)(ns[2 * w + 1] - ns[2 * w]) / 1e3;
      cnt[h->warp_kind[w]]++;
    }
  }
  const double c_us = 10.0;   // duration a warp gains per co-resident warp less
  const int nmax = VICGPU_HRU_BLOCK_MAX / 32, nmin = 6;
  const int nb_max = std::max(1, h->sm_count - h->bal_reserve);  // (the rest is left to the dependent cell-output grid)
  std::vector<double> t((size_t)nk, 0.0);
  double lo = 1e30, hi = 0;
  for (int k = 0; k < nk; k++) {
    if (!cnt[k]) continue;
    t[k] = sum[k] / cnt[k];
    lo = std::min(lo, t[k] - c_us * nmax);
    hi = std::max(hi, t[k] + c_us * nmax);
  }
  if (hi <= 0) return VICGPU_OK;  // nothing measured
  auto n_of = [&](int k, double T) {
    if (!cnt[k]) return h->kind_n[k];
    int n = (int)floor(h->kind_n[k] + (T - t[k]) / c_us);
    n = std::max(h->kind_n[k] - 2, std::min(h->kind_n[k] + 2, n));  // damped: at most two warps per update
    return std::max(nmin, std::min(nmax, n));
  };
  auto blocks = [&](double T) {
    int nb = 0;
    for (int k = 0; k < nk; k++) nb += (wk[k] + n_of(k, T) - 1) / n_of(k, T);
    return nb;
  };
  for (int it = 0; it < 40; it++) {  // smallest common duration whose blocks fit
    const double mid = 0.5 * (lo + hi);
    if (blocks(mid) <= nb_max) hi = mid; else lo = mid;
  }
  if (blocks(hi) > nb_max) return VICGPU_OK;  // (cannot happen: at T = hi every kind is at 16 warps, the initial layout)
  for (int k = 0; k < nk; k++) h->kind_n[k] = n_of(k, hi);
  return upload_block_layout(h);
}

static int create_on_device(vicgpu_handle* h, const vicgpu_options* opt, const Opts& o, int device);

extern "C" {

int vicgpu_abi_version(void) { return VICGPU_ABI_VERSION; }
const char* vicgpu_last_error(void) { return vicgpu_err.c_str(); }

int vicgpu_create(vicgpu_handle** out, const vicgpu_options* opt, int device) {
  if (!out || !opt) return fail(VICGPU_EINVAL, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return fail(VICGPU_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(VICGPU_ENODEV, "device ordinal out of range");
  Opts o;
  const char* why = "";
  int rc = opts_from_abi(*opt, o, &why);
  if (rc != VICGPU_OK) return fail(rc, why);
  CK(cudaSetDevice(device));
  vicgpu_handle* h = new vicgpu_handle();
  int crc = create_on_device(h, opt, o, device);
  if (crc != VICGPU_OK) {
    vicgpu_destroy(h);
    return crc;
  }
  *out = h;
  return VICGPU_OK;
}

}  // extern "C"

static int create_on_device(vicgpu_handle* h, const vicgpu_options* opt, const Opts& o, int device) {
  h->device = device;
  h->abi = *opt;
  h->o = o;
  h->nout = o.L.out_off[VICGPU_N_OUTVARS];
  memset(&h->t, 0, sizeof(h->t));
  // the HRU step owns the machine: its stream has the highest priority; host <-> device copies run on their own stream
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
  CK(cudaStreamCreateWithPriority(&h->stream_copy, cudaStreamNonBlocking, prio_lo));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  for (int k = 0; k < 2; k++) {
    CK(cudaEventCreateWithFlags(&h->ev_stage_full[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_stage_free[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->fwin[k].ready, cudaEventDisableTiming));
  }
  // tuning / A-B knobs (environment, read once per handle; every combination computes the same bits, tests/test_gpu.py)
  const char* wt = getenv("VICGPU_WARPTIME");  // per-warp timers in profiled launches (vicgpu_get_warp_times); they slow the kernel
  h->warp_timing = wt && atoi(wt) != 0;
  const char* noov = getenv("VICGPU_NOOVERLAP");  // 1: the cell output of record r finishes before step r + 1 starts (no dependent launch)
  h->pdl = !(noov && atoi(noov) != 0);
  const char* sl = getenv("VICGPU_SYNC");  // most clock cycles a warp waits for its block at a phase boundary of the step (0: no rendezvous)
  const char* sm = getenv("VICGPU_SYNCMASK");  // phase boundaries at which the warps of a block wait for each other (bit per phase; A/B knob)
  h->sync_limit = ((sl ? atoll(sl) : 1000000) & 0x00ffffffffffffffll) | ((long long)((sm ? atoi(sm) : 0x7f) & 0x7f) << 56);  // 0.5 ms: a safety bound, not a tuning parameter (waits end when the group has arrived)
  const char* nobin = getenv("VICGPU_NOBIN");  // keep the caller's row order (no binning at all)
  h->binned = !(nobin && atoi(nobin) != 0);
  const char* rb = getenv("VICGPU_REBIN");  // records between re-sorts of the rows by (kind, snow); 0: bin by kind once (set_cells)
  h->rebin_every = rb ? atoi(rb) : 24;
  h->rebin = h->binned && h->rebin_every > 0;
  const char* pw = getenv("VICGPU_PDLWAIT");  // 0: the dependent cell-output grid does not end with griddepcontrol.wait (A/B only)
  h->pdl_wait = !(pw && atoi(pw) == 0);
  // 1: round a step grid smaller than the machine up to one block per SM.  Off by default: the step kernel's duration is set by its
  // slowest warp, not by how many SMs share the warps, and the SMs a 118-block grid leaves idle are where the dependent cell-output
  // grid runs (measured at 10,000 cells: record 588 us with 118 blocks, 835 us = step + output back to back with 148)
  const char* ev = getenv("VICGPU_EVEN");
  h->even = ev && atoi(ev) != 0;
  // rows binned by the cost of their last step as well (k_bin_keys): 0 off, 1 inside a (kind, regime) bin, 2 cost first; -1 (default):
  // decided in vicgpu_set_cells from the configuration and the domain size
  const char* bc = getenv("VICGPU_BINCOST");
  h->bin_cost_env = bc ? atoi(bc) : -1;
  h->rebin_env = rb != nullptr;
  const char* lp = getenv("VICGPU_L2PERSIST");  // MB of L2 set aside for the cell-parameter table (persisting access-policy window), 0: off
  h->l2_persist_mb = lp ? atoi(lp) : 0;
#if defined(VIC_NO_AERO_TABLE)
  h->aero_cache = false;  // A/B build: the step kernel evaluates the aerodynamic geometry every record
#else
  h->aero_cache = true;
#endif
  const char* bl = getenv("VICGPU_BALANCE");  // SMs left to the cell-output grid by the balanced step grid; 0: blocks of equal warp counts (A/B)
  h->bal_reserve = bl ? atoi(bl) : 28;
  h->balance = h->bal_reserve > 0;
  const char* bf = getenv("VICGPU_BINFINE");  // 1: bin by pack regime and canopy snow as well (k_bin_keys)
  h->bin_fine = bf && atoi(bf) != 0;
  CK(cudaMalloc(&h->d_o, sizeof(Opts)));
  CK(cudaMemcpy(h->d_o, &h->o, sizeof(Opts), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_aggtype, VICGPU_N_OUTVARS * sizeof(int)));
  int agg[VICGPU_N_OUTVARS];
  vicgpu_default_aggtypes(agg);
  CK(cudaMemcpy(h->d_aggtype, agg, sizeof(agg), cudaMemcpyHostToDevice));
  // The step kernel keeps its working set in thread-local memory; the limit is per device, not per handle: it is raised, never
  // lowered (another handle or another library in the process may have asked for more), and stays raised after vicgpu_destroy.
  size_t stack_now = 0;
  CK(cudaDeviceGetLimit(&stack_now, cudaLimitStackSize));
  if (stack_now < 24 * 1024) CK(cudaDeviceSetLimit(cudaLimitStackSize, 24 * 1024));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  const char* blk = getenv("VICGPU_BLOCK");  // threads per block of the per-HRU step kernel (multiple of 32, <= VICGPU_HRU_BLOCK_MAX)
  if (blk && atoi(blk) >= 32 && atoi(blk) <= VICGPU_HRU_BLOCK_MAX && atoi(blk) % 32 == 0) {
    h->hru_block = atoi(blk);
    h->hru_block_fixed = true;
  }
  return VICGPU_OK;
}

extern "C" {

int vicgpu_destroy(vicgpu_handle* h) {
  if (!h) return VICGPU_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream_copy) cudaStreamSynchronize(h->stream_copy);
  cudaFree(h->d_aero);
  cudaFree(h->d_o); cudaFree(h->d_veglib); cudaFree(h->d_cellpar); cudaFree(h->d_cellder); cudaFree(h->d_gmb_cum); cudaFree(h->d_gmb);
  for (int b = 0; b < 2; b++) {
    free_half(h->half[b]);
    free_order(h->order[b]);
    cudaFree(h->d_ostage[b]);
    cudaFree(h->fwin[b].d);
    if (h->ev_stage_full[b]) cudaEventDestroy(h->ev_stage_full[b]);
    if (h->ev_stage_free[b]) cudaEventDestroy(h->ev_stage_free[b]);
    if (h->fwin[b].ready) cudaEventDestroy(h->fwin[b].ready);
  }
  cudaFree(h->d_fail_rec); cudaFree(h->d_keys[0]); cudaFree(h->d_keys[1]); cudaFree(h->d_oldslot[0]); cudaFree(h->d_oldslot[1]);
  cudaFree(h->d_sort_tmp);
  cudaFree(h->d_carry); cudaFree(h->d_out); cudaFree(h->d_agg); cudaFree(h->d_stage); cudaFree(h->d_fstage);
  cudaFree(h->d_cell_h0); cudaFree(h->d_status); cudaFree(h->d_aggtype);
  cudaFree(h->d_warp_ns); cudaFree(h->d_cost); cudaFree(h->d_work); cudaFree(h->d_block_w0);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (cudaEvent_t e : h->pev) cudaEventDestroy(e);
  if (h->stream_copy) cudaStreamDestroy(h->stream_copy);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VICGPU_OK;
}

int vicgpu_get_layout(const vicgpu_handle* h, vicgpu_layout* L) {
  if (!h || !L) return fail(VICGPU_EINVAL, "null argument");
  *L = h->o.L;
  return VICGPU_OK;
}

int vicgpu_set_veglib(vicgpu_handle* h, int nclass, const double* veglib) {
  if (!h || !veglib || nclass < h->o.NVegLibTypes + 4) return fail(VICGPU_EINVAL, "veglib must hold NVegLibTypes + 4 rows");
  if (h->have_cells && nclass < h->t.nclass) return fail(VICGPU_ESTATE, "a smaller vegetation library after set_cells: call set_cells again");
  CK(cudaSetDevice(h->device));
  cudaFree(h->d_veglib);
  const size_t n = (size_t)nclass * h->o.L.vl_stride;
  CK(cudaMalloc(&h->d_veglib, n * sizeof(double)));
  CK(cudaMemcpy(h->d_veglib, veglib, n * sizeof(double), cudaMemcpyHostToDevice));
  h->t.veglib = h->d_veglib;
  h->t.nclass = nclass;
  h->aero_month = -1;
  return VICGPU_OK;
}

int vicgpu_set_cells(vicgpu_handle* h, int ncell, const double* cellpar, int nhru, const double* hrupar) {
  if (!h || !cellpar || !hrupar || ncell <= 0 || nhru <= 0) return fail(VICGPU_EINVAL, "bad argument");
  if (!h->d_veglib) return fail(VICGPU_ESTATE, "set_veglib before set_cells (vegetation indices are checked against it)");
  if ((long long)ncell > 65535LL * 32) return fail(VICGPU_EUNSUPPORTED, "more than 2,097,120 cells per device (shard the domain)");
  CK(cudaSetDevice(h->device));
  const vicgpu_layout& L = h->o.L;
  // HRU -> cell map must be ascending (hruList order grouped by cell)
  std::vector<int> h0(ncell + 1, 0);
  int prev = 0;
  for (int k = 0; k < nhru; k++) {
    const double cd = hrupar[(size_t)k * HP_N + HP_cell];
    const int c = (int)cd;
    if (cd != (double)c || c < prev || c >= ncell) return fail(VICGPU_EINVAL, "hrupar[HP_cell] must be ascending cell indices in [0, ncell)");
    const int vi = (int)hrupar[(size_t)k * HP_N + HP_vegIndex], b = (int)hrupar[(size_t)k * HP_N + HP_band];
    if (vi < 0 || vi >= h->t.nclass) return fail(VICGPU_EINVAL, "hrupar[HP_vegIndex] out of range");
    if (b < 0 || b >= h->o.Nbands) return fail(VICGPU_EINVAL, "hrupar[HP_band] out of range");
    prev = c;
    h0[c + 1]++;
  }
  for (int c = 0; c < ncell; c++) h0[c + 1] += h0[c];
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->stream_copy));
  // initial row order: binned by kind (bin_hrus)
  std::vector<int> hru_of_slot, slot_of_hru;
  if (h->binned) bin_hrus(hrupar, nhru, hru_of_slot, slot_of_hru);
  // from here on the old domain is gone: a failure below (out of memory on a large domain) must leave a handle that refuses to step
  h->have_cells = h->have_state = false;
  h->fwin[0].nrec = h->fwin[1].nrec = 0;
  for (int k = 0; k < 2; k++) {
    cudaFree(h->d_ostage[k]);
    h->d_ostage[k] = nullptr;
  }
  {
    const double* vl = h->t.veglib;
    const int nc = h->t.nclass;
    memset(&h->t, 0, sizeof(h->t));
    h->t.veglib = vl;
    h->t.nclass = nc;
  }
  h->d_state_cur = nullptr;
  cudaFree(h->d_gmb_cum); cudaFree(h->d_gmb);
  h->d_gmb_cum = h->d_gmb = nullptr;
  cudaFree(h->d_aero); h->d_aero = nullptr;
  cudaFree(h->d_cellpar); cudaFree(h->d_cellder); cudaFree(h->d_carry); cudaFree(h->d_out); cudaFree(h->d_agg); cudaFree(h->d_cell_h0); cudaFree(h->d_status);
  h->d_cellder = nullptr;
  cudaFree(h->d_fail_rec); cudaFree(h->d_keys[0]); cudaFree(h->d_keys[1]); cudaFree(h->d_oldslot[0]); cudaFree(h->d_oldslot[1]); cudaFree(h->d_warp_ns);
  cudaFree(h->d_cost);
  h->d_cost = nullptr;
  h->d_cellpar = h->d_carry = h->d_out = h->d_agg = nullptr;
  h->d_cell_h0 = h->d_status = h->d_fail_rec = nullptr;
  h->d_keys[0] = h->d_keys[1] = nullptr;
  h->d_oldslot[0] = h->d_oldslot[1] = nullptr;
  h->d_warp_ns = nullptr;
  for (int b = 0; b < 2; b++) {
    free_half(h->half[b]);
    free_order(h->order[b]);
  }
  const size_t state_bytes = hr_rows(nhru) * L.hr_stride * sizeof(double);  // whole 32-row tiles
  CK(cudaMalloc(&h->d_cellpar, (size_t)ncell * L.cp_stride * sizeof(double)));
  if (h->aero_cache) CK(cudaMalloc(&h->d_aero, (size_t)nhru * VIC_AERO_NCOL * sizeof(double)));
  h->aero_month = -1;
  CK(cudaMalloc(&h->d_cellder, (size_t)ncell * VIC_NCELLDER * sizeof(double)));
  CK(cudaMalloc(&h->d_cost, (size_t)nhru * sizeof(int)));
  CK(cudaMemset(h->d_cost, 0, (size_t)nhru * sizeof(int)));
  {
    // per-thread cost counters of the step kernel (vic_frozen.cuh): one slot per thread of the largest grid the launcher makes
    cudaFree(h->d_work);
    h->d_work = nullptr;
    const size_t slots = ((size_t)nhru / 32 + (size_t)h->sm_count + 2) * VICGPU_HRU_BLOCK_MAX;
    CK(cudaMalloc(&h->d_work, slots * sizeof(int)));
    CK(cudaMemset(h->d_work, 0, slots * sizeof(int)));
    const int rcw = vic_node_width(h->o) == 3 ? vicgpu_set_work_buffer_nn3(h->d_work) : vic_node_width(h->o) == 10 ? vicgpu_set_work_buffer_nn10(h->d_work) : vicgpu_set_work_buffer_nn32(h->d_work);
    if (rcw != 0) return fail(VICGPU_ECUDA, "cudaMemcpyToSymbol(vic_work_buf)");
  }
  CK(cudaMalloc(&h->d_gmb_cum, (size_t)nhru * sizeof(double)));
  CK(cudaMemset(h->d_gmb_cum, 0, (size_t)nhru * sizeof(double)));
  CK(cudaMalloc(&h->d_gmb, (size_t)ncell * 4 * sizeof(double)));
  {
    std::vector<double> g0((size_t)ncell * 4, 0.0);
    for (int c = 0; c < ncell; c++) g0[(size_t)3 * ncell + c] = -1;  // GraphingEquation(): fitError -1
    CK(cudaMemcpy(h->d_gmb, g0.data(), g0.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  for (int b = 0; b < 2; b++) {
    StateHalf& s = h->half[b];
    CK(cudaMalloc(&s.in, state_bytes));
    CK(cudaMalloc(&s.snap, state_bytes));
    CK(cudaMemset(s.in, 0, state_bytes));
    CK(cudaMemset(s.snap, 0, state_bytes));
    CK(cudaMalloc(&s.hdiag, (size_t)nhru * 3 * sizeof(double)));
    CK(cudaMemset(s.hdiag, 0, (size_t)nhru * 3 * sizeof(double)));
    RowOrder& r = h->order[b];
    if (b == 0 || h->rebin) {
      CK(cudaMalloc(&r.hrupar, (size_t)nhru * HP_N * sizeof(double)));
      if (h->binned) {
        CK(cudaMalloc(&r.slot_of_hru, (size_t)nhru * sizeof(int)));
        CK(cudaMalloc(&r.hru_of_slot, (size_t)nhru * sizeof(int)));
      }
    }
  }
  if (h->binned) {
    CK(cudaMemcpy(h->order[0].slot_of_hru, slot_of_hru.data(), (size_t)nhru * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->order[0].hru_of_slot, hru_of_slot.data(), (size_t)nhru * sizeof(int), cudaMemcpyHostToDevice));
  }
  h->recs_since_rebin = 1 << 30;
  if (h->rebin) {
    for (int b = 0; b < 2; b++) {
      CK(cudaMalloc(&h->d_keys[b], (size_t)nhru * sizeof(unsigned long long)));
      CK(cudaMalloc(&h->d_oldslot[b], (size_t)nhru * sizeof(int)));
    }
  }
  CK(cudaMalloc(&h->d_carry, (size_t)ncell * CC_N * sizeof(double)));
  CK(cudaMalloc(&h->d_out, (size_t)ncell * h->nout * sizeof(double)));
  CK(cudaMalloc(&h->d_agg, (size_t)ncell * h->nout * sizeof(double)));
  CK(cudaMalloc(&h->d_cell_h0, (size_t)(ncell + 1) * sizeof(int)));
  CK(cudaMalloc(&h->d_status, (size_t)ncell * sizeof(int)));
  CK(cudaMalloc(&h->d_fail_rec, (size_t)ncell * sizeof(int)));
  {
    std::vector<int> never((size_t)ncell, INT_MAX);
    CK(cudaMemcpy(h->d_fail_rec, never.data(), (size_t)ncell * sizeof(int), cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(h->d_carry, 0, (size_t)ncell * CC_N * sizeof(double)));
  CK(cudaMemset(h->d_out, 0, (size_t)ncell * h->nout * sizeof(double)));
  CK(cudaMemset(h->d_agg, 0, (size_t)ncell * h->nout * sizeof(double)));
  CK(cudaMemset(h->d_status, 0, (size_t)ncell * sizeof(int)));
  CK(cudaMemcpy(h->d_cell_h0, h0.data(), (size_t)(ncell + 1) * sizeof(int), cudaMemcpyHostToDevice));
  int rc = upload_transposed(h, cellpar, h->d_cellpar, ncell, L.cp_stride);
  if (rc) return rc;
  k_derive_cells<<<(ncell + 127) / 128, 128, 0, h->stream>>>(h->d_o, h->d_cellpar, h->d_cellder, ncell);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  rc = upload_transposed(h, hrupar, h->order[0].hrupar, nhru, HP_N, h->order[0].hru_of_slot);
  if (rc) return rc;
  // Cost binning pays only when the domain runs in many waves (the step kernel of a one-wave domain ends when its slowest HRU does,
  // however the rows are packed: measured 60 -> 74 ms per record at 10,000 frozen-soil cells) and only where the cost varies by orders
  // of magnitude between HRUs: the soil-thermal-profile configurations, 100,000 cells 0.287 -> 0.353 M cell-timesteps/s with a re-sort
  // every other record.  For the quick-flux configurations it is neutral (profiles/r02_summary.md).
  h->bin_cost = h->bin_cost_env >= 0 ? h->bin_cost_env : ((!h->o.QUICK_FLUX && (long long)nhru > 8LL * h->sm_count * VICGPU_HRU_BLOCK_MAX) ? 2 : 0);
  if (h->bin_cost == 2 && !h->rebin_env) h->rebin_every = 2;
  if (!h->hru_block_fixed) h->hru_block = ((long long)nhru <= (long long)h->sm_count * VICGPU_HRU_BLOCK) ? VICGPU_HRU_BLOCK : VICGPU_HRU_BLOCK_MAX;
  // balanced blocks: rows sorted by kind first (so that a warp's kind never changes), re-sorted regularly (the layout is revised then),
  // the whole domain resident at once with SMs to spare, default block size, no per-warp profiling requested
  h->bal_active = h->balance && h->binned && h->rebin && h->bin_cost != 2 && !h->warp_timing && !h->even && h->hru_block == VICGPU_HRU_BLOCK_MAX &&
                  (long long)nhru <= (long long)(h->sm_count - h->bal_reserve) * VICGPU_HRU_BLOCK_MAX * 15 / 16;
  h->warp_kind.clear();
  h->kind_n.clear();
  cudaFree(h->d_block_w0);
  h->d_block_w0 = nullptr;
  if (h->bal_active) {
    const int nw = (nhru + 31) / 32;
    std::vector<long long> keys;
    h->warp_kind.resize((size_t)nw);
    for (int w = 0; w < nw; w++) {
      const double* pr = hrupar + (size_t)hru_of_slot[(size_t)w * 32] * HP_N;
      long long kind = (long long)pr[HP_vegIndex];
      if (pr[HP_isArtBare] != 0.0) kind += 1000000;
      if (pr[HP_isGlacier] != 0.0) kind += 2000000;
      size_t k = 0;
      while (k < keys.size() && keys[k] != kind) k++;
      if (k == keys.size()) keys.push_back(kind);
      h->warp_kind[(size_t)w] = (int)k;
    }
    h->kind_n.assign(keys.size(), VICGPU_HRU_BLOCK_MAX / 32);
    cudaFree(h->d_warp_ns);
    h->d_warp_ns = nullptr;
    CK(cudaMalloc(&h->d_warp_ns, 2 * (size_t)nw * sizeof(unsigned long long)));
    int rcb = upload_block_layout(h);
    if (rcb) return rcb;
  }
  h->t.ncell = ncell; h->t.nhru = nhru;
  h->t.gmb_cum = h->d_gmb_cum; h->t.gmb = h->d_gmb; h->t.cost = h->d_cost; h->t.aero = nullptr;
  h->t.cellpar = h->d_cellpar; h->t.cellder = h->d_cellder; h->t.cell_h0 = h->d_cell_h0; h->t.fail_rec = h->d_fail_rec;
  h->t.status = h->d_status; h->t.carry = h->d_carry; h->t.out = h->d_out; h->t.agg = h->d_agg; h->t.aggtype = h->d_aggtype;
  if (h->l2_persist_mb > 0) {
    // The cell-parameter table is read column by column all through the step (227 columns, 18 MB at 10,000 cells) while 300+ MB of
    // state and stack stream through L2 per record: a persisting window keeps it resident (measurement knob, off by default).
    int maxwin = 0;
    CK(cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, h->device));
    const size_t want = (size_t)h->l2_persist_mb << 20, bytes = (size_t)ncell * L.cp_stride * sizeof(double);
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    cudaStreamAttrValue av = {};
    av.accessPolicyWindow.base_ptr = h->d_cellpar;
    av.accessPolicyWindow.num_bytes = bytes < (size_t)maxwin ? bytes : (size_t)maxwin;
    av.accessPolicyWindow.hitRatio = (float)(want >= av.accessPolicyWindow.num_bytes ? 1.0 : (double)want / (double)av.accessPolicyWindow.num_bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av));
  }
  h->cur_half = 0;
  h->d_state_cur = h->half[0].in;
  h->have_cells = true;
  h->have_state = false;
  h->step_count = 0;
  h->glac_started = false;
  return VICGPU_OK;
}

int vicgpu_set_output_spec(vicgpu_handle* h, const int* aggtype) {
  if (!h) return fail(VICGPU_EINVAL, "null handle");
  CK(cudaSetDevice(h->device));
  int agg[VICGPU_N_OUTVARS];
  if (aggtype) memcpy(agg, aggtype, sizeof(agg));
  else vicgpu_default_aggtypes(agg);
  for (int v = 0; v < VICGPU_N_OUTVARS; v++)
    if (agg[v] != VICGPU_AGG_AVG && agg[v] != VICGPU_AGG_END && agg[v] != VICGPU_AGG_SUM)
      return fail(VICGPU_EUNSUPPORTED, "only AVG / END / SUM aggregation is implemented (as in put_data.c:664-680)");
  CK(cudaMemcpy(h->d_aggtype, agg, sizeof(agg), cudaMemcpyHostToDevice));
  return VICGPU_OK;
}

int vicgpu_set_cell_status(vicgpu_handle* h, const int* status) {
  if (!h || !status || !h->have_cells) return fail(VICGPU_ESTATE, "set_cells first");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(h->d_status, status, (size_t)h->t.ncell * sizeof(int), cudaMemcpyHostToDevice));
  std::vector<int> fr((size_t)h->t.ncell);
  for (int c = 0; c < h->t.ncell; c++) fr[c] = status[c] != 0 ? -1 : INT_MAX;
  CK(cudaMemcpy(h->d_fail_rec, fr.data(), fr.size() * sizeof(int), cudaMemcpyHostToDevice));
  return VICGPU_OK;
}

int vicgpu_set_state(vicgpu_handle* h, const double* hrurec) {
  if (!h || !hrurec) return fail(VICGPU_EINVAL, "null argument");
  if (!h->have_cells) return fail(VICGPU_ESTATE, "set_cells before set_state");
  h->aero_month = -1;  // the rows return to the order of the first half
  CK(cudaSetDevice(h->device));
  StateHalf& s = h->half[h->cur_half];
  {
    const int rows = h->t.nhru, cols = h->o.L.hr_stride;
    const size_t n = (size_t)rows * cols;
    int rc = ensure_stage(h, n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_stage, hrurec, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_scatter_state<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_stage, s.in, rows, cols, h->order[s.ord].hru_of_slot);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  h->d_state_cur = s.in;
  h->have_state = true;
  h->recs_since_rebin = 1 << 30;
  return VICGPU_OK;
}

int vicgpu_get_state(vicgpu_handle* h, double* hrurec) {
  if (!h || !hrurec) return fail(VICGPU_EINVAL, "null argument");
  if (!h->have_state) return fail(VICGPU_ESTATE, "no state set");
  CK(cudaSetDevice(h->device));
  const int rows = h->t.nhru, cols = h->o.L.hr_stride;
  const size_t n = (size_t)rows * cols;
  int rc = ensure_stage(h, n);
  if (rc) return rc;
  k_gather_state<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_state_cur, h->d_stage, rows, cols, h->order[h->half[h->cur_half].ord].slot_of_hru);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(hrurec, h->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return VICGPU_OK;
}

// The device keeps TWO forcing windows.  vicgpu_set_forcing fills the one that was filled less recently and returns as soon as the
// copy is queued on the copy stream (the host buffer must stay untouched until the next call into the library returns; pinned host
// memory makes the copy truly asynchronous), so that the upload of the next block of records overlaps the step over the current one:
//     set_forcing(block b + 1);  step(block b);  set_forcing(block b + 2);  step(block b + 1); ...
// vicgpu_step waits (on the device) for the window that holds its records.
int vicgpu_set_forcing(vicgpu_handle* h, int rec0, int nrec, const double* forcing) {
  if (!h || !forcing || nrec <= 0 || rec0 < 0) return fail(VICGPU_EINVAL, "bad argument");
  if (!h->have_cells) return fail(VICGPU_ESTATE, "set_cells before set_forcing");
  CK(cudaSetDevice(h->device));
  const size_t per = (size_t)h->t.ncell * h->o.L.f_stride;
  ForcingWindow& w = h->fwin[h->fwin_next];
  h->fwin_next ^= 1;
  w.nrec = 0;
  { int rcf = vicgpu_ensure_window(h, w, per * nrec); if (rcf) return rcf; }
  // staged in chunks of at most 256 MiB so that the staging buffer stays small next to the window
  const size_t chunk_recs = std::min<size_t>(65535, std::max<size_t>(1, (size_t)(256u << 20) / (per * sizeof(double))));
  const size_t stage_need = per * std::min<size_t>(chunk_recs, (size_t)nrec);
  if (stage_need > h->fstage_cap) {
    CK(cudaStreamSynchronize(h->stream_copy));
    cudaFree(h->d_fstage);
    h->d_fstage = nullptr;
    h->fstage_cap = 0;
    CK(cudaMalloc(&h->d_fstage, stage_need * sizeof(double)));
    h->fstage_cap = stage_need;
  }
  for (size_t r = 0; r < (size_t)nrec; r += chunk_recs) {
    const int nr = (int)std::min<size_t>(chunk_recs, (size_t)nrec - r);
    CK(cudaMemcpyAsync(h->d_fstage, forcing + r * per, per * nr * sizeof(double), cudaMemcpyHostToDevice, h->stream_copy));
    int rct = vicgpu_transpose(h, h->d_fstage, w.d + r * per, h->t.ncell, h->o.L.f_stride, nr, h->stream_copy);
    if (rct) return rct;
  }
  CK(cudaEventRecord(w.ready, h->stream_copy));
  w.rec0 = rec0;
  w.nrec = nrec;
  return VICGPU_OK;
}

}  // extern "C"

// One set of output rows [nout][ncell] on the device -> host [ncell][nout] (double, or float32 as the NetCDF writer narrows it), through
// one of two staging buffers: the transpose runs in the step's stream (the next cell-output launch overwrites the rows), the copy on
// the copy stream, and the staging buffer is handed back with an event, so that the copy of record r overlaps the kernels of record r + 1.
static int stage_out(vicgpu_handle* h, const double* d_rows, void* host, bool f32) {
  const int k = h->stage_idx;
  h->stage_idx ^= 1;
  const size_t rowsz = (size_t)h->t.ncell * h->nout;
  CK(cudaStreamWaitEvent(h->stream, h->ev_stage_free[k], 0));
  if (f32) {
    dim3 b(32, 8), g((h->t.ncell + 31) / 32, (h->nout + 31) / 32, 1);
    k_transpose_f32<<<g, b, 0, h->stream>>>(d_rows, (float*)h->d_ostage[k], h->nout, h->t.ncell);
    h->last_launches++;
    CK(cudaGetLastError());
  } else {
    int rc = vicgpu_transpose(h, d_rows, h->d_ostage[k], h->nout, h->t.ncell, 1, h->stream);
    if (rc) return rc;
  }
  CK(cudaEventRecord(h->ev_stage_full[k], h->stream));
  CK(cudaStreamWaitEvent(h->stream_copy, h->ev_stage_full[k], 0));
  CK(cudaMemcpyAsync(host, h->d_ostage[k], rowsz * (f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, h->stream_copy));
  CK(cudaEventRecord(h->ev_stage_free[k], h->stream_copy));
  return VICGPU_OK;
}

// ---- one record per launch, cell output as a programmatic dependent launch ---------------------------------------------------------
// Everything runs in ONE stream: step(r), then output(r-1) launched with programmatic stream serialization.  The output kernel does
// not wait for step(r) to finish (it reads the state step(r-1) wrote), only for every block of step(r) to be resident
// (griddepcontrol.launch_dependents at the top of the step kernel); its one-warp blocks then fit beside the step blocks.  Launched
// from a second stream instead, the output blocks reach the SMs first and the step blocks, which need a whole SM, wait for them:
// measured 0.3-0.4 ms per record.  step(r+1), an ordinary launch, starts when both have finished (k_cell_output ends with
// griddepcontrol.wait), which also protects the state buffer output(r-1) reads.
static int step_impl(vicgpu_handle* h, int rec0, int nrec, const int* dmy, void* out_data, void* out_agg, bool f32) {
  if (!h || !dmy || nrec <= 0) return fail(VICGPU_EINVAL, "bad argument");
  if (!h->have_cells || !h->have_state || !h->d_veglib) return fail(VICGPU_ESTATE, "set_veglib, set_cells and set_state before step");
  const ForcingWindow* fw = nullptr;
  for (int k = 0; k < 2; k++)
    if (h->fwin[k].nrec > 0 && rec0 >= h->fwin[k].rec0 && rec0 + nrec <= h->fwin[k].rec0 + h->fwin[k].nrec) fw = &h->fwin[k];
  if (!fw) return fail(VICGPU_ESTATE, "records outside the resident forcing windows");
  CK(cudaSetDevice(h->device));
  const vicgpu_layout& L = h->o.L;
  const int nhru = h->t.nhru;
  const size_t per = (size_t)h->t.ncell * L.f_stride;
  const size_t rowsz = (size_t)h->t.ncell * h->nout;
  const size_t esz = f32 ? sizeof(float) : sizeof(double);
  // cell output: three warps (the three variable groups) per 32 cells
  const int B = 96;
  const int cgrid = (h->t.ncell + 31) / 32;
  const bool one = h->o.NF == 1;
  h->last_launches = 0;
  int nagg = 0;
  if (out_data || out_agg) {
    for (int k = 0; k < 2; k++)
      if (!h->d_ostage[k]) CK(cudaMalloc(&h->d_ostage[k], rowsz * sizeof(double)));
  }
  if (h->profiling) {
    while ((int)h->pev.size() < 2 * nrec) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->pev.push_back(e);
    }
  }
  CK(cudaStreamWaitEvent(h->stream, fw->ready, 0));
  struct Pending { bool valid; Tables t; const double* frec; int rec, step_count, idx; } pend = {false, h->t, nullptr, 0, 0, 0};
  auto launch_output = [&](const Pending& p, bool dependent) -> int {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cgrid);
    cfg.blockDim = dim3(B);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dependent ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_cell_output, h->o, p.t, p.frec, p.rec, p.step_count, h->pdl_wait ? 1 : 0));
    h->last_launches++;
    if (out_data) {
      int rc = stage_out(h, h->d_out, (char*)out_data + (size_t)p.idx * rowsz * esz, f32);
      if (rc) return rc;
    }
    if (p.step_count == h->o.out_step_ratio) {
      if (out_agg) {
        int rc = stage_out(h, h->d_agg, (char*)out_agg + (size_t)nagg * rowsz * esz, f32);
        if (rc) return rc;
      }
      nagg++;
      CK(cudaMemsetAsync(h->d_agg, 0, rowsz * sizeof(double), h->stream));
    }
    return VICGPU_OK;
  };
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < nrec; i++) {
    const int rec = rec0 + i;
    StateHalf& S = h->half[h->cur_half];
    StateHalf& D = h->half[h->cur_half ^ 1];
    const double* input = h->d_state_cur;
    if (h->rebin && h->recs_since_rebin >= h->rebin_every) {
      // the pending output still needs the old row order and the state it was launched for: run it now, undeferred
      if (pend.valid) {
        int rc = launch_output(pend, false);
        if (rc) return rc;
        pend.valid = false;
      }
      if (h->bal_active && h->bal_measured) {  // the layout is revised when the rows are re-sorted: once a day
        int rcb = balance_blocks(h);
        if (rcb) return rcb;
      }
      int rc = rebin_rows(h, h->order[S.ord], h->order[S.ord ^ 1], h->d_state_cur, D.in);
      if (rc) return rc;
      D.ord = S.ord ^ 1;
      input = D.in;
      h->recs_since_rebin = 0;
      h->aero_month = -1;  // the table is in row order
    } else {
      D.ord = S.ord;
    }
    h->recs_since_rebin++;
    const RowOrder& R = h->order[D.ord];
    Tables t = h->t;
    t.hrupar = R.hrupar;
    t.slot_of_hru = R.slot_of_hru;
    t.hrurec = input;
    t.hrurec_out = D.snap;
    t.hdiag_out = D.hdiag;
    const int* d = &dmy[i * 5];
    const Dmy dm = {d[0], d[1], d[2], d[3], d[4]};
    if (h->d_aero) {
      if (h->aero_month != dm.month) {
        k_hru_aero<<<(nhru + 127) / 128, 128, 0, h->stream>>>(h->d_o, t, h->d_aero, dm.month - 1);
        h->last_launches++;
        h->aero_month = dm.month;
      }
      t.aero = h->d_aero;
    } else t.aero = nullptr;
    const GlacAccum ga = glacier_accum_flags(h->o, d, d + 5, rec, &h->glac_started);
    if (rec == 0) {
      // storage terms of the initial state: put_data(rec = -nrecs), vicNl.c:524-541
      Tables t0 = t;
      t0.hrurec_out = const_cast<double*>(input);
      k_cell_output<<<cgrid, B, 0, h->stream>>>(h->o, t0, nullptr, -1, h->step_count + 1, 0);
      h->last_launches++;
    }
    const double* frec = fw->d + (size_t)(rec - fw->rec0) * per;
    unsigned long long* wns = nullptr;
    if (h->profiling && h->warp_timing) {
      const size_t nw = ((size_t)nhru + 31) / 32;
      if (!h->d_warp_ns) CK(cudaMalloc(&h->d_warp_ns, 2 * nw * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(h->d_warp_ns, 0, 2 * nw * sizeof(unsigned long long), h->stream));
      wns = h->d_warp_ns;
    }
    const bool bal_time = h->bal_active && h->bal_measure_next && !wns;
    if (bal_time) {
      CK(cudaMemsetAsync(h->d_warp_ns, 0, 2 * (((size_t)nhru + 31) / 32) * sizeof(unsigned long long), h->stream));
      wns = h->d_warp_ns;
      h->bal_measure_next = false;
      h->bal_measured = true;
    }
    const int* bw0 = h->bal_active ? h->d_block_w0 : nullptr;
    const int nbw0 = h->nb_balanced;
    if (h->profiling) CK(cudaEventRecord(h->pev[2 * i], h->stream));
    if (vic_node_width(h->o) == 3) vicgpu_launch_hru_step_nn3(h->d_o, one, t, frec, dm, rec, ga, h->hru_block, h->stream, wns, h->sync_limit, h->even ? h->sm_count : 0, L.f_stride, bw0, nbw0);
    else if (vic_node_width(h->o) == 10) vicgpu_launch_hru_step_nn10(h->d_o, one, t, frec, dm, rec, ga, h->hru_block, h->stream, wns, h->sync_limit, h->even ? h->sm_count : 0, L.f_stride, bw0, nbw0);
    else vicgpu_launch_hru_step_nn32(h->d_o, one, t, frec, dm, rec, ga, h->hru_block, h->stream, wns, h->sync_limit, h->even ? h->sm_count : 0, L.f_stride, bw0, nbw0);
    h->last_launches++;
    // the previous record's output rides on this step
    if (pend.valid) {
      int rc = launch_output(pend, h->pdl);
      if (rc) return rc;
    }
    if (ga.enabled && ga.reset_after) {  // end of a glacier accumulation interval: the cells' mass-balance curves
      k_cell_gmb<<<(h->t.ncell + 127) / 128, 128, 0, h->stream>>>(h->d_o, t);
      h->last_launches++;
    }
    if (h->profiling) CK(cudaEventRecord(h->pev[2 * i + 1], h->stream));  // step(r) and the output riding on it
    h->step_count++;
    pend.valid = true;
    pend.t = t;
    pend.frec = frec;
    pend.rec = rec;
    pend.step_count = h->step_count;
    pend.idx = i;
    if (h->step_count == h->o.out_step_ratio) h->step_count = 0;
    h->d_state_cur = D.snap;
    h->cur_half ^= 1;
  }
  if (pend.valid) {
    int rc = launch_output(pend, false);
    if (rc) return rc;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->stream_copy));
  float msp = 0;
  CK(cudaEventElapsedTime(&msp, h->ev0, h->ev1));
  h->last_ms = msp;
  if (h->profiling) {
    for (int i = 0; i < nrec; i++) {
      float k = 0;
      CK(cudaEventElapsedTime(&k, h->pev[2 * i], h->pev[2 * i + 1]));
      h->prof_hru_ms += k;
      h->prof_hru_launches++;
    }
  }
  return VICGPU_OK;
}

extern "C" {

int vicgpu_step(vicgpu_handle* h, int rec0, int nrec, const int* dmy, double* out_data, double* out_agg) {
  return step_impl(h, rec0, nrec, dmy, out_data, out_agg, false);
}
int vicgpu_step_f32(vicgpu_handle* h, int rec0, int nrec, const int* dmy, float* out_data, float* out_agg) {
  return step_impl(h, rec0, nrec, dmy, out_data, out_agg, true);
}

int vicgpu_get_cell_status(vicgpu_handle* h, int* status) {
  if (!h || !status || !h->have_cells) return fail(VICGPU_ESTATE, "set_cells first");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(status, h->d_status, (size_t)h->t.ncell * sizeof(int), cudaMemcpyDeviceToHost));
  return VICGPU_OK;
}

int vicgpu_get_balance_errors(vicgpu_handle* h, double* err) {
  if (!h || !err || !h->have_cells) return fail(VICGPU_ESTATE, "set_cells first");
  CK(cudaSetDevice(h->device));
  std::vector<double> c((size_t)CC_N * h->t.ncell);
  CK(cudaMemcpy(c.data(), h->d_carry, c.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->t.ncell; i++)
    for (int k = 0; k < 5; k++) err[(size_t)i * 5 + k] = c[(size_t)(CC_water_last_storage + k) * h->t.ncell + i];
  return VICGPU_OK;
}

int vicgpu_set_profiling(vicgpu_handle* h, int on) {
  if (!h) return fail(VICGPU_EINVAL, "null handle");
  h->profiling = on != 0;
  h->prof_hru_ms = 0;
  h->prof_hru_launches = 0;
  return VICGPU_OK;
}

int vicgpu_get_kernel_profile(vicgpu_handle* h, double* hru_step_ms_total, long long* hru_step_launches) {
  if (!h) return fail(VICGPU_EINVAL, "null handle");
  if (hru_step_ms_total) *hru_step_ms_total = h->prof_hru_ms;
  if (hru_step_launches) *hru_step_launches = h->prof_hru_launches;
  return VICGPU_OK;
}

int vicgpu_get_warp_times(vicgpu_handle* h, double* times, double* kind, int max_warps) {
  if (!h || !times || !h->have_cells) return fail(VICGPU_EINVAL, "bad argument");
  if (!h->d_warp_ns) return fail(VICGPU_ESTATE, "no profiled launch yet (vicgpu_set_profiling)");
  CK(cudaSetDevice(h->device));
  const int nw = (h->t.nhru + 31) / 32;
  std::vector<unsigned long long> ns((size_t)2 * nw);
  CK(cudaMemcpy(ns.data(), h->d_warp_ns, ns.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<double> hp((size_t)h->t.nhru * HP_N);
  CK(cudaMemcpy(hp.data(), h->order[h->half[h->cur_half].ord].hrupar, hp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  for (int w = 0; w < nw; w++) if (ns[2 * w] && ns[2 * w] < t0) t0 = ns[2 * w];
  const size_t n = (size_t)h->t.nhru;
  for (int w = 0; w < nw && w < max_warps; w++) {
    times[2 * w] = (double)(ns[2 * w] - t0);
    times[2 * w + 1] = (double)(ns[2 * w + 1] - t0);
    if (kind) {
      const size_t s = (size_t)w * 32;
      kind[w] = hp[(size_t)HP_vegIndex * n + s] + (hp[(size_t)HP_isArtBare * n + s] != 0.0 ? 1e6 : 0) + (hp[(size_t)HP_isGlacier * n + s] != 0.0 ? 2e6 : 0);
    }
  }
  return nw;
}

int vicgpu_get_glacier_fit(vicgpu_handle* h, double* gmb) {
  if (!h || !gmb || !h->have_cells) return fail(VICGPU_ESTATE, "set_cells first");
  CK(cudaSetDevice(h->device));
  std::vector<double> g((size_t)4 * h->t.ncell);
  CK(cudaMemcpy(g.data(), h->d_gmb, g.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int c = 0; c < h->t.ncell; c++)
    for (int k = 0; k < 4; k++) gmb[(size_t)c * 4 + k] = g[(size_t)k * h->t.ncell + c];
  return VICGPU_OK;
}

int vicgpu_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(VICGPU_EINVAL, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(VICGPU_ENODEV, "no such CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CK(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 1 << 16, threads = 256, blocks = prop.multiProcessorCount * 8;
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(d, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;  // the first launch warms up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return VICGPU_OK;
}

int vicgpu_measure_phase_tax(vicgpu_handle* h, int nframe, int reps, double* us_per_pass) {
  if (!h || !us_per_pass || nframe < 0 || reps < 1) return fail(VICGPU_EINVAL, "bad argument");
  if (!h->have_state) return fail(VICGPU_ESTATE, "no state set");
  CK(cudaSetDevice(h->device));
  double* frame = nullptr;
  CK(cudaMalloc(&frame, std::max<size_t>(1, (size_t)nframe * h->t.nhru) * sizeof(double)));
  CK(cudaMemsetAsync(frame, 0, std::max<size_t>(1, (size_t)nframe * h->t.nhru) * sizeof(double), h->stream));
  Tables t = h->t;
  t.hrupar = h->order[h->half[h->cur_half].ord].hrupar;
  t.hrurec = h->d_state_cur;
  t.hrurec_out = h->half[h->cur_half ^ 1].in;  // scratch between steps: a re-sort overwrites it completely before anything reads it
  auto launch = [&]() {
    if (vic_node_width(h->o) == 3) vicgpu_launch_hru_pass_nn3(h->d_o, t, frame, nframe, h->hru_block, h->stream);
    else if (vic_node_width(h->o) == 10) vicgpu_launch_hru_pass_nn10(h->d_o, t, frame, nframe, h->hru_block, h->stream);
    else vicgpu_launch_hru_pass_nn32(h->d_o, t, frame, nframe, h->hru_block, h->stream);
  };
  launch();  // warm-up
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; r++) launch();
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  cudaFree(frame);
  *us_per_pass = (double)ms * 1e3 / reps;
  return VICGPU_OK;
}

int vicgpu_get_last_step_timing(vicgpu_handle* h, double* kernel_ms, long long* launches) {
  if (!h) return fail(VICGPU_EINVAL, "null handle");
  if (kernel_ms) *kernel_ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return VICGPU_OK;
}

}  // extern "C"

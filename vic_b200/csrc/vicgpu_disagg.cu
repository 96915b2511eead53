// vicgpu_disagg.cu -- forcing disaggregation on the device (vicgpu_disagg): one kernel per stage of
// vic_disagg.cuh, cells fastest in every grid so that scratch and forcing accesses are coalesced.
#include <algorithm>
#include "vicgpu_internal.h"
#include "vic_disagg.cuh"

using namespace vic;

namespace {

struct DisArgs {
  const Opts* o;          // device copy (layout)
  const double* cellpar;  // [cp_stride][ncell]
  const double* daily;    // [Ndays*4][ncell]
  double* forcing;        // [nrecs][f_stride][ncell]
  DisaggOpts d;
  DisaggScratch s;
  int nchunk;
};

__device__ __forceinline__ CellPar cellpar_of(const DisArgs& a, int cell) { return CellPar{Col{a.cellpar + cell, a.s.ntotal}, &a.o->L}; }

// stage kernels: tid -> (item, cell), cell fastest
__global__ void __launch_bounds__(128) k_dis_solar(DisArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)a.nchunk * 365) return;
  const int cell = a.s.cell0 + (int)(tid % a.nchunk), item = (int)(tid / a.nchunk);
  disagg_solar(cellpar_of(a, cell), a.d, a.s, cell, item);
}
__global__ void __launch_bounds__(128) k_dis_daily(DisArgs a) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= a.nchunk) return;
  const int cell = a.s.cell0 + tid;
  disagg_daily(cellpar_of(a, cell), a.d, a.s, a.daily, cell);
}
template <int STAGE>
__global__ void __launch_bounds__(128) k_dis_items(DisArgs a, int nitem) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)a.nchunk * nitem) return;
  const int cell = a.s.cell0 + (int)(tid % a.nchunk), item = (int)(tid / a.nchunk);
  const CellPar cp = cellpar_of(a, cell);
  if (STAGE == 0) disagg_day_radiation(cp, a.d, a.s, cell, item);
  else if (STAGE == 1) disagg_day_maxmin(cp, a.d, a.s, cell, item);
  else if (STAGE == 2) disagg_knot_coeff(cp, a.d, a.s, cell, item);
  else if (STAGE == 3) disagg_day_hourly(cp, a.d, a.s, cell, item);
  else disagg_record(cp, a.d, a.s, a.daily, a.forcing + (size_t)item * a.o->L.f_stride * a.s.ntotal, cell, item);
}

inline unsigned blocks_for(size_t n) { return (unsigned)((n + 127) / 128); }

}  // namespace

static int disagg_impl(vicgpu_handle* h, const vicgpu_disagg_options* dopt, const double* daily, double* forcing_out, bool time_major);
extern "C" int vicgpu_disagg(vicgpu_handle* h, const vicgpu_disagg_options* dopt, const double* daily, double* forcing_out) {
  return disagg_impl(h, dopt, daily, forcing_out, false);
}
extern "C" int vicgpu_disagg_tm(vicgpu_handle* h, const vicgpu_disagg_options* dopt, const double* daily_tm, double* forcing_out) {
  return disagg_impl(h, dopt, daily_tm, forcing_out, true);
}
static int disagg_impl(vicgpu_handle* h, const vicgpu_disagg_options* dopt, const double* daily, double* forcing_out, bool time_major) {
  if (!h || !dopt || !daily) return vicgpu_fail(VICGPU_EINVAL, "null argument");
  if (dopt->abi_version != VICGPU_ABI_VERSION) return vicgpu_fail(VICGPU_EINVAL, "abi_version mismatch");
  if (!h->have_cells) return vicgpu_fail(VICGPU_ESTATE, "set_cells before disagg");
  if (h->abi.COMPUTE_TREELINE) return vicgpu_fail(VICGPU_EUNSUPPORTED, "COMPUTE_TREELINE with device-side disaggregation: compute_treeline() is not implemented on the device");
  if (dopt->Ndays < 1 || h->abi.nrecs < 1) return vicgpu_fail(VICGPU_EINVAL, "Ndays and nrecs must be positive");
  if (24 % h->abi.dt != 0) return vicgpu_fail(VICGPU_EINVAL, "dt must divide 24");
  if (h->abi.SNOW_STEP < 1 || h->abi.dt % h->abi.SNOW_STEP != 0 || h->abi.NF != h->abi.dt / h->abi.SNOW_STEP)
    return vicgpu_fail(VICGPU_EINVAL, "SNOW_STEP must divide dt and NF must be dt / SNOW_STEP");
  // the per-cell scratch holds (Ndays + 1) local days of hourly values; the records (shifted by the start hour and by up to one
  // day of time-zone offset, initialize_atmos.c:125-156) must stay inside it
  if ((long long)h->abi.nrecs * h->abi.dt + dopt->starthour + 24 > ((long long)dopt->Ndays + 1) * 24)
    return vicgpu_fail(VICGPU_EINVAL, "nrecs * dt + starthour exceeds the Ndays of daily input");
  CK(cudaSetDevice(h->device));
  const vicgpu_layout& L = h->o.L;
  const int ncell = h->t.ncell, nrecs = h->abi.nrecs, Ndays = dopt->Ndays;
  DisArgs a;
  a.o = h->d_o;
  a.cellpar = h->d_cellpar;
  a.d = disagg_opts_from_abi(h->abi, *dopt, L.f_nslot);
  a.s.ntotal = (size_t)ncell;
  a.s.Ndl = Ndays + 1;
  // forcing window [0, nrecs)
  const size_t per = (size_t)ncell * L.f_stride;
  // the whole run's forcing becomes window 0; a window uploaded earlier is dropped
  CK(cudaStreamSynchronize(h->stream_copy));
  ForcingWindow& w = h->fwin[0];
  w.nrec = 0;
  h->fwin[1].nrec = 0;
  int rc = vicgpu_ensure_window(h, w, per * nrecs);
  if (rc) return rc;
  a.forcing = w.d;
  // daily input: copy, then transpose to [Ndays*4][ncell]
  struct DevBuf {  // freed on every return path
    double* p = nullptr;
    ~DevBuf() { cudaFree(p); }
  } b_in, b_daily, b_scratch, b_t;
  double *&d_in = b_in.p, *&d_daily = b_daily.p, *&d_scratch = b_scratch.p;
  const size_t nd4 = (size_t)Ndays * 4;
  CK(cudaMalloc(&d_daily, nd4 * ncell * sizeof(double)));
  if (time_major) {  // [Ndays][4][ncell] is the device layout: straight copy
    CK(cudaMemcpyAsync(d_daily, daily, nd4 * ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    CK(cudaMalloc(&d_in, nd4 * ncell * sizeof(double)));
    CK(cudaMemcpyAsync(d_in, daily, nd4 * ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = vicgpu_transpose(h, d_in, d_daily, ncell, (int)nd4, 1);
    if (rc) return rc;
  }
  a.daily = d_daily;
  // scratch for a chunk of cells: bounded to ~8 GiB
  const size_t per_cell = a.s.per_cell();
  size_t chunk = std::min<size_t>((size_t)ncell, std::max<size_t>(1024, ((size_t)8 << 30) / (per_cell * sizeof(double))));
  CK(cudaMalloc(&d_scratch, per_cell * chunk * sizeof(double)));
  a.s.base = d_scratch;
  a.s.ncell = chunk;
  for (int c0 = 0; c0 < ncell; c0 += (int)chunk) {
    a.s.cell0 = c0;
    a.nchunk = (int)std::min<size_t>(chunk, (size_t)(ncell - c0));
    const size_t n = (size_t)a.nchunk;
    k_dis_solar<<<blocks_for(n * 365), 128, 0, h->stream>>>(a);
    k_dis_daily<<<blocks_for(n), 128, 0, h->stream>>>(a);
    k_dis_items<0><<<blocks_for(n * a.s.Ndl), 128, 0, h->stream>>>(a, a.s.Ndl);
    k_dis_items<1><<<blocks_for(n * a.s.Ndl), 128, 0, h->stream>>>(a, a.s.Ndl);
    k_dis_items<2><<<blocks_for(n * (2 * a.s.Ndl + 2)), 128, 0, h->stream>>>(a, 2 * a.s.Ndl + 2);
    k_dis_items<3><<<blocks_for(n * a.s.Ndl), 128, 0, h->stream>>>(a, a.s.Ndl);
    k_dis_items<4><<<blocks_for(n * nrecs), 128, 0, h->stream>>>(a, nrecs);
    CK(cudaGetLastError());
  }
  h->last_launches = 7 * ((ncell + (int)chunk - 1) / (int)chunk) + 1;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventRecord(w.ready, h->stream));
  w.rec0 = 0;
  w.nrec = nrecs;
  h->fwin_next = 1;
  if (forcing_out) {
    // [nrecs][f_stride][ncell] -> [nrecs][ncell][f_stride], staged in slabs of records
    const size_t slab_recs = std::min<size_t>((size_t)nrecs, std::min<size_t>(65535, std::max<size_t>(1, ((size_t)256 << 20) / (per * sizeof(double)))));
    double*& d_t = b_t.p;
    CK(cudaMalloc(&d_t, per * slab_recs * sizeof(double)));
    for (size_t r = 0; r < (size_t)nrecs; r += slab_recs) {
      const int nr = (int)std::min<size_t>(slab_recs, (size_t)nrecs - r);
      rc = vicgpu_transpose(h, w.d + r * per, d_t, L.f_stride, ncell, nr);
      if (rc) return rc;
      CK(cudaMemcpyAsync(forcing_out + r * per, d_t, per * nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
  }
  return VICGPU_OK;
}

// vicgpu_lakeice.cu -- vicgpu_ice_melt: the lake-ice surface solve (vic_lakeice.cuh) for a batch of independent columns.
// Records cross the ABI row-major; a block stages its 128 rows through shared memory so that global loads and stores are coalesced
// (row stride padded by one double against bank conflicts) and every thread then works on its own row.
#include "vicgpu.h"
#include "vicgpu_internal.h"
#include "vic_lakeice.cuh"

namespace {
constexpr int ROWS = 128;

__global__ void __launch_bounds__(ROWS) k_ice_melt(int n, int delta_t, int tfallback, const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double s[ROWS * (VICGPU_ICE_NIN + 1)];
  const int row0 = blockIdx.x * ROWS;
  const int rows = min(ROWS, n - row0);
  for (int k = threadIdx.x; k < rows * VICGPU_ICE_NIN; k += ROWS) s[(k / VICGPU_ICE_NIN) * (VICGPU_ICE_NIN + 1) + k % VICGPU_ICE_NIN] = in[(size_t)row0 * VICGPU_ICE_NIN + k];
  __syncthreads();
  double o[VICGPU_ICE_NOUT];
  if ((int)threadIdx.x < rows) {
    const double* a = &s[threadIdx.x * (VICGPU_ICE_NIN + 1)];
    vic::IceSnow snow;
    snow.swq = a[ICEIN_swq]; snow.surf_temp = a[ICEIN_surf_temp]; snow.pack_temp = a[ICEIN_pack_temp]; snow.pack_water = a[ICEIN_pack_water];
    snow.surf_water = a[ICEIN_surf_water]; snow.vapor_flux = a[ICEIN_vapor_flux]; snow.blowing_flux = 0; snow.surface_flux = a[ICEIN_surface_flux];
    snow.surf_temp_fbflag = a[ICEIN_surf_temp_fbflag]; snow.surf_temp_fbcount = a[ICEIN_surf_temp_fbcount];
    snow.coverage = 0; snow.mass_error = 0; snow.coldcontent = 0;
    vic::IceLake lake{a[ICEIN_ice_water_eq], a[ICEIN_areai], a[ICEIN_hice], a[ICEIN_volume]};
    vic::IceMeltOut r = {};
    const int rc = vic::ice_melt(a[ICEIN_z2], a[ICEIN_aero_resist], a[ICEIN_latent_heat_Le], snow, lake, delta_t, a[ICEIN_Z0], a[ICEIN_rainfall], a[ICEIN_snowfall],
                                 a[ICEIN_wind], a[ICEIN_Tcutoff], a[ICEIN_air_temp], a[ICEIN_net_short], a[ICEIN_longwave], a[ICEIN_density], a[ICEIN_pressure],
                                 a[ICEIN_vpd], a[ICEIN_vp], tfallback != 0, r);
    o[ICEOUT_rc] = rc; o[ICEOUT_aero_resist_used] = r.aero_resist_used; o[ICEOUT_melt] = r.melt; o[ICEOUT_advection] = r.advection; o[ICEOUT_deltaCC] = r.deltaCC;
    o[ICEOUT_SnowFlux] = r.SnowFlux; o[ICEOUT_latent] = r.latent; o[ICEOUT_sensible] = r.sensible; o[ICEOUT_Qnet] = r.Qnet;
    o[ICEOUT_refreeze_energy] = r.refreeze_energy; o[ICEOUT_LWnet] = r.LWnet; o[ICEOUT_swq] = snow.swq; o[ICEOUT_surf_temp] = snow.surf_temp;
    o[ICEOUT_pack_temp] = snow.pack_temp; o[ICEOUT_pack_water] = snow.pack_water; o[ICEOUT_surf_water] = snow.surf_water; o[ICEOUT_vapor_flux] = snow.vapor_flux;
    o[ICEOUT_blowing_flux] = snow.blowing_flux; o[ICEOUT_surface_flux] = snow.surface_flux; o[ICEOUT_surf_temp_fbflag] = snow.surf_temp_fbflag;
    o[ICEOUT_surf_temp_fbcount] = snow.surf_temp_fbcount; o[ICEOUT_coverage] = snow.coverage; o[ICEOUT_mass_error] = snow.mass_error;
    o[ICEOUT_coldcontent] = snow.coldcontent; o[ICEOUT_ice_water_eq] = lake.ice_water_eq; o[ICEOUT_volume] = lake.volume;
  }
  __syncthreads();  // everyone has read its inputs: the staging buffer is reused for the results (NOUT <= NIN)
  static_assert(VICGPU_ICE_NOUT <= VICGPU_ICE_NIN, "staging buffer sized for the inputs");
  if ((int)threadIdx.x < rows)
    for (int k = 0; k < VICGPU_ICE_NOUT; k++) s[threadIdx.x * (VICGPU_ICE_NOUT + 1) + k] = o[k];
  __syncthreads();
  for (int k = threadIdx.x; k < rows * VICGPU_ICE_NOUT; k += ROWS) out[(size_t)row0 * VICGPU_ICE_NOUT + k] = s[(k / VICGPU_ICE_NOUT) * (VICGPU_ICE_NOUT + 1) + k % VICGPU_ICE_NOUT];
}
}  // namespace

extern "C" int vicgpu_ice_melt(int device, int n, int delta_t, int tfallback, const double* in, double* out) {
  if (n < 0 || delta_t < 1 || (n > 0 && (!in || !out))) return vicgpu_fail(VICGPU_EINVAL, "bad argument");
  if (n == 0) return VICGPU_OK;
  int ndev = 0;
  const cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return vicgpu_fail(VICGPU_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return vicgpu_fail(VICGPU_ENODEV, "device ordinal out of range");
  CK(cudaSetDevice(device));
  struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { cudaFree(p); }
  } d_in, d_out;
  CK(cudaMalloc(&d_in.p, (size_t)n * VICGPU_ICE_NIN * sizeof(double)));
  CK(cudaMalloc(&d_out.p, (size_t)n * VICGPU_ICE_NOUT * sizeof(double)));
  CK(cudaMemcpy(d_in.p, in, (size_t)n * VICGPU_ICE_NIN * sizeof(double), cudaMemcpyHostToDevice));
  k_ice_melt<<<(n + ROWS - 1) / ROWS, ROWS>>>(n, delta_t, tfallback, d_in.p, d_out.p);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out.p, (size_t)n * VICGPU_ICE_NOUT * sizeof(double), cudaMemcpyDeviceToHost));
  return VICGPU_OK;
}

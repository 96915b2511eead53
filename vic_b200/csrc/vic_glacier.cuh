// vic_glacier.cuh -- placeholder until the glacier tile step is written
#ifndef VIC_GLACIER_CUH
#define VIC_GLACIER_CUH
#include "vic_surface.cuh"
namespace vic {
template <int NN>
VIC_HDI int surface_fluxes_glac(double, double, double, Hru<NN>&, const AeroState&, const double*, int, const Ctx&, int, SurfaceFluxOut&) {
  return ERROR_I;
}
}  // namespace vic
#endif

// vicgpu_nc.cu -- C-ABI of the NetCDF forcing slab reader (host code only; include/vicgpu.h, vic_b200/host/vicgpu_ncslab.h)
#include "vicgpu.h"
#include "vicgpu_internal.h"
#include "../host/vicgpu_ncslab.h"

struct vicgpu_ncfile {
  vicgpu_nc::File file;
  explicit vicgpu_ncfile(const char* path) : file(path) {}
};

namespace {
int nc_error(const std::exception& e) {
  const std::string m = e.what();
  return vicgpu_fail(m.find("NetCDF-4") != std::string::npos || m.find("CDF-5") != std::string::npos ? VICGPU_EUNSUPPORTED : VICGPU_EINVAL, m);
}
}  // namespace

extern "C" int vicgpu_nc_open(vicgpu_ncfile** nc, const char* path) {
  if (!nc || !path) return vicgpu_fail(VICGPU_EINVAL, "null argument");
  *nc = nullptr;
  try {
    *nc = new vicgpu_ncfile(path);
  } catch (const std::exception& e) {
    return nc_error(e);
  }
  return VICGPU_OK;
}

extern "C" int vicgpu_nc_close(vicgpu_ncfile* nc) {
  delete nc;
  return VICGPU_OK;
}

extern "C" int vicgpu_nc_dims(vicgpu_ncfile* nc, long long* ntime, long long* nlat, long long* nlon) {
  if (!nc || !ntime || !nlat || !nlon) return vicgpu_fail(VICGPU_EINVAL, "null argument");
  try {
    const char* names[3] = {"time", "lat", "lon"};
    long long* out[3] = {ntime, nlat, nlon};
    for (int k = 0; k < 3; k++) {
      const vicgpu_nc::Var* v = nc->file.var(names[k]);
      if (!v || v->dimids.size() != 1) return vicgpu_fail(VICGPU_EINVAL, std::string("no one-dimensional variable '") + names[k] + "'");
      *out[k] = (long long)nc->file.dimlen(v->dimids[0]);
    }
  } catch (const std::exception& e) {
    return nc_error(e);
  }
  return VICGPU_OK;
}

extern "C" int vicgpu_nc_read_slab(vicgpu_ncfile* nc, int nvar, const char* const* varnames, long long t0, long long nt, int ncell, const double* lat,
                                   const double* lng, double* out) {
  if (!nc || !varnames || !lat || !lng || !out) return vicgpu_fail(VICGPU_EINVAL, "null argument");
  try {
    vicgpu_nc::read_slab(nc->file, nvar, varnames, t0, nt, ncell, lat, lng, out);
  } catch (const std::exception& e) {
    return nc_error(e);
  }
  return VICGPU_OK;
}

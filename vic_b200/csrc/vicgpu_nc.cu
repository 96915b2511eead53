// vicgpu_nc.cu -- C-ABI of the NetCDF forcing slab reader (host code only; include/vicgpu.h, vic_b200/host/vicgpu_ncslab.h)
#include "vicgpu.h"
#include "vicgpu_internal.h"
#include "../host/vicgpu_ncslab.h"
#include "../host/vicgpu_ncwrite.h"

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

// ---- output writer -------------------------------------------------------------------------------------------------------------
struct vicgpu_ncout {
  vicgpu_nc::Writer* w = nullptr;
  ~vicgpu_ncout() { delete w; }
};

extern "C" int vicgpu_ncout_create(vicgpu_ncout** out, const char* path, const vicgpu_ncout_spec* sp) {
  if (!out || !path || !sp || !sp->vars || !sp->time_units || !sp->lat_index || !sp->lon_index || sp->nvar < 1 || sp->ncell < 0)
    return vicgpu_fail(VICGPU_EINVAL, "null or empty argument");
  *out = nullptr;
  try {
    auto str = [](const char* s) { return std::string(s ? s : ""); };
    std::vector<vicgpu_nc::OutVar> vars((size_t)sp->nvar);
    for (int v = 0; v < sp->nvar; v++) {
      const vicgpu_ncout_var& a = sp->vars[v];
      if (!a.name) return vicgpu_fail(VICGPU_EINVAL, "variable without a name");
      vars[(size_t)v].name = a.name;
      vars[(size_t)v].nelem = a.nelem;
      vars[(size_t)v].text_atts = {{"long_name", str(a.long_name)}, {"units", str(a.units)}, {"standard_name", str(a.standard_name)}, {"cell_methods", str(a.cell_methods)},
                                   {"internal_vic_name", str(a.internal_vic_name)}, {"category", str(a.category)}};
    }
    std::vector<std::pair<std::string, std::string>> gtext;
    for (int k = 0; k < sp->ntext; k++) gtext.emplace_back(str(sp->text_keys[k]), str(sp->text_values[k]));
    std::vector<std::pair<std::string, int>> gint;
    for (int k = 0; k < sp->nint; k++) gint.emplace_back(str(sp->int_keys[k]), sp->int_values[k]);
    std::vector<int> li(sp->lat_index, sp->lat_index + sp->ncell), lo(sp->lon_index, sp->lon_index + sp->ncell);
    vicgpu_ncout* h = new vicgpu_ncout;
    try {
      h->w = new vicgpu_nc::Writer(path, sp->nlat, sp->lat0, sp->dlat, sp->nlon, sp->lon0, sp->dlon, sp->depth, sp->time_units, sp->time_step, vars, gtext, gint, li, lo);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  } catch (const std::exception& e) {
    return vicgpu_fail(VICGPU_EINVAL, e.what());
  }
  return VICGPU_OK;
}

extern "C" int vicgpu_ncout_write_step(vicgpu_ncout* w, const float* rows, long long row_stride, const int* col_of_var) {
  if (!w || !w->w || !rows || !col_of_var || row_stride < 1) return vicgpu_fail(VICGPU_EINVAL, "bad argument");
  try {
    w->w->write_step(rows, (size_t)row_stride, col_of_var);
  } catch (const std::exception& e) {
    return vicgpu_fail(VICGPU_EINVAL, e.what());
  }
  return VICGPU_OK;
}

extern "C" int vicgpu_ncout_close(vicgpu_ncout* w) {
  if (!w) return VICGPU_OK;
  int rc = VICGPU_OK;
  try {
    if (w->w) w->w->close();
  } catch (const std::exception& e) {
    rc = vicgpu_fail(VICGPU_EINVAL, e.what());
  }
  delete w;
  return rc;
}

// vicgpu_ncslab.h -- time-major slabs out of a NetCDF forcing file, host side, no library needed.
//
// The reference reads its NetCDF forcing one cell at a time: for every cell and every variable one strided nc_get_varm_* of the
// cell's whole time series out of a (time, lat, lon) variable (read_atmos_data.c:109-338) -- Ncell x Nvar passes over the file, each
// touching one value per (lat, lon) grid.  Here the file is read the way it is laid out: for every time step and variable ONE
// contiguous (lat, lon) grid, from which the modelled cells are gathered, giving [time][variable][cell] -- the layout the device
// disaggregation takes as it is (vicgpu_disagg_tm, include/vicgpu.h) -- so that the hourly forcing never exists on the host.
//
// What is mirrored from the reference, value for value:
//   * variables are (time, lat, lon); "time", "lat", "lon" are one-dimensional float or double coordinate variables (:147-167)
//   * a cell is the FIRST index whose coordinate, read as double, equals the cell's (double)(float) latitude / longitude exactly
//     (:176-189); a cell without a match is an error (the reference asserts)
//   * NC_SHORT: (double)v / inverse_scale_factor when the variable has that attribute, else (double)v * scale_factor, the attribute
//     taken as float (nc_get_att_float) (:226-252); NC_FLOAT: (double)v (:283-296); NC_DOUBLE: v (:297-310); other types: error (:311)
//
// Container: the NetCDF classic formats, CDF-1 and CDF-2 (64-bit offsets), parsed here from their published layout (big-endian
// header: magic, numrecs, dim_list, gatt_list, var_list; fixed variables contiguous, record variables interleaved per record).
// NetCDF-4 files are HDF5 containers and need libnetcdf/libhdf5, which this image does not have: they are refused with a message
// saying so (`nccopy -k classic` converts them).  NC_USHORT exists only in CDF-5 / NetCDF-4 and is refused with them.
#ifndef VICGPU_NCSLAB_H
#define VICGPU_NCSLAB_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace vicgpu_nc {

enum NcType { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };
inline int type_size(int t) { return t == NC_BYTE || t == NC_CHAR ? 1 : t == NC_SHORT ? 2 : t == NC_INT || t == NC_FLOAT ? 4 : t == NC_DOUBLE ? 8 : 0; }

struct Attr {
  int type = 0;
  std::vector<unsigned char> raw;  // big-endian values
  size_t nelems = 0;
};
struct Var {
  std::string name;
  std::vector<int> dimids;
  std::map<std::string, Attr> atts;
  int type = 0;
  uint64_t vsize = 0, begin = 0;
  bool record = false;
};

template <class T>
inline T from_be(const unsigned char* p) {
  unsigned char b[sizeof(T)];
  for (size_t i = 0; i < sizeof(T); i++) b[i] = p[sizeof(T) - 1 - i];
  T v;
  memcpy(&v, b, sizeof(T));
  return v;
}
// one value of an attribute / variable as double, whatever its external type
inline double be_to_double(int type, const unsigned char* p) {
  switch (type) {
    case NC_BYTE: return (double)(signed char)p[0];
    case NC_CHAR: return (double)p[0];
    case NC_SHORT: return (double)from_be<int16_t>(p);
    case NC_INT: return (double)from_be<int32_t>(p);
    case NC_FLOAT: return (double)from_be<float>(p);
    case NC_DOUBLE: return from_be<double>(p);
  }
  throw std::runtime_error("unknown NetCDF external type");
}

class File {
 public:
  explicit File(const std::string& path) : path_(path) {
    f_ = fopen(path.c_str(), "rb");
    if (!f_) throw std::runtime_error("cannot open " + path);
    try {
      parse_header();
    } catch (...) {
      fclose(f_);
      throw;
    }
  }
  ~File() { if (f_) fclose(f_); }
  File(const File&) = delete;
  File& operator=(const File&) = delete;

  const Var* var(const std::string& name) const {
    auto it = byname_.find(name);
    return it == byname_.end() ? nullptr : &vars_[it->second];
  }
  uint64_t dimlen(int dimid) const { return dimid == recdim_ ? numrecs_ : dimlen_[dimid]; }
  int version() const { return version_; }

  // the whole of a one-dimensional coordinate variable as doubles (nc_get_vara_double converts float the same way)
  std::vector<double> coordinate(const std::string& name) const {
    const Var* v = var(name);
    if (!v) throw std::runtime_error("no variable '" + name + "' in " + path_);
    if (v->dimids.size() != 1) throw std::runtime_error("'" + name + "' is not one-dimensional");
    if (v->type != NC_FLOAT && v->type != NC_DOUBLE) throw std::runtime_error("'" + name + "' must be float or double (read_atmos_data.c:148)");
    const uint64_t n = dimlen(v->dimids[0]);
    const int es = type_size(v->type);
    std::vector<double> out(n);
    std::vector<unsigned char> buf;
    for (uint64_t i = 0; i < n; i++) {
      // a record coordinate variable ("time" on the unlimited dimension) has one value per record
      const uint64_t off = v->record ? v->begin + i * recsize_ : v->begin + i * es;
      read_at(off, es, buf);
      out[i] = be_to_double(v->type, buf.data());
    }
    return out;
  }

  // float value of a numeric attribute as nc_get_att_float gives it; false when the variable has no such attribute
  bool att_float(const Var& v, const std::string& name, float* out) const {
    auto it = v.atts.find(name);
    if (it == v.atts.end() || it->second.nelems < 1 || it->second.type == NC_CHAR) return false;
    *out = (float)be_to_double(it->second.type, it->second.raw.data());
    return true;
  }

  // the (lat, lon) grid of time step t of a (time, lat, lon) variable, raw big-endian
  void read_grid(const Var& v, uint64_t t, std::vector<unsigned char>& buf) const {
    const uint64_t grid = dimlen(v.dimids[1]) * dimlen(v.dimids[2]) * type_size(v.type);
    const uint64_t off = v.record ? v.begin + t * recsize_ : v.begin + t * grid;
    read_at(off, grid, buf);
  }

 private:
  std::string path_;
  FILE* f_ = nullptr;
  int version_ = 0;
  uint64_t numrecs_ = 0, recsize_ = 0;
  int recdim_ = -1;
  std::vector<uint64_t> dimlen_;
  std::vector<std::string> dimname_;
  std::vector<Var> vars_;
  std::unordered_map<std::string, size_t> byname_;

  void read_at(uint64_t off, uint64_t n, std::vector<unsigned char>& buf) const {
    buf.resize(n);
    if (fseeko(f_, (off_t)off, SEEK_SET) != 0 || fread(buf.data(), 1, n, f_) != n) throw std::runtime_error("short read in " + path_);
  }
  uint32_t u32() {
    unsigned char b[4];
    if (fread(b, 1, 4, f_) != 4) throw std::runtime_error("truncated NetCDF header in " + path_);
    return from_be<uint32_t>(b);
  }
  uint64_t offset() { return version_ == 2 ? ((uint64_t)u32() << 32) | u32() : (uint64_t)u32(); }
  std::string name() {
    const uint32_t n = u32();
    std::string s(n, '\0');
    if (n && fread(&s[0], 1, n, f_) != n) throw std::runtime_error("truncated NetCDF header in " + path_);
    skip_pad(n);
    return s;
  }
  void skip_pad(uint64_t n) {
    const uint64_t pad = (4 - n % 4) % 4;
    if (pad) fseeko(f_, (off_t)pad, SEEK_CUR);
  }
  void att_list(std::map<std::string, Attr>& atts) {
    const uint32_t tag = u32(), n = u32();
    if (tag == 0 && n == 0) return;
    if (tag != 0x0C) throw std::runtime_error("bad attribute list tag in " + path_);
    for (uint32_t i = 0; i < n; i++) {
      const std::string nm = name();
      Attr a;
      a.type = (int)u32();
      a.nelems = u32();
      const int es = type_size(a.type);
      if (!es) throw std::runtime_error("attribute '" + nm + "' has a type outside the classic format in " + path_);
      a.raw.resize((size_t)a.nelems * es);
      if (!a.raw.empty() && fread(a.raw.data(), 1, a.raw.size(), f_) != a.raw.size()) throw std::runtime_error("truncated NetCDF header in " + path_);
      skip_pad(a.raw.size());
      atts[nm] = std::move(a);
    }
  }
  void parse_header() {
    unsigned char magic[4];
    if (fread(magic, 1, 4, f_) != 4) throw std::runtime_error("not a NetCDF file: " + path_);
    if (memcmp(magic, "\x89HDF", 4) == 0)
      throw std::runtime_error(path_ + " is a NetCDF-4 / HDF5 container; this reader serves the classic formats (CDF-1, CDF-2) -- convert with `nccopy -k classic`, "
                                       "or link libnetcdf on the host side of the ABI");
    if (memcmp(magic, "CDF", 3) != 0) throw std::runtime_error("not a NetCDF file: " + path_);
    version_ = magic[3];
    if (version_ != 1 && version_ != 2) throw std::runtime_error(path_ + ": NetCDF classic version " + std::to_string(version_) + " (CDF-5) is not served; CDF-1 and CDF-2 are");
    const uint32_t nr = u32();
    bool streaming = nr == 0xFFFFFFFFu;
    numrecs_ = nr;
    {  // dim_list
      const uint32_t tag = u32(), n = u32();
      if (!(tag == 0 && n == 0)) {
        if (tag != 0x0A) throw std::runtime_error("bad dimension list tag in " + path_);
        for (uint32_t i = 0; i < n; i++) {
          dimname_.push_back(name());
          dimlen_.push_back(u32());
          if (dimlen_.back() == 0) recdim_ = (int)i;
        }
      }
    }
    std::map<std::string, Attr> gatts;
    att_list(gatts);
    {  // var_list
      const uint32_t tag = u32(), n = u32();
      if (!(tag == 0 && n == 0)) {
        if (tag != 0x0B) throw std::runtime_error("bad variable list tag in " + path_);
        for (uint32_t i = 0; i < n; i++) {
          Var v;
          v.name = name();
          const uint32_t nd = u32();
          for (uint32_t d = 0; d < nd; d++) v.dimids.push_back((int)u32());
          att_list(v.atts);
          v.type = (int)u32();
          v.vsize = u32();
          v.begin = offset();
          v.record = !v.dimids.empty() && v.dimids[0] == recdim_;
          for (int d : v.dimids)
            if (d < 0 || d >= (int)dimlen_.size()) throw std::runtime_error("variable '" + v.name + "' names an unknown dimension in " + path_);
          byname_[v.name] = vars_.size();
          vars_.push_back(std::move(v));
        }
      }
    }
    // record size: the sum of the record variables' padded slab sizes -- computed from the shapes, not from the vsize fields
    // (vsize is 2^32 - 1 for slabs beyond 4 GiB).  A single record variable is stored without padding.
    size_t nrecvar = 0;
    uint64_t sum = 0, only = 0;
    for (const Var& v : vars_)
      if (v.record) {
        uint64_t b = type_size(v.type);
        for (size_t d = 1; d < v.dimids.size(); d++) b *= dimlen_[v.dimids[d]];
        only = b;
        sum += (b + 3) / 4 * 4;
        nrecvar++;
      }
    recsize_ = nrecvar == 1 ? only : sum;
    if (streaming) {  // numrecs not written: take it from the file length
      fseeko(f_, 0, SEEK_END);
      const uint64_t len = (uint64_t)ftello(f_);
      uint64_t first = UINT64_MAX;
      for (const Var& v : vars_)
        if (v.record && v.begin < first) first = v.begin;
      numrecs_ = recsize_ && first != UINT64_MAX && len > first ? (len - first) / recsize_ : 0;
    }
  }
};

// first index whose coordinate equals x exactly (read_atmos_data.c:176-189), -1 if none
struct FirstMatch {
  std::unordered_map<double, long long> first;
  explicit FirstMatch(const std::vector<double>& coord) {
    for (size_t i = 0; i < coord.size(); i++) first.emplace(coord[i], (long long)i);
  }
  long long operator()(double x) const {
    auto it = first.find(x);
    return it == first.end() ? -1 : it->second;
  }
};

// out[(t * nvar + v) * ncell + c] = variable varnames[v] of cell c at time step t0 + t, converted as the reference converts it
inline void read_slab(const File& nc, int nvar, const char* const* varnames, long long t0, long long nt, long long ncell, const double* lat, const double* lng,
                      double* out) {
  if (nvar < 1 || nt < 0 || ncell < 0 || t0 < 0) throw std::runtime_error("bad slab request");
  const FirstMatch ilat(nc.coordinate("lat")), ilon(nc.coordinate("lon"));
  const Var* tv = nc.var("time");
  if (!tv || tv->dimids.size() != 1 || (tv->type != NC_FLOAT && tv->type != NC_DOUBLE)) throw std::runtime_error("no one-dimensional float / double variable 'time'");
  const Var* latv = nc.var("lat");
  const Var* lonv = nc.var("lon");
  const uint64_t ntime = nc.dimlen(tv->dimids[0]), nlon = nc.dimlen(lonv->dimids[0]);
  if ((uint64_t)(t0 + nt) > ntime) throw std::runtime_error("time steps " + std::to_string(t0) + ".." + std::to_string(t0 + nt - 1) + " requested, the file has " + std::to_string(ntime));
  std::vector<uint64_t> gidx((size_t)ncell);
  for (long long c = 0; c < ncell; c++) {
    const long long i = ilat(lat[c]), j = ilon(lng[c]);
    if (i < 0 || j < 0) throw std::runtime_error("cell " + std::to_string(c) + " (lat " + std::to_string(lat[c]) + ", lon " + std::to_string(lng[c]) + ") has no exactly matching grid point in the forcing file");
    gidx[(size_t)c] = (uint64_t)i * nlon + (uint64_t)j;
  }
  struct Plan {
    const Var* v;
    bool has_inv;
    float scale;
  };
  std::vector<Plan> plan((size_t)nvar);
  for (int k = 0; k < nvar; k++) {
    const Var* v = nc.var(varnames[k]);
    if (!v) throw std::runtime_error(std::string("no variable '") + varnames[k] + "' in the forcing file");
    if (v->dimids.size() != 3 || v->dimids[0] != tv->dimids[0] || v->dimids[1] != latv->dimids[0] || v->dimids[2] != lonv->dimids[0])
      throw std::runtime_error(std::string("variable '") + varnames[k] + "' is not (time, lat, lon) (read_atmos_data.c:213-216)");
    Plan p{v, false, 0.f};
    if (v->type == NC_SHORT) {
      p.has_inv = nc.att_float(*v, "inverse_scale_factor", &p.scale);
      if (!p.has_inv && !nc.att_float(*v, "scale_factor", &p.scale))
        throw std::runtime_error(std::string("short variable '") + varnames[k] + "' has neither inverse_scale_factor nor scale_factor (read_atmos_data.c:226-231)");
    } else if (v->type != NC_FLOAT && v->type != NC_DOUBLE) {
      throw std::runtime_error(std::string("variable '") + varnames[k] + "': type not supported (read_atmos_data.c:311-316)");
    }
    plan[(size_t)k] = p;
  }
  std::vector<unsigned char> grid;
  for (long long t = 0; t < nt; t++)
    for (int k = 0; k < nvar; k++) {
      const Plan& p = plan[(size_t)k];
      nc.read_grid(*p.v, (uint64_t)(t0 + t), grid);
      double* o = out + ((size_t)t * nvar + k) * (size_t)ncell;
      const unsigned char* g = grid.data();
      if (p.v->type == NC_SHORT) {
        if (p.has_inv)
          for (long long c = 0; c < ncell; c++) o[c] = (double)from_be<int16_t>(g + gidx[(size_t)c] * 2) / p.scale;
        else
          for (long long c = 0; c < ncell; c++) o[c] = (double)from_be<int16_t>(g + gidx[(size_t)c] * 2) * p.scale;
      } else if (p.v->type == NC_FLOAT) {
        for (long long c = 0; c < ncell; c++) o[c] = (double)from_be<float>(g + gidx[(size_t)c] * 4);
      } else {
        for (long long c = 0; c < ncell; c++) o[c] = from_be<double>(g + gidx[(size_t)c] * 8);
      }
    }
}

}  // namespace vicgpu_nc
#endif

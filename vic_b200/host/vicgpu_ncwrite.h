// vicgpu_ncwrite.h -- the model output as a NetCDF file, one contiguous record per output step, host side, no library.
//
// The reference's writer (WriteOutputNetCDF.c) builds, per output step and per variable, a (depth, lat, lon) float grid -- the modelled
// cells' aggregates narrowed to float32, NETCDF_FILL_VALUE elsewhere (:386-452) -- and hands each to the NetCDF library with its own
// putVar: 40-184 library calls per step, the wall-clock bottleneck once the physics takes microseconds per cell.  Here
// `time` is the record dimension, so ONE record holds the grids of all variables of a step: the step's float32 rows (exactly what
// vicgpu_step_f32 brings back) are scattered into the record buffer and the record goes out with one write.
//
// Same logical content as the reference's file (WriteOutputNetCDF.c:163-299): dimensions lat, lon, bnds, time, depth (MAX_BANDS); double
// lat / lon and float time / depth coordinate variables with the reference's attributes and values (lat_i = gridStartLat + i * gridStepLat,
// time_i = i * out_dt hours or i days since the start date); one float variable per output variable, (time, lat, lon) or, for variables
// with more than one element, (time, depth, lat, lon) with the elements beyond nelem left at the fill value; the seven per-variable
// attributes long_name, units, standard_name, cell_methods, _FillValue, internal_vic_name, category; global attributes as given.
// Differences, all of the container: NetCDF classic with 64-bit offsets (CDF-2) instead of NetCDF-4/HDF5 (this image has no libnetcdf;
// any NetCDF tool reads both), `time` unlimited instead of fixed, no internal compression.
#ifndef VICGPU_NCWRITE_H
#define VICGPU_NCWRITE_H
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vicgpu_nc {

struct OutVar {
  std::string name;  // NetCDF variable name (VariableMetaData::name)
  int nelem = 1;     // > 1: (time, depth, lat, lon)
  std::vector<std::pair<std::string, std::string>> text_atts;  // long_name, units, standard_name, cell_methods, internal_vic_name, category
};

class Writer {
 public:
  // lat_index / lon_index [ncell]: grid position of every modelled cell (latitudeToIndex / longitudeToIndex of the reference)
  Writer(const std::string& path, int nlat, double lat0, double dlat, int nlon, double lon0, double dlon, int depth, const std::string& time_units, double time_step,
         const std::vector<OutVar>& vars, const std::vector<std::pair<std::string, std::string>>& global_text, const std::vector<std::pair<std::string, int>>& global_int,
         const std::vector<int>& lat_index, const std::vector<int>& lon_index, float fill = 1e20f)
      : nlat_(nlat), nlon_(nlon), depth_(depth), time_step_(time_step), vars_(vars), fill_(fill) {
    if (nlat < 1 || nlon < 1 || depth < 1) throw std::runtime_error("bad grid");
    if (lat_index.size() != lon_index.size()) throw std::runtime_error("lat_index and lon_index differ in length");
    for (size_t k = 0; k < lat_index.size(); k++) {
      if (lat_index[k] < 0 || lat_index[k] >= nlat || lon_index[k] < 0 || lon_index[k] >= nlon) throw std::runtime_error("cell outside the grid");
      cell_pos_.push_back((size_t)lat_index[k] * nlon + lon_index[k]);
    }
    for (const OutVar& v : vars)
      if (v.nelem < 1 || v.nelem > depth) throw std::runtime_error("variable '" + v.name + "': nelem outside 1..depth");
    f_ = fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("cannot create " + path);
    write_header(lat0, dlat, lon0, dlon, time_units, global_text, global_int);
  }
  ~Writer() {
    try {
      close();
    } catch (...) {
    }
  }
  Writer(const Writer&) = delete;
  Writer& operator=(const Writer&) = delete;

  // rows [ncell][row_stride] float32; the elements of variable v of a cell start at column col_of_var[v]
  void write_step(const float* rows, size_t row_stride, const int* col_of_var) {
    if (!f_) throw std::runtime_error("writer is closed");
    const size_t grid = (size_t)nlat_ * nlon_;
    unsigned char* p = rec_.data();
    put_be(p, (float)(nrec_ * time_step_));  // the time coordinate of this record
    p += 4;
    // (everything but the modelled cells' positions of the used planes holds the fill value since the record buffer was set up:
    // the same positions are overwritten at every step)
    // blocks of cells: a block's rows stay in cache while its values go to the planes of all variables (cells in grid order write
    // each plane sequentially)
    const size_t ncell = cell_pos_.size(), block = 512;
    for (size_t k0 = 0; k0 < ncell; k0 += block) {
      const size_t k1 = std::min(ncell, k0 + block);
      unsigned char* pv = p;
      for (size_t v = 0; v < vars_.size(); v++) {
        const size_t planes = vars_[v].nelem > 1 ? (size_t)depth_ : 1;
        for (int e = 0; e < vars_[v].nelem; e++) {
          unsigned char* plane = pv + (size_t)e * grid * 4;
          const float* col = rows + col_of_var[v] + e;
          for (size_t k = k0; k < k1; k++) put_be(plane + cell_pos_[k] * 4, col[k * row_stride]);
        }
        pv += planes * grid * 4;
      }
    }
    if (fseeko(f_, (off_t)(rec_begin_ + nrec_ * rec_.size()), SEEK_SET) != 0 || fwrite(rec_.data(), 1, rec_.size(), f_) != rec_.size())
      throw std::runtime_error("short write");
    nrec_++;
  }
  uint64_t records() const { return nrec_; }
  void close() {
    if (!f_) return;
    unsigned char b[4];
    put_be(b, (uint32_t)nrec_);
    const bool ok = fseeko(f_, 4, SEEK_SET) == 0 && fwrite(b, 1, 4, f_) == 4;  // numrecs
    const bool closed = fclose(f_) == 0;
    f_ = nullptr;
    if (!ok || !closed) throw std::runtime_error("cannot finish the NetCDF file");
  }

 private:
  FILE* f_ = nullptr;
  int nlat_, nlon_, depth_;
  double time_step_;
  std::vector<OutVar> vars_;
  float fill_;
  std::vector<size_t> cell_pos_;
  std::vector<unsigned char> hdr_, rec_;
  uint64_t rec_begin_ = 0, nrec_ = 0;

  template <class T>
  static void put_be(unsigned char* p, T v) {
    unsigned char b[sizeof(T)];
    memcpy(b, &v, sizeof(T));
    for (size_t i = 0; i < sizeof(T); i++) p[i] = b[sizeof(T) - 1 - i];
  }
  void u32(uint32_t v) {
    unsigned char b[4];
    put_be(b, v);
    hdr_.insert(hdr_.end(), b, b + 4);
  }
  void u64(uint64_t v) {
    unsigned char b[8];
    put_be(b, v);
    hdr_.insert(hdr_.end(), b, b + 8);
  }
  void pad() {
    while (hdr_.size() % 4) hdr_.push_back(0);
  }
  void name(const std::string& s) {
    u32((uint32_t)s.size());
    hdr_.insert(hdr_.end(), s.begin(), s.end());
    pad();
  }
  void att_text(const std::string& k, const std::string& v) {
    name(k);
    u32(2);  // NC_CHAR
    u32((uint32_t)v.size());
    hdr_.insert(hdr_.end(), v.begin(), v.end());
    pad();
  }
  void att_int(const std::string& k, int v) {
    name(k);
    u32(4);  // NC_INT
    u32(1);
    u32((uint32_t)v);
  }
  void att_float(const std::string& k, float v) {
    name(k);
    u32(5);  // NC_FLOAT
    u32(1);
    unsigned char b[4];
    put_be(b, v);
    hdr_.insert(hdr_.end(), b, b + 4);
  }
  struct VarDef {
    std::string name;
    std::vector<int> dims;
    int type;  // 5 float, 6 double
    std::vector<std::pair<std::string, std::string>> text;
    bool has_fill;
    uint64_t bytes;  // per record for record variables, whole variable otherwise
    size_t begin_at;  // position of the `begin` field in the header
  };

  void write_header(double lat0, double dlat, double lon0, double dlon, const std::string& time_units,
                    const std::vector<std::pair<std::string, std::string>>& gtext, const std::vector<std::pair<std::string, int>>& gint) {
    enum { D_LAT = 0, D_LON = 1, D_BNDS = 2, D_TIME = 3, D_DEPTH = 4 };
    hdr_.clear();
    hdr_.insert(hdr_.end(), {'C', 'D', 'F', 2});
    u32(0);  // numrecs, filled in at close
    u32(0x0A);
    u32(5);
    name("lat"); u32((uint32_t)nlat_);
    name("lon"); u32((uint32_t)nlon_);
    name("bnds"); u32(2);
    name("time"); u32(0);  // the record dimension
    name("depth"); u32((uint32_t)depth_);
    if (gtext.empty() && gint.empty()) {
      u32(0); u32(0);
    } else {
      u32(0x0C);
      u32((uint32_t)(gtext.size() + gint.size()));
      for (auto& a : gtext) att_text(a.first, a.second);
      for (auto& a : gint) att_int(a.first, a.second);
    }
    const uint64_t grid = (uint64_t)nlat_ * nlon_;
    std::vector<VarDef> defs;
    defs.push_back({"lat", {D_LAT}, 6, {{"axis", "Y"}, {"units", "degrees_north"}, {"standard name", "latitude"}, {"long name", "latitude"}, {"bounds", "lat_bnds"}}, false, (uint64_t)nlat_ * 8, 0});
    defs.push_back({"lon", {D_LON}, 6, {{"axis", "X"}, {"units", "degrees_east"}, {"standard name", "longitude"}, {"long name", "longitude"}, {"bounds", "lon_bnds"}}, false, (uint64_t)nlon_ * 8, 0});
    defs.push_back({"depth", {D_DEPTH}, 5, {{"standard name", "z_dim"}, {"long name", "array values"}, {"units", "z_dim"}}, false, (uint64_t)depth_ * 4, 0});
    defs.push_back({"time", {D_TIME}, 5, {{"axis", "T"}, {"standard name", "time"}, {"long name", "time"}, {"units", time_units}, {"bounds", "time_bnds"}, {"calendar", "gregorian"}}, false, 4, 0});
    for (const OutVar& v : vars_) {
      VarDef d{v.name, v.nelem > 1 ? std::vector<int>{D_TIME, D_DEPTH, D_LAT, D_LON} : std::vector<int>{D_TIME, D_LAT, D_LON}, 5, v.text_atts, true,
               (v.nelem > 1 ? (uint64_t)depth_ : 1) * grid * 4, 0};
      defs.push_back(d);
    }
    u32(0x0B);
    u32((uint32_t)defs.size());
    for (VarDef& d : defs) {
      name(d.name);
      u32((uint32_t)d.dims.size());
      for (int k : d.dims) u32((uint32_t)k);
      u32(0x0C);
      u32((uint32_t)(d.text.size() + (d.has_fill ? 1 : 0)));
      for (auto& a : d.text) att_text(a.first, a.second);
      if (d.has_fill) att_float("_FillValue", fill_);
      u32((uint32_t)d.type);
      u32((uint32_t)std::min<uint64_t>(d.bytes, 0xFFFFFFFFull));  // vsize (a multiple of 4 for every variable here)
      d.begin_at = hdr_.size();
      u64(0);
    }
    // data: the three fixed coordinate variables, then the records (time, then every output variable)
    uint64_t pos = hdr_.size();
    uint64_t rec_bytes = 0;
    for (VarDef& d : defs) {
      const bool record = d.dims[0] == D_TIME;
      if (record) continue;
      put_be(&hdr_[d.begin_at], pos);
      pos += d.bytes;
    }
    rec_begin_ = pos;
    for (VarDef& d : defs) {
      if (d.dims[0] != D_TIME) continue;
      put_be(&hdr_[d.begin_at], rec_begin_ + rec_bytes);
      rec_bytes += d.bytes;
    }
    rec_.assign(rec_bytes, 0);
    for (size_t i = 4; i + 4 <= rec_bytes; i += 4) put_be(&rec_[i], fill_);  // after the 4 bytes of the time coordinate: grids of fill values
    std::vector<unsigned char> fixed;
    auto push_d = [&](double v) { unsigned char b[8]; put_be(b, v); fixed.insert(fixed.end(), b, b + 8); };
    auto push_f = [&](float v) { unsigned char b[4]; put_be(b, v); fixed.insert(fixed.end(), b, b + 4); };
    for (int i = 0; i < nlat_; i++) push_d(lat0 + (i * dlat));  // WriteOutputNetCDF.c:204
    for (int i = 0; i < nlon_; i++) push_d(lon0 + (i * dlon));  // :217
    for (int i = 0; i < depth_; i++) push_f((float)i);          // :249-252
    if (fwrite(hdr_.data(), 1, hdr_.size(), f_) != hdr_.size() || fwrite(fixed.data(), 1, fixed.size(), f_) != fixed.size()) throw std::runtime_error("short write");
  }
};

}  // namespace vicgpu_nc
#endif

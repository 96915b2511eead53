// oracle/casefile.h -- tiny tagged-array container used to move test cases between the
// reference harness (C++), the host-compiled port (C++) and the Python tests.
// TEST INFRASTRUCTURE ONLY.
//
// File = magic "VICCASE1" followed by records:
//   char name[32]; int32 dtype (0 = float64, 1 = int32); int32 ndim; int64 dims[4]; raw data
#ifndef VIC_ORACLE_CASEFILE_H
#define VIC_ORACLE_CASEFILE_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

struct CaseArray {
  int dtype;
  std::vector<int64_t> dims;
  std::vector<double> f64;
  std::vector<int32_t> i32;
  size_t count() const { size_t n = 1; for (size_t i = 0; i < dims.size(); i++) n *= (size_t)dims[i]; return n; }
};

class CaseWriter {
 public:
  explicit CaseWriter(const char *path) { fp = fopen(path, "wb"); if (fp) fwrite("VICCASE1", 1, 8, fp); }
  ~CaseWriter() { if (fp) fclose(fp); }
  bool ok() const { return fp != NULL; }
  void header(const char *name, int dtype, int ndim, const int64_t *dims) {
    char nm[32]; memset(nm, 0, sizeof(nm)); strncpy(nm, name, 31);
    int32_t dt = dtype, nd = ndim; int64_t d4[4] = {1, 1, 1, 1};
    for (int i = 0; i < ndim; i++) d4[i] = dims[i];
    fwrite(nm, 1, 32, fp); fwrite(&dt, 4, 1, fp); fwrite(&nd, 4, 1, fp); fwrite(d4, 8, 4, fp);
  }
  void f64(const char *name, const double *data, int ndim, const int64_t *dims) {
    header(name, 0, ndim, dims); size_t n = 1; for (int i = 0; i < ndim; i++) n *= (size_t)dims[i];
    fwrite(data, 8, n, fp);
  }
  void i32(const char *name, const int32_t *data, int ndim, const int64_t *dims) {
    header(name, 1, ndim, dims); size_t n = 1; for (int i = 0; i < ndim; i++) n *= (size_t)dims[i];
    fwrite(data, 4, n, fp);
  }
  // streaming variant: header now, rows appended later with raw()
  void raw(const void *data, size_t bytes) { fwrite(data, 1, bytes, fp); }
 private:
  FILE *fp;
};

inline bool case_read(const char *path, std::map<std::string, CaseArray> &out) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return false;
  char magic[8];
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "VICCASE1", 8) != 0) { fclose(fp); return false; }
  for (;;) {
    char nm[32]; int32_t dt, nd; int64_t d4[4];
    if (fread(nm, 1, 32, fp) != 32) break;
    if (fread(&dt, 4, 1, fp) != 1 || fread(&nd, 4, 1, fp) != 1 || fread(d4, 8, 4, fp) != 4) { fclose(fp); return false; }
    CaseArray a; a.dtype = dt; a.dims.assign(d4, d4 + nd);
    size_t n = a.count();
    if (dt == 0) { a.f64.resize(n); if (fread(a.f64.data(), 8, n, fp) != n) { fclose(fp); return false; } }
    else { a.i32.resize(n); if (fread(a.i32.data(), 4, n, fp) != n) { fclose(fp); return false; } }
    nm[31] = 0;
    out[std::string(nm)] = a;
  }
  fclose(fp);
  return true;
}
#endif

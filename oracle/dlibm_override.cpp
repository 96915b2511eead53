// oracle/dlibm_override.cpp -- TEST INFRASTRUCTURE ONLY.
// Linked into oracle/_ref/vic_ref_harness_dl: the reference's own objects then resolve exp / log / log10 / pow /
// sin / cos / sincos / acos (the complete list of libm entry points the reference imports, `nm -u`) to the portable
// implementations of vic_b200/csrc/vic_math.cuh instead of glibc's.  The reference's algorithm is untouched; only
// the platform's elementary functions are replaced by the ones the CUDA library uses, which makes the reference's
// answers reproducible bit for bit on the GPU (see the header of vic_math.cuh for why that matters).
#include "vic_math.cuh"
extern "C" {
double exp(double x) { return vic::dl::exp(x); }
double log(double x) { return vic::dl::log(x); }
double log10(double x) { return vic::dl::log10(x); }
double pow(double x, double y) { return vic::dl::pow(x, y); }
double sin(double x) { return vic::dl::sin(x); }
double cos(double x) { return vic::dl::cos(x); }
double acos(double x) { return vic::dl::acos(x); }
void sincos(double x, double *s, double *c) { *s = vic::dl::sin(x); *c = vic::dl::cos(x); }
}

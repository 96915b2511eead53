// vic_common.cuh -- constants, math wrappers and small helpers shared by every physics
// header of the vic-b200 step kernels.
//
// All physics headers are written once and compiled twice:
//   * by nvcc for sm_100a into libvicgpu.so (the product; vicgpu_kernels.cu), and
//   * by g++ into oracle/_ref/libvicport.so (the CPU restatement used ONLY by tests/ to
//     check the algorithm against the reference build without a GPU).
// FP64 throughout; no fast-math: NaN is the reference's INVALID sentinel
// (vicNl_def.h:150-159) and its tests (IS_VALID / IS_INVALID) must keep working.
#ifndef VIC_COMMON_CUH
#define VIC_COMMON_CUH

#include <math.h>
#include <limits.h>
#include "vicgpu.h"
#include "vic_glibm.cuh"

// VIC_HD : small leaf relations, always inlined.
// VIC_HDI: the larger routines and the residual functors' operator(): real calls (one copy of
//          each in the kernel), because the Brent solver evaluates a residual from a dozen call
//          sites and residuals nest (surface -> soil profile -> per-node solve); inlining all of
//          that multiplies code size and compile time by orders of magnitude.
#if defined(__CUDACC__)
#define VIC_HD __host__ __device__ __forceinline__
#define VIC_HDI inline __host__ __device__ __noinline__
#else
#define VIC_HD inline
#define VIC_HDI inline
#endif

namespace vic {

// ---- model constants (vicNl_def.h:138-302, snow.h:33-79) -------------------------------
constexpr double HUGE_RESIST = 1.e20;
constexpr double SMALL = 1.e-12;
constexpr int ERROR_I = -999;
constexpr double ERROR_D = -999.0;
constexpr int INVALID_INT = INT_MIN;
constexpr double ice_density = 917.0;
constexpr double von_K = 0.40;
constexpr double KELVIN = 273.15;
constexpr double STEFAN_B = 5.6696e-8;
constexpr double Lf = 3.337e5;
constexpr double RHO_W = 999.842594;
constexpr double Cp = 1013.0;
constexpr double CH_ICE = 2100.0e3;
constexpr double CH_WATER = 4186.8e3;
constexpr double K_SNOW = 2.9302e-6;
constexpr double EPS = 0.62196351;
constexpr double G_GRAV = 9.81;
constexpr double JOULESPCAL = 4.1868;
constexpr double GRAMSPKG = 1000.0;
constexpr double SEC_PER_DAY = 86400.;
constexpr int SECPHOUR = 3600;
constexpr double GLAC_TEMP = 0.0;
constexpr double GLAC_K_ICE = 2.14;
constexpr double SNOW_SURF_DENSITY = 350;
constexpr double CUTOFF_DENSITY = 830;
constexpr double A_SVP = 0.61078;
constexpr double B_SVP = 17.269;
constexpr double C_SVP = 237.3;
constexpr double CP_PM = 1013;
constexpr double PS_PM = 101300;
constexpr double LAPSE_PM = -0.006;
constexpr double SNOW_DT = 5.0;
constexpr double SURF_DT = 1.0;
constexpr double SOIL_DT = 0.25;
constexpr double CANOPY_DT = 1.0;
// snow.h
constexpr double LIQUID_WATER_CAPACITY = 0.035;
constexpr double LAI_SNOW_MULTIPLIER = 0.0005;
constexpr double MIN_INTERCEPTION_STORAGE = 0.005;
constexpr double MAX_SURFACE_SWE = 0.125;
constexpr double NEW_SNOW_DENSITY = 50.;
constexpr double SNDENS_DMLIMIT = 100.;
constexpr double SNDENS_ETA0 = 3.6e6;
constexpr double SNDENS_C1 = 0.04;
constexpr double SNDENS_C2 = 2.778e-6;
constexpr double SNDENS_C5 = 0.08;
constexpr double SNDENS_C6 = 0.021;
constexpr double SNDENS_F = 0.6;
constexpr double MIN_SWQ_EB_THRES = 0.0010;
constexpr double TraceSnow = 0.03;

// option codes (vicNl_def.h:166-214)
enum { AR_406 = 0, AR_406_LS, AR_406_FULL, AR_410, AR_COMBO };
enum { GF_406 = 0, GF_410, GF_FULL };
enum { USACE = 0, SUN1999 };
enum { DENS_BRAS = 0, DENS_SNTHRM };
enum { VIC_412 = 0, KIENZLE };
enum { N_PET_TYPES = 6, N_PET_TYPES_NON_NAT = 4, PET_VEGNOCR = 5 };

// surface types of the aerodynamic tables (VegConditions.h)
enum Surf { SNOW_FREE = 0, CANOPY_OVER = 1, SNOW_COVERED = 2, GLACIER_SURF = 3, SURF_UNSET = 4 };

// Elementary functions: vic_glibm.cuh, the operation-by-operation restatement of glibc 2.39's exp/log/log10/pow/sin/cos/acos
// (bit-identical to the libm the reference is linked against, on the device and on the host).
// -DVIC_USE_LIBM swaps in the platform's libm itself; it exists only so that the host port can be run both ways and the two
// compared (oracle/Makefile builds both flavours; tests/test_cpu.py); libvicgpu.so never uses it.
#if defined(VIC_USE_LIBM) && !defined(__CUDACC__)
inline double vpow(double a, double b) { return pow(a, b); }
inline double vexp(double a) { return exp(a); }
inline double vlog(double a) { return log(a); }
inline double vlog10(double a) { return log10(a); }
inline double vsin(double a) { return sin(a); }
inline double vcos(double a) { return cos(a); }
inline double vacos(double a) { return acos(a); }
#else
VIC_HD double vpow(double a, double b) { return gl::pow(a, b); }
VIC_HD double vexp(double a) { return gl::exp(a); }
VIC_HD double vlog(double a) { return gl::log(a); }
VIC_HD double vlog10(double a) { return gl::log10(a); }
VIC_HD double vsin(double a) { return gl::sin(a); }
VIC_HD double vcos(double a) { return gl::cos(a); }
VIC_HD double vacos(double a) { return gl::acos(a); }
#endif

// a / b where b is known to be positive and finite (a time step, a density, a count of sub-steps, a resistance).  The device's
// IEEE double division branches to a ~70-instruction subroutine when the numerator is zero or tiny (|a| < 2^-967); 0 / b is the
// numerator itself (sign included), so that case never reaches the divider: it is handed 1.0 instead and the quotient discarded
// (a plain `a == 0 ? a : a / b` does not help -- the compiler evaluates the division speculatively and selects afterwards).
VIC_HD double div_pos(double a, double b) {
  const bool z = (a == 0.0);
  double n = z ? 1.0 : a;
#if defined(__CUDA_ARCH__)
  // opaque to the optimiser: otherwise it proves that n == a on the path that uses the quotient and divides a itself again
  asm volatile("" : "+d"(n));
#endif
  const double q = n / b;
  return z ? a : q;
}

// a / b for any b, exactly, when the numerator is often an exact zero (dry canopy, no ice, no rain): +-0 / b is +-0 with the sign of
// the operands for every b except 0 and NaN, which go through the divider as they are
VIC_HD double div_zn(double a, double b) {
#if defined(__CUDA_ARCH__)
  const bool z = (a == 0.0) && (b > 0.0 || b < 0.0);
  double n = z ? 1.0 : a;
  asm volatile("" : "+d"(n));
  const double q = n / b;
  return z ? (b > 0.0 ? a : -a) : q;
#else
  return a / b;
#endif
}

VIC_HD double vnan() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return NAN;
#endif
}
VIC_HD bool is_invalid(double a) { return a != a; }
VIC_HD bool is_valid(double a) { return a == a; }
VIC_HD double vmin(double a, double b) { return (b < a) ? b : a; }   // std::min semantics
VIC_HD double vmax(double a, double b) { return (a < b) ? b : a; }   // std::max semantics
VIC_HD bool result_is_error(double r) { return r <= -998; }           // root_brent.h:11

// Four quantities per surface type (snow-free ground / canopy / snow / glacier ice).
struct Surf4 {
  double v[4];
  VIC_HD double& operator[](int i) { return v[i]; }
  VIC_HD const double& operator[](int i) const { return v[i]; }
  VIC_HD void set_invalid() { v[0] = v[1] = v[2] = v[3] = vnan(); }
};

struct RaUsed {  // AeroResistUsed, vicNl_def.h:602-606
  double surface, overstory;
};

}  // namespace vic
#endif

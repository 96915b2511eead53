// oracle/icemeltcheck.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's own ice_melt() (ice_melt.c, with IceEnergyBalance.c, lakes.eb.c icerad, root_brent.c, latent_heat_from_snow.c,
// StabilityCorrection.c, svp.c: objects compiled from /root/reference by oracle/Makefile) against the host build of the product's
// restatement (vic_b200/csrc/vic_lakeice.cuh through oracle/icemeltport.cpp) on seeded random lake-ice columns: every output of
// every column is compared bit for bit (NaN, the reference's INVALID, equals NaN).  With -o the inputs and the REFERENCE's outputs
// are written as a case file: the golden vectors of the GPU operator test (tests/golden/make_ice_melt_golden.py).
// Usage: icemeltcheck [-n N] [--seed S] [--dt HOURS] [--tfallback 0|1] [-o case.bin]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "vicNl.h"
#include "casefile.h"
#include "vicgpu.h"

extern "C" void port_ice_melt(int n, int delta_t, int tfallback, const double* in, double* out);

int main(int argc, char** argv) {
  int n = 20000, dt = 1, tfallback = 1;
  unsigned seed = 1;
  const char* out_path = NULL;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-n" && i + 1 < argc) n = atoi(argv[++i]);
    else if (a == "--seed" && i + 1 < argc) seed = (unsigned)atoi(argv[++i]);
    else if (a == "--dt" && i + 1 < argc) dt = atoi(argv[++i]);
    else if (a == "--tfallback" && i + 1 < argc) tfallback = atoi(argv[++i]);
    else if (a == "-o" && i + 1 < argc) out_path = argv[++i];
    else { fprintf(stderr, "icemeltcheck: bad argument\n"); return 2; }
  }
  if (!freopen("/dev/null", "w", stderr)) {}
  ProgramState state;
  state.initialize_global();
  state.options.BLOWING = FALSE;
  state.options.TFALLBACK = tfallback ? TRUE : FALSE;

  std::mt19937_64 rng(seed);
  auto U = [&](double lo, double hi) { return lo + (hi - lo) * std::uniform_real_distribution<double>(0.0, 1.0)(rng); };
  std::vector<double> in((size_t)n * VICGPU_ICE_NIN), ref((size_t)n * VICGPU_ICE_NOUT), port((size_t)n * VICGPU_ICE_NOUT);
  for (int i = 0; i < n; i++) {
    double* a = &in[(size_t)i * VICGPU_ICE_NIN];
    const double tair = U(0, 1) < 0.25 ? U(-2, 8) : U(-35, 1);
    a[ICEIN_z2] = U(2, 10);
    a[ICEIN_aero_resist] = U(20, 300);
    a[ICEIN_latent_heat_Le] = (2.501 - 0.002361 * tair) * 1.0e6;
    a[ICEIN_Z0] = U(0.001, 0.03);
    a[ICEIN_rainfall] = (tair > 0 && U(0, 1) < 0.4) ? U(0, 5) : 0.0;
    a[ICEIN_snowfall] = (tair <= 1 && U(0, 1) < 0.4) ? (U(0, 1) < 0.1 ? U(100, 200) : U(0, 10)) : 0.0;
    a[ICEIN_wind] = U(0, 1) < 0.05 ? 0.0 : U(0.1, 15);
    a[ICEIN_Tcutoff] = U(0, 1) < 0.8 ? 0.0 : U(-0.5, 0.5);
    a[ICEIN_air_temp] = tair;
    a[ICEIN_net_short] = U(0, 1) < 0.4 ? 0.0 : U(0, 400);
    a[ICEIN_longwave] = U(150, 350);
    a[ICEIN_density] = U(1.1, 1.45);
    a[ICEIN_pressure] = U(70, 102);
    a[ICEIN_vp] = U(0.05, 1.2);
    a[ICEIN_vpd] = U(0, 1) < 0.15 ? 0.0 : U(0, 0.8);
    const double swq = U(0, 1) < 0.3 ? 0.0 : (U(0, 1) < 0.3 ? U(0, 0.004) : U(0, 0.6));
    a[ICEIN_surf_water] = swq * (U(0, 1) < 0.5 ? 0.0 : U(0, 0.03));
    a[ICEIN_pack_water] = swq * (U(0, 1) < 0.5 ? 0.0 : U(0, 0.03));
    a[ICEIN_swq] = swq;
    a[ICEIN_surf_temp] = U(0, 1) < 0.2 ? 0.0 : U(-25, 0);
    a[ICEIN_pack_temp] = U(0, 1) < 0.2 ? 0.0 : U(-15, 0);
    a[ICEIN_vapor_flux] = U(-1e-5, 1e-5);
    a[ICEIN_surface_flux] = a[ICEIN_vapor_flux];
    a[ICEIN_surf_temp_fbflag] = 0;
    a[ICEIN_surf_temp_fbcount] = (double)(int)U(0, 4);
    a[ICEIN_areai] = U(1e4, 1e8);
    a[ICEIN_hice] = U(0, 1) < 0.2 ? U(0.001, 0.02) : U(0.02, 1.5);
    a[ICEIN_ice_water_eq] = a[ICEIN_hice] * a[ICEIN_areai] * 0.917 * U(0.5, 1.0);
    a[ICEIN_volume] = U(1e6, 1e10);
  }
  const auto t_ref0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) {
    const double* a = &in[(size_t)i * VICGPU_ICE_NIN];
    double* o = &ref[(size_t)i * VICGPU_ICE_NOUT];
    snow_data_struct snow;
    memset(&snow, 0, sizeof(snow));
    lake_var_struct lake;
    memset(&lake, 0, sizeof(lake));
    snow.swq = a[ICEIN_swq]; snow.surf_temp = a[ICEIN_surf_temp]; snow.pack_temp = a[ICEIN_pack_temp]; snow.pack_water = a[ICEIN_pack_water];
    snow.surf_water = a[ICEIN_surf_water]; snow.vapor_flux = a[ICEIN_vapor_flux]; snow.surface_flux = a[ICEIN_surface_flux];
    snow.surf_temp_fbflag = (char)a[ICEIN_surf_temp_fbflag]; snow.surf_temp_fbcount = (int)a[ICEIN_surf_temp_fbcount];
    lake.ice_water_eq = a[ICEIN_ice_water_eq]; lake.areai = a[ICEIN_areai]; lake.hice = a[ICEIN_hice]; lake.volume = a[ICEIN_volume];
    double ra_used = 0, melt = 0, adv = 0, dcc = 0, sflux = 0, lat = 0, sens = 0, qnet = 0, refr = 0, lwnet = 0;
    const int rc = ice_melt(a[ICEIN_z2], a[ICEIN_aero_resist], &ra_used, a[ICEIN_latent_heat_Le], &snow, &lake, dt, 0.0, a[ICEIN_Z0], 1.0, a[ICEIN_rainfall],
                            a[ICEIN_snowfall], a[ICEIN_wind], a[ICEIN_Tcutoff], a[ICEIN_air_temp], a[ICEIN_net_short], a[ICEIN_longwave], a[ICEIN_density],
                            a[ICEIN_pressure], a[ICEIN_vpd], a[ICEIN_vp], &melt, &adv, &dcc, &sflux, &lat, &sens, &qnet, &refr, &lwnet, 0.0, &state);
    o[ICEOUT_rc] = rc; o[ICEOUT_aero_resist_used] = ra_used; o[ICEOUT_melt] = melt; o[ICEOUT_advection] = adv; o[ICEOUT_deltaCC] = dcc; o[ICEOUT_SnowFlux] = sflux;
    o[ICEOUT_latent] = lat; o[ICEOUT_sensible] = sens; o[ICEOUT_Qnet] = qnet; o[ICEOUT_refreeze_energy] = refr; o[ICEOUT_LWnet] = lwnet;
    o[ICEOUT_swq] = snow.swq; o[ICEOUT_surf_temp] = snow.surf_temp; o[ICEOUT_pack_temp] = snow.pack_temp; o[ICEOUT_pack_water] = snow.pack_water;
    o[ICEOUT_surf_water] = snow.surf_water; o[ICEOUT_vapor_flux] = snow.vapor_flux; o[ICEOUT_blowing_flux] = snow.blowing_flux;
    o[ICEOUT_surface_flux] = snow.surface_flux; o[ICEOUT_surf_temp_fbflag] = snow.surf_temp_fbflag; o[ICEOUT_surf_temp_fbcount] = snow.surf_temp_fbcount;
    o[ICEOUT_coverage] = snow.coverage; o[ICEOUT_mass_error] = snow.mass_error; o[ICEOUT_coldcontent] = snow.coldcontent;
    o[ICEOUT_ice_water_eq] = lake.ice_water_eq; o[ICEOUT_volume] = lake.volume;
  }
  const double t_ref = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ref0).count();
  const auto t_port0 = std::chrono::steady_clock::now();
  port_ice_melt(n, dt, tfallback, in.data(), port.data());
  const double t_port = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_port0).count();

  static const char* names[] = {
#define X(nm) #nm,
      VICGPU_ICE_OUT(X)
#undef X
  };
  long bad = 0, errors = 0, solved = 0, invalid = 0, melting = 0;
  std::vector<long> badcol(VICGPU_ICE_NOUT, 0);
  for (int i = 0; i < n; i++) {
    const double *r = &ref[(size_t)i * VICGPU_ICE_NOUT], *p = &port[(size_t)i * VICGPU_ICE_NOUT];
    if (r[ICEOUT_rc] != 0) {
      errors++;
      if (p[ICEOUT_rc] != r[ICEOUT_rc]) { bad++; badcol[ICEOUT_rc]++; }
      continue;  // outputs undefined after an ERROR return
    }
    if (r[ICEOUT_surf_temp] != r[ICEOUT_surf_temp]) invalid++;
    else if (r[ICEOUT_Qnet] == 0.0) melting++;
    else solved++;
    bool rowbad = false;
    for (int k = 0; k < VICGPU_ICE_NOUT; k++) {
      const bool same = (r[k] != r[k] && p[k] != p[k]) || memcmp(&r[k], &p[k], 8) == 0 || (r[k] == 0.0 && p[k] == 0.0);
      if (!same) { rowbad = true; badcol[k]++; }
    }
    if (rowbad) {
      if (bad < 3) {
        printf("column %d differs:", i);
        for (int k = 0; k < VICGPU_ICE_NOUT; k++)
          if (memcmp(&r[k], &p[k], 8) != 0 && !(r[k] != r[k] && p[k] != p[k])) printf(" %s ref %.17g port %.17g;", names[k], r[k], p[k]);
        printf("\n");
      }
      bad++;
    }
  }
  printf("n %d dt %d tfallback %d: balance at 0 C %ld, surface solved %ld, thin pack (INVALID surface) %ld, ERROR returns %ld\n", n, dt, tfallback, melting, solved, invalid, errors);
  for (int k = 0; k < VICGPU_ICE_NOUT; k++)
    if (badcol[k]) printf("  %s: %ld rows differ\n", names[k], badcol[k]);
  printf("reference ice_melt() %.3f s = %.2f M columns/s on one host thread; host build of vic_lakeice.cuh %.3f s\n", t_ref, n / t_ref / 1e6, t_port);
  printf(bad ? "DIFFERENT (%ld)\n" : "identical\n", bad);
  if (out_path) {
    CaseWriter w(out_path);
    int64_t d[2] = {n, VICGPU_ICE_NIN};
    w.f64("in", in.data(), 2, d);
    d[1] = VICGPU_ICE_NOUT;
    w.f64("out_ref", ref.data(), 2, d);
    int32_t meta[2] = {dt, tfallback};
    int64_t dm[1] = {2};
    w.i32("meta", meta, 1, dm);
  }
  return bad ? 1 : 0;
}

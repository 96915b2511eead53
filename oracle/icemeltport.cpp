// oracle/icemeltport.cpp -- TEST INFRASTRUCTURE ONLY: host (g++) build of vic_b200/csrc/vic_lakeice.cuh behind a flat interface,
// in its own translation unit because the reference's headers #define names the port uses as constants (Lf, Cp, CH_ICE ...).
#include "vic_lakeice.cuh"

extern "C" void port_ice_melt(int n, int delta_t, int tfallback, const double* in, double* out) {
  for (int i = 0; i < n; i++) {
    const double* a = in + (size_t)i * VICGPU_ICE_NIN;
    double* o = out + (size_t)i * VICGPU_ICE_NOUT;
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
}

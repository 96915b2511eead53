// vic_aero.cuh -- aerodynamic resistances and wind profile for the four surface types
// (snow-free ground, canopy top, snow surface, glacier ice).
// Reproduces CalcAerodynamic() (CalcAerodynamic.c:64-272): logarithmic profile above the
// canopy, exponential profile inside it, logarithmic again in the trunk space.
//
// In/out convention of the reference (full_energy.c:302-354): displacement / ref_height /
// roughness / wind_speed are shared by the seven successive calls and overwritten by each;
// what the LAST call (current vegetation) leaves behind feeds surface_fluxes.
#ifndef VIC_AERO_CUH
#define VIC_AERO_CUH
#include "vic_common.cuh"

namespace vic {

// CalcAerodynamic() is a geometry part -- the profile of a unit wind over the land cover -- followed by a scaling with the wind at
// the reference height (CalcAerodynamic.c:240-270).  The geometry depends only on the land cover's monthly parameters and the
// cell's roughness lengths, not on the forcing or the state: calc_aerodynamic_geom() is evaluated once per HRU and month on the
// device (AeroGeom, vic_step.cuh) and calc_aerodynamic_wind() every record; calc_aerodynamic() is the two in sequence, the
// reference's operations in the reference's order.
// returns 0 or ERROR_I
VIC_HDI int calc_aerodynamic_geom(bool OverStory, double Height, double Trunk, double Z0_SNOW, double Z0_SOIL, double n,
                                  Surf4& aero_resist, Surf4& wind_speed, Surf4& displacement, Surf4& ref_height, Surf4& roughness) {
  double d_Lower, d_Upper, Uh, Ut, Uw, Z0_Lower, Z0_Upper, Zt, Zw;
  const double K2 = von_K * von_K;
  if (!OverStory) {
    Z0_Lower = roughness[SNOW_FREE];
    d_Lower = displacement[SNOW_FREE];
    const double l2 = vlog((2. + Z0_Lower) / Z0_Lower);
    const double lr = vlog((ref_height[SNOW_FREE] - d_Lower) / Z0_Lower);
    wind_speed[SNOW_FREE] = l2 / lr;
    aero_resist[SNOW_FREE] = l2 * lr / K2;
    ref_height[CANOPY_OVER] = ref_height[SNOW_FREE];
    roughness[CANOPY_OVER] = roughness[SNOW_FREE];
    displacement[CANOPY_OVER] = displacement[SNOW_FREE];
    wind_speed[CANOPY_OVER] = wind_speed[SNOW_FREE];
    aero_resist[CANOPY_OVER] = aero_resist[SNOW_FREE];
    ref_height[SNOW_COVERED] = ref_height[SNOW_FREE];
    roughness[SNOW_COVERED] = Z0_SNOW;
    displacement[SNOW_COVERED] = 0.;
    const double s2 = vlog((2. + Z0_SNOW) / Z0_SNOW);
    const double sr = vlog(ref_height[SNOW_COVERED] / Z0_SNOW);
    wind_speed[SNOW_COVERED] = s2 / sr;
    aero_resist[SNOW_COVERED] = s2 * sr / K2;
    ref_height[SNOW_COVERED] = 2. + Z0_SNOW;
    ref_height[GLACIER_SURF] = ref_height[SNOW_FREE];
    roughness[GLACIER_SURF] = Z0_Lower;
    displacement[GLACIER_SURF] = 0.;
    const double gr = vlog(ref_height[GLACIER_SURF] / Z0_Lower);
    wind_speed[GLACIER_SURF] = l2 / gr;
    aero_resist[GLACIER_SURF] = l2 * gr / K2;
    ref_height[GLACIER_SURF] = 2. + Z0_Lower;
  } else {
    Z0_Upper = roughness[SNOW_FREE];
    d_Upper = displacement[SNOW_FREE];
    Z0_Lower = Z0_SOIL;
    d_Lower = 0;
    Zw = 1.5 * Height - 0.5 * d_Upper;
    Zt = Trunk * Height;
    if (Zt < (Z0_Lower + d_Lower)) return ERROR_I;
    const double lru = vlog((ref_height[SNOW_FREE] - d_Upper) / Z0_Upper);
    aero_resist[CANOPY_OVER] = lru / K2 *
        (Height / (n * (Zw - d_Upper)) * (vexp(n * (1 - (d_Upper + Z0_Upper) / Height)) - 1) + (Zw - Height) / (Zw - d_Upper) +
         vlog((ref_height[SNOW_FREE] - d_Upper) / (Zw - d_Upper)));
    Uw = vlog((Zw - d_Upper) / Z0_Upper) / lru;
    Uh = Uw - (1 - (Height - d_Upper) / (Zw - d_Upper)) / lru;
    wind_speed[CANOPY_OVER] = Uh * vexp(n * ((Z0_Upper + d_Upper) / Height - 1.));
    Ut = Uh * vexp(n * (Zt / Height - 1.));
    const double l2 = vlog((2. + Z0_Lower) / Z0_Lower);
    const double lt = vlog(Zt / Z0_Lower);
    wind_speed[SNOW_FREE] = Ut * l2 / lt;
    aero_resist[SNOW_FREE] = l2 * lt / (K2 * Ut);
    if (Zt > (2. + Z0_SNOW)) {
      const double s2 = vlog((2. + Z0_SNOW) / Z0_SNOW), st = vlog(Zt / Z0_SNOW);
      wind_speed[SNOW_COVERED] = Ut * s2 / st;
      aero_resist[SNOW_COVERED] = s2 * st / (K2 * Ut);
    } else if (Height > (2. + Z0_SNOW)) {
      const double st = vlog(Zt / Z0_SNOW);
      wind_speed[SNOW_COVERED] = Uh * vexp(n * ((2. + Z0_SNOW) / Height - 1.));
      aero_resist[SNOW_COVERED] = st * st / (K2 * Ut) +
          Height * lru / (n * K2 * (Zw - d_Upper)) * (vexp(n * (1 - Zt / Height)) - vexp(n * (1 - (Z0_SNOW + 2.) / Height)));
    } else {
      const double st = vlog(Zt / Z0_SNOW);
      wind_speed[SNOW_COVERED] = Uh;
      aero_resist[SNOW_COVERED] = st * st / (K2 * Ut) + Height * lru / (n * K2 * (Zw - d_Upper)) * (vexp(n * (1 - Zt / Height)) - 1);
    }
    ref_height[CANOPY_OVER] = ref_height[SNOW_FREE];
    roughness[CANOPY_OVER] = roughness[SNOW_FREE];
    displacement[CANOPY_OVER] = displacement[SNOW_FREE];
    ref_height[SNOW_FREE] = 2. + Z0_Lower;
    roughness[SNOW_FREE] = Z0_Lower;
    displacement[SNOW_FREE] = d_Lower;
    ref_height[SNOW_COVERED] = 2. + Z0_SNOW;
    roughness[SNOW_COVERED] = Z0_SNOW;
    displacement[SNOW_COVERED] = 0.;
    ref_height[GLACIER_SURF] = 2. + Z0_Lower;
    roughness[GLACIER_SURF] = Z0_Lower;
    displacement[GLACIER_SURF] = 0.;
  }
  return 0;
}

// the wind at the reference height scales the unit-wind profile (CalcAerodynamic.c:240-270)
VIC_HD void calc_aerodynamic_wind(double tmp_wind, Surf4& aero_resist, Surf4& wind_speed) {
  if (tmp_wind > 0.) {
    wind_speed[SNOW_FREE] *= tmp_wind;
    aero_resist[SNOW_FREE] /= tmp_wind;
    if (is_valid(wind_speed[CANOPY_OVER])) { wind_speed[CANOPY_OVER] *= tmp_wind; aero_resist[CANOPY_OVER] /= tmp_wind; }
    if (is_valid(wind_speed[SNOW_COVERED])) { wind_speed[SNOW_COVERED] *= tmp_wind; aero_resist[SNOW_COVERED] /= tmp_wind; }
    if (is_valid(wind_speed[GLACIER_SURF])) { wind_speed[GLACIER_SURF] *= tmp_wind; aero_resist[GLACIER_SURF] /= tmp_wind; }
  } else {
    wind_speed[SNOW_FREE] *= tmp_wind;
    aero_resist[SNOW_FREE] = HUGE_RESIST;
    if (is_valid(wind_speed[CANOPY_OVER])) wind_speed[CANOPY_OVER] *= tmp_wind;
    aero_resist[CANOPY_OVER] = HUGE_RESIST;
    if (is_valid(wind_speed[SNOW_COVERED])) wind_speed[SNOW_COVERED] *= tmp_wind;
    aero_resist[SNOW_COVERED] = HUGE_RESIST;
    if (is_valid(wind_speed[GLACIER_SURF])) wind_speed[GLACIER_SURF] *= tmp_wind;
    aero_resist[GLACIER_SURF] = HUGE_RESIST;
  }
}

VIC_HDI int calc_aerodynamic(bool OverStory, double Height, double Trunk, double Z0_SNOW, double Z0_SOIL, double n,
                             Surf4& aero_resist, Surf4& wind_speed, Surf4& displacement, Surf4& ref_height, Surf4& roughness) {
  const double tmp_wind = wind_speed[SNOW_FREE];
  if (calc_aerodynamic_geom(OverStory, Height, Trunk, Z0_SNOW, Z0_SOIL, n, aero_resist, wind_speed, displacement, ref_height, roughness) == ERROR_I)
    return ERROR_I;
  calc_aerodynamic_wind(tmp_wind, aero_resist, wind_speed);
  return 0;
}

}  // namespace vic
#endif

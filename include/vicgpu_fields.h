/* vicgpu_fields.h -- flat field tables of the vic-b200 C-ABI.
 *
 * The reference keeps one grid cell as a tree of C++ structs
 * (cell_info_struct -> soil_con_struct / std::vector<HRU> -> energy_bal_struct,
 * snow_data_struct, hru_data_struct, veg_var_struct, glac_data_struct;
 * /root/reference/vicNl_def.h:951-1100 soil_con, :1134-1209 energy, :1224-1258 snow,
 * :1114-1131 cell, :1215-1219 veg_var, :1342-1366 glacier, :1375-1389 HRU).
 * The GPU library never sees those types.  Across the C-ABI every cell is a row of
 * doubles ("cell parameter record") and every HRU (veg tile x snow band) is a row of
 * doubles ("HRU record").  The tables below name every column once; the second macro
 * argument is the reference member the column mirrors, so that the host-side packer
 * (vic_b200/host/vicgpu_pack.h, compiled against the reference's own headers) and the
 * device-side loader are generated from the same list and cannot drift apart.
 *
 * Integer members (counters, flags, last_snow) are carried as doubles: every value
 * that occurs (|x| <= 2^31) is exactly representable, so integer bookkeeping stays
 * bit-exact.
 *
 * Column classes (third argument):
 *   VG_C  carried between time steps (read and written by the step kernel)
 *   VG_D  diagnostic of the current step (written by the step kernel, read by the
 *         output reduction that stands in for put_data.c)
 *   VG_U  not touched on the hot path (kept so state files round-trip unchanged)
 */
#ifndef VICGPU_FIELDS_H
#define VICGPU_FIELDS_H

#define VICGPU_NLAYER 3       /* MAX_LAYERS, user_def.h:36 ; FULL_ENERGY requires exactly 3 */
#define VICGPU_MAX_NODES 32   /* device-side cap (reference MAX_NODES is 50, user_def.h) */
#define VICGPU_MAX_BANDS 30   /* MAX_BANDS, user_def.h */
#define VICGPU_NFRONTS 3      /* MAX_FRONTS */
#define VICGPU_NPET 6         /* N_PET_TYPES, vicNl_def.h:221 */
#define VICGPU_NZWT 11        /* MAX_ZWTVMOIST */

/* ------------------------------------------------------------------ HRU record */
/* scalars: X(column, reference member relative to HRU, class) */
/* Scalar columns, one list per sub-structure of the HRU.  X(P##name, reference member, class):
 * invoke with a prefix (E_, S_, C_, V_, G_) to get the column name, with an empty prefix to get the
 * bare member name used by the device-side structs (vic_b200/csrc/vic_types.cuh). */
#define VICGPU_HRU_ENERGY(X, P) \
  X(P##AlbedoLake, energy.AlbedoLake, VG_U) \
  X(P##AlbedoOver, energy.AlbedoOver, VG_C) \
  X(P##AlbedoUnder, energy.AlbedoUnder, VG_C) \
  X(P##Cs0, energy.Cs[0], VG_C) \
  X(P##Cs1, energy.Cs[1], VG_C) \
  X(P##frozen, energy.frozen, VG_C) \
  X(P##kappa0, energy.kappa[0], VG_C) \
  X(P##kappa1, energy.kappa[1], VG_C) \
  X(P##Nfrost, energy.Nfrost, VG_C) \
  X(P##Nthaw, energy.Nthaw, VG_C) \
  X(P##T1_index, energy.T1_index, VG_U) \
  X(P##Tcanopy, energy.Tcanopy, VG_C) \
  X(P##Tcanopy_fbflag, energy.Tcanopy_fbflag, VG_C) \
  X(P##Tcanopy_fbcount, energy.Tcanopy_fbcount, VG_C) \
  X(P##Tfoliage, energy.Tfoliage, VG_C) \
  X(P##Tfoliage_fbflag, energy.Tfoliage_fbflag, VG_C) \
  X(P##Tfoliage_fbcount, energy.Tfoliage_fbcount, VG_C) \
  X(P##Tsurf, energy.Tsurf, VG_C) \
  X(P##Tsurf_fbflag, energy.Tsurf_fbflag, VG_C) \
  X(P##Tsurf_fbcount, energy.Tsurf_fbcount, VG_C) \
  X(P##unfrozen, energy.unfrozen, VG_U) \
  X(P##advected_sensible, energy.advected_sensible, VG_C) \
  X(P##advection, energy.advection, VG_C) \
  X(P##AtmosError, energy.AtmosError, VG_C) \
  X(P##AtmosLatent, energy.AtmosLatent, VG_C) \
  X(P##AtmosLatentSub, energy.AtmosLatentSub, VG_C) \
  X(P##AtmosSensible, energy.AtmosSensible, VG_C) \
  X(P##canopy_advection, energy.canopy_advection, VG_C) \
  X(P##canopy_latent, energy.canopy_latent, VG_C) \
  X(P##canopy_latent_sub, energy.canopy_latent_sub, VG_C) \
  X(P##canopy_refreeze, energy.canopy_refreeze, VG_C) \
  X(P##canopy_sensible, energy.canopy_sensible, VG_C) \
  X(P##deltaCC, energy.deltaCC, VG_C) \
  X(P##deltaH, energy.deltaH, VG_C) \
  X(P##error, energy.error, VG_C) \
  X(P##fusion, energy.fusion, VG_C) \
  X(P##grnd_flux, energy.grnd_flux, VG_C) \
  X(P##latent, energy.latent, VG_C) \
  X(P##latent_sub, energy.latent_sub, VG_C) \
  X(P##longwave, energy.longwave, VG_C) \
  X(P##LongOverIn, energy.LongOverIn, VG_C) \
  X(P##LongUnderIn, energy.LongUnderIn, VG_C) \
  X(P##LongUnderOut, energy.LongUnderOut, VG_C) \
  X(P##melt_energy, energy.melt_energy, VG_C) \
  X(P##NetLongAtmos, energy.NetLongAtmos, VG_C) \
  X(P##NetLongOver, energy.NetLongOver, VG_C) \
  X(P##NetLongUnder, energy.NetLongUnder, VG_C) \
  X(P##NetShortAtmos, energy.NetShortAtmos, VG_C) \
  X(P##NetShortGrnd, energy.NetShortGrnd, VG_C) \
  X(P##NetShortOver, energy.NetShortOver, VG_C) \
  X(P##NetShortUnder, energy.NetShortUnder, VG_C) \
  X(P##out_long_canopy, energy.out_long_canopy, VG_C) \
  X(P##out_long_surface, energy.out_long_surface, VG_C) \
  X(P##refreeze_energy, energy.refreeze_energy, VG_C) \
  X(P##sensible, energy.sensible, VG_C) \
  X(P##shortwave, energy.shortwave, VG_C) \
  X(P##ShortOverIn, energy.ShortOverIn, VG_C) \
  X(P##ShortUnderIn, energy.ShortUnderIn, VG_C) \
  X(P##snow_flux, energy.snow_flux, VG_C) \
  X(P##glacier_flux, energy.glacier_flux, VG_C) \
  X(P##deltaCC_glac, energy.deltaCC_glac, VG_C) \
  X(P##glacier_melt_energy, energy.glacier_melt_energy, VG_C)

#define VICGPU_HRU_SNOW(X, P) \
  X(P##albedo, snow.albedo, VG_C) \
  X(P##canopy_albedo, snow.canopy_albedo, VG_C) \
  X(P##coldcontent, snow.coldcontent, VG_C) \
  X(P##coverage, snow.coverage, VG_C) \
  X(P##density, snow.density, VG_C) \
  X(P##depth, snow.depth, VG_C) \
  X(P##last_snow, snow.last_snow, VG_C) \
  X(P##max_swq, snow.max_swq, VG_C) \
  X(P##MELTING, snow.MELTING, VG_C) \
  X(P##pack_temp, snow.pack_temp, VG_C) \
  X(P##pack_water, snow.pack_water, VG_C) \
  X(P##snow, snow.snow, VG_C) \
  X(P##snow_canopy, snow.snow_canopy, VG_C) \
  X(P##store_coverage, snow.store_coverage, VG_C) \
  X(P##store_snow, snow.store_snow, VG_C) \
  X(P##store_swq, snow.store_swq, VG_C) \
  X(P##surf_temp, snow.surf_temp, VG_C) \
  X(P##surf_temp_fbcount, snow.surf_temp_fbcount, VG_C) \
  X(P##surf_temp_fbflag, snow.surf_temp_fbflag, VG_C) \
  X(P##surf_water, snow.surf_water, VG_C) \
  X(P##swq, snow.swq, VG_C) \
  X(P##swq_slope, snow.swq_slope, VG_C) \
  X(P##tmp_int_storage, snow.tmp_int_storage, VG_C) \
  X(P##blowing_flux, snow.blowing_flux, VG_C) \
  X(P##canopy_vapor_flux, snow.canopy_vapor_flux, VG_C) \
  X(P##mass_error, snow.mass_error, VG_C) \
  X(P##melt, snow.melt, VG_C) \
  X(P##Qnet, snow.Qnet, VG_C) \
  X(P##surface_flux, snow.surface_flux, VG_C) \
  X(P##transport, snow.transport, VG_C) \
  X(P##vapor_flux, snow.vapor_flux, VG_C)

#define VICGPU_HRU_CELL(X, P) \
  X(P##aero_surface, cell[0].aero_resist.surface, VG_C) \
  X(P##aero_overstory, cell[0].aero_resist.overstory, VG_C) \
  X(P##asat, cell[0].asat, VG_C) \
  X(P##baseflow, cell[0].baseflow, VG_C) \
  X(P##inflow, cell[0].inflow, VG_C) \
  X(P##excess_moist, cell[0].excess_moist, VG_C) \
  X(P##runoff, cell[0].runoff, VG_C) \
  X(P##rootmoist, cell[0].rootmoist, VG_C) \
  X(P##wetness, cell[0].wetness, VG_C) \
  X(P##zwt, cell[0].zwt, VG_C) \
  X(P##zwt2, cell[0].zwt2, VG_C) \
  X(P##zwt3, cell[0].zwt3, VG_C)

#define VICGPU_HRU_VEG(X, P) \
  X(P##canopyevap, veg_var[0].canopyevap, VG_C) \
  X(P##throughfall, veg_var[0].throughfall, VG_C) \
  X(P##Wdew, veg_var[0].Wdew, VG_C)

#define VICGPU_HRU_GLAC(X, P) \
  X(P##cold_content, glacier.cold_content, VG_C) \
  X(P##surf_temp, glacier.surf_temp, VG_C) \
  X(P##surf_temp_fbcount, glacier.surf_temp_fbcount, VG_C) \
  X(P##surf_temp_fbflag, glacier.surf_temp_fbflag, VG_C) \
  X(P##Qnet, glacier.Qnet, VG_C) \
  X(P##mass_balance, glacier.mass_balance, VG_C) \
  X(P##ice_mass_balance, glacier.ice_mass_balance, VG_C) \
  X(P##cum_mass_balance, glacier.cum_mass_balance, VG_C) \
  X(P##accumulation, glacier.accumulation, VG_C) \
  X(P##melt, glacier.melt, VG_C) \
  X(P##vapor_flux, glacier.vapor_flux, VG_C) \
  X(P##water_storage, glacier.water_storage, VG_C) \
  X(P##outflow, glacier.outflow, VG_C) \
  X(P##outflow_coef, glacier.outflow_coef, VG_C) \
  X(P##inflow, glacier.inflow, VG_C)

#define VICGPU_HRU_SCALARS(X) \
  VICGPU_HRU_ENERGY(X, E_) VICGPU_HRU_SNOW(X, S_) VICGPU_HRU_CELL(X, C_) VICGPU_HRU_VEG(X, V_) \
  VICGPU_HRU_GLAC(X, G_) X(H_mu, mu, VG_C)

/* per-soil-layer columns: index i in [0, VICGPU_NLAYER) */
#define VICGPU_HRU_LAYER(X, P) \
  X(P##Cs, cell[0].layer[i].Cs, VG_C) \
  X(P##T, cell[0].layer[i].T, VG_C) \
  X(P##evap, cell[0].layer[i].evap, VG_C) \
  X(P##soil_ice, cell[0].layer[i].soil_ice, VG_C) \
  X(P##kappa, cell[0].layer[i].kappa, VG_C) \
  X(P##moist, cell[0].layer[i].moist, VG_C) \
  X(P##phi, cell[0].layer[i].phi, VG_C) \
  X(P##zwt, cell[0].layer[i].zwt, VG_C)

/* freeze/thaw front columns: index i in [0, VICGPU_NFRONTS) */
#define VICGPU_HRU_FRONT(X, P) \
  X(P##fdepth, energy.fdepth[i], VG_C) \
  X(P##tdepth, energy.tdepth[i], VG_C)

/* potential-evaporation columns cell[0].pot_evap[i], i in [0, VICGPU_NPET): handled explicitly */

/* per-thermal-node columns: index i in [0, Nnode) */
#define VICGPU_HRU_NODE(X, P) \
  X(P##Cs, energy.Cs_node[i], VG_C) \
  X(P##ice, energy.ice_content[i], VG_C) \
  X(P##kappa, energy.kappa_node[i], VG_C) \
  X(P##moist, energy.moist[i], VG_C) \
  X(P##T, energy.T[i], VG_C) \
  X(P##T_fbflag, energy.T_fbflag[i], VG_C) \
  X(P##T_fbcount, energy.T_fbcount[i], VG_C)

/* ------------------------------------------------------------------ HRU parameters (constant) */
#define VICGPU_HPAR_SCALARS(X) \
  X(HP_cell, /* index of owning cell */, 0) \
  X(HP_Cv, veg_con.Cv, 0) \
  X(HP_root0, veg_con.root[0], 0) \
  X(HP_root1, veg_con.root[1], 0) \
  X(HP_root2, veg_con.root[2], 0) \
  X(HP_vegIndex, veg_con.vegIndex, 0) \
  X(HP_vegClass, veg_con.vegClass, 0) \
  X(HP_band, bandIndex, 0) \
  X(HP_isGlacier, isGlacier, 0) \
  X(HP_isArtBare, isArtificialBareSoil, 0) \
  X(HP_sigma_slope, veg_con.sigma_slope, 0) \
  X(HP_lag_one, veg_con.lag_one, 0) \
  X(HP_fetch, veg_con.fetch, 0)

/* ------------------------------------------------------------------ cell parameter record */
/* X(column, reference member relative to soil_con_struct) */
#define VICGPU_CPAR_SCALARS(X) \
  X(CP_FS_ACTIVE, FS_ACTIVE) \
  X(CP_Ds, Ds) \
  X(CP_Dsmax, Dsmax) \
  X(CP_Ws, Ws) \
  X(CP_c, c) \
  X(CP_b_infilt, b_infilt) \
  X(CP_dp, dp) \
  X(CP_avg_temp, avg_temp) \
  X(CP_rough, rough) \
  X(CP_snow_rough, snow_rough) \
  X(CP_elevation, elevation) \
  X(CP_lat, lat) \
  X(CP_lng, lng) \
  X(CP_time_zone_lng, time_zone_lng) \
  X(CP_annual_prec, annual_prec) \
  X(CP_avgJulyAirTemp, avgJulyAirTemp) \
  X(CP_max_infil, max_infil) \
  X(CP_cell_area, cell_area) \
  X(CP_slope, slope) \
  X(CP_aspect, aspect) \
  X(CP_ehoriz, ehoriz) \
  X(CP_whoriz, whoriz) \
  X(CP_NEW_SNOW_ALB, NEW_SNOW_ALB) \
  X(CP_SNOW_ALB_ACCUM_A, SNOW_ALB_ACCUM_A) \
  X(CP_SNOW_ALB_ACCUM_B, SNOW_ALB_ACCUM_B) \
  X(CP_SNOW_ALB_THAW_A, SNOW_ALB_THAW_A) \
  X(CP_SNOW_ALB_THAW_B, SNOW_ALB_THAW_B) \
  X(CP_MIN_RAIN_TEMP, MIN_RAIN_TEMP) \
  X(CP_MAX_SNOW_TEMP, MAX_SNOW_TEMP) \
  X(CP_PADJ_R, PADJ_R) \
  X(CP_PADJ_S, PADJ_S) \
  X(CP_T_LAPSE, T_LAPSE) \
  X(CP_PGRAD, PGRAD) \
  X(CP_GLAC_SURF_THICK, GLAC_SURF_THICK) \
  X(CP_GLAC_SURF_WE, GLAC_SURF_WE) \
  X(CP_GLAC_KMIN, GLAC_KMIN) \
  X(CP_GLAC_DK, GLAC_DK) \
  X(CP_GLAC_A, GLAC_A) \
  X(CP_GLAC_ALBEDO, GLAC_ALBEDO) \
  X(CP_GLAC_ROUGH, GLAC_ROUGH)

#define VICGPU_CPAR_LAYER(X) \
  X(CL_Ksat, Ksat[i]) \
  X(CL_Wcr, Wcr[i]) \
  X(CL_Wpwp, Wpwp[i]) \
  X(CL_expt, expt[i]) \
  X(CL_bubble, bubble[i]) \
  X(CL_bulk_density, bulk_density[i]) \
  X(CL_bulk_dens_min, bulk_dens_min[i]) \
  X(CL_soil_density, soil_density[i]) \
  X(CL_soil_dens_min, soil_dens_min[i]) \
  X(CL_organic, organic[i]) \
  X(CL_depth, depth[i]) \
  X(CL_max_moist, max_moist[i]) \
  X(CL_porosity, porosity[i]) \
  X(CL_quartz, quartz[i]) \
  X(CL_resid_moist, resid_moist[i]) \
  X(CL_init_moist, init_moist[i])

#define VICGPU_CPAR_NODE(X) \
  X(CN_alpha, alpha[i]) \
  X(CN_beta, beta[i]) \
  X(CN_gamma, gamma[i]) \
  X(CN_dz_node, dz_node[i]) \
  X(CN_Zsum_node, Zsum_node[i]) \
  X(CN_max_moist_node, max_moist_node[i]) \
  X(CN_expt_node, expt_node[i]) \
  X(CN_bubble_node, bubble_node[i])

/* zwt-vs-moisture tables: (VICGPU_NLAYER+2) curves x VICGPU_NZWT points, index i = curve*NZWT + point */
#define VICGPU_CPAR_ZWT(X) \
  X(CZ_zwt, zwtvmoist_zwt[i / VICGPU_NZWT][i % VICGPU_NZWT]) \
  X(CZ_moist, zwtvmoist_moist[i / VICGPU_NZWT][i % VICGPU_NZWT])

#define VICGPU_CPAR_BAND(X) \
  X(CB_AreaFract, AreaFract[i]) \
  X(CB_BandElev, BandElev[i]) \
  X(CB_Pfactor, Pfactor[i]) \
  X(CB_Tfactor, Tfactor[i]) \
  X(CB_AboveTreeLine, AboveTreeLine[i])

/* ------------------------------------------------------------------ vegetation library row */
#define VICGPU_VEGLIB_SCALARS(X) \
  X(VL_overstory, overstory) \
  X(VL_rad_atten, rad_atten) \
  X(VL_rarc, rarc) \
  X(VL_rmin, rmin) \
  X(VL_trunk_ratio, trunk_ratio) \
  X(VL_wind_atten, wind_atten) \
  X(VL_wind_h, wind_h) \
  X(VL_RGL, RGL) \
  X(VL_veg_class, veg_class)

#define VICGPU_VEGLIB_MONTHLY(X) \
  X(VM_LAI, LAI[i]) \
  X(VM_Wdmax, Wdmax[i]) \
  X(VM_albedo, albedo[i]) \
  X(VM_displacement, displacement[i]) \
  X(VM_roughness, roughness[i])

/* ------------------------------------------------------------------ forcing record (per cell, per model step) */
/* one value per sub-step slot; slots = NF+1 when NF>1 (index NF is the step mean), 1 when NF==1
 * (atmos_data_struct, vicNl_def.h:1061-1077; NR/NF get_global_param.c:969-973) */
#define VICGPU_FORCING(X) \
  X(FV_air_temp, air_temp) \
  X(FV_density, density) \
  X(FV_longwave, longwave) \
  X(FV_prec, prec) \
  X(FV_pressure, pressure) \
  X(FV_shortwave, shortwave) \
  X(FV_tskc, tskc) \
  X(FV_vp, vp) \
  X(FV_vpd, vpd) \
  X(FV_wind, wind) \
  X(FV_snowflag, snowflag)

#endif /* VICGPU_FIELDS_H */

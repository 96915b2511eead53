// vic_soil.cuh -- soil column of one HRU: variable-infiltration runoff, hourly Brooks-Corey
// drainage between layers and ARNO baseflow (runoff.c:7-813), water-table diagnostics
// (compute_zwt.c:45-112), redistribution of layer moisture to the thermal nodes
// (soil_conduction.c:304-440), layer temperature / ice from node temperatures
// (frozen_soil.c:12-103, soil_conduction.c:444-828) and the top-two-layer thermal
// properties prepared before the surface energy balance (prepare_full_energy.c:8-94).
#ifndef VIC_SOIL_CUH
#define VIC_SOIL_CUH
#include "vic_leaf.cuh"

namespace vic {

// derived per-(cell, layer) constants: [layer][VIC_NKPRE] of SoilKPre
#define VIC_NCELLDER (VICGPU_NLAYER * VIC_NKPRE)
VIC_HD SoilKPre cell_kpre(const CellPar& cp, int l) {
  if (cp.d.p) return SoilKPre{cp.d(l * VIC_NKPRE + 0), cp.d(l * VIC_NKPRE + 1), cp.d(l * VIC_NKPRE + 2), cp.d(l * VIC_NKPRE + 3)};
  return soil_k_pre(cp.layer(CL_soil_dens_min, l), cp.layer(CL_bulk_dens_min, l), cp.layer(CL_quartz, l), cp.layer(CL_soil_density, l),
                    cp.layer(CL_bulk_density, l), cp.layer(CL_organic, l));
}
// fills out[k * stride], k < VIC_NCELLDER, for the cell cp views (device: one thread per cell at vicgpu_set_cells; host port: a loop)
VIC_HD void derive_cell_constants(const CellPar& cp, double* out, size_t stride) {
  for (int l = 0; l < VICGPU_NLAYER; l++) {
    CellPar raw = cp;
    raw.d = Col{nullptr, 0};
    const SoilKPre p = cell_kpre(raw, l);
    out[(size_t)(l * VIC_NKPRE + 0) * stride] = p.Kdry;
    out[(size_t)(l * VIC_NKPRE + 1) * stride] = p.porosity;
    out[(size_t)(l * VIC_NKPRE + 2) * stride] = p.Ks_pow;
    out[(size_t)(l * VIC_NKPRE + 3) * stride] = p.Ksat_unfrozen;
  }
}

// runoff.c:773-813
VIC_HDI void compute_runoff_and_asat(const CellPar& cp, const double* moist, double inflow, double* A, double* runoff) {
  double top_moist = 0., top_max_moist = 0.;
  for (int l = 0; l < VICGPU_NLAYER - 1; l++) {
    top_moist += moist[l];
    top_max_moist += cp.layer(CL_max_moist, l);
  }
  if (top_moist > top_max_moist) top_moist = top_max_moist;
  const double b = cp(CP_b_infilt);
  double ex = b / (1.0 + b);
  *A = 1.0 - vpow((1.0 - top_moist / top_max_moist), ex);
  double max_infil = (1.0 + b) * top_max_moist;
  double i_0 = max_infil * (1.0 - vpow((1.0 - *A), (1.0 / b)));
  if (inflow == 0.0) *runoff = 0.0;
  else if (max_infil == 0.0) *runoff = inflow;
  else if ((i_0 + inflow) > max_infil) *runoff = inflow - top_max_moist + top_moist;
  else {
    double basis = 1.0 - (i_0 + inflow) / max_infil;
    *runoff = (inflow - top_max_moist + top_moist + top_max_moist * vpow(basis, 1.0 * (1.0 + b)));
  }
  if (*runoff < 0.) *runoff = 0.;
}

// compute_zwt.c:45-112
VIC_HDI void wrap_compute_zwt(const CellPar& cp, SoilCol& cell) {
  const int NL = VICGPU_NLAYER;
  double total_depth = 0;
  for (int l = 0; l < NL; l++) total_depth += cp.layer(CL_depth, l);
  for (int l = 0; l < NL; l++) cell.layer[l].zwt = compute_zwt(cp, l, cell.layer[l].moist);
  if (is_invalid(cell.layer[NL - 1].zwt)) cell.layer[NL - 1].zwt = -total_depth * 100;
  int l = NL - 1;
  double tmp_depth = total_depth;
  while (l >= 0 && cp.layer(CL_max_moist, l) - cell.layer[l].moist <= SMALL) {
    tmp_depth -= cp.layer(CL_depth, l);
    l--;
  }
  if (l < 0) cell.zwt = 0;
  else if (l < NL - 1) {
    if (is_valid(cell.layer[l].zwt)) cell.zwt = cell.layer[l].zwt;
    else cell.zwt = -tmp_depth * 100;
  } else cell.zwt = cell.layer[l].zwt;
  double tmp_moist = 0;
  for (int i = 0; i < NL - 1; i++) tmp_moist += cell.layer[i].moist;
  cell.zwt2 = compute_zwt(cp, NL, tmp_moist);
  if (is_invalid(cell.zwt2)) cell.zwt2 = cell.layer[NL - 1].zwt;
  tmp_moist = 0;
  for (int i = 0; i < NL; i++) tmp_moist += cell.layer[i].moist;
  cell.zwt3 = compute_zwt(cp, NL + 1, tmp_moist);
  if (is_invalid(cell.zwt3)) cell.zwt3 = -total_depth * 100;
}

// soil_conduction.c:304-440: node moisture, ice, conductivity and heat capacity from the
// layer moistures (moist[] in mm)
template <int NN>
VIC_HDI void distribute_node_moisture_properties(EnergyBal<NN>& energy, const CellPar& cp, const double* moist, const Opts& o) {
  const int NL = VICGPU_NLAYER;
  int lidx = 0;
  double Lsum = 0.;
  bool PAST_BOTTOM = false;
  const bool fs = (cp(CP_FS_ACTIVE) != 0.0) && o.FROZEN_SOIL;
  for (int n = 0; n < NN; n++) {
    if (n >= o.Nnode) break;
    const double zs = cp.node(CN_Zsum_node, n);
    const double d = cp.layer(CL_depth, lidx);
    if (zs == Lsum + d && n != 0 && lidx != NL - 1)
      energy.moist[n] = (moist[lidx] / d + moist[lidx + 1] / cp.layer(CL_depth, lidx + 1)) / 1000 / 2.;
    else
      energy.moist[n] = moist[lidx] / d / 1000;
    const double mmn = cp.node(CN_max_moist_node, n);
    if (energy.moist[n] - mmn > 0) energy.moist[n] = mmn;
    const double sd = cp.layer(CL_soil_density, lidx), bd = cp.layer(CL_bulk_density, lidx), org = cp.layer(CL_organic, lidx);
    const SoilKPre kp = cell_kpre(cp, lidx);
    if (energy.T[n] < 0 && fs) {
      energy.ice[n] = energy.moist[n] - maximum_unfrozen_water(energy.T[n], mmn, cp.node(CN_bubble_node, n), cp.node(CN_expt_node, n));
      if (energy.ice[n] < 0) energy.ice[n] = 0;
      energy.kappa_node[n] = soil_conductivity_pre(energy.moist[n], energy.moist[n] - energy.ice[n], kp);
    } else {
      energy.ice[n] = 0;
      energy.kappa_node[n] = soil_conductivity_pre(energy.moist[n], energy.moist[n], kp);
    }
    energy.Cs_node[n] = volumetric_heat_capacity(bd / sd, energy.moist[n] - energy.ice[n], energy.ice[n], org);
    if (zs > Lsum + d && !PAST_BOTTOM) {
      Lsum += d;
      lidx++;
      if (lidx == NL) {
        PAST_BOTTOM = true;
        lidx = NL - 1;
      }
    }
  }
}

// runoff.c:7-771 for the wet fraction (mu == 1, FROST_SUBAREAS == 1); ppt [mm] is the water
// reaching the soil surface during the model step.
template <int NN>
VIC_HDI int runoff(SoilCol& cell, EnergyBal<NN>& energy, const CellPar& cp, double ppt, const Opts& o) {
  const int NL = VICGPU_NLAYER;
  double resid_moist[NL], liq[NL], ice[NL], max_moist[NL], Ksat[NL], Q12[NL - 1], evap[NL], expt[NL], mm_tmp[NL];
  for (int i = 0; i < NL; i++) resid_moist[i] = cp.layer(CL_resid_moist, i) * cp.layer(CL_depth, i) * 1000.;
  cell.runoff = 0;
  cell.baseflow = 0;
  cell.asat = 0;
  double baseflow = 0, runoff_v, A;
  const int dt = o.dt;
  for (int l = 0; l < NL; l++) evap[l] = div_pos(cell.layer[l].evap, (double)dt);  // (dt >= 1)
  double inflow = ppt;
  for (int l = 0; l < NL; l++) {
    Ksat[l] = cp.layer(CL_Ksat, l) / 24.;
    liq[l] = cell.layer[l].moist - cell.layer[l].soil_ice;
    ice[l] = cell.layer[l].soil_ice;
    max_moist[l] = cp.layer(CL_max_moist, l);
    expt[l] = cp.layer(CL_expt, l);
  }
  for (int l = 0; l < NL; l++) mm_tmp[l] = (liq[l] + ice[l]);
  compute_runoff_and_asat(cp, mm_tmp, inflow, &A, &runoff_v);
  const double tmp_dt_runoff = div_pos(runoff_v, (double)dt);
  const double dt_inflow = div_pos(inflow, (double)dt);
  const double Dsmax = cp(CP_Dsmax) / 24.;
  const double Ds = cp(CP_Ds), Ws = cp(CP_Ws), c_exp = cp(CP_c);
  for (int time_step = 0; time_step < dt; time_step++) {
    inflow = dt_inflow;
    // drainage between layers (Brooks & Corey)
    for (int l = 0; l < NL - 1; l++) {
      double tmp_liq = liq[l] - evap[l];
      if (tmp_liq < resid_moist[l]) tmp_liq = resid_moist[l];
      if (liq[l] > resid_moist[l]) Q12[l] = Ksat[l] * vpow(((tmp_liq - resid_moist[l]) / (max_moist[l] - resid_moist[l])), expt[l]);
      else Q12[l] = 0.;
    }
    for (int l = 0; l < NL - 1; l++) {
      const double dt_runoff = (l == 0) ? tmp_dt_runoff : 0;
      double tmp_inflow = 0.;
      liq[l] = liq[l] + (inflow - dt_runoff) - (Q12[l] + evap[l]);
      if ((liq[l] + ice[l]) > max_moist[l]) {
        tmp_inflow = (liq[l] + ice[l]) - max_moist[l];
        liq[l] = max_moist[l] - ice[l];
        if (l == 0) {
          Q12[l] += tmp_inflow;
          tmp_inflow = 0;
        } else {
          int tl = l;
          while (tmp_inflow > 0) {
            tl--;
            if (tl < 0) {
              runoff_v += tmp_inflow;
              tmp_inflow = 0;
            } else {
              liq[tl] += tmp_inflow;
              if ((liq[tl] + ice[tl]) > max_moist[tl]) {
                tmp_inflow = ((liq[tl] + ice[tl]) - max_moist[tl]);
                liq[tl] = max_moist[tl] - ice[tl];
              } else tmp_inflow = 0;
            }
          }
        }
      }
      if ((liq[l] + ice[l]) < resid_moist[l]) {
        Q12[l] += (liq[l] + ice[l]) - resid_moist[l];
        liq[l] = resid_moist[l] - ice[l];
      }
      inflow = (Q12[l] + tmp_inflow);
      Q12[l] += tmp_inflow;
    }
    // ARNO baseflow from the bottom layer
    const int l = NL - 1;
    double rel_moist = (liq[l] - resid_moist[l]) / (max_moist[l] - resid_moist[l]);
    double frac = Dsmax * Ds / Ws;
    double dt_baseflow = frac * rel_moist;
    if (rel_moist > Ws) {
      frac = (rel_moist - Ws) / (1 - Ws);
      dt_baseflow += Dsmax * (1 - Ds / Ws) * vpow(frac, c_exp);
    }
    if (dt_baseflow < 0) dt_baseflow = 0;
    liq[l] += Q12[l - 1] - (evap[l] + dt_baseflow);
    if ((liq[l] + ice[l]) < resid_moist[l]) {
      dt_baseflow += (liq[l] + ice[l]) - resid_moist[l];
      liq[l] = resid_moist[l] - ice[l];
    }
    if ((liq[l] + ice[l]) > max_moist[l]) {
      double tmp_moist = ((liq[l] + ice[l]) - max_moist[l]);
      liq[l] = max_moist[l] - ice[l];
      int tl = l;
      while (tmp_moist > 0) {
        tl--;
        if (tl < 0) {
          runoff_v += tmp_moist;
          tmp_moist = 0;
        } else {
          liq[tl] += tmp_moist;
          if ((liq[tl] + ice[tl]) > max_moist[tl]) {
            tmp_moist = ((liq[tl] + ice[tl]) - max_moist[tl]);
            liq[tl] = max_moist[tl] - ice[tl];
          } else tmp_moist = 0;
        }
      }
    }
    baseflow += dt_baseflow;
  }
  if (baseflow < 0) {  // runoff.c:706-709 (bottom layer)
    cell.layer[NL - 1].evap += baseflow;
    baseflow = 0;
  }
  for (int l = 0; l < NL; l++) mm_tmp[l] = (liq[l] + ice[l]);
  double tmp_runoff;
  compute_runoff_and_asat(cp, mm_tmp, 0, &A, &tmp_runoff);
  for (int l = 0; l < NL; l++) cell.layer[l].moist = liq[l] + ice[l];
  cell.asat += A;
  cell.runoff += runoff_v;
  cell.baseflow += baseflow;
  wrap_compute_zwt(cp, cell);
  if (o.FULL_ENERGY || o.FROZEN_SOIL) {
    double moist[NL];
    for (int l = 0; l < NL; l++) moist[l] = cell.layer[l].moist;
    distribute_node_moisture_properties<NN>(energy, cp, moist, o);
  }
  return 0;
}

// soil_conduction.c:775-828
template <int NN>
VIC_HDI void find_0_degree_fronts(EnergyBal<NN>& energy, const CellPar& cp, const double* T, int Nnodes) {
  int Nthaw = 0, Nfrost = 0;
  double tdepth[VICGPU_NFRONTS], fdepth[VICGPU_NFRONTS];
  for (int f = 0; f < VICGPU_NFRONTS; f++) fdepth[f] = tdepth[f] = vnan();
  for (int n = Nnodes - 2; n >= 0; n--) {
    if (T[n] > 0 && T[n + 1] <= 0 && Nthaw < VICGPU_NFRONTS) {
      tdepth[Nthaw] = linear_interp(0, T[n], T[n + 1], cp.node(CN_Zsum_node, n), cp.node(CN_Zsum_node, n + 1));
      Nthaw++;
    } else if (T[n] < 0 && T[n + 1] >= 0 && Nfrost < VICGPU_NFRONTS) {
      fdepth[Nfrost] = linear_interp(0, T[n], T[n + 1], cp.node(CN_Zsum_node, n), cp.node(CN_Zsum_node, n + 1));
      Nfrost++;
    }
  }
  for (int f = 0; f < VICGPU_NFRONTS; f++) {
    energy.tdepth[f] = tdepth[f];
    energy.fdepth[f] = fdepth[f];
  }
  energy.Nthaw = Nthaw;
  energy.Nfrost = Nfrost;
}

// soil_conduction.c:617-723
VIC_HDI void estimate_layer_ice_content_quick_flux(SoilLayer* layer, double Tsurf, double T1, const CellPar& cp, const Opts& o) {
  const int NL = VICGPU_NLAYER;
  double Lsum[NL + 1];
  Lsum[0] = 0;
  for (int l = 1; l <= NL; l++) Lsum[l] = cp.layer(CL_depth, l - 1) + Lsum[l - 1];
  const double avg_temp = cp(CP_avg_temp), dp = cp(CP_dp);
  layer[0].T = 0.5 * (Tsurf + T1);
  for (int l = 1; l < NL; l++)
    layer[l].T = avg_temp - dp / (cp.layer(CL_depth, l)) * (T1 - avg_temp) * (vexp(-(Lsum[l + 1] - Lsum[1]) / dp) - vexp(-(Lsum[l] - Lsum[1]) / dp));
  const bool fs = o.FROZEN_SOIL && (cp(CP_FS_ACTIVE) != 0.0);
  for (int l = 0; l < NL; l++) {
    layer[l].soil_ice = 0;
    if (fs) {
      layer[l].soil_ice = layer[l].moist - maximum_unfrozen_water(layer[l].T, cp.layer(CL_max_moist, l), cp.layer(CL_bubble, l), cp.layer(CL_expt, l));
      if (layer[l].soil_ice < 0) layer[l].soil_ice = 0;
      if (layer[l].soil_ice > layer[l].moist) layer[l].soil_ice = layer[l].moist;
    }
  }
}

// soil_conduction.c:444-614; returns 0 or ERROR_I
template <int NN>
VIC_HDI int estimate_layer_ice_content(SoilLayer* layer, const double* T, int Nnodes, const CellPar& cp, const Opts& o) {
  const int NL = VICGPU_NLAYER;
  double Lsum[NL + 1], tmp_ice[NN], tmpT[NN], tmpZ[NN];
  Lsum[0] = 0;
  for (int l = 1; l <= NL; l++) Lsum[l] = cp.layer(CL_depth, l - 1) + Lsum[l - 1];
  const bool fs = o.FROZEN_SOIL && (cp(CP_FS_ACTIVE) != 0.0);
  for (int l = 0; l < NL; l++) {
    layer[l].T = 0.;
    layer[l].soil_ice = 0.;
    int min_n = Nnodes - 2;
    while (Lsum[l] < cp.node(CN_Zsum_node, min_n) && min_n > 0) min_n--;
    int max_n = 1;
    while (max_n < Nnodes && Lsum[l + 1] > cp.node(CN_Zsum_node, max_n)) max_n++;
    if (max_n >= Nnodes) return ERROR_I;
    if (cp.node(CN_Zsum_node, min_n) < Lsum[l])
      tmpT[min_n] = linear_interp(Lsum[l], cp.node(CN_Zsum_node, min_n), cp.node(CN_Zsum_node, min_n + 1), T[min_n], T[min_n + 1]);
    else tmpT[min_n] = T[min_n];
    tmpZ[min_n] = Lsum[l];
    for (int n = min_n + 1; n < max_n; n++) {
      tmpT[n] = T[n];
      tmpZ[n] = cp.node(CN_Zsum_node, n);
    }
    if (cp.node(CN_Zsum_node, max_n) > Lsum[l + 1])
      tmpT[max_n] = linear_interp(Lsum[l + 1], cp.node(CN_Zsum_node, max_n - 1), cp.node(CN_Zsum_node, max_n), T[max_n - 1], T[max_n]);
    else tmpT[max_n] = T[max_n];
    tmpZ[max_n] = Lsum[l + 1];
    for (int n = min_n; n <= max_n; n++) {
      if (fs) {
        tmp_ice[n] = layer[l].moist - maximum_unfrozen_water(tmpT[n], cp.layer(CL_max_moist, l), cp.layer(CL_bubble, l), cp.layer(CL_expt, l));
        if (tmp_ice[n] < 0) tmp_ice[n] = 0.;
      } else tmp_ice[n] = 0;
    }
    for (int n = min_n; n < max_n; n++) {
      layer[l].soil_ice += (tmpZ[n + 1] - tmpZ[n]) * (tmp_ice[n + 1] + tmp_ice[n]) / 2.;
      layer[l].T += (tmpZ[n + 1] - tmpZ[n]) * (tmpT[n + 1] + tmpT[n]) / 2.;
    }
    layer[l].soil_ice /= cp.layer(CL_depth, l);
    layer[l].T /= cp.layer(CL_depth, l);
  }
  return 0;
}

// frozen_soil.c:12-103: copy the new node temperatures into the energy record and derive
// layer temperature / ice.  T has Nnodes entries.
template <int NN>
VIC_HDI int calc_layer_average_thermal_props(EnergyBal<NN>& energy, SoilLayer* layer, const CellPar& cp, int Nnodes, const double* T, const Opts& o) {
  if (o.FROZEN_SOIL && (cp(CP_FS_ACTIVE) != 0.0)) find_0_degree_fronts<NN>(energy, cp, T, Nnodes);
  else energy.Nfrost = 0;
  for (int i = 0; i < NN; i++) if (i < Nnodes) energy.T[i] = T[i];
  energy.frozen = (energy.Nfrost > 0) ? 1.0 : 0.0;
  if (o.QUICK_FLUX) estimate_layer_ice_content_quick_flux(layer, energy.T[0], energy.T[1], cp, o);
  else return estimate_layer_ice_content<NN>(layer, energy.T, Nnodes, cp, o);
  return 0;
}

// prepare_full_energy.c:8-94 (+ compute_soil_layer_thermal_properties, soil_conduction.c:725-773)
template <int NN>
VIC_HDI void prepare_full_energy(Hru<NN>& h, const CellPar& cp, double AreaFract_band, const Opts& o, double* moist0, double* ice0) {
  if (AreaFract_band > 0.0) {
    const double d0 = cp.layer(CL_depth, 0);
    *moist0 = h.cell.layer[0].moist / (d0 * 1000.);
    if (o.FROZEN_SOIL && (cp(CP_FS_ACTIVE) != 0.0)) {
      if ((h.energy.T[0] + h.energy.T[1]) / 2. < 0.) {
        *ice0 = *moist0 - maximum_unfrozen_water((h.energy.T[0] + h.energy.T[1]) / 2., cp.layer(CL_max_moist, 0) / (d0 * 1000.),
                                                 cp.layer(CL_bubble, 0), cp.layer(CL_expt, 0));
        if (*ice0 < 0.) *ice0 = 0.;
      } else *ice0 = 0.;
    } else *ice0 = 0.;
    // only the top two layers are used afterwards (energy.kappa[0..1], energy.Cs[0..1])
    for (int l = 0; l < 2; l++) {
      const double dl = cp.layer(CL_depth, l);
      const double moist = h.cell.layer[l].moist / dl / 1000;
      const double ice = div_pos(div_zn(h.cell.layer[l].soil_ice, dl), 1000);
      const double kappa = soil_conductivity_pre(moist, moist - ice, cell_kpre(cp, l));
      const double Cs = volumetric_heat_capacity(cp.layer(CL_bulk_density, l) / cp.layer(CL_soil_density, l), moist - ice, ice, cp.layer(CL_organic, l));
      if (l == 0) { h.energy.kappa0 = kappa; h.energy.Cs0 = Cs; }
      else { h.energy.kappa1 = kappa; h.energy.Cs1 = Cs; }
    }
  } else {
    *ice0 = 0.;
  }
}

}  // namespace vic
#endif

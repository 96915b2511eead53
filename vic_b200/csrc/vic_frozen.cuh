// vic_frozen.cuh -- soil thermal-node profile for QUICK_FLUX = FALSE:
//   explicit finite-difference heat equation, Gauss-Seidel sweeps to 0.01 C (<= 1000 sweeps),
//   per-node Brent solve where the node is below freezing and frozen soil is active.
//   solve_T_profile            frozen_soil.c:105-225
//   calc_soil_thermal_fluxes   frozen_soil.c:305-505
//   SoilThermalEqn::calculate  soil_thermal_eqn.c:8-92
//
// Two reference behaviours are reproduced on purpose (the oracle is the reference build):
//  (1) the coefficient arrays A..E (frozen_soil.c:150-154) are only filled on the first solve of a
//      surface-temperature search; they depend only on quantities that are constant during that
//      search, so this code recomputes them on every call (same values).
//  (2) calc_soil_thermal_fluxes is handed the per-LAYER arrays soil_con->max_moist / bubble / expt
//      (frozen_soil.c:219) but indexes them per NODE; with MAX_LAYERS == 3 element j >= 3 of those
//      arrays is element j-3 of the member that follows them in soil_con_struct
//      (vicNl_def.h:920-941: bubble -> bubble_node, expt -> expt_node, max_moist -> max_moist_node).
#ifndef VIC_FROZEN_CUH
#define VIC_FROZEN_CUH
#include "vic_brent.cuh"
#include "vic_leaf.cuh"

namespace vic {

// Cost estimate for the row binning of the soil-thermal-profile configurations (vicgpu_api.cu k_bin_keys): the number of frozen-node
// solves of the thread's current HRU-step, counted into a per-thread slot of a device buffer that the translation unit of the step
// kernel owns (vicgpu_step.inc defines VIC_WORK_BUFFER for the 10- and 32-node kernels only; hru_work moves the count to
// Tables::cost).  A side channel on purpose: threading a counter through the signatures of the surface solve cost the 3-node kernel
// 11 % (profiles/r02_summary.md).  No physics reads it.
#if defined(VIC_WORK_BUFFER) && defined(__CUDACC__)
static __device__ int* vic_work_buf = nullptr;
#endif
VIC_HD void vic_count_work(int n) {
#if defined(VIC_WORK_BUFFER) && defined(__CUDA_ARCH__)
  if (vic_work_buf) vic_work_buf[blockIdx.x * blockDim.x + threadIdx.x] += n;
#else
  (void)n;
#endif
}

struct SoilThermalEqn {
  double TL, TU, T0, moist, max_moist, bubble, expt, ice0, A, B, C, D, E;
  int EXP_TRANS, node;
  VIC_HDI double operator()(double T) { return eval(T); }
  // (inlined into the single call site of the lane-asynchronous sweeps, solve_T_profile)
  VIC_HD double eval(double T) {
    double ice;
    if (T < 0.) {
      ice = moist - maximum_unfrozen_water(T, max_moist, bubble, expt);
      if (ice < 0.) ice = 0.;
      if (ice > max_moist) ice = max_moist;
    } else ice = 0.;
    double value;
    if (!EXP_TRANS) {
      value = -A * (T - T0) + B * (TL - TU) + C * (TL - T) - D * (T - TU) + E * (ice - ice0);
      const double flux_term1 = B * (TL - TU);
      const double flux_term2 = C * (TL - T) - D * (T - TU);
      if (node == 1 && fabs(TL - TU) > 5. && (T < TL && T < TU) && (flux_term1 < 0 && flux_term2 > 0) && fabs(flux_term1) > fabs(flux_term2))
        value = -A * (T - T0) + C * (TL - T) - D * (T - TU) + E * (ice - ice0);  // cold-nose fix
    } else {
      value = -A * (T - T0) + B * (TL - TU) + C * (TL - 2. * T + TU) - D * (TL - TU) + E * (ice - ice0);
      const double flux_term1 = B * (TL - TU);
      const double flux_term2 = C * (TL - 2. * T + TU) - D * (TL - TU);
      if (node == 1 && fabs(TL - TU) > 5. && (T < TL && T < TU) && (flux_term1 < 0 && flux_term2 > 0) && fabs(flux_term1) > fabs(flux_term2))
        value = -A * (T - T0) + C * (TL - 2. * T + TU) - D * (TL - TU) + E * (ice - ice0);
    }
    return value;
  }
};

// the three mis-indexed per-layer arrays, see (2) above
VIC_HD double layer_array_as_node(const CellPar& cp, int layer_field, int node_field, int j) {
  return (j < VICGPU_NLAYER) ? cp.layer(layer_field, j) : cp.node(node_field, j - VICGPU_NLAYER);
}

// Returns 0 or ERROR_I.  T (out) and T0 (in) have Nnodes entries; T0[0] is the trial surface
// temperature.  Tfbflag / Tfbcount are reset on every call, as in the reference.
template <int NN>
VIC_HDI int solve_T_profile(double* T, const double* T0, double* Tfbflag, double* Tfbcount, const double* kappa, const double* Cs,
                            const double* moist, double deltat, const double* ice, double Dp, int Nnodes, int* FIRST_SOLN, int NOFLUX,
                            int EXP_TRANS, const CellPar& cp, const Opts& o) {
  const int MAXIT = 1000;
  double A[NN], B[NN], C[NN], D[NN], E[NN], Tlast[NN];
  FIRST_SOLN[0] = 0;
  {
    double Bexp = 0;
    if (EXP_TRANS) Bexp = vlog(Dp + 1.) / (double)(Nnodes - 1);
    const int jend = NOFLUX ? Nnodes : Nnodes - 1;
    for (int j = 1; j < jend; j++) {
      const double kup = (j == Nnodes - 1) ? kappa[j] : kappa[j + 1];
      if (!EXP_TRANS) {
        const double al = cp.node(CN_alpha, j - 1);
        A[j] = Cs[j] * al * al;
        B[j] = (kup - kappa[j - 1]) * deltat;
        C[j] = 2 * deltat * kappa[j] * al / cp.node(CN_gamma, j - 1);
        D[j] = 2 * deltat * kappa[j] * al / cp.node(CN_beta, j - 1);
        E[j] = ice_density * Lf * al * al;
      } else {
        const double z1 = (cp.node(CN_Zsum_node, j) + 1);
        A[j] = 4 * Bexp * Bexp * Cs[j] * z1 * z1;
        B[j] = (kup - kappa[j - 1]) * deltat;
        C[j] = 4 * deltat * kappa[j];
        D[j] = 2 * deltat * kappa[j] * Bexp;
        E[j] = 4 * Bexp * Bexp * ice_density * Lf * z1 * z1;
      }
    }
  }
  for (int j = 0; j < Nnodes; j++) T[j] = T0[j];

  // ---- calc_soil_thermal_fluxes
  const bool frozen_on = (cp(CP_FS_ACTIVE) != 0.0) && o.FROZEN_SOIL;
  bool Done = false;
  int ItCount = 0;
  const double threshold = 1.e-2;
  for (int j = 0; j < Nnodes; j++) {
    Tlast[j] = T[j];
    Tfbflag[j] = 0;
    Tfbcount[j] = 0;
  }
  // A profile that does not converge does not wander: the Gauss-Seidel sweep is a deterministic map of the node temperatures, and
  // within a dozen sweeps the iterates repeat EXACTLY with period 2 or 3.  From an exact repeat on, every later sweep is known --
  // it can neither converge (the periodic sweeps did not) nor differ -- so the remaining sweeps up to MAXIT are not executed: the
  // temperatures after sweep MAXIT are read off the cycle and the per-node fallback counts of the skipped sweeps are added from the
  // cycle's own counts.  Bit-identical to running all 1000 sweeps (which cost ~0.25 s of a warp per residual evaluation and,
  // because a whole grid waits for its slowest warp, were the entire run time of the frozen-soil configuration).
  const int HIST = 8;
  double Th[HIST][NN];     // temperatures after the last HIST sweeps
  unsigned fbh[HIST];      // nodes whose Brent solve fell back, per sweep
#if !defined(VIC_FROZEN_ASYNC)
  while (!Done && ItCount < MAXIT) {
    ItCount++;
    unsigned fbmask = 0;
    double maxdiff = threshold;
    const int jend = NOFLUX ? Nnodes : Nnodes - 1;
    for (int j = 1; j < jend; j++) {
      const bool bottom = (j == Nnodes - 1);  // only reached with NOFLUX: the node below is the node itself
      const double oldT = T[j];
      const double Tdn = bottom ? T[j] : T[j + 1];
      if (T[j] >= 0 || !frozen_on) {
        if (!EXP_TRANS) T[j] = (A[j] * T0[j] + B[j] * (Tdn - T[j - 1]) + C[j] * Tdn + D[j] * T[j - 1] + E[j] * (0. - ice[j])) / (A[j] + C[j] + D[j]);
        else T[j] = (A[j] * T0[j] + B[j] * (Tdn - T[j - 1]) + C[j] * (Tdn + T[j - 1]) - D[j] * (Tdn - T[j - 1]) + E[j] * (0. - ice[j])) / (A[j] + 2. * C[j]);
      } else {
        SoilThermalEqn eq;
        eq.TL = Tdn; eq.TU = T[j - 1]; eq.T0 = T0[j]; eq.moist = moist[j];
        eq.max_moist = layer_array_as_node(cp, CL_max_moist, CN_max_moist_node, j);
        eq.bubble = layer_array_as_node(cp, CL_bubble, CN_bubble_node, j);
        eq.expt = layer_array_as_node(cp, CL_expt, CN_expt_node, j);
        eq.ice0 = ice[j]; eq.A = A[j]; eq.B = B[j]; eq.C = C[j]; eq.D = D[j]; eq.E = E[j]; eq.EXP_TRANS = EXP_TRANS; eq.node = j;
        T[j] = root_brent(T0[j] - (SOIL_DT), T0[j] + (SOIL_DT), eq);
        vic_count_work(1);
        if (result_is_error(T[j])) {
          if (o.TFALLBACK) {
            T[j] = T0[j];
            Tfbflag[j] = 1;
            Tfbcount[j] += 1;
            fbmask |= 1u << j;
          } else return ERROR_I;
        }
      }
      const double diff = fabs(oldT - T[j]);
      if (diff > maxdiff) maxdiff = diff;
    }
    if (maxdiff <= threshold) Done = true;
    if (!Done && ItCount >= 6 && ItCount < MAXIT) {
      // exact repeat of an earlier iterate?  (period p: the state after this sweep equals the state p sweeps ago)
      int period = 0;
      for (int p = 1; p < HIST && p < ItCount && period == 0; p++) {
        bool same = true;
        for (int j = 0; j < Nnodes && same; j++) same = (T[j] == Th[(ItCount - p) % HIST][j]) && !(T[j] != T[j]);
        if (same) period = p;
      }
      if (period > 0) {
        Th[ItCount % HIST][0] = T[0];
        for (int j = 0; j < Nnodes; j++) Th[ItCount % HIST][j] = T[j];
        fbh[ItCount % HIST] = fbmask;
        // sweeps ItCount+1 .. MAXIT repeat sweeps ItCount-period+1 .. ItCount: position q of the cycle comes up once for every
        // k = q, q + period, ... below the number of skipped sweeps (the counts are small integers held in doubles: adding the
        // number of visits at once is the same as adding 1 that many times)
        const int skipped = MAXIT - ItCount;
        for (int q = 0; q < period && q < skipped; q++) {
          const unsigned m = fbh[(ItCount - period + 1 + q) % HIST];
          if (m) {
            const int visits = (skipped - q + period - 1) / period;
            for (int j = 0; j < Nnodes; j++)
              if (m & (1u << j)) {
                Tfbflag[j] = 1;
                Tfbcount[j] += visits;
              }
          }
        }
        const int last = ItCount - period + 1 + (MAXIT - ItCount - 1) % period;
        for (int j = 0; j < Nnodes; j++) T[j] = Th[last % HIST][j];
        ItCount = MAXIT;
        break;
      }
    }
    for (int j = 0; j < Nnodes; j++) Th[ItCount % HIST][j] = T[j];
    fbh[ItCount % HIST] = fbmask;
  }
#else
  // Lane-asynchronous form of the same sweeps (-DVIC_FROZEN_ASYNC; measured, NOT the default).  In lock-step form (above) the threads
  // of a warp walk the nodes together and at every node wait for the lanes whose node is frozen to finish a Brent solve of 5-15
  // residual evaluations: 6.8 of 32 lanes are active per instruction on the 100,000-cell frozen-soil workload.  Here every
  // thread keeps its own cursor (sweep, node, state of the node's solve: BrentStep) and the loop body is ONE residual evaluation --
  // the single expensive call site (maximum_unfrozen_water's pow) -- preceded by whatever cheap work brings the thread to its next
  // evaluation (explicit nodes, end-of-sweep bookkeeping).  A thread is busy as long as it has evaluations left, wherever the other
  // lanes are; the operations of each thread and their order are exactly those of the lock-step form (bit-identical on the frozen-soil
  // parity tests, CPU port and GPU).  Result (profiles/r02_summary.md): the residual's pow runs at 12.6 lanes instead of 6.9 and half
  // as often per warp, but the machine's own bookkeeping now runs at 5.6 lanes -- the lanes of a warp do not differ in WHERE their
  // frozen nodes are but in WHETHER they have expensive (non-converging) profiles this record at all -- and the record takes 316 ms
  // instead of 282 ms at 100,000 cells (58 instead of 60 ms at 10,000).
  {
    const int jend = NOFLUX ? Nnodes : Nnodes - 1;
    int j = jend;          // cursor; jend = "between sweeps"
    bool started = false, in_solve = false;
    unsigned fbmask = 0;
    double maxdiff = threshold, oldT = 0;
    BrentStep bs;
    SoilThermalEqn eq;
    for (;;) {
      while (!in_solve) {
        if (j >= jend) {
          if (started) {  // end of a sweep
            if (maxdiff <= threshold) Done = true;
            bool cycle = false;
            if (!Done && ItCount >= 6 && ItCount < MAXIT) {
              // exact repeat of an earlier iterate?  (period p: the state after this sweep equals the state p sweeps ago)
              int period = 0;
              for (int p = 1; p < HIST && p < ItCount && period == 0; p++) {
                bool same = true;
                for (int k = 0; k < Nnodes && same; k++) same = (T[k] == Th[(ItCount - p) % HIST][k]) && !(T[k] != T[k]);
                if (same) period = p;
              }
              if (period > 0) {
                for (int k = 0; k < Nnodes; k++) Th[ItCount % HIST][k] = T[k];
                fbh[ItCount % HIST] = fbmask;
                // sweeps ItCount+1 .. MAXIT repeat sweeps ItCount-period+1 .. ItCount: position q of the cycle comes up once for every
                // k = q, q + period, ... below the number of skipped sweeps (the counts are small integers held in doubles: adding the
                // number of visits at once is the same as adding 1 that many times)
                const int skipped = MAXIT - ItCount;
                for (int q = 0; q < period && q < skipped; q++) {
                  const unsigned m = fbh[(ItCount - period + 1 + q) % HIST];
                  if (m) {
                    const int visits = (skipped - q + period - 1) / period;
                    for (int k = 0; k < Nnodes; k++)
                      if (m & (1u << k)) {
                        Tfbflag[k] = 1;
                        Tfbcount[k] += visits;
                      }
                  }
                }
                const int last = ItCount - period + 1 + (MAXIT - ItCount - 1) % period;
                for (int k = 0; k < Nnodes; k++) T[k] = Th[last % HIST][k];
                ItCount = MAXIT;
                cycle = true;
              }
            }
            if (!cycle) {
              for (int k = 0; k < Nnodes; k++) Th[ItCount % HIST][k] = T[k];
              fbh[ItCount % HIST] = fbmask;
            }
          }
          if (Done || ItCount >= MAXIT) goto sweeps_done;
          started = true;
          ItCount++;
          fbmask = 0;
          maxdiff = threshold;
          j = 1;
        }
        const bool bottom = (j == Nnodes - 1);  // only reached with NOFLUX: the node below is the node itself
        oldT = T[j];
        const double Tdn = bottom ? T[j] : T[j + 1];
        if (T[j] >= 0 || !frozen_on) {
          if (!EXP_TRANS) T[j] = (A[j] * T0[j] + B[j] * (Tdn - T[j - 1]) + C[j] * Tdn + D[j] * T[j - 1] + E[j] * (0. - ice[j])) / (A[j] + C[j] + D[j]);
          else T[j] = (A[j] * T0[j] + B[j] * (Tdn - T[j - 1]) + C[j] * (Tdn + T[j - 1]) - D[j] * (Tdn - T[j - 1]) + E[j] * (0. - ice[j])) / (A[j] + 2. * C[j]);
          const double diff = fabs(oldT - T[j]);
          if (diff > maxdiff) maxdiff = diff;
          j++;
        } else {
          eq.TL = Tdn; eq.TU = T[j - 1]; eq.T0 = T0[j]; eq.moist = moist[j];
          eq.max_moist = layer_array_as_node(cp, CL_max_moist, CN_max_moist_node, j);
          eq.bubble = layer_array_as_node(cp, CL_bubble, CN_bubble_node, j);
          eq.expt = layer_array_as_node(cp, CL_expt, CN_expt_node, j);
          eq.ice0 = ice[j]; eq.A = A[j]; eq.B = B[j]; eq.C = C[j]; eq.D = D[j]; eq.E = E[j]; eq.EXP_TRANS = EXP_TRANS; eq.node = j;
          brent_begin(bs, T0[j] - (SOIL_DT), T0[j] + (SOIL_DT));
          in_solve = true;
        }
      }
      // one residual evaluation of this thread's current node solve
      if (brent_advance(bs, eq.eval(bs.x))) {
        T[j] = bs.res;
        vic_count_work(1);
        if (result_is_error(T[j])) {
          if (o.TFALLBACK) {
            T[j] = T0[j];
            Tfbflag[j] = 1;
            Tfbcount[j] += 1;
            fbmask |= 1u << j;
          } else return ERROR_I;
        }
        const double diff = fabs(oldT - T[j]);
        if (diff > maxdiff) maxdiff = diff;
        j++;
        in_solve = false;
      }
    }
  sweeps_done:;
  }
#endif
  if (o.TFALLBACK) {
    // "cold nose" repair (frozen_soil.c:470-484)
    for (int j = 1; j < Nnodes - 1; j++) {
      if (Tlast[j - 1] - Tlast[j] > 0 && Tlast[j + 1] - T[j] > 0 && (T[j - 1] - T[j]) - (Tlast[j - 1] - Tlast[j]) > 0 &&
          (T[j + 1] - T[j]) - (Tlast[j + 1] - Tlast[j]) > 0) {
        T[j] = 0.5 * (T[j - 1] + T[j + 1]);
        Tfbflag[j] = 1;
        Tfbcount[j] += 1;
      }
    }
  }
  if (!Done) {
    if (o.TFALLBACK) {
      for (int j = 0; j < Nnodes; j++) {
        T[j] = T0[j];
        Tfbflag[j] = 1;
        Tfbcount[j] += 1;
      }
    } else return ERROR_I;
  }
  return 0;
}


// ---- IMPLICIT: Newton-Raphson solution of the whole profile (frozen_soil.c:229-301 solve_T_profile_implicit, :540-803
// NewtonRaphsonMethod::fda_heat_eqn; newt_raph_func_fast.c:17-220 compute / fdjac3 / tridiag) ------------------------------------
// The residual of the implicit finite-difference heat equation at the n unknown nodes.  focus == -1 evaluates all of them; the
// Jacobian (fdjac3) perturbs one unknown and re-evaluates only that node and its neighbours (focus >= 0), reading what the last
// all-node evaluation left in ice_new / Cs_new / kappa_new -- static arrays in upstream VIC 4.1.2, members here (the oracle gives
// the reference's copies `static thread_local`, oracle/Makefile patch 3): an all-node evaluation rewrites every entry a later
// single-node evaluation reads except kappa_new[n + 1] without NOFLUX, which nobody ever writes (zero, as a static is).
// max_moist / bubble / expt are the per-LAYER arrays indexed per NODE, as in the explicit scheme (header, (2)).
template <int NN>
struct FdaHeat {
  const double *T0, *moist, *ice, *kappa, *Cs;
  const CellPar* cp;
  double deltat, Ts, Tb, Bexp;
  int NOFLUX, EXP_TRANS;
  double ice_new[NN + 2], Cs_new[NN + 2], kappa_new[NN + 2];

  VIC_HD void node_props(int i, int lidx) {
    kappa_new[i] = soil_conductivity_pre(moist[i], moist[i] - ice_new[i], cell_kpre(*cp, lidx));
    Cs_new[i] = volumetric_heat_capacity(cp->layer(CL_bulk_density, lidx) / cp->layer(CL_soil_density, lidx), moist[i] - ice_new[i], ice_new[i],
                                         cp->layer(CL_organic, lidx));
  }
  VIC_HD double new_ice(double T, int i) const {
    if (T < 0) {
      double v = moist[i] - maximum_unfrozen_water(T, layer_array_as_node(*cp, CL_max_moist, CN_max_moist_node, i),
                                                   layer_array_as_node(*cp, CL_bubble, CN_bubble_node, i),
                                                   layer_array_as_node(*cp, CL_expt, CN_expt_node, i));
      if (v < 0) v = 0;
      return v;
    }
    return 0;
  }
  // the equation of unknown i (node i + 1); first: the i == 0 || i == 1 restriction of the cold-nose fix in single-node mode
  VIC_HD double node_residual(const double* T_2, int n, int i, bool nose_everywhere) const {
    double DT, DT_up, DT_down, T_up, Dkappa;
    if (i == 0) {
      DT = T_2[i + 1] - Ts; DT_up = T_2[i] - Ts; DT_down = T_2[i + 1] - T_2[i]; T_up = Ts;
    } else if (i == n - 1) {
      DT = Tb - T_2[i - 1]; DT_up = T_2[i] - T_2[i - 1]; DT_down = Tb - T_2[i]; T_up = T_2[i - 1];
    } else {
      DT = T_2[i + 1] - T_2[i - 1]; DT_up = T_2[i] - T_2[i - 1]; DT_down = T_2[i + 1] - T_2[i]; T_up = T_2[i - 1];
    }
    if (i < n - 1) Dkappa = kappa_new[i + 2] - kappa_new[i];
    else if (!NOFLUX) Dkappa = kappa_new[i + 2] - kappa_new[i];
    else Dkappa = kappa_new[i + 1] - kappa_new[i];
    const double storage_term = Cs_new[i + 1] * (T_2[i] - T0[i + 1]) / deltat + T_2[i] * (Cs_new[i + 1] - Cs[i + 1]) / deltat;
    double flux_term1, flux_term2;
    if (!EXP_TRANS) {
      const double al = cp->node(CN_alpha, i);
      flux_term1 = Dkappa / al * DT / al;
      flux_term2 = kappa_new[i + 1] * (DT_down / cp->node(CN_gamma, i) - DT_up / cp->node(CN_beta, i)) / (0.5 * al);
    } else {
      const double z1 = cp->node(CN_Zsum_node, i + 1) + 1.;
      flux_term1 = Dkappa / 2. * DT / 2. / (Bexp * z1) / (Bexp * z1);
      flux_term2 = kappa_new[i + 1] * ((DT_down - DT_up) / (Bexp * z1) / (Bexp * z1) - DT / 2. / (Bexp * z1 * z1));
    }
    if (nose_everywhere || i == 0 || i == 1) {
      if (fabs(DT) > 5. && (T_2[i] < T_2[i + 1] && T_2[i] < T_up)) {  // cold nose
        if ((flux_term1 < 0 && flux_term2 > 0) && fabs(flux_term1) > fabs(flux_term2)) flux_term1 = 0;
      }
    }
    const double flux_term = flux_term1 + flux_term2;
    const double phase_term = ice_density * Lf * (ice_new[i + 1] - ice[i + 1]) / deltat;
    return flux_term + phase_term - storage_term;
  }
  VIC_HD void eval(const double* T_2, double* res, int n, int focus) {
    const int Nlayers = VICGPU_NLAYER;
    int lidx = 0;
    double Lsum = 0.;
    bool PAST_BOTTOM = false;
    if (focus == -1) {
      for (int i = 0; i < n + 1; i++) {
        kappa_new[i] = kappa[i];
        if (i >= 1) {
          ice_new[i] = new_ice(T_2[i - 1], i);
          Cs_new[i] = Cs[i];
          if (ice_new[i] != ice[i]) node_props(i, lidx);
        }
        if (cp->node(CN_Zsum_node, i) > Lsum + cp->layer(CL_depth, lidx) && !PAST_BOTTOM) {
          Lsum += cp->layer(CL_depth, lidx);
          lidx++;
          if (lidx == Nlayers) {
            PAST_BOTTOM = true;
            lidx = Nlayers - 1;
          }
        }
      }
      // (the cold-nose test of the all-node evaluation looks at T_2[i + 1] of the last unknown too: one past the unknowns, i.e. the
      // caller's array element after them -- T[Nnodes - 1] without NOFLUX; see solve_T_profile_implicit)
      for (int i = 0; i < n; i++) res[i] = node_residual(T_2, n, i, true);
    } else {
      const int left = (focus == 0) ? 0 : focus - 1, right = (focus == n - 1) ? n - 1 : focus + 1;
      for (int i = left; i <= right; i++) ice_new[i + 1] = new_ice(T_2[i], i + 1);
      for (int i = 0; i <= right + 1; i++) {
        if (i >= left + 1) {
          if (ice_new[i] != ice[i]) node_props(i, lidx);
        }
        if (cp->node(CN_Zsum_node, i) > Lsum + cp->layer(CL_depth, lidx) && !PAST_BOTTOM) {
          Lsum += cp->layer(CL_depth, lidx);
          lidx++;
          if (lidx == Nlayers) {
            PAST_BOTTOM = true;
            lidx = Nlayers - 1;
          }
        }
      }
      for (int i = left; i <= right; i++) res[i] = node_residual(T_2, n, i, false);
    }
  }
};

// tridiagonal solve, newt_raph_func_fast.c:180-220 (a: sub-, b: main, c: super-diagonal; r: right-hand side in, solution out)
VIC_HD void nr_tridiag(double* a, double* b, double* c, double* r, int n) {
  double factor = b[0];
  b[0] = 1.0;
  c[0] = c[0] / factor;
  r[0] = r[0] / factor;
  for (int j = 1; j < n; j++) {
    factor = a[j];
    a[j] = a[j] - b[j - 1] * factor;
    b[j] = b[j] - c[j - 1] * factor;
    r[j] = r[j] - r[j - 1] * factor;
    factor = b[j];
    b[j] = 1.0;
    c[j] = c[j] / factor;
    r[j] = r[j] / factor;
  }
  for (int j = n - 2; j >= 0; j--) {
    factor = c[j];
    c[j] = c[j] - b[j + 1] * factor;
    r[j] = r[j] - r[j + 1] * factor;
    factor = b[j];
    r[j] = r[j] / factor;
  }
}

// Returns 0 (T holds the new profile) or 1 (no convergence in 150 trials: the caller falls back to the explicit scheme,
// func_surf_energy_bal.c:212-221).
template <int NN>
VIC_HDI int solve_T_profile_implicit(double* T, const double* T0, const double* kappa, const double* Cs, const double* moist, double deltat,
                                     const double* ice, double Dp, int Nnodes, int* FIRST_SOLN, int NOFLUX, int EXP_TRANS, const CellPar& cp) {
  const int MAXTRIAL = 150;
  const double TOLX = 1e-4, TOLF = 1e-1, R_MAX = 2.0, R_MIN = -5.0, RELAX1 = 0.9, RELAX2 = 0.7, RELAX3 = 0.2, EPS2 = 1e-4;
  if (FIRST_SOLN[0]) FIRST_SOLN[0] = 0;
  const int n = NOFLUX ? Nnodes - 1 : Nnodes - 2;
  FdaHeat<NN> fda;
  fda.T0 = T0; fda.moist = moist; fda.ice = ice; fda.kappa = kappa; fda.Cs = Cs; fda.cp = &cp;
  fda.deltat = deltat; fda.NOFLUX = NOFLUX; fda.EXP_TRANS = EXP_TRANS;
  fda.Bexp = 0;
  if (EXP_TRANS) fda.Bexp = NOFLUX ? vlog(Dp + 1.) / (double)(n) : vlog(Dp + 1.) / (double)(n + 1);
  fda.Ts = T0[0];
  fda.Tb = NOFLUX ? T0[n] : T0[n + 1];
  for (int i = 0; i < NN + 2; i++) fda.ice_new[i] = fda.Cs_new[i] = fda.kappa_new[i] = 0;
  double* x = &T[1];
  for (int i = 0; i < n; i++) x[i] = T0[i + 1];
  double fvec[NN], f[NN], p[NN], a[NN], b[NN], c[NN];
  for (int i = 0; i < NN; i++) fvec[i] = f[i] = p[i] = a[i] = b[i] = c[i] = 0;
  int Error = 1;
  for (int k = 0; k < MAXTRIAL; k++) {
    fda.eval(x, fvec, n, -1);
    double errf = 0.0;
    for (int i = 0; i < n; i++) errf += fabs(fvec[i]);
    if (errf <= TOLF) {
      Error = 0;
      break;
    }
    // forward-difference Jacobian, tridiagonal part only (fdjac3)
    for (int j = 0; j < n; j++) {
      const double temp = x[j];
      double h = EPS2 * fabs(temp);
      if (h == 0) h = EPS2;
      x[j] = temp + h;
      h = x[j] - temp;
      fda.eval(x, f, n, j);
      x[j] = temp;
      b[j] = (f[j] - fvec[j]) / h;
      if (j != 0) c[j - 1] = (f[j - 1] - fvec[j - 1]) / h;
      if (j != n - 1) a[j + 1] = (f[j + 1] - fvec[j + 1]) / h;
    }
    for (int i = 0; i < n; i++) p[i] = -fvec[i];
    nr_tridiag(a, b, c, p, n);
    double errx = 0.0;
    for (int i = 0; i < n; i++) {
      errx += fabs(p[i]);
      if (k > 10 && k <= 20 && x[i] < R_MAX && x[i] > R_MIN) x[i] += p[i] * RELAX1;
      else if (k > 20 && k <= 60 && x[i] < R_MAX && x[i] > R_MIN) x[i] += p[i] * RELAX2;
      else if (k > 60 && x[i] < R_MAX && x[i] > R_MIN) x[i] += p[i] * RELAX3;
      else x[i] += p[i];
    }
    if (errx <= TOLX) {
      Error = 0;
      break;
    }
  }
  if (Error == 0) {
    T[0] = T0[0];
    if (!NOFLUX) T[Nnodes - 1] = T0[Nnodes - 1];
  }
  return Error;
}

}  // namespace vic
#endif

// vic_gmb.cuh -- the glacier mass-balance curve of a cell at the end of an accumulation interval: cumulative mass balance of the
// cell's glacier HRUs against band elevation, fitted by b0 + b1 x + b2 x^2 (GlacierMassBalanceResult.c:35-72,
// GraphingEquation.c:7-126); the four numbers go to the state file (write_model_state.c:153-156).  Once per cell per interval --
// not hot, but it belongs to the time loop (vicNl.c:563) and consumes the balance the step kernels accumulate.
#ifndef VIC_GMB_CUH
#define VIC_GMB_CUH
#include "vic_types.cuh"

namespace vic {

// points (x[k], y[k]), k < n, in the order the reference builds them; out = {b0, b1, b2, fitError}
VIC_HD void gmb_fit(int n, const double* x, const double* y, double* out) {
  double b0 = 0, b1 = 0, b2 = 0;
  if (n == 1) {  // a horizontal line through the point
    b0 = y[0];
  } else if (n == 2) {
    const double slope = (y[1] - y[0]) / (x[1] - x[0]);
    b0 = y[0] - slope * x[0];
    b1 = slope;
  } else {
    // normal equations with the closed-form adjugate of X^T X, evaluated in the reference's order
    double sumx4 = 0, sumx3 = 0, sumx2 = 0, sumx1 = 0;
    for (int i = 0; i < n; i++) {
      sumx4 += x[i] * x[i] * x[i] * x[i];
      sumx3 += x[i] * x[i] * x[i];
      sumx2 += x[i] * x[i];
      sumx1 += x[i];
    }
    const int size = n;
    const double det = (sumx4 * sumx2 * size) + (sumx3 * sumx1 * sumx2) + (sumx2 * sumx3 * sumx1) - (sumx2 * sumx2 * sumx2) - (sumx1 * sumx1 * sumx4) -
                       (size * sumx3 * sumx3);
    const double inverse[3][3] = {{size * sumx2 - sumx1 * sumx1, -(size * sumx3 - sumx1 * sumx2), sumx1 * sumx3 - sumx2 * sumx2},
                                  {-(size * sumx3 - sumx2 * sumx1), size * sumx4 - sumx2 * sumx2, -(sumx1 * sumx4 - sumx3 * sumx2)},
                                  {sumx1 * sumx3 - sumx2 * sumx2, -(sumx1 * sumx4 - sumx2 * sumx3), sumx2 * sumx4 - sumx3 * sumx3}};
    double a[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < n; j++) {
        const double stuff = inverse[i][0] * (x[j] * x[j]) + inverse[i][1] * x[j] + inverse[i][2] * 1;
        a[i] += stuff * y[j];
      }
      a[i] /= det;
    }
    b0 = a[2];
    b1 = a[1];
    b2 = a[0];
  }
  double err = 0;
  for (int k = 0; k < n; k++) {
    const double curve = b0 + b1 * (x[k]) + b2 * (x[k] * x[k]);
    err += fabs(curve - y[k]);
  }
  out[0] = b0;
  out[1] = b1;
  out[2] = b2;
  out[3] = err;
}

}  // namespace vic
#endif

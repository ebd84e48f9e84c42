/* oracle/port_dist.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 *
 * Plain-C restatement of the encoder's perceptual block distortion (reference src/encode.c:
 * od_compute_var_4x4 :1081, od_compute_dist_8x8 :1111, od_compute_dist :1180).  Oracle for
 * SURVEY.md 8(f) rank 2 (the metric the block-size and deringing RDO loops evaluate); no CUDA twin yet.
 * Pinned against the reference build by tests/test_oracle_dering.py.
 *
 * For the flat QM the metric is the plain squared error.  Otherwise the error x - y is low-passed by
 * the separable kernel [1 5 1] (edge samples: [5 2]), and per 8x8 block
 *     dist = a^2 * (0.92/7^4 * sum(lowpassed error^2) + sum over the nine overlapping 4x4 windows
 *                   of (sqrt(var_x) - sqrt(var_y))^2),
 * a = activity masking factor from the windows' variances of the source; the block sum is scaled by a
 * quantiser-dependent constant.  Double accumulations keep the reference's order.
 */
#include <limits.h>
#include <math.h>
#include "port.h"

/* variance (times 16) of a 4x4 window of samples pre-shifted by 2 bits */
static int window_var(const od_coeff *p, int stride) {
  int i, j, s = 0, s2 = 0;
  for (i = 0; i < 4; i++) {
    for (j = 0; j < 4; j++) {
      int t = p[i*stride + j] >> 2;
      s += t;
      s2 += t*t;
    }
  }
  return s2 - (s*s >> 4);
}

static double dist_8x8(const od_coeff *x, const od_coeff *y, const od_coeff *err_lp, int stride,
 int use_activity_masking) {
  double inv_sum = 0, texture = 0, energy = 0, activity, stat;
  int lowest = INT_MAX, i, j;
  for (i = 0; i < 3; i++) {
    for (j = 0; j < 3; j++) {
      int vx = window_var(x + 2*i*stride + 2*j, stride);
      int vy = window_var(y + 2*i*stride + 2*j, stride);
      if (vx < lowest) lowest = vx;
      inv_sum += 1./(1 + vx);
      texture += vx - 2*sqrt(vx*(double)vy) + vy;
    }
  }
  stat = use_activity_masking ? 9./inv_sum : lowest;
  activity = (use_activity_masking ? 1.95 : 1.62)*pow(.25 + stat/(1 << 2*4), -1./6);   /* OD_COEFF_SHIFT = 4 */
  for (i = 0; i < 8; i++)
    for (j = 0; j < 8; j++) energy += err_lp[i*stride + j]*(double)err_lp[i*stride + j];
  energy *= 0.92/(7*7*7*7);
  return activity*activity*(energy + texture);
}

/* x, y: packed n x n blocks (n = 8..64 for the HVS QM path; any n for flat). */
double port_compute_dist(const od_coeff *x, const od_coeff *y, int n, int qm_is_flat, int use_activity_masking,
 int coded_quantizer) {
  static od_coeff err[64*64], rows[64*64], lp[64*64];
  double total = 0;
  int i, j;
  if (qm_is_flat) {
    for (i = 0; i < n*n; i++) {
      double d = x[i] - y[i];
      total += d*d;
    }
    return total;
  }
  for (i = 0; i < n*n; i++) err[i] = x[i] - y[i];
  for (i = 0; i < n; i++) {               /* horizontal [1 5 1], mirrored weight at the two ends */
    const od_coeff *e = err + i*n;
    rows[i*n] = 5*e[0] + 2*e[1];
    rows[i*n + n - 1] = 5*e[n - 1] + 2*e[n - 2];
    for (j = 1; j < n - 1; j++) rows[i*n + j] = 5*e[j] + e[j - 1] + e[j + 1];
  }
  for (j = 0; j < n; j++) {               /* vertical */
    lp[j] = 5*rows[j] + 2*rows[n + j];
    lp[(n - 1)*n + j] = 5*rows[(n - 1)*n + j] + 2*rows[(n - 2)*n + j];
  }
  for (i = 1; i < n - 1; i++)
    for (j = 0; j < n; j++) lp[i*n + j] = 5*rows[i*n + j] + rows[(i - 1)*n + j] + rows[(i + 1)*n + j];
  for (i = 0; i < n; i += 8)
    for (j = 0; j < n; j += 8) total += dist_8x8(x + i*n + j, y + i*n + j, lp + i*n + j, n, use_activity_masking);
  total *= coded_quantizer >= 47 ? 1.2 : coded_quantizer <= 36 ? 1.7 : 1.7 + (1.2 - 1.7)*(coded_quantizer - 36)/(47 - 36);
  return total;
}

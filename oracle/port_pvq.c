/* oracle/port_pvq.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 *
 * Plain-C restatement of the fixed-point PVQ of xiph/daala (OD_FLOAT_PVQ off):
 * src/pvq.c (gain companding, theta, K, Householder, synthesis) and the
 * encoder search of src/pvq_encoder.c (pvq_search_rdo_double :93, od_pvq_rate
 * :247 closed-form branch, pvq_theta :333).  All integer steps are bit-exact
 * restatements; the double-precision search keeps the reference's operation
 * order (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"
#include "port_pvq.h"

/* ---- fixed-point helpers (src/odintrin.h:157-190, src/pvq.h:73-111) ------ */
#define CGAIN_SHIFT 8            /* OD_CGAIN_SHIFT */
#define CGAIN_ONE (1 << CGAIN_SHIFT)
#define COMPAND_SHIFT 12         /* OD_COMPAND_SHIFT = 8 + OD_COEFF_SHIFT */
#define BETA_SHIFT 12            /* OD_BETA_SHIFT */
#define BETA_1 4096              /* OD_BETA(1.0) */
#define BETA_1_5 6144            /* OD_BETA(1.5) */
#define QM_SHIFT 11              /* OD_QM_SHIFT */
#define QM_INV_SHIFT 12          /* OD_QM_INV_SHIFT */
#define THETA_SHIFT 15           /* OD_THETA_SHIFT */
#define PVQ_SKIP_ZERO 1
#define PVQ_SKIP_COPY 2
#define MAX_PVQ_SIZE 128         /* OD_MAX_PVQ_SIZE */

static int ilog32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }   /* OD_ILOG */
static int32_t shl32(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
static int32_t shr_round32(int32_t x, int s) { return (x + ((1 << s) >> 1)) >> s; }
static int32_t shr_round64(int64_t x, int s) { return (int32_t)((x + ((1 << s) >> 1)) >> s); }
static int32_t vshr32(int32_t x, int s) { return s > 0 ? x >> s : shl32(x, -s); }
static int32_t vshr_round32(int32_t x, int s) { return s > 0 ? shr_round32(x, s) : shl32(x, -s); }
static int32_t vshr_round64(int64_t x, int s) { return s > 0 ? shr_round64(x, s) : shl32((int32_t)x, -s); }
static int32_t mul16_q15(int32_t a, int32_t b) { return ((int16_t)a*(int32_t)(int16_t)b) >> 15; }
static int32_t mul16_q16(int32_t a, int32_t b) { return ((int16_t)a*(int32_t)(int16_t)b) >> 16; }
static int32_t mul16(int32_t a, int32_t b) { return (int32_t)(int16_t)a*(int32_t)(int16_t)b; }
static int32_t mul16_qbeta(int32_t a, int32_t b) { return ((int16_t)a*(int32_t)(int16_t)b) >> BETA_SHIFT; }
static int32_t round32(double x) { return (int32_t)floor(.5 + x); }

/* src/pvq.c:408 */
static int16_t cos_pi_2(int16_t x) {
  int16_t x2 = (int16_t)mul16_q15(x, x);
  int32_t v = (1073758164 - x*x + x2*(-7654 + mul16_q16(x2, 16573 + mul16_q16(-2529, x2)))) >> 15;
  return (int16_t)(v < 32767 ? v : 32767);
}

/* src/pvq.c:428: cos of an angle in units of pi/2 / 32768. */
int port_pvq_cos(int32_t x) {
  x &= 0x1ffff;
  if (x > (1 << 16)) x = (1 << 17) - x;
  if (x & 0x7fff) {
    if (x < (1 << 15)) return cos_pi_2((int16_t)x);
    return (int16_t)-cos_pi_2((int16_t)(65536 - x));
  }
  if (x & 0xffff) return 0;
  if (x & 0x1ffff) return -32767;
  return 32767;
}

/* src/pvq.c:461 */
int port_pvq_sin(int32_t x) { return port_pvq_cos(32768 - x); }

/* src/pvq.c:472 */
int port_vector_log_mag(const od_coeff *x, int n) {
  int32_t sum = 0;
  int i;
  for (i = 0; i < n; i++) {
    int16_t t = (int16_t)(x[i] >> 8);
    sum += t*(int32_t)t;
  }
  return 8 + 1 + ilog32((uint32_t)(n + sum))/2;
}

/* src/pvq.c:526: Q15 reciprocal, two Newton steps. */
static int16_t rcp16(int16_t x) {
  int i = ilog32((uint32_t)(int32_t)x) - 1;
  int16_t n = (int16_t)(vshr_round32(x, i - 15) - 32768);
  int16_t r = (int16_t)(30840 + mul16_q15(-15420, n));
  r = (int16_t)(r - mul16_q15(r, mul16_q15(r, n) + r - 32768));
  r = (int16_t)(r - (1 + mul16_q15(r, mul16_q15(r, n) + r - 32768)));
  return (int16_t)vshr_round32(r, i - 14);
}

/* src/pvq.c:498 */
int port_compute_householder(int16_t *r, int n, int32_t gr, int *sign, int shift) {
  int m = 0;
  int i;
  int s;
  int16_t maxr = 0;
  for (i = 0; i < n; i++) {
    if (abs(r[i]) > maxr) {
      maxr = (int16_t)abs(r[i]);
      m = i;
    }
  }
  s = r[m] > 0 ? 1 : -1;
  r[m] = (int16_t)(r[m] + shr_round32(gr*s, shift));
  *sign = s;
  return m;
}

/* src/pvq.c:560 */
void port_apply_householder(int16_t *out, const int16_t *x, const int16_t *r, int n) {
  int32_t l2r = 0;
  int32_t proj = 0;
  int l2r_shift;
  int proj_shift;
  int outshift;
  int16_t l2r_norm;
  int16_t rcp;
  int16_t proj_norm;
  int16_t proj_1;
  int i;
  for (i = 0; i < n; i++) l2r += mul16(r[i], r[i]);
  for (i = 0; i < n; i++) proj += mul16(r[i], x[i]);
  l2r_shift = (ilog32((uint32_t)l2r) - 1) - 14;
  l2r_norm = (int16_t)vshr_round32(l2r, l2r_shift);
  rcp = rcp16(l2r_norm);
  proj_shift = (ilog32((uint32_t)abs(proj)) - 1) - 14;
  proj_norm = (int16_t)vshr_round32(proj, proj_shift);
  proj_1 = (int16_t)mul16_q15(proj_norm, rcp);
  outshift = 14 - proj_shift - 1 + l2r_shift;
  if (outshift > 30) outshift = 30;
  if (outshift >= 0) {
    for (i = 0; i < n; i++) out[i] = (int16_t)(x[i] - shr_round32(mul16(r[i], proj_1), outshift));
  }
  else {
    for (i = 0; i < n; i++) out[i] = (int16_t)(x[i] - shl32(mul16(r[i], proj_1), -outshift));
  }
}

/* src/pvq.c:625 */
static int16_t beta_rcp(int16_t beta) {
  if (beta == BETA_1) return BETA_1;
  if (beta == BETA_1_5) return 2731;  /* OD_BETA(1./1.5) */
  return (int16_t)shr_round32(rcp16((int16_t)(beta << (15 - 1 - BETA_SHIFT))), 14 + 1 - BETA_SHIFT);
}

/* src/pvq.c:638-662 */
static int32_t exp2_q15(int32_t x) {
  static const int32_t C[5] = {32768, 22709, 7913, 1704, 443};
  int integer = x >> 15;
  int32_t f;
  int32_t frac;
  if (integer > 14) return 0x7f000000;
  if (integer < -15) return 0;
  f = x - shl32(integer, 15);
  frac = mul16_q15(f, C[1] + mul16_q15(f, C[2] + mul16_q15(f, C[3] + mul16_q15(f, C[4]))));
  return vshr_round32(C[0] + frac, -integer) + 1;
}

/* src/pvq.c:668 */
static int16_t log2_q15(int16_t x) {
  return (int16_t)(x + mul16_q15(x, 14482 + mul16_q15(x, -23234 + mul16_q15(x, 13643
   + mul16_q15(x, -6403 + mul16_q15(x, 1515))))));
}

/* src/pvq.c:675 */
static int32_t pow_q(int32_t x, int16_t beta) {
  int log2_x;
  int16_t t;
  int32_t logr;
  if (x == 0) return 0;
  log2_x = ilog32((uint32_t)x) - 1;
  t = (int16_t)(vshr32(x, log2_x - 15) - 32768);
  logr = log2_q15(t) + (log2_x - COMPAND_SHIFT)*32768;
  logr = (int32_t)(((int16_t)beta*(int64_t)logr) >> BETA_SHIFT);
  return exp2_q15(logr);
}

/* src/pvq.c:967: 1/sqrt on [0.25, 1) in Q16 -> Q14. */
static int16_t rsqrt_norm(int16_t t) {
  int16_t n = (int16_t)(t - 32768);
  int32_t r = 23565 + mul16_q15(n, -13481 + mul16_q15(n, 6711));
  int32_t r2 = r*r;
  int32_t y = (((r2 >> 15)*n + r2) >> 12) - 131077;
  int32_t ry = r*y;
  return (int16_t)(r + ((((ry >> 16)*(3*y) >> 3) - ry) >> 18));
}

/* src/pvq.c:998 */
static int16_t rsqrt32(int32_t x, int *shift) {
  int k = (ilog32((uint32_t)x) - 1) >> 1;
  int s = 2*k - 14;
  int16_t t = (int16_t)vshr32(x, s);
  *shift = 14 + ((s + 16) >> 1);
  return rsqrt_norm(t);
}

/* src/pvq.c:726-756 */
static int16_t sqrt32(int32_t x, int *shift) {
  int k;
  int s;
  int32_t t;
  int32_t v;
  if (x == 0) {
    *shift = 0;
    return 0;
  }
  k = (ilog32((uint32_t)x) - 1) >> 1;
  s = 2*k - 14;
  t = vshr32(x, s);
  *shift = 15 - ((s + 16) >> 1);
  v = shr_round32(t*rsqrt_norm((int16_t)t), 15);
  return (int16_t)(v < 32767 ? v : 32767);
}

/* src/pvq.c:706 */
static int32_t gain_compand(int32_t g, int q0, int16_t beta) {
  int32_t e;
  if (beta == BETA_1) return (CGAIN_ONE*g + (q0 >> 1))/q0;
  e = pow_q(g, beta_rcp(beta));
  e <<= CGAIN_SHIFT + COMPAND_SHIFT - 15;
  return (e + (q0 >> 1))/q0;
}

/* src/pvq.c:766 */
int32_t port_gain_expand(int32_t cg0, int q0, int beta) {
  if (beta == BETA_1) return shr_round32(cg0*q0, CGAIN_SHIFT);
  if (beta == BETA_1_5) {
    int outshift;
    int32_t irt = sqrt32(cg0*q0, &outshift);
    int64_t tmp = cg0*q0*(int64_t)irt;
    return vshr_round64(tmp, CGAIN_SHIFT + outshift + ((CGAIN_SHIFT + COMPAND_SHIFT) >> 1));
  }
  return shr_round32(pow_q(shr_round32(cg0*q0, CGAIN_SHIFT), (int16_t)beta), 15 - COMPAND_SHIFT);
}

/* src/pvq.c:824 */
int32_t port_pvq_compute_gain(const int16_t *x, int n, int q0, int32_t *g, int beta, int bshift) {
  int32_t acc = 0;
  int sqrt_shift;
  int32_t irt;
  int i;
  for (i = 0; i < n; i++) acc += x[i]*(int32_t)x[i];
  irt = sqrt32(acc, &sqrt_shift);
  *g = vshr_round32(irt, sqrt_shift - bshift);
  return gain_compand(*g, q0, (int16_t)beta);
}

/* src/pvq.c:855 */
int port_pvq_compute_max_theta(int32_t qcg, int beta) {
  int ts = shr_round32(qcg*mul16_qbeta(402 /* OD_QCONST32(M_PI/2, 8) */, beta_rcp((int16_t)beta)),
   CGAIN_SHIFT*2);
  if (qcg < 358 /* OD_QCONST32(1.4, 8) */) ts = 1;
  return ts;
}

/* src/pvq.c:874 */
int32_t port_pvq_compute_theta(int t, int max_theta) {
  if (max_theta == 0) return 0;
  return ((1 << THETA_SHIFT)*(t < max_theta - 1 ? t : max_theta - 1) + (max_theta >> 1))/max_theta;
}

/* src/pvq.c:902 (nodesync == 1 always: OD_ROBUST_STREAM, src/internal.h:118) */
int port_pvq_compute_k(int32_t qcg, int itheta, int noref, int n, int beta) {
  static const int16_t sqrt_tbl[2][13] = {
    {0, 0, 0, 0, 2290, 2985, 4222, 0, 8256, 0, 16416, 0, 32767},
    {0, 0, 0, 0, 2401, 3072, 4284, 0, 8287, 0, 16432, 0, 32767}};
  int k;
  if (noref) {
    if (qcg == 0) return 0;
    if (n == 15 && qcg == CGAIN_ONE && beta > 5120 /* OD_BETA(1.25) */) return 1;
    k = shr_round64((int64_t)((qcg - (int64_t)51 /* QCONST32(.2, 8) */)
     *mul16_qbeta(beta_rcp((int16_t)beta), sqrt_tbl[1][ilog32((uint32_t)(n + 1))])), CGAIN_SHIFT + 10);
    return k > 1 ? k : 1;
  }
  if (itheta == 0) return 0;
  k = vshr_round64((shl32(itheta, 15) - 6554 /* QCONST32(.2, 15) */)
   *(int64_t)sqrt_tbl[0][ilog32((uint32_t)(n + 1))], 10 + 15);
  return k > 1 ? k : 1;
}

/* src/pvq.c:1037 */
void port_pvq_synthesis_partial(od_coeff *xcoeff, const od_coeff *ypulse, const int16_t *r16, int n,
 int noref, int32_t g, int32_t theta, int m, int s, const int16_t *qm_inv) {
  int nn = n - !noref;
  int yy = 0;
  int gshift;
  int qshift;
  int32_t scale;
  int i;
  for (i = 0; i < nn; i++) yy += ypulse[i]*(int32_t)ypulse[i];
  gshift = ilog32((uint32_t)g) - 14;
  if (gshift < 0) gshift = 0;
  if (yy == 0) scale = 0;
  else {
    int rshift;
    int16_t rs = rsqrt32(yy, &rshift);
    scale = vshr_round64(rs*(int64_t)g, rshift + gshift - 16);
  }
  qshift = QM_INV_SHIFT - gshift;
  if (noref) {
    for (i = 0; i < n; i++) {
      int32_t x = (int32_t)(((int16_t)ypulse[i]*(int64_t)scale) >> 16);
      xcoeff[i] = shr_round32(x*qm_inv[i], qshift);
    }
  }
  else {
    int16_t x[MAX_PVQ_SIZE];
    scale = round32(scale*(1./32768)*port_pvq_sin(theta));
    for (i = 0; i < m; i++) x[i] = (int16_t)(((int16_t)ypulse[i]*(int64_t)scale) >> 16);
    x[m] = (int16_t)floor(.5 + -s*(shr_round32(g, gshift))*(1./32768)*port_pvq_cos(theta));
    for (i = m; i < nn; i++) x[i + 1] = (int16_t)(((int16_t)ypulse[i]*(int64_t)scale) >> 16);
    port_apply_householder(x, x, r16, n);
    for (i = 0; i < n; i++) xcoeff[i] = shr_round32(x[i]*qm_inv[i], qshift);
  }
}

/* ---- encoder search (src/pvq_encoder.c) ---------------------------------- */

/* src/pvq_encoder.c:53 */
static double rsqrt_small(int i) {
  static const double tbl[16] = {
    1.000000, 0.707107, 0.577350, 0.500000, 0.447214, 0.408248, 0.377964, 0.353553,
    0.333333, 0.316228, 0.301511, 0.288675, 0.277350, 0.267261, 0.258199, 0.250000};
  if (i <= 16) return tbl[i - 1];
  return 1./sqrt(i);
}

/* src/pvq_encoder.c:93 */
double port_pvq_search_rdo_double(const int16_t *xcoeff, int n, int k, od_coeff *ypulse, double g2,
 double pvq_norm_lambda, int prev_k) {
  double x[MAX_PVQ_SIZE];
  double xx = 0;
  double xy = 0;
  double yy = 0;
  double norm_1;
  double lambda;
  double delta_rate;
  double accel_rate;
  int rdo_pulses;
  int i;
  int j;
  for (j = 0; j < n; j++) {
    x[j] = fabs((float)xcoeff[j]);
    xx += x[j]*x[j];
  }
  norm_1 = 1./sqrt(1e-30 + xx);
  lambda = pvq_norm_lambda/(1e-30 + g2);
  i = 0;
  if (prev_k > 0 && prev_k <= k) {
    for (j = 0; j < n; j++) {
      ypulse[j] = abs(ypulse[j]);
      xy += x[j]*ypulse[j];
      yy += ypulse[j]*ypulse[j];
      i += ypulse[j];
    }
  }
  else if (k > 2) {
    double l1_norm = 0;
    double l1_inv;
    for (j = 0; j < n; j++) l1_norm += x[j];
    l1_inv = 1./(l1_norm > 1e-100 ? l1_norm : 1e-100);
    for (j = 0; j < n; j++) {
      double tmp = k*x[j]*l1_inv;
      int f = (int)floor(tmp);
      ypulse[j] = f > 0 ? f : 0;
      xy += x[j]*ypulse[j];
      yy += ypulse[j]*ypulse[j];
      i += ypulse[j];
    }
  }
  else memset(ypulse, 0, n*sizeof(*ypulse));
  rdo_pulses = 1 + k/4;
  delta_rate = 3./n;
  accel_rate = 0.;
  if (k == 1) {
    if (n == 15) {
      accel_rate = -8./n;
      delta_rate = 4.5/n - accel_rate;
    }
    else if (n == 8) {
      accel_rate = 5.7/n;
      delta_rate = 9.3/n - accel_rate;
    }
  }
  for (; i < k - rdo_pulses; i++) {
    int pos = 0;
    double best_xy = -10;
    double best_yy = 1;
    for (j = 0; j < n; j++) {
      double tmp_xy = xy + x[j];
      double tmp_yy = yy + 2*ypulse[j] + 1;
      tmp_xy *= tmp_xy;
      if (j == 0 || tmp_xy*best_yy > best_xy*tmp_yy) {
        best_xy = tmp_xy;
        best_yy = tmp_yy;
        pos = j;
      }
    }
    xy = xy + x[pos];
    yy = yy + 2*ypulse[pos] + 1;
    ypulse[pos]++;
  }
  for (; i < k; i++) {
    double tbl[4];
    int pos = 0;
    double best_cost = -1e5;
    for (j = 0; j < 4; j++) tbl[j] = rsqrt_small((int)(yy + 2*j + 1));
    for (j = 0; j < n; j++) {
      double tmp_xy = xy + x[j];
      double tmp_yy = ypulse[j] < 4 ? tbl[ypulse[j]] : rsqrt_small((int)(yy + 2*ypulse[j] + 1));
      tmp_xy = 2*tmp_xy*norm_1*tmp_yy - lambda*j*(delta_rate + j*accel_rate);
      if (j == 0 || tmp_xy > best_cost) {
        best_cost = tmp_xy;
        pos = j;
      }
    }
    xy = xy + x[pos];
    yy = yy + 2*ypulse[pos] + 1;
    ypulse[pos]++;
  }
  for (i = 0; i < n; i++) if (xcoeff[i] < 0) ypulse[i] = -ypulse[i];
  return xy/(1e-100 + sqrt(xx*yy));
}

/* src/pvq_encoder.c:247, `speed > 0` closed form only (the speed == 0 branch
   runs the adaptive entropy coder and stays on the reference's host side). */
double port_pvq_rate(int qg, int icgr, int theta, int ts, const od_coeff *y0, int k, int n,
 int is_keyframe, int pli) {
  double rate;
  if (k == 0) rate = 0;
  else {
    int sum = 0;
    int i;
    double f;
    double t;
    for (i = 0; i < n - (theta != -1); i++) sum += i*abs(y0[i]);
    f = sum/(double)(k*n);
    t = log(n*2*(1*f + .025))*k/n;
    rate = (1 + .4*f)*n*(M_LOG2E*log(1 + (0 > t ? 0 : t))) + 3;
  }
  if (qg > 0 && theta >= 0) {
    rate += .9*(M_LOG2E*log(ts));
    if (is_keyframe && pli == 0) rate += 6;
    if (qg == icgr) rate -= .5;
  }
  return rate;
}

/* src/pvq_encoder.c:236 */
static int neg_interleave(int x, int ref) {
  if (x < ref) return -2*(x - ref) - 1;
  if (x < 2*ref) return 2*(x - ref);
  return x - 1;
}

typedef struct {
  int gain;
  int k;
  int32_t qtheta;
  int theta;
  int ts;
  int32_t qcg;
} cand;

/* src/pvq_encoder.c:333.  speed must be > 0 (closed-form rate). */
int port_pvq_theta(od_coeff *out, const od_coeff *x0, const od_coeff *r0, int n, int q0, od_coeff *y,
 int *itheta, int *max_theta, int *vk, int beta, double *skip_diff, int is_keyframe, int pli,
 const int16_t *qm, const int16_t *qm_inv, double pvq_norm_lambda) {
  const double gain_weight = 1.4;
  const double cgain_1 = 1./CGAIN_ONE;
  const double cgain_2 = cgain_1*cgain_1;  /* OD_CGAIN_SCALE_2 */
  const double theta_scale = (1 << THETA_SHIFT)*2./M_PI;  /* OD_THETA_SCALE */
  const double theta_scale_1 = 1./theta_scale;
  const double trig_1 = 1./32768;
  int32_t g;
  int32_t gr;
  int32_t cg;
  int32_t cgr;
  int32_t gain_offset;
  int32_t theta = 0;
  int32_t best_qtheta = 0;
  od_coeff y_tmp[MAX_PVQ_SIZE];
  int16_t x16[MAX_PVQ_SIZE];
  int16_t r16[MAX_PVQ_SIZE];
  int icgr;
  int qg = 0;
  int best_k = 0;
  int noref = 1;
  int m = 0;
  int s = 1;
  int skip = 0;
  int cfl_enabled;
  int xshift;
  int rshift;
  int i;
  int k;
  int r_is_null = 1;
  double corr = 0;
  double best_cost;
  double best_dist;
  double dist0;
  double dist;
  double skip_dist;
  xshift = port_vector_log_mag(x0, n) - 15;
  if (xshift < 0) xshift = 0;
  rshift = port_vector_log_mag(r0, n) - 14;
  if (rshift < 0) rshift = 0;
  for (i = 0; i < n; i++) {
    x16[i] = (int16_t)shr_round32(x0[i]*qm[i], QM_SHIFT + xshift);
    r16[i] = (int16_t)shr_round32(r0[i]*qm[i], QM_SHIFT + rshift);
    corr += mul16(x16[i], r16[i]);
    if (r0[i]) r_is_null = 0;
  }
  cfl_enabled = is_keyframe && pli != 0;
  cg = port_pvq_compute_gain(x16, n, q0, &g, beta, xshift);
  cgr = port_pvq_compute_gain(r16, n, q0, &gr, beta, rshift);
  if (cfl_enabled) cgr = CGAIN_ONE;
  icgr = shr_round32(cgr, CGAIN_SHIFT);
  gain_offset = cgr - shl32(icgr, CGAIN_SHIFT);
  dist = gain_weight*cg*cg*cgain_2;
  best_dist = dist;
  best_cost = dist + pvq_norm_lambda*port_pvq_rate(0, 0, -1, 0, NULL, 0, n, is_keyframe, pli);
  *itheta = -1;
  *max_theta = 0;
  memset(y, 0, n*sizeof(*y));
  corr = corr/(1e-100 + g*(double)gr/shl32(1, xshift + rshift));
  corr = corr < 1. ? corr : 1.;
  corr = corr > -1. ? corr : -1.;
  if (is_keyframe) skip_dist = gain_weight*cg*cg*cgain_2;
  else {
    skip_dist = gain_weight*(cg - cgr)*(cg - cgr) + cgr*(double)cg*(2 - 2*corr);
    skip_dist *= cgain_2;
  }
  if (!is_keyframe) {
    int32_t scgr = gain_offset > 0 ? gain_offset : 0;
    if (icgr == 0) {
      best_dist = gain_weight*(cg - scgr)*(cg - scgr) + scgr*(double)cg*(2 - 2*corr);
      best_dist *= cgain_2;
    }
    best_cost = best_dist + pvq_norm_lambda*port_pvq_rate(0, icgr, 0, 0, NULL, 0, n, is_keyframe, pli);
    best_qtheta = 0;
    *itheta = 0;
    *max_theta = 0;
    noref = 0;
  }
  dist0 = best_dist;
  if (n <= MAX_PVQ_SIZE && !r_is_null && corr > 0) {
    int16_t xr[MAX_PVQ_SIZE];
    cand items[20];
    int nitems = 0;
    int gain_bound = (cg - gain_offset) >> CGAIN_SHIFT;
    int prev_k = 0;
    int idx;
    double cos_dist = 0;
    theta = round32(theta_scale*acos(corr));
    m = port_compute_householder(r16, n, gr, &s, rshift);
    port_apply_householder(xr, x16, r16, n);
    for (i = m; i < n - 1; i++) xr[i] = xr[i + 1];
    for (i = gain_bound - 1 > 1 ? gain_bound - 1 : 1; i <= gain_bound + 1; i++) {
      int32_t qcg = shl32(i, CGAIN_SHIFT) + gain_offset;
      int ts = port_pvq_compute_max_theta(qcg, beta);
      int lo = (int)floor(.5 + theta*theta_scale_1*2/M_PI*ts) - 2;
      int hi = (int)ceil(theta*theta_scale_1*2/M_PI*ts);
      int j;
      if (lo < 0) lo = 0;
      if (hi > ts - 1) hi = ts - 1;
      for (j = lo; j <= hi; j++) {
        int32_t qtheta = port_pvq_compute_theta(j, ts);
        cand *c = &items[nitems++];
        c->gain = i;
        c->theta = j;
        c->k = port_pvq_compute_k(qcg, j, 0, n, beta);
        c->qcg = qcg;
        c->qtheta = qtheta;
        c->ts = ts;
      }
    }
    /* Stable insertion sort by (k, gain): what glibc's merge-sort qsort gives
       for src/pvq_encoder.c:504. */
    for (i = 1; i < nitems; i++) {
      cand c = items[i];
      int j = i - 1;
      while (j >= 0 && (items[j].k > c.k || (items[j].k == c.k && items[j].gain > c.gain))) {
        items[j + 1] = items[j];
        j--;
      }
      items[j + 1] = c;
    }
    for (idx = 0; idx < nitems; idx++) {
      int32_t qcg = items[idx].qcg;
      int j = items[idx].theta;
      int ts = items[idx].ts;
      int32_t qtheta = items[idx].qtheta;
      double cost;
      double dist_theta;
      double sin_prod;
      i = items[idx].gain;
      k = items[idx].k;
      dist_theta = 2 - 2.*port_pvq_cos(theta - qtheta)*trig_1;
      dist = gain_weight*(qcg - cg)*(qcg - cg) + qcg*(double)cg*dist_theta;
      dist *= cgain_2;
      if (dist > dist0 + 1.0*pvq_norm_lambda && k != 0) continue;
      sin_prod = port_pvq_sin(theta)*trig_1*port_pvq_sin(qtheta)*trig_1;
      if (k == 0) {
        cos_dist = 0;
        memset(y_tmp, 0, (n - 1)*sizeof(*y_tmp));
      }
      else if (k != prev_k) {
        cos_dist = port_pvq_search_rdo_double(xr, n - 1, k, y_tmp, qcg*(double)cg*sin_prod*cgain_2,
         pvq_norm_lambda, prev_k);
      }
      prev_k = k;
      dist_theta = 2 - 2.*port_pvq_cos(theta - qtheta)*trig_1 + sin_prod*(2 - 2*cos_dist);
      dist = gain_weight*(qcg - cg)*(qcg - cg) + qcg*(double)cg*dist_theta;
      dist *= cgain_2;
      cost = dist + pvq_norm_lambda*port_pvq_rate(i, icgr, j, ts, y_tmp, k, n, is_keyframe, pli);
      if (cost < best_cost) {
        best_cost = cost;
        best_dist = dist;
        qg = i;
        best_k = k;
        best_qtheta = qtheta;
        *itheta = j;
        *max_theta = ts;
        noref = 0;
        memcpy(y, y_tmp, (n - 1)*sizeof(*y));
      }
    }
  }
  if (n <= MAX_PVQ_SIZE && ((is_keyframe && pli == 0) || corr < .5 || cg < (int32_t)shl32(2, CGAIN_SHIFT))) {
    int gain_bound = cg >> CGAIN_SHIFT;
    int prev_k = 0;
    for (i = gain_bound > 1 ? gain_bound : 1; i <= gain_bound + 1; i++) {
      double cos_dist;
      double cost;
      int32_t qcg = shl32(i, CGAIN_SHIFT);
      k = port_pvq_compute_k(qcg, -1, 1, n, beta);
      dist = gain_weight*(qcg - cg)*(qcg - cg);
      dist *= cgain_2;
      if (dist > dist0 && k != 0) continue;
      cos_dist = port_pvq_search_rdo_double(x16, n, k, y_tmp, qcg*(double)cg*cgain_2, pvq_norm_lambda,
       prev_k);
      prev_k = k;
      dist = gain_weight*(qcg - cg)*(qcg - cg) + qcg*(double)cg*(2 - 2*cos_dist);
      dist *= cgain_2;
      cost = dist + pvq_norm_lambda*port_pvq_rate(i, 0, -1, 0, y_tmp, k, n, is_keyframe, pli);
      if (cost <= best_cost) {
        best_cost = cost;
        best_dist = dist;
        qg = i;
        noref = 1;
        best_k = k;
        *itheta = -1;
        *max_theta = 0;
        memcpy(y, y_tmp, n*sizeof(*y));
      }
    }
  }
  k = best_k;
  theta = best_qtheta;
  if (noref) {
    if (qg == 0) skip = PVQ_SKIP_ZERO;
  }
  else {
    if (!is_keyframe && qg == 0) skip = icgr ? PVQ_SKIP_ZERO : PVQ_SKIP_COPY;
    if (qg == icgr && *itheta == 0 && !cfl_enabled) skip = PVQ_SKIP_COPY;
  }
  if (skip) {
    if (skip == PVQ_SKIP_COPY) memcpy(out, r0, n*sizeof(*out));
    else memset(out, 0, n*sizeof(*out));
  }
  else {
    if (noref) gain_offset = 0;
    g = port_gain_expand(shl32(qg, CGAIN_SHIFT) + gain_offset, q0, beta);
    port_pvq_synthesis_partial(out, y, r16, n, noref, g, theta, m, s, qm_inv);
  }
  *vk = k;
  *skip_diff += skip_dist - best_dist;
  if (is_keyframe) return noref ? qg : neg_interleave(qg, icgr);
  return noref ? qg - 1 : neg_interleave(qg + 1, icgr + 1);
}

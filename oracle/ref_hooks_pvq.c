/* oracle/ref_hooks_pvq.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the reference's src/pvq_encoder.c *in place* (by #include, nothing
 * is copied) so that its file-static functions can be reached from the parity
 * tests, exactly as the reference's own src/tests/test_coef_coder.c:25-34
 * reaches statics.  Every oracle_ref_* symbol below is a one-line forwarder.
 */
#include "pvq_encoder.c"

/* src/pvq_encoder.c:93 */
double oracle_ref_pvq_search_rdo_double(const od_val16 *xcoeff, int n, int k,
 od_coeff *ypulse, double g2, double pvq_norm_lambda, int prev_k) {
  return pvq_search_rdo_double(xcoeff, n, k, ypulse, g2, pvq_norm_lambda,
   prev_k);
}

/* src/pvq_encoder.c:247 -- adapt may be NULL when speed > 0 or k == 0. */
double oracle_ref_pvq_rate(int qg, int icgr, int theta, int ts,
 const od_adapt_ctx *adapt, const od_coeff *y0, int k, int n,
 int is_keyframe, int pli, int speed) {
  return od_pvq_rate(qg, icgr, theta, ts, adapt, y0, k, n, is_keyframe, pli,
   speed);
}

/* src/pvq_encoder.c:333 */
int oracle_ref_pvq_theta(od_coeff *out, const od_coeff *x0, const od_coeff *r0,
 int n, int q0, od_coeff *y, int *itheta, int *max_theta, int *vk,
 int beta, double *skip_diff, int nodesync, int is_keyframe, int pli,
 const od_adapt_ctx *adapt, const int16_t *qm, const int16_t *qm_inv,
 double pvq_norm_lambda, int speed) {
  return pvq_theta(out, x0, r0, n, q0, y, itheta, max_theta, vk,
   (od_val16)beta, skip_diff, nodesync, is_keyframe, pli, adapt, qm, qm_inv,
   pvq_norm_lambda, speed);
}

/* Size of the adaptation context, so that ctypes callers can allocate one. */
int oracle_ref_sizeof_adapt_ctx(void) { return (int)sizeof(od_adapt_ctx); }

/* oracle/ref_hooks_encode.c -- TEST INFRASTRUCTURE ONLY.
 * Compiles the reference's src/encode.c in place (see ref_hooks_pvq.c). */
#include "encode.c"

/* od_compute_dist (static, src/encode.c:1180) on a minimal encoder context: only enc->qm,
   enc->use_activity_masking and enc->state.coded_quantizer are read. */
double oracle_ref_compute_dist(od_coeff *x, od_coeff *y, int n, int qm, int use_activity_masking,
 int coded_quantizer) {
  daala_enc_ctx *enc;
  double d;
  enc = (daala_enc_ctx *)calloc(1, sizeof(*enc));
  enc->qm = qm;
  enc->use_activity_masking = use_activity_masking;
  enc->state.coded_quantizer = coded_quantizer;
  d = od_compute_dist(enc, x, y, n);
  free(enc);
  return d;
}

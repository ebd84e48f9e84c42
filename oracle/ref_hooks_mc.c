/* oracle/ref_hooks_mc.c -- TEST INFRASTRUCTURE ONLY.
 * Forwarders that give the reference's MC entry points the od_state they
 * expect (only opt_vtbl.od_copy_nxn is touched, src/mc.c:206) without
 * building a whole encoder. */
#include <string.h>
#include "state.h"
#include "mc.h"
#include "util.h"

void od_mc_predict1fmv8_c(od_state *state, unsigned char *dst,
 const unsigned char *src, int systride, int32_t mvx, int32_t mvy,
 int log_xblk_sz, int log_yblk_sz);
void od_mc_blend_full8_c(unsigned char *dst, int dystride,
 const unsigned char *src[4], int log_xblk_sz, int log_yblk_sz);
void od_mc_blend_full_split8_c(unsigned char *dst, int dystride,
 const unsigned char *src[4], int oc, int s, int log_xblk_sz, int log_yblk_sz);

static od_state *fake_state(void) {
  static od_state st;
  static int init;
  if (!init) {
    int i;
    memset(&st, 0, sizeof(st));
    for (i = 0; i <= OD_LOG_COPYBSIZE_MAX; i++) st.opt_vtbl.od_copy_nxn[i] = OD_COPY_NXN_8_C[i];
    st.opt_vtbl.mc_predict1fmv = od_mc_predict1fmv8_c;
    st.opt_vtbl.mc_blend_full = od_mc_blend_full8_c;
    st.opt_vtbl.mc_blend_full_split = od_mc_blend_full_split8_c;
    for (i = 0; i < 5; i++) st.mc_buf[i] = (unsigned char *)malloc(64*64*2);
    init = 1;
  }
  return &st;
}

/* src/mc.c:94 */
void oracle_ref_mc_predict1fmv8(unsigned char *dst, const unsigned char *src, int systride,
 int32_t mvx, int32_t mvy, int log_xblk_sz, int log_yblk_sz) {
  od_mc_predict1fmv8_c(fake_state(), dst, src, systride, mvx, mvy, log_xblk_sz, log_yblk_sz);
}

/* src/mc.c:2006 (single reference image): up to four predictions + OBMC blend. */
void oracle_ref_mc_predict(unsigned char *dst, int dystride, const unsigned char *src, int systride,
 const int32_t mvx[4], const int32_t mvy[4], int oc, int s, int log_xblk_sz, int log_yblk_sz) {
  od_mc_predict_singleref(fake_state(), dst, dystride, src, systride, mvx, mvy, oc, s,
   log_xblk_sz, log_yblk_sz);
}

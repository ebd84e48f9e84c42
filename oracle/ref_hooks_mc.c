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

/* src/state.c:932 od_state_mc_predict on a caller-supplied MV grid: builds a real od_state
   (od_state_init), loads the reference planes into ref_imgs[PREV] (with od_img_edge_ext),
   fills mv_grid[vy][vx] = {mv, valid, ref = OD_FRAME_PREV} and predicts into ref_imgs[SELF].
   planes: padded frame size (multiple of 64); valid/mv: (nvmvbs+1) x (nhmvbs+1) row-major. */
#include "../include/daala/codec.h"
int oracle_ref_state_mc_predict(int pic_w, int pic_h, const unsigned char *ref_y, const unsigned char *ref_u,
 const unsigned char *ref_v, const unsigned char *valid, const int32_t *mv, unsigned char *out_y,
 unsigned char *out_u, unsigned char *out_v) {
  od_state st;
  daala_info info;
  const unsigned char *refp[3];
  unsigned char *outp[3];
  int pli;
  int vx;
  int vy;
  int y;
  daala_info_init(&info);
  info.pic_width = pic_w;
  info.pic_height = pic_h;
  info.nplanes = 3;
  info.plane_info[0].xdec = info.plane_info[0].ydec = 0;
  info.plane_info[1].xdec = info.plane_info[1].ydec = 1;
  info.plane_info[2].xdec = info.plane_info[2].ydec = 1;
  info.bitdepth_mode = OD_BITDEPTH_MODE_8;
  info.full_precision_references = 0;
  info.timebase_numerator = 30; info.timebase_denominator = 1; info.frame_duration = 1;
  info.pixel_aspect_numerator = 1; info.pixel_aspect_denominator = 1;
  info.keyframe_rate = 256;
  if (od_state_init(&st, &info) < 0) return -1;
  st.opt_vtbl.mc_predict1fmv = od_mc_predict1fmv8_c;
  st.opt_vtbl.mc_blend_full = od_mc_blend_full8_c;
  st.opt_vtbl.mc_blend_full_split = od_mc_blend_full_split8_c;
  st.ref_imgi[OD_FRAME_PREV] = 0;
  st.ref_imgi[OD_FRAME_SELF] = 1;
  refp[0] = ref_y; refp[1] = ref_u; refp[2] = ref_v;
  outp[0] = out_y; outp[1] = out_u; outp[2] = out_v;
  for (pli = 0; pli < 3; pli++) {
    daala_image_plane *ip = st.ref_imgs[0].planes + pli;
    int w = st.frame_width >> ip->xdec;
    int h = st.frame_height >> ip->ydec;
    for (y = 0; y < h; y++) memcpy(ip->data + y*ip->ystride, refp[pli] + y*w, w);
  }
  od_img_edge_ext(st.ref_imgs + 0);
  for (vy = 0; vy <= st.nvmvbs; vy++) {
    for (vx = 0; vx <= st.nhmvbs; vx++) {
      od_mv_grid_pt *g = st.mv_grid[vy] + vx;
      int i = vy*(st.nhmvbs + 1) + vx;
      g->valid = valid[i];
      g->mv[0] = mv[2*i];
      g->mv[1] = mv[2*i + 1];
      g->ref = OD_FRAME_PREV;
    }
  }
  od_state_mc_predict(&st, st.ref_imgs + 1);
  for (pli = 0; pli < 3; pli++) {
    daala_image_plane *ip = st.ref_imgs[1].planes + pli;
    int w = st.frame_width >> ip->xdec;
    int h = st.frame_height >> ip->ydec;
    for (y = 0; y < h; y++) memcpy(outp[pli] + y*w, ip->data + y*ip->ystride, w);
  }
  od_state_clear(&st);
  return 0;
}

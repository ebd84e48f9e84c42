/* oracle/ref_hooks_mcenc.c -- TEST INFRASTRUCTURE ONLY.
 * Compiles the reference's src/mcenc.c in place (see ref_hooks_pvq.c). */
#include "mcenc.c"

/* od_mv_est_bma_sad (static, src/mcenc.c:2224) for a list of candidates: a real od_state inside a zeroed
   encoder context (od_state_init), the reference planes loaded into ref_imgs[PREV] with od_img_edge_ext,
   the current picture as enc->curr_img.  planes: padded frame size (multiples of 64), row stride = plane
   width.  jobs: count x {bx, by, mvx, mvy, log_mvb_sz}. */
#include <stdlib.h>
#include <string.h>
int oracle_ref_bma_sad(int pic_w, int pic_h, const unsigned char *cur_y, const unsigned char *cur_u,
 const unsigned char *cur_v, const unsigned char *ref_y, const unsigned char *ref_u, const unsigned char *ref_v,
 int use_chroma, const int32_t *jobs, int count, int32_t *out) {
  daala_enc_ctx *enc;
  daala_info info;
  daala_image cur;
  od_mv_est_ctx est;
  const unsigned char *refp[3];
  const unsigned char *curp[3];
  int pli;
  int y;
  int i;
  enc = (daala_enc_ctx *)calloc(1, sizeof(*enc));
  if (enc == NULL) return -1;
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
  if (od_state_init(&enc->state, &info) < 0) { free(enc); return -1; }
  od_enc_opt_vtbl_init_c(enc);
  enc->state.ref_imgi[OD_FRAME_PREV] = 0;
  refp[0] = ref_y; refp[1] = ref_u; refp[2] = ref_v;
  curp[0] = cur_y; curp[1] = cur_u; curp[2] = cur_v;
  memset(&cur, 0, sizeof(cur));
  cur.nplanes = 3;
  cur.width = enc->state.frame_width;
  cur.height = enc->state.frame_height;
  for (pli = 0; pli < 3; pli++) {
    daala_image_plane *ip = enc->state.ref_imgs[0].planes + pli;
    int w = enc->state.frame_width >> ip->xdec;
    int h = enc->state.frame_height >> ip->ydec;
    for (y = 0; y < h; y++) memcpy(ip->data + y*ip->ystride, refp[pli] + y*w, w);
    cur.planes[pli].data = (unsigned char *)curp[pli];
    cur.planes[pli].xdec = ip->xdec;
    cur.planes[pli].ydec = ip->ydec;
    cur.planes[pli].xstride = 1;
    cur.planes[pli].ystride = w;
    cur.planes[pli].bitdepth = 8;
  }
  od_img_edge_ext(enc->state.ref_imgs + 0);
  enc->curr_img = &cur;
  memset(&est, 0, sizeof(est));
  est.enc = enc;
  est.flags = use_chroma ? OD_MC_USE_CHROMA : 0;
  for (i = 0; i < count; i++) {
    const int32_t *j = jobs + 5*i;
    out[i] = od_mv_est_bma_sad(&est, OD_FRAME_PREV, j[0], j[1], j[2], j[3], j[4]);
  }
  od_state_clear(&enc->state);
  free(enc);
  return 0;
}

/* od_mv_est_sad (static, src/mcenc.c:2267): the OBMC-based cost of MV-grid blocks -- od_state_pred_block_from_setup
   of every plane from a caller-supplied MV grid + od_enc_sad, chroma >> OD_MC_CHROMA_SCALE.  valid / mv:
   (nvmvbs+1) x (nhmvbs+1) row-major (mv in 1/8 luma pixel); jobs: count x {vx, vy, oc, s, log_mvb_sz}. */
int oracle_ref_mv_est_sad(int pic_w, int pic_h, const unsigned char *cur_y, const unsigned char *cur_u,
 const unsigned char *cur_v, const unsigned char *ref_y, const unsigned char *ref_u, const unsigned char *ref_v,
 const unsigned char *valid, const int32_t *mv, int use_chroma, const int32_t *jobs, int count, int32_t *out) {
  daala_enc_ctx *enc;
  daala_info info;
  daala_image cur;
  od_mv_est_ctx est;
  const unsigned char *refp[3];
  const unsigned char *curp[3];
  int pli;
  int y;
  int i;
  int vx;
  int vy;
  enc = (daala_enc_ctx *)calloc(1, sizeof(*enc));
  if (enc == NULL) return -1;
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
  if (od_state_init(&enc->state, &info) < 0) { free(enc); return -1; }
  od_enc_opt_vtbl_init_c(enc);
  enc->state.ref_imgi[OD_FRAME_PREV] = 0;
  refp[0] = ref_y; refp[1] = ref_u; refp[2] = ref_v;
  curp[0] = cur_y; curp[1] = cur_u; curp[2] = cur_v;
  memset(&cur, 0, sizeof(cur));
  cur.nplanes = 3;
  cur.width = enc->state.frame_width;
  cur.height = enc->state.frame_height;
  for (pli = 0; pli < 3; pli++) {
    daala_image_plane *ip = enc->state.ref_imgs[0].planes + pli;
    int w = enc->state.frame_width >> ip->xdec;
    int h = enc->state.frame_height >> ip->ydec;
    for (y = 0; y < h; y++) memcpy(ip->data + y*ip->ystride, refp[pli] + y*w, w);
    cur.planes[pli].data = (unsigned char *)curp[pli];
    cur.planes[pli].xdec = ip->xdec;
    cur.planes[pli].ydec = ip->ydec;
    cur.planes[pli].xstride = 1;
    cur.planes[pli].ystride = w;
    cur.planes[pli].bitdepth = 8;
  }
  od_img_edge_ext(enc->state.ref_imgs + 0);
  enc->curr_img = &cur;
  for (vy = 0; vy <= enc->state.nvmvbs; vy++) {
    for (vx = 0; vx <= enc->state.nhmvbs; vx++) {
      od_mv_grid_pt *g = enc->state.mv_grid[vy] + vx;
      int k = vy*(enc->state.nhmvbs + 1) + vx;
      g->valid = valid[k];
      g->mv[0] = mv[2*k];
      g->mv[1] = mv[2*k + 1];
      g->ref = OD_FRAME_PREV;
    }
  }
  memset(&est, 0, sizeof(est));
  est.enc = enc;
  est.flags = use_chroma ? OD_MC_USE_CHROMA : 0;
  est.compute_distortion = od_enc_sad;
  for (i = 0; i < count; i++) {
    const int32_t *j = jobs + 5*i;
    out[i] = od_mv_est_sad(&est, j[0], j[1], j[2], j[3], j[4]);
  }
  od_state_clear(&enc->state);
  free(enc);
  return 0;
}

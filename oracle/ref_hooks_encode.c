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

/* The WHOLE reference encoder on one 4:2:0 8-bit picture coded as a keyframe, through the public API
   (include/daala/daalaenc.h), exporting what its RDO decided: the block-size map (one byte per 8x8
   luma unit, [nvsb*8][nhsb*8]), the deringing level per superblock ([nvsb][nhsb]) and the size and a
   checksum of the coded packet.  Gives tests real block-size decisions instead of synthetic quadtrees. */
int oracle_ref_encode_keyframe(int w, int h, unsigned char *y, unsigned char *u, unsigned char *v,
 int quant, int complexity, unsigned char *bsize_out, unsigned char *dering_out, long *packet_bytes,
 unsigned *packet_sum) {
  daala_info info;
  daala_enc_ctx *enc;
  daala_image img;
  daala_packet op;
  int pli;
  int i;
  int j;
  int ret;
  daala_info_init(&info);
  info.pic_width = w;
  info.pic_height = h;
  info.timebase_numerator = 30;
  info.timebase_denominator = 1;
  info.frame_duration = 1;
  info.pixel_aspect_numerator = 1;
  info.pixel_aspect_denominator = 1;
  info.nplanes = 3;
  info.plane_info[0].xdec = info.plane_info[0].ydec = 0;
  info.plane_info[1].xdec = info.plane_info[1].ydec = 1;
  info.plane_info[2].xdec = info.plane_info[2].ydec = 1;
  info.keyframe_rate = 1;
  enc = daala_encode_create(&info);
  if (enc == NULL) return -1;
  daala_encode_ctl(enc, OD_SET_QUANT, &quant, sizeof(quant));
  daala_encode_ctl(enc, OD_SET_COMPLEXITY, &complexity, sizeof(complexity));
  img.nplanes = 3;
  img.width = w;
  img.height = h;
  for (pli = 0; pli < 3; pli++) {
    img.planes[pli].data = pli == 0 ? y : pli == 1 ? u : v;
    img.planes[pli].xdec = img.planes[pli].ydec = pli > 0;
    img.planes[pli].xstride = 1;
    img.planes[pli].ystride = pli == 0 ? w : (w + 1) >> 1;
    img.planes[pli].bitdepth = 8;
  }
  ret = daala_encode_img_in(enc, &img, 1);
  if (ret < 0) {
    daala_encode_free(enc);
    return ret;
  }
  ret = daala_encode_packet_out(enc, 1, &op);
  if (ret <= 0) {
    daala_encode_free(enc);
    return -2;
  }
  *packet_bytes = op.bytes;
  *packet_sum = 0;
  for (i = 0; i < op.bytes; i++) *packet_sum = *packet_sum*31 + op.packet[i];
  for (i = 0; i < enc->state.nvsb*8; i++) {
    for (j = 0; j < enc->state.nhsb*8; j++) {
      bsize_out[i*enc->state.nhsb*8 + j] = enc->state.bsize[i*enc->state.bstride + j];
    }
  }
  for (i = 0; i < enc->state.nvsb*enc->state.nhsb; i++) {
    dering_out[i] = enc->state.dering_level ? enc->state.dering_level[i] : 0;
  }
  daala_encode_free(enc);
  return 0;
}

/* Multi-frame variant (keyframe + P frames) returning every packet's size and checksum: the
   reference side of the drop-in link test (dropin_main.c). */
#define ENCODE_FRAMES_NAME oracle_ref_encode_frames
#include "encode_frames.inc"

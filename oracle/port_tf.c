/* oracle/port_tf.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 * Restates src/tf.c (time/frequency resolution switching built on the 2x2 Haar
 * kernel of src/tf.h:35) and src/intra.c (od_hv_intra_pred :37,
 * od_resample_luma_coeffs :72 with OD_CFL_SCALING4 :65). */
#include "port.h"

#define RSH1(a) ((int)((a) + (int)((unsigned)(a) >> 31)) >> 1)  /* OD_DCT_RSHIFT(a, 1) */

/* OD_HAAR_KERNEL(ll, lh, hl, hh) */
static void haar4(od_coeff *ll, od_coeff *lh, od_coeff *hl, od_coeff *hh) {
  od_coeff t;
  *ll += *hl;
  *hh -= *lh;
  t = (*ll - *hh) >> 1;
  *lh = t - *lh;
  *hl = t - *hl;
  *ll -= *lh;
  *hh += *hl;
}

/* src/tf.c:38 */
void port_tf_up_h_lp(od_coeff *dst, int dstride, const od_coeff *src, int sstride, int dx, int n) {
  int x;
  int y;
  for (y = 0; y < n; y++) {
    for (x = 0; x < n >> 1; x++) {
      od_coeff ll = src[y*sstride + x];
      od_coeff lh = ll - src[y*sstride + x + dx];
      int sw = x & 1;
      ll -= RSH1(lh);
      dst[y*dstride + 2*x + sw] = ll;
      dst[y*dstride + 2*x + 1 - sw] = lh;
    }
  }
}

/* src/tf.c:60 */
void port_tf_up_v_lp(od_coeff *dst, int dstride, const od_coeff *src, int sstride, int dy, int n) {
  int x;
  int y;
  for (y = 0; y < n >> 1; y++) {
    int sw = y & 1;
    for (x = 0; x < n; x++) {
      od_coeff ll = src[y*sstride + x];
      od_coeff hl = ll - src[(y + dy)*sstride + x];
      ll -= RSH1(hl);
      dst[(2*y + sw)*dstride + x] = ll;
      dst[(2*y + 1 - sw)*dstride + x] = hl;
    }
  }
}

/* src/tf.c:82 (dx = dy = n gives od_tf_up_hv's 2x2 merge restricted to the LF quarter) */
void port_tf_up_hv_lp(od_coeff *dst, int dstride, const od_coeff *src, int sstride, int dx, int dy, int n) {
  int x;
  int y;
  for (y = 0; y < n >> 1; y++) {
    int vs = y & 1;
    for (x = 0; x < n >> 1; x++) {
      od_coeff ll = src[y*sstride + x];
      od_coeff lh = src[y*sstride + x + dx];
      od_coeff hl = src[(y + dy)*sstride + x];
      od_coeff hh = src[(y + dy)*sstride + x + dx];
      int hs = x & 1;
      haar4(&ll, &hl, &lh, &hh);
      dst[(2*y + vs)*dstride + 2*x + hs] = ll;
      dst[(2*y + vs)*dstride + 2*x + 1 - hs] = lh;
      dst[(2*y + 1 - vs)*dstride + 2*x + hs] = hl;
      dst[(2*y + 1 - vs)*dstride + 2*x + 1 - hs] = hh;
    }
  }
}

/* src/tf.c:112 */
void port_tf_up_hv(od_coeff *dst, int dstride, const od_coeff *src, int sstride, int n) {
  int x;
  int y;
  for (y = 0; y < n; y++) {
    int vs = y & 1;
    for (x = 0; x < n; x++) {
      od_coeff ll = src[y*sstride + x];
      od_coeff lh = src[y*sstride + x + n];
      od_coeff hl = src[(y + n)*sstride + x];
      od_coeff hh = src[(y + n)*sstride + x + n];
      int hs = x & 1;
      haar4(&ll, &hl, &lh, &hh);
      dst[(2*y + vs)*dstride + 2*x + hs] = ll;
      dst[(2*y + vs)*dstride + 2*x + 1 - hs] = lh;
      dst[(2*y + 1 - vs)*dstride + 2*x + hs] = hl;
      dst[(2*y + 1 - vs)*dstride + 2*x + 1 - hs] = hh;
    }
  }
}

/* src/tf.c:142 */
void port_tf_down_hv(od_coeff *dst, int dstride, const od_coeff *src, int sstride, int n) {
  int x;
  int y;
  n >>= 1;
  for (y = 0; y < n; y++) {
    int vs = y & 1;
    for (x = 0; x < n; x++) {
      int hs = x & 1;
      od_coeff ll = src[(2*y + vs)*sstride + 2*x + hs];
      od_coeff lh = src[(2*y + vs)*sstride + 2*x + 1 - hs];
      od_coeff hl = src[(2*y + 1 - vs)*sstride + 2*x + hs];
      od_coeff hh = src[(2*y + 1 - vs)*sstride + 2*x + 1 - hs];
      haar4(&ll, &lh, &hl, &hh);
      dst[y*dstride + x] = ll;
      dst[y*dstride + x + n] = lh;
      dst[(y + n)*dstride + x] = hl;
      dst[(y + n)*dstride + x + n] = hh;
    }
  }
}

/* src/intra.c:37.  bsize: one byte per 8x8 luma unit; bx, by in 4x4 units. */
void port_hv_intra_pred(od_coeff *pred, const od_coeff *d, int w, int bx, int by,
 const unsigned char *bsize, int bstride, int bs) {
  int n = 4 << bs;
  const od_coeff *t = d + (by << 2)*w + (bx << 2);
  int top = by > 0 && bsize[((by - 1) >> 1)*bstride + (bx >> 1)] == bs;
  int left = bx > 0 && bsize[(by >> 1)*bstride + ((bx - 1) >> 1)] == bs;
  double g1 = 0;
  double g2 = 0;
  int i;
  if (top) for (i = 1; i < 4; i++) g1 += t[-n*w + i]*(double)t[-n*w + i];
  if (left) for (i = 1; i < 4; i++) g2 += t[-n + i*w]*(double)t[-n + i*w];
  if (top) for (i = 4; i < n; i++) pred[i] = t[-n*w + i];
  if (left) for (i = 4; i < n; i++) pred[i*n] = t[-n + i*w];
  if (g1 > g2) {
    if (top) for (i = 1; i < 4; i++) pred[i] = t[-n*w + i];
  }
  else {
    if (left) for (i = 1; i < 4; i++) pred[i*n] = t[-n + i*w];
  }
}

/* src/intra.c:72 for 4:2:0 (xdec == ydec == 1).  bs: chroma log2(n) - 2;
   luma_is_4x4: the luma block size entry of this area is OD_BLOCK_4X4. */
void port_resample_luma_coeffs_420(od_coeff *chroma_pred, int cpstride, const od_coeff *decoded_luma,
 int dlstride, int bs, int luma_is_4x4) {
  static const int scaling4[4][4] = {
    {128, 128, 100, 36}, {128, 80, 71, 35}, {100, 71, 35, 31}, {36, 35, 31, 18}};
  int n = 4 << bs;
  int i;
  int j;
  if (luma_is_4x4) {
    port_tf_up_hv_lp(chroma_pred, cpstride, decoded_luma, dlstride, n, n, n);
    for (i = 0; i < 4; i++) {
      for (j = 0; j < 4; j++) {
        chroma_pred[i*cpstride + j] = (scaling4[j][i]*chroma_pred[i*cpstride + j] + 64) >> 7;
      }
    }
  }
  else {
    for (i = 0; i < n; i++) for (j = 0; j < n; j++) chroma_pred[i*cpstride + j] = decoded_luma[i*dlstride + j];
  }
}

/* oracle/port_mc.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 * Restates the 8-bit motion-compensation primitives of src/mc.c
 * (od_mc_predict1fmv8_c :94, od_mc_blend_full8_c :352,
 * od_mc_blend_full_split8_c :1104 with od_mc_setup_s_split :1056,
 * od_mc_predict_singleref :1965) and the block-matching metrics of
 * src/mcenc.c (od_mc_compute_sad8_c :1333, od_mc_compute_satd8 :1467 and the
 * sum-of-8x8 rule :1517). */
#include <stdlib.h>
#include <string.h>
#include "port.h"

/* 6-tap windowed-sinc bank, 1/8-pel phases, 7-bit coefficients (src/mc.c:66-77). */
static const int16_t SUBPEL[8][6] = {
  {0, 0, 128, 0, 0, 0}, {1, -9, 122, 18, -5, 1}, {3, -15, 112, 37, -11, 2}, {3, -18, 97, 58, -15, 3},
  {4, -20, 80, 80, -20, 4}, {3, -15, 58, 97, -18, 3}, {2, -11, 37, 112, -15, 3}, {1, -5, 18, 122, -9, 1}};

static unsigned char clamp255(int v) { return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* dst: n_x * n_y bytes, row stride n_x.  mv in 1/8 pel. */
void port_mc_predict1fmv8(unsigned char *dst, const unsigned char *src, int systride, int32_t mvx,
 int32_t mvy, int log_xblk_sz, int log_yblk_sz) {
  int nx = 1 << log_xblk_sz;
  int ny = 1 << log_yblk_sz;
  int fxi = mvx & 7;
  int fyi = mvy & 7;
  const unsigned char *p = src + (mvx >> 3) + (mvy >> 3)*systride;
  int16_t buf[(64 + 5)*64];
  int i;
  int j;
  int k;
  if (!fxi && !fyi) {
    for (j = 0; j < ny; j++) memcpy(dst + j*nx, p + j*systride, nx);
    return;
  }
  /* horizontal stage over rows -2 .. ny+2, biased by -128 << 7 */
  for (j = -2; j < ny + 3; j++) {
    const unsigned char *row = p + j*systride;
    int16_t *b = buf + (j + 2)*nx;
    for (i = 0; i < nx; i++) {
      if (fxi) {
        int sum = 0;
        for (k = 0; k < 6; k++) sum += row[i + k - 2]*SUBPEL[fxi][k];
        b[i] = (int16_t)(sum - (128 << 7));
      }
      else b[i] = (int16_t)((row[i] << 7) - (128 << 7));
    }
  }
  for (j = 0; j < ny; j++) {
    const int16_t *b = buf + (j + 2)*nx;
    for (i = 0; i < nx; i++) {
      if (fyi) {
        int sum = 0;
        for (k = 0; k < 6; k++) sum += b[i + (k - 2)*nx]*SUBPEL[fyi][k];
        dst[j*nx + i] = clamp255((sum + (1 << 13) + (128 << 14)) >> 14);
      }
      else dst[j*nx + i] = clamp255((b[i] + (1 << 6) + (128 << 7)) >> 7);
    }
  }
}

/* Bilinear OBMC blend of four predictions (corner order is rotational: 0 top-left,
   1 top-right, 2 bottom-right, 3 bottom-left). */
void port_mc_blend_full8(unsigned char *dst, int dystride, const unsigned char *src[4],
 int log_xblk_sz, int log_yblk_sz) {
  int nx = 1 << log_xblk_sz;
  int ny = 1 << log_yblk_sz;
  int l2 = log_xblk_sz + log_yblk_sz;
  int i;
  int j;
  for (j = 0; j < ny; j++) {
    for (i = 0; i < nx; i++) {
      int a = src[0][j*nx + i];
      int b = src[3][j*nx + i];
      a = (a << log_xblk_sz) + (src[1][j*nx + i] - a)*i;
      b = (b << log_xblk_sz) + (src[2][j*nx + i] - b)*i;
      dst[j*dystride + i] = (unsigned char)(((a << log_yblk_sz) + (b - a)*j + (1 << (l2 - 1))) >> l2);
    }
  }
}

/* Weights for blocks with unsplit edges: the weight of the vertex missing on an
   unsplit edge is halved and the half moved to the outside corner oc. */
void port_mc_blend_full_split8(unsigned char *dst, int dystride, const unsigned char *src[4], int oc,
 int s, int log_xblk_sz, int log_yblk_sz) {
  int nx = 1 << log_xblk_sz;
  int ny = 1 << log_yblk_sz;
  int l2p1 = log_xblk_sz + log_yblk_sz + 1;
  int s0[4];
  int dsdi[4];
  int dsdj[4];
  int dd[4];
  int e;
  int i;
  int j;
  s0[0] = 2 << (l2p1 - 1); s0[1] = s0[2] = s0[3] = 0;
  dsdi[0] = -(2 << log_xblk_sz); dsdi[1] = 2 << log_xblk_sz; dsdi[2] = dsdi[3] = 0;
  dsdj[0] = -(2 << log_yblk_sz); dsdj[1] = dsdj[2] = 0; dsdj[3] = 2 << log_yblk_sz;
  dd[0] = dd[2] = 2; dd[1] = dd[3] = -2;
  for (e = 0; e < 2; e++) {
    if (!(s & (1 << e))) {
      int k = (oc + (e ? 3 : 1)) & 3;
      s0[k] >>= 1; s0[oc] += s0[k];
      dsdi[k] >>= 1; dsdi[oc] += dsdi[k];
      dsdj[k] >>= 1; dsdj[oc] += dsdj[k];
      dd[k] >>= 1; dd[oc] += dd[k];
    }
  }
  for (j = 0; j < ny; j++) {
    for (i = 0; i < nx; i++) {
      int a = src[0][j*nx + i];
      int acc = a << l2p1;
      int k;
      for (k = 1; k < 4; k++) acc += (src[k][j*nx + i] - a)*(s0[k] + j*dsdj[k] + i*(dsdi[k] + j*dd[k]));
      dst[j*dystride + i] = (unsigned char)((acc + (1 << (l2p1 - 1))) >> l2p1);
    }
  }
}

/* od_mc_predict_singleref + od_mc_blend (s == 3: plain bilinear). */
void port_mc_predict(unsigned char *dst, int dystride, const unsigned char *src, int systride,
 const int32_t mvx[4], const int32_t mvy[4], int oc, int s, int log_xblk_sz, int log_yblk_sz) {
  unsigned char buf[4][64*64];
  const unsigned char *pred[4];
  int k;
  for (k = 0; k < 4; k++) {
    port_mc_predict1fmv8(buf[k], src, systride, mvx[k], mvy[k], log_xblk_sz, log_yblk_sz);
    pred[k] = buf[k];
  }
  if (s == 3) port_mc_blend_full8(dst, dystride, pred, log_xblk_sz, log_yblk_sz);
  else port_mc_blend_full_split8(dst, dystride, pred, oc, s, log_xblk_sz, log_yblk_sz);
}

int32_t port_mc_compute_sad8(const unsigned char *src, int systride, const unsigned char *ref,
 int dystride, int w, int h) {
  int32_t ret = 0;
  int i;
  int j;
  for (j = 0; j < h; j++) for (i = 0; i < w; i++) ret += abs(ref[j*dystride + i] - src[j*systride + i]);
  return ret;
}

static int32_t satd_square(int ln, const unsigned char *src, int systride, const unsigned char *ref,
 int rystride) {
  int n = 1 << ln;
  int32_t w[64];
  int32_t satd = 0;
  int i;
  int j;
  int len;
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) w[i*n + j] = src[i*systride + j] - ref[i*rystride + j];
  /* separable Walsh-Hadamard butterflies; the sum of magnitudes does not
     depend on the output ordering */
  for (len = 1; len < n; len <<= 1) {
    for (i = 0; i < n; i++) {
      for (j = 0; j < n; j++) {
        if (!(j & len)) {
          int32_t a = w[i*n + j];
          int32_t b = w[i*n + j + len];
          w[i*n + j] = a + b;
          w[i*n + j + len] = a - b;
        }
      }
    }
  }
  for (len = 1; len < n; len <<= 1) {
    for (j = 0; j < n; j++) {
      for (i = 0; i < n; i++) {
        if (!(i & len)) {
          int32_t a = w[i*n + j];
          int32_t b = w[(i + len)*n + j];
          w[i*n + j] = a + b;
          w[(i + len)*n + j] = a - b;
        }
      }
    }
  }
  for (i = 0; i < n*n; i++) satd += abs(w[i]);
  return (satd + (1 << ln >> 1)) >> ln;
}

/* 4x4: one transform; larger: sum of the 8x8 SATDs (src/mcenc.c:1517-1537). */
int32_t port_mc_compute_satd8(int ln, const unsigned char *src, int systride, const unsigned char *ref,
 int rystride) {
  int n = 1 << ln;
  int32_t satd = 0;
  int i;
  int j;
  if (ln == 2) return satd_square(2, src, systride, ref, rystride);
  for (i = 0; i < n; i += 8) {
    for (j = 0; j < n; j += 8) satd += satd_square(3, src + i*systride + j, systride, ref + i*rystride + j, rystride);
  }
  return satd;
}

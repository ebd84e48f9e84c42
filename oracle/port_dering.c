/* oracle/port_dering.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 *
 * Plain-C restatement of Daala's directional deringing filter for one superblock
 * (reference src/dering.c: od_dir_find8 :61, od_filter_dering_direction_c :132,
 * od_filter_dering_orthogonal_c :172, od_compute_thresh :237, od_dering :252 in its
 * DAALA_ODINTRIN form).  Oracle for SURVEY.md 8(f) rank 1 -- the next row of the hot path; no CUDA
 * twin exists yet.  Pinned against the reference build by tests/test_oracle_dering.py.
 *
 * Shape of the algorithm: every 8x8 luma block gets the direction (of 8) along which its pixels vary
 * least; the superblock is smoothed along that direction with taps 3,2,1 on either side, then across
 * it with four unit taps, each tap gated by a threshold on its difference to the centre sample.  The
 * threshold grows with the block's directional contrast and is zero for blocks whose whole
 * lapped neighbourhood was skipped.  Samples outside the frame read as a huge value, so their
 * differences never pass a threshold.
 */
#include <stdlib.h>
#include "port.h"

#define DERING_BORDER 3
#define DERING_PITCH (64 + 2*DERING_BORDER)
#define DERING_OUTSIDE 30000
#define DERING_MAXB 8   /* 8x8 (luma) or 4x4 (4:2:0 chroma) blocks per superblock side */

/* Step k = 1..3 along direction d as (rows, columns); the opposite side is the negation.
   (OD_DIRECTION_OFFSETS_TABLE, src/dering.c:39-48: 0 = 45 degrees up-right, 2 = horizontal, 6 = vertical.) */
static const signed char kStep[8][3][2] = {
  {{-1, 1}, {-2, 2}, {-3, 3}}, {{0, 1}, {-1, 2}, {-1, 3}}, {{0, 1}, {0, 2}, {0, 3}}, {{0, 1}, {1, 2}, {1, 3}},
  {{1, 1}, {2, 2}, {3, 3}},    {{1, 0}, {2, 1}, {3, 1}},   {{1, 0}, {2, 0}, {3, 0}}, {{1, 0}, {2, -1}, {3, -1}},
};

/* Which line of direction d the pixel (i, j) of an 8x8 block lies on (src/dering.c:79-86). */
static int line_of(int d, int i, int j) {
  switch (d) {
    case 0: return i + j;
    case 1: return i + j/2;
    case 2: return i;
    case 3: return 3 + i - j/2;
    case 4: return 7 + i - j;
    case 5: return 3 - i/2 + j;
    case 6: return j;
    default: return i/2 + j;
  }
}

/* Direction whose lines explain the block best: maximise sum over lines of (line sum)^2 / (line length),
   scaled by 840 = lcm(1..8) so that it stays integral; *var = contrast against the orthogonal direction. */
int port_dering_find_direction(const int16_t *img, int stride, int32_t *var, int coeff_shift) {
  int32_t cost[8];
  int d, i, j, best = 0;
  int32_t best_cost = 0;
  for (d = 0; d < 8; d++) {
    int sum[15] = {0}, len[15] = {0};
    for (i = 0; i < 8; i++) {
      for (j = 0; j < 8; j++) {
        int l = line_of(d, i, j);
        sum[l] += img[i*stride + j] >> coeff_shift;
        len[l]++;
      }
    }
    cost[d] = 0;
    for (i = 0; i < 15; i++) if (len[i]) cost[d] += sum[i]*sum[i]*(840/len[i]);
  }
  for (d = 0; d < 8; d++) {
    if (cost[d] > best_cost) {
      best_cost = cost[d];
      best = d;
    }
  }
  *var = (best_cost - cost[(best + 4) & 7]) >> 10;
  return best;
}

/* round(256 * clamp(1.08 * (sqrt(2) * 2^(k+8) / 65536)^0.16, 0.5, 3)), k = log2 of the variance
   (OD_THRESH_TABLE_Q8, src/dering.c:225) */
static const int16_t kThreshQ8[18] = {128, 134, 150, 168, 188, 210, 234, 262, 292, 327, 365, 408, 455, 509, 569,
                                      635, 710, 768};

static int ilog_u(unsigned v) {
  int n = 0;
  while (v) { n++; v >>= 1; }
  return n;
}

/* y, x: superblock origin inside 16-bit planes.  nhb x nvb blocks of (8 >> xdec)^2 samples.  dir[8][8] is
   written for pli == 0 and read for chroma.  bskip: one flag per 4x4 luma block, origin at this
   superblock.  Same contract as od_dering (src/dering.h:63-68) without the vtable. */
void port_dering(int16_t *y, int ystride, const int16_t *x, int xstride, int nhb, int nvb, int sbx, int sby,
 int nhsb, int nvsb, int xdec, int dir[8][8], int pli, const unsigned char *bskip, int skip_stride, int threshold,
 int overlap, int coeff_shift) {
  static int16_t padded[DERING_PITCH*DERING_PITCH];
  int16_t *in = padded + DERING_BORDER*DERING_PITCH + DERING_BORDER;
  int thresh[DERING_MAXB][DERING_MAXB];
  const int lb = 3 - xdec, n = 1 << lb;
  const int rows = nvb << lb, cols = nhb << lb;
  int i, j, k, bx, by;
  /* window with a 3-sample apron; the apron exists only towards neighbouring superblocks */
  for (i = 0; i < DERING_PITCH*DERING_PITCH; i++) padded[i] = DERING_OUTSIDE;
  for (i = sby ? -DERING_BORDER : 0; i < rows + (sby != nvsb - 1 ? DERING_BORDER : 0); i++)
    for (j = sbx ? -DERING_BORDER : 0; j < cols + (sbx != nhsb - 1 ? DERING_BORDER : 0); j++)
      in[i*DERING_PITCH + j] = x[i*xstride + j];
  for (by = 0; by < nvb; by++) {
    for (bx = 0; bx < nhb; bx++) {
      if (pli == 0) {
        int32_t var;
        int v;
        dir[by][bx] = port_dering_find_direction(x + 8*by*xstride + 8*bx, xstride, &var, coeff_shift);
        v = var >> 6;
        if (v > 32767) v = 32767;
        thresh[by][bx] = (threshold*kThreshQ8[ilog_u((unsigned)v)] + 128) >> 8;
      }
      else thresh[by][bx] = threshold;
    }
  }
  /* a block whose own 4x4 units -- and, with overlap, the ring of units lapped into it -- were all skipped
     is left untouched */
  for (by = 0; by < nvb; by++) {
    for (bx = 0; bx < nhb; bx++) {
      int u0 = 0, v0 = 0, u1 = 2 >> xdec, v1 = 2 >> xdec, all = 1;
      if (overlap) {
        u0 -= sbx != 0;
        v0 -= sby != 0;
        u1 += sbx != nhsb - 1;
        v1 += sby != nvsb - 1;
      }
      for (i = v0; i < v1; i++)
        for (j = u0; j < u1; j++)
          all = all && bskip[((by << 1 >> xdec) + i)*skip_stride + (bx << 1 >> xdec) + j];
      if (all) thresh[by][bx] = 0;
    }
  }
  /* pass 1: along the direction */
  for (by = 0; by < nvb; by++) {
    for (bx = 0; bx < nhb; bx++) {
      const int t = thresh[by][bx], d = dir[by][bx];
      for (i = by*n; i < (by + 1)*n; i++) {
        for (j = bx*n; j < (bx + 1)*n; j++) {
          const int16_t c = in[i*DERING_PITCH + j];
          int16_t acc = 0;
          for (k = 0; k < 3; k++) {
            const int o = kStep[d][k][0]*DERING_PITCH + kStep[d][k][1];
            const int16_t a = (int16_t)(in[i*DERING_PITCH + j + o] - c);
            const int16_t b = (int16_t)(in[i*DERING_PITCH + j - o] - c);
            if (abs(a) < t) acc = (int16_t)(acc + (3 - k)*a);
            if (abs(b) < t) acc = (int16_t)(acc + (3 - k)*b);
          }
          y[i*ystride + j] = (int16_t)(c + ((acc + 8) >> 4));
        }
      }
    }
  }
  /* pass 2: across it, on the result of pass 1 (the apron keeps the unfiltered input) */
  for (i = 0; i < rows; i++) for (j = 0; j < cols; j++) in[i*DERING_PITCH + j] = y[i*ystride + j];
  for (by = 0; by < nvb; by++) {
    for (bx = 0; bx < nhb; bx++) {
      const int t = thresh[by][bx], d = dir[by][bx];
      const int o = (d > 0 && d < 4) ? DERING_PITCH : 1;
      for (i = by*n; i < (by + 1)*n; i++) {
        for (j = bx*n; j < (bx + 1)*n; j++) {
          const int16_t c = in[i*DERING_PITCH + j];
          int moved = abs(c - x[i*xstride + j]);
          int16_t lim = (int16_t)(t/3 + moved < t ? t/3 + moved : t);
          int16_t acc = 0;
          for (k = 0; k < 4; k++) {
            static const int kTap[4] = {1, -1, 2, -2};
            const int16_t p = (int16_t)(in[i*DERING_PITCH + j + kTap[k]*o] - c);
            if (abs(p) < lim) acc = (int16_t)(acc + p);
          }
          y[i*ystride + j] = (int16_t)(c + ((3*acc + 8) >> 4));
        }
      }
    }
  }
}

/* oracle/port_dct.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 * Restates src/dct.c: 1-D transforms come from gen/dct_port.inc (generated
 * from the lifting IR), the separable 2-D wrappers follow od_bin_fdctNxN /
 * od_bin_idctNxN (src/dct.c:151-163, :351-363, :792-806, :4890-4920). */
#include "port.h"
#include "gen/dct_port.inc"

/* ln = log2(n), n in {4,8,16,32,64}.  Reference table: OD_FDCT_1D src/dct.c:70. */
void port_bin_fdct(int ln, od_coeff *y, const od_coeff *x, int xstride) {
  switch (ln) {
    case 2: port_fdct4(y, x, xstride); break;
    case 3: port_fdct8(y, x, xstride); break;
    case 4: port_fdct16(y, x, xstride); break;
    case 5: port_fdct32(y, x, xstride); break;
    default: port_fdct64(y, x, xstride); break;
  }
}

/* Reference table: OD_IDCT_1D src/dct.c:78. */
void port_bin_idct(int ln, od_coeff *x, int xstride, const od_coeff *y) {
  switch (ln) {
    case 2: port_idct4(x, xstride, y); break;
    case 3: port_idct8(x, xstride, y); break;
    case 4: port_idct16(x, xstride, y); break;
    case 5: port_idct32(x, xstride, y); break;
    default: port_idct64(x, xstride, y); break;
  }
}

/* Columns of x -> rows of a transposed scratch z; columns of z -> rows of y
   (src/dct.c:151-156 and the larger twins). */
void port_bin_fdct2d(int ln, od_coeff *y, int ystride, const od_coeff *x, int xstride) {
  od_coeff z[64*64];
  int n = 1 << ln;
  int i;
  for (i = 0; i < n; i++) port_bin_fdct(ln, z + n*i, x + i, xstride);
  for (i = 0; i < n; i++) port_bin_fdct(ln, y + ystride*i, z + i, n);
}

/* Rows of y -> columns of z; rows of z -> columns of x (src/dct.c:158-163). */
void port_bin_idct2d(int ln, od_coeff *x, int xstride, const od_coeff *y, int ystride) {
  od_coeff z[64*64];
  int n = 1 << ln;
  int i;
  for (i = 0; i < n; i++) port_bin_idct(ln, z + i, n, y + ystride*i);
  for (i = 0; i < n; i++) port_bin_idct(ln, x + i, xstride, z + n*i);
}

/* od_haar / od_haar_inv (src/dct.c:4822 / :4861) written level-synchronously: per level every 2x2
   group is read from a snapshot and then written -- what the reference's serial in-place loops amount
   to (forward: LL(i, j) feeds group (i/2, j/2), visited no later; inverse: descending visiting order).
   Lifting steps: OD_HAAR_KERNEL, src/tf.h:35-46, with (ll, lh, hl, hh) = (sample, below, right, diagonal). */
static void port_haar_kernel(od_coeff *ll, od_coeff *lh, od_coeff *hl, od_coeff *hh) {
  od_coeff half;
  *ll += *hl;
  *hh -= *lh;
  half = (*ll - *hh) >> 1;
  *lh = half - *lh;
  *hl = half - *hl;
  *ll -= *lh;
  *hh += *hl;
}

void port_haar(od_coeff *y, int ystride, const od_coeff *x, int xstride, int ln) {
  static od_coeff cur[64*64], nxt[64*64];
  int n = 1 << ln, size, i, j;
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) cur[i*n + j] = x[i*xstride + j];
  for (size = n; size > 1; size >>= 1) {
    int half = size >> 1;
    for (i = 0; i < half; i++) {
      for (j = 0; j < half; j++) {
        od_coeff a = cur[2*i*n + 2*j], below = cur[(2*i + 1)*n + 2*j];
        od_coeff right = cur[2*i*n + 2*j + 1], diag = cur[(2*i + 1)*n + 2*j + 1];
        port_haar_kernel(&a, &below, &right, &diag);
        nxt[i*n + j] = a;
        y[i*ystride + j + half] = below;
        y[(i + half)*ystride + j] = right;
        y[(i + half)*ystride + j + half] = diag;
      }
    }
    for (i = 0; i < half; i++) for (j = 0; j < half; j++) cur[i*n + j] = nxt[i*n + j];
  }
  y[0] = cur[0];
}

void port_haar_inv(od_coeff *x, int xstride, const od_coeff *y, int ystride, int ln) {
  static od_coeff cur[64*64], nxt[64*64];
  int n = 1 << ln, half, i, j;
  cur[0] = y[0];
  for (half = 1; half < n; half <<= 1) {
    for (i = 0; i < half; i++) {
      for (j = 0; j < half; j++) {
        od_coeff a = cur[i*n + j], below = y[i*ystride + j + half];
        od_coeff right = y[(i + half)*ystride + j], diag = y[(i + half)*ystride + j + half];
        port_haar_kernel(&a, &below, &right, &diag);
        nxt[2*i*n + 2*j] = a;
        nxt[(2*i + 1)*n + 2*j] = below;
        nxt[2*i*n + 2*j + 1] = right;
        nxt[(2*i + 1)*n + 2*j + 1] = diag;
      }
    }
    for (i = 0; i < 2*half; i++) for (j = 0; j < 2*half; j++) cur[i*n + j] = nxt[i*n + j];
  }
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) x[i*xstride + j] = cur[i*n + j];
}

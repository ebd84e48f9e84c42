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

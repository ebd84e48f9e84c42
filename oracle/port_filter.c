/* oracle/port_filter.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 * Restates the live part of src/filter.c: the 4-point lapped pre/post filter
 * (OD_FILT_SIZE() == 0 everywhere, src/filter.h:77) and its appliers. */
#include "port.h"

/* Lifting parameters, src/filter.c:137-140. */
#define S0 85
#define S1 75
#define P0 (-15)
#define U0 33

/* src/filter.c:147-193.  Works in place (y may alias x). */
void port_pre_filter4(od_coeff y[4], const od_coeff x[4]) {
  int d3 = x[0] - x[3];
  int d2 = x[1] - x[2];
  int m1 = x[1] - (d2 >> 1);
  int m0 = x[0] - (d3 >> 1);
  /* scale, then bump positive values by one so the decoder can invert it */
  d2 = d2*S0 >> 6;
  d2 += (-d2 >> 31) & 1;
  d3 = d3*S1 >> 6;
  d3 += (-d3 >> 31) & 1;
  d3 += (d2*P0 + 32) >> 6;
  d2 += (d3*U0 + 32) >> 6;
  m0 += d3 >> 1;
  m1 += d2 >> 1;
  y[0] = m0;
  y[1] = m1;
  y[2] = m1 - d2;
  y[3] = m0 - d3;
}

/* src/filter.c:195-225.  Truncating C division undoes the scale exactly. */
void port_post_filter4(od_coeff x[4], const od_coeff y[4]) {
  int d3 = y[0] - y[3];
  int d2 = y[1] - y[2];
  int m1 = y[1] - (d2 >> 1);
  int m0 = y[0] - (d3 >> 1);
  d2 -= (d3*U0 + 32) >> 6;
  d3 -= (d2*P0 + 32) >> 6;
  d3 = d3*64/S1;
  d2 = d2*64/S0;
  m0 += d3 >> 1;
  m1 += d2 >> 1;
  x[0] = m0;
  x[1] = m1;
  x[2] = m1 - d2;
  x[3] = m0 - d3;
}

/* 8/16/32-point filters (generated from the lifting IR like the DCTs). */
#include "gen/dct_port.inc"
void port_pre_filter_n(int n, od_coeff *y, const od_coeff *x) {
  if (n == 4) port_pre_filter4(y, x);
  else if (n == 8) port_pre_filter8_impl(y, x);
  else if (n == 16) port_pre_filter16_impl(y, x);
  else port_pre_filter32_impl(y, x);
}
void port_post_filter_n(int n, od_coeff *x, const od_coeff *y) {
  if (n == 4) port_post_filter4(x, y);
  else if (n == 8) port_post_filter8_impl(x, y);
  else if (n == 16) port_post_filter16_impl(x, y);
  else port_post_filter32_impl(x, y);
}

static void filt_col(od_coeff *c, int stride, int post) {
  od_coeff t[4];
  int k;
  for (k = 0; k < 4; k++) t[k] = c[k*stride];
  if (post) port_post_filter4(t, t); else port_pre_filter4(t, t);
  for (k = 0; k < 4; k++) c[k*stride] = t[k];
}

static void filt_row(od_coeff *c, int post) {
  if (post) port_post_filter4(c, c); else port_pre_filter4(c, c);
}

/* src/filter.c:1459-1483: interior cross of a (4<<bs)-square node; first the
   horizontal edge (vertical taps, "hfilter"), then the vertical edge. */
void port_prefilter_split(od_coeff *c0, int stride, int bs, int hfilter, int vfilter) {
  int n = 4 << bs;
  int i;
  if (hfilter) for (i = 0; i < n; i++) filt_col(c0 + (n/2 - 2)*stride + i, stride, 0);
  if (vfilter) for (i = 0; i < n; i++) filt_row(c0 + i*stride + n/2 - 2, 0);
}

/* src/filter.c:1485-1527 (non-deblocking branch): exact reverse order. */
void port_postfilter_split(od_coeff *c0, int stride, int bs, int hfilter, int vfilter) {
  int n = 4 << bs;
  int i;
  if (vfilter) for (i = 0; i < n; i++) filt_row(c0 + i*stride + n/2 - 2, 1);
  if (hfilter) for (i = 0; i < n; i++) filt_col(c0 + (n/2 - 2)*stride + i, stride, 1);
}

/* src/filter.c:1529-1559: every horizontal superblock edge first (vertical
   taps over the full plane width), then every vertical edge. */
void port_apply_prefilter_frame_sbs(od_coeff *c0, int stride, int nhsb, int nvsb, int xdec, int ydec) {
  int sbw = 64 >> xdec;
  int sbh = 64 >> ydec;
  int w = nhsb*sbw;
  int h = nvsb*sbh;
  int e;
  int i;
  for (e = 1; e < nvsb; e++) for (i = 0; i < w; i++) filt_col(c0 + (e*sbh - 2)*stride + i, stride, 0);
  for (e = 1; e < nhsb; e++) for (i = 0; i < h; i++) filt_row(c0 + i*stride + e*sbw - 2, 0);
}

/* src/filter.c:1561-1620 (non-deblocking branch): vertical edges first. */
void port_apply_postfilter_frame_sbs(od_coeff *c0, int stride, int nhsb, int nvsb, int xdec, int ydec) {
  int sbw = 64 >> xdec;
  int sbh = 64 >> ydec;
  int w = nhsb*sbw;
  int h = nvsb*sbh;
  int e;
  int i;
  for (e = 1; e < nhsb; e++) for (i = 0; i < h; i++) filt_row(c0 + i*stride + e*sbw - 2, 1);
  for (e = 1; e < nvsb; e++) for (i = 0; i < w; i++) filt_col(c0 + (e*sbh - 2)*stride + i, stride, 1);
}

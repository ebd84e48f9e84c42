/* oracle/port_partition.c -- TEST INFRASTRUCTURE ONLY (see port.h).
 * Restates src/partition.c:123-184 (od_raster_to_coding_order /
 * od_coding_order_to_raster): the 4x4 stage fills positions 1..15, the 8x8
 * stage 16..63, the 16x16 stage 64..255, and the 32x32 stage only its two
 * coded bands 256..511 (OD_LAYOUT32 has 2 bands, OD_LAYOUT64 none,
 * src/partition.c:40-55). */
#include "port.h"
#include "gen/coding_order_port.inc"

static int stage_len(int stage, int n) {
  static const int size[4] = {4, 8, 16, 32};
  static const int len[4] = {15, 48, 192, 256};
  return n >= size[stage] ? len[stage] : 0;
}

static const unsigned short *stage_tbl(int stage) {
  switch (stage) {
    case 0: return kScan4;
    case 1: return kScan8;
    case 2: return kScan16;
    default: return kScan32;
  }
}

void port_raster_to_coding_order(od_coeff *dst, int n, const od_coeff *src, int stride) {
  static const int base[4] = {1, 16, 64, 256};
  int stage;
  for (stage = 0; stage < 4; stage++) {
    const unsigned short *t = stage_tbl(stage);
    int sn = 4 << stage;
    int len = stage_len(stage, n);
    int i;
    for (i = 0; i < len; i++) dst[base[stage] + i] = src[(t[i]/sn)*stride + (t[i]%sn)];
  }
  dst[0] = src[0];
}

void port_coding_order_to_raster(od_coeff *dst, int stride, const od_coeff *src, int n) {
  static const int base[4] = {1, 16, 64, 256};
  int stage;
  for (stage = 0; stage < 4; stage++) {
    const unsigned short *t = stage_tbl(stage);
    int sn = 4 << stage;
    int len = stage_len(stage, n);
    int i;
    for (i = 0; i < len; i++) dst[(t[i]/sn)*stride + (t[i]%sn)] = src[base[stage] + i];
  }
  dst[0] = src[0];
}

/* oracle/ref_pipeline.c -- TEST INFRASTRUCTURE ONLY.
 * pipeline_driver.inc bound to the REAL reference functions (linked from the
 * unmodified xiph/daala objects in oracle/_ref/). */
#include "filter.h"
#include "dct.h"
#define PIPE(name) oracle_ref_##name
#define X_FDCT2D(ln, y, ys, x, xs) (*OD_FDCT_2D_C[(ln) - 2])(y, ys, x, xs)
#define X_IDCT2D(ln, x, xs, y, ys) (*OD_IDCT_2D_C[(ln) - 2])(x, xs, y, ys)
#define X_PRE_SPLIT(c, stride, bs, h, v) od_prefilter_split(c, stride, bs, 0, h, v)
#define X_POST_SPLIT(c, stride, bs, h, v) od_postfilter_split(c, stride, bs, 0, 0, NULL, 0, h, v)
#define X_PRE_SBS(c, stride, nhsb, nvsb, xdec) od_apply_prefilter_frame_sbs(c, stride, nhsb, nvsb, xdec, xdec)
#define X_POST_SBS(c, stride, nhsb, nvsb, xdec) \
  od_apply_postfilter_frame_sbs(c, stride, nhsb, nvsb, xdec, xdec, 0, NULL, 0)
#include "pipeline_driver.inc"

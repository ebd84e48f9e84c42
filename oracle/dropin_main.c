/* oracle/dropin_main.c -- TEST INFRASTRUCTURE ONLY.
 * Link test of the drop-in boundary: this program is the reference encoder's own objects MINUS
 * filter.o and dct.o, with src/state.c and src/encode.c compiled with the reference's own
 * -DOD_X86ASM hook, plus shim/cudastate.o, linked against libdaala_b200.so.  Everything the reference
 * calls by symbol (od_apply_prefilter_frame_sbs, od_prefilter_split, od_haar, OD_FDCT_2D_C ...) or
 * through its vtables (fdct_2d / idct_2d / MC / SAD slots) therefore runs on the GPU.
 *
 *   1. dcttest-style table check (reference src/dct.c:8259-8328, 8379-8409): random blocks through the
 *      linked OD_FDCT_2D_C / OD_IDCT_2D_C tables (= the CUDA entry points): exact round trip, and
 *      equality with the pure-C reference build loaded from libdaala_ref.so.
 *   2. A keyframe (and optionally a P frame) through daala_encode_* with quant 20 / complexity 7: the
 *      coded packets must be byte-identical to the pure-C reference build's.
 *   3. The same packets through the reference DECODER (daala_decode_*) of each build: the pictures decoded on
 *      the CUDA back end must equal the pure-C decoder's (decoder reuse).
 * Exit status 0 = all equal. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "daala/daalaenc.h"
#include "dct.h"
#include "filter.h"

typedef int (*encode_frames_fn)(int w, int h, int nframes, int quant, int complexity, long *bytes, unsigned *sums,
 unsigned *decsums);

/* Encodes `nframes` synthetic frames (first one a keyframe, the rest P frames) and returns each
   packet's size and checksum.  Compiled into this program (GPU back end) AND into libdaala_ref.so
   (ref_hooks_encode.c includes this file's twin, oracle_ref_encode_frames). */
int oracle_dropin_encode_frames(int w, int h, int nframes, int quant, int complexity, long *bytes, unsigned *sums,
 unsigned *decsums);

static unsigned lcg(unsigned *s) { *s = *s*1103515245u + 12345u; return (*s >> 16) & 0x7fff; }

int main(int argc, char **argv) {
  const char *refpath = argc > 1 ? argv[1] : "libdaala_ref.so";
  int w = argc > 2 ? atoi(argv[2]) : 128;
  int h = argc > 3 ? atoi(argv[3]) : 128;
  int nframes = argc > 4 ? atoi(argv[4]) : 1;
  void *ref = dlopen(refpath, RTLD_NOW | RTLD_LOCAL);
  const od_dct_func_2d *ref_fdct;
  const od_dct_func_2d *ref_idct;
  encode_frames_fn ref_encode;
  unsigned seed = 1;
  int bs;
  int fails = 0;
  if (ref == NULL) { fprintf(stderr, "dlopen %s: %s\n", refpath, dlerror()); return 2; }
  ref_fdct = (const od_dct_func_2d *)dlsym(ref, "OD_FDCT_2D_C");
  ref_idct = (const od_dct_func_2d *)dlsym(ref, "OD_IDCT_2D_C");
  ref_encode = (encode_frames_fn)dlsym(ref, "oracle_ref_encode_frames");
  if (!ref_fdct || !ref_idct || !ref_encode) { fprintf(stderr, "missing reference symbols\n"); return 2; }
  if (ref_fdct == OD_FDCT_2D_C) { fprintf(stderr, "tables were not replaced\n"); return 2; }
  for (bs = 0; bs < OD_NBSIZES; bs++) {
    int n = 4 << bs;
    int trial;
    od_coeff *x = (od_coeff *)malloc(sizeof(od_coeff)*n*n*4);
    od_coeff *y = x + n*n, *y2 = y + n*n, *x2 = y2 + n*n;
    for (trial = 0; trial < 8; trial++) {
      int i;
      /* ieee1180-style ranges (src/dct.c:8379): (-256, 255), (-5, 5), (-300, 300), scaled by 16 */
      int range = trial%3 == 0 ? 256 : trial%3 == 1 ? 5 : 300;
      for (i = 0; i < n*n; i++) x[i] = ((int)(lcg(&seed)%(2*range + 1)) - range)*16*(trial & 4 ? -1 : 1);
      (*OD_FDCT_2D_C[bs])(y, n, x, n);
      (*ref_fdct[bs])(y2, n, x, n);
      if (memcmp(y, y2, sizeof(od_coeff)*n*n)) { fprintf(stderr, "fdct %dx%d differs from the reference\n", n, n); fails++; }
      (*OD_IDCT_2D_C[bs])(x2, n, y, n);
      if (memcmp(x, x2, sizeof(od_coeff)*n*n)) { fprintf(stderr, "idct(fdct(x)) != x for %dx%d\n", n, n); fails++; }
      (*ref_idct[bs])(y2, n, y, n);
      if (memcmp(x2, y2, sizeof(od_coeff)*n*n)) { fprintf(stderr, "idct %dx%d differs from the reference\n", n, n); fails++; }
    }
    free(x);
  }
  {
    /* lapped filter tables through the replaced symbols: post(pre(x)) == x (src/filter.c -DTEST main) */
    int f;
    for (f = 0; f < 4; f++) {
      int n = 4 << f;
      od_coeff a[32], b[32], c[32];
      int i;
      for (i = 0; i < n; i++) a[i] = ((int)(lcg(&seed)%601) - 300)*16;
      (*OD_PRE_FILTER[f])(b, a);
      (*OD_POST_FILTER[f])(c, b);
      if (memcmp(a, c, sizeof(od_coeff)*n)) { fprintf(stderr, "post(pre(x)) != x for %d\n", n); fails++; }
    }
  }
  printf("table checks: %s\n", fails ? "FAILED" : "ok");
  {
    long b0[8], b1[8];
    unsigned s0[8], s1[8];
    unsigned d0[8], d1[8];
    int i;
    if (nframes > 8) nframes = 8;
    if (oracle_dropin_encode_frames(w, h, nframes, 20, 7, b0, s0, d0)) { fprintf(stderr, "GPU-backed encode failed\n"); return 3; }
    if ((*ref_encode)(w, h, nframes, 20, 7, b1, s1, d1)) { fprintf(stderr, "reference encode failed\n"); return 3; }
    for (i = 0; i < nframes; i++) {
      printf("frame %d: packet %ld bytes sum %08x (GPU back end) vs %ld bytes sum %08x (pure C reference)%s\n", i,
       b0[i], s0[i], b1[i], s1[i], b0[i] == b1[i] && s0[i] == s1[i] ? "" : "  <-- DIFFERENT");
      if (b0[i] != b1[i] || s0[i] != s1[i]) fails++;
      printf("frame %d: decoded picture sum %08x (GPU back end) vs %08x (pure C reference)%s\n", i, d0[i], d1[i],
       d0[i] == d1[i] ? "" : "  <-- DIFFERENT");
      if (d0[i] != d1[i]) fails++;
    }
  }
  printf("%s\n", fails ? "DROP-IN LINK TEST FAILED" : "drop-in link test ok");
  return fails ? 1 : 0;
}

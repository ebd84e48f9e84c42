/* shim/cudastate.c -- the reference-side binding of libdaala_b200.so: vtable initialisers of the CUDA
 * back end.  This is the file a maintainer adds to libdaalabase / libdaalaenc next to
 * src/x86/x86state.c and src/x86/x86enc.c; it is written against the reference's OWN headers
 * (compile with -I<daala>/src -I<daala>/include) and therefore builds only where the reference tree is
 * present (oracle/Makefile target `dropin`).
 *
 * Pattern (reference src/x86/x86state.c:39-97, src/x86/x86enc.c:37-86): call the C initialiser, then
 * override the slots the accelerated back end provides.
 *
 *   od_state_opt_vtbl (src/state.h:112-131):  fdct_2d[], idct_2d[]          -> OD_FDCT_2D_CUDA / OD_IDCT_2D_CUDA
 *                                             mc_predict1fmv                 -> od_mc_predict1fmv8_cuda
 *                                             mc_blend_full(_split)          -> od_mc_blend_full(_split)8_cuda
 *   od_enc_opt_vtbl   (src/encint.h:77-98):   mc_compute_sad_NxN / satd_NxN  -> od_mc_compute_{sad,satd}8_NxN_cuda
 *
 * The lapped filter, Haar and frame-level appliers are not vtable slots in the reference: they are plain
 * external symbols (od_apply_prefilter_frame_sbs, od_prefilter_split, od_haar, OD_FDCT_2D_C ...), which
 * libdaala_b200.so exports under the reference's names -- linking it INSTEAD of filter.o / dct.o replaces
 * them (oracle/Makefile `dropin` does exactly that).
 *
 * Hook: the reference selects its initialiser at compile time (src/state.c:346-352,
 * src/encode.c:187-193: `#if defined(OD_X86ASM) od_*_opt_vtbl_init_x86`).  A CUDA build compiles those
 * two files with -DOD_X86ASM and supplies the two *_x86 entry points below, so no reference source
 * changes; a maintainer would rather add an `#elif defined(OD_CUDA)` branch calling the *_cuda names.
 *
 * 16-bit (full_precision_references) contexts keep the C kernels: the CUDA back end implements the
 * 8-bit reference path.  The deringing slots keep the C kernels too (the CUDA deringing filter works
 * on whole planes: daala_b200_dering_plane). */
#include "state.h"
#include "encint.h"

extern const od_dct_func_2d OD_FDCT_2D_CUDA[OD_NBSIZES + 1];
extern const od_dct_func_2d OD_IDCT_2D_CUDA[OD_NBSIZES + 1];
void od_mc_predict1fmv8_cuda(od_state *state, unsigned char *dst, const unsigned char *src, int systride,
 int32_t mvx, int32_t mvy, int log_xblk_sz, int log_yblk_sz);
void od_mc_blend_full8_cuda(unsigned char *dst, int dystride, const unsigned char *src[4], int log_xblk_sz,
 int log_yblk_sz);
void od_mc_blend_full_split8_cuda(unsigned char *dst, int dystride, const unsigned char *src[4], int c, int s,
 int log_xblk_sz, int log_yblk_sz);
#define OD_CUDA_MATCH(name) \
  int32_t name(const unsigned char *src, int systride, const unsigned char *ref, int dystride)
OD_CUDA_MATCH(od_mc_compute_sad8_4x4_cuda);
OD_CUDA_MATCH(od_mc_compute_sad8_8x8_cuda);
OD_CUDA_MATCH(od_mc_compute_sad8_16x16_cuda);
OD_CUDA_MATCH(od_mc_compute_sad8_32x32_cuda);
OD_CUDA_MATCH(od_mc_compute_sad8_64x64_cuda);
OD_CUDA_MATCH(od_mc_compute_satd8_4x4_cuda);
OD_CUDA_MATCH(od_mc_compute_satd8_8x8_cuda);
OD_CUDA_MATCH(od_mc_compute_satd8_16x16_cuda);
OD_CUDA_MATCH(od_mc_compute_satd8_32x32_cuda);
OD_CUDA_MATCH(od_mc_compute_satd8_64x64_cuda);

void od_state_opt_vtbl_init_cuda(od_state *state) {
  od_state_opt_vtbl_init_c(state);
  OD_COPY(state->opt_vtbl.fdct_2d, OD_FDCT_2D_CUDA, OD_NBSIZES + 1);
  OD_COPY(state->opt_vtbl.idct_2d, OD_IDCT_2D_CUDA, OD_NBSIZES + 1);
  if (!state->info.full_precision_references) {
    state->opt_vtbl.mc_predict1fmv = od_mc_predict1fmv8_cuda;
    state->opt_vtbl.mc_blend_full = od_mc_blend_full8_cuda;
    state->opt_vtbl.mc_blend_full_split = od_mc_blend_full_split8_cuda;
  }
}

void od_enc_opt_vtbl_init_cuda(od_enc_ctx *enc) {
  od_enc_opt_vtbl_init_c(enc);
  if (!enc->state.info.full_precision_references) {
    enc->opt_vtbl.mc_compute_sad_4x4 = od_mc_compute_sad8_4x4_cuda;
    enc->opt_vtbl.mc_compute_sad_8x8 = od_mc_compute_sad8_8x8_cuda;
    enc->opt_vtbl.mc_compute_sad_16x16 = od_mc_compute_sad8_16x16_cuda;
    enc->opt_vtbl.mc_compute_sad_32x32 = od_mc_compute_sad8_32x32_cuda;
    enc->opt_vtbl.mc_compute_sad_64x64 = od_mc_compute_sad8_64x64_cuda;
    enc->opt_vtbl.mc_compute_satd_4x4 = od_mc_compute_satd8_4x4_cuda;
    enc->opt_vtbl.mc_compute_satd_8x8 = od_mc_compute_satd8_8x8_cuda;
    enc->opt_vtbl.mc_compute_satd_16x16 = od_mc_compute_satd8_16x16_cuda;
    enc->opt_vtbl.mc_compute_satd_32x32 = od_mc_compute_satd8_32x32_cuda;
    enc->opt_vtbl.mc_compute_satd_64x64 = od_mc_compute_satd8_64x64_cuda;
  }
}

/* The names an OD_X86ASM build of src/state.c / src/encode.c calls. */
void od_state_opt_vtbl_init_x86(od_state *state) { od_state_opt_vtbl_init_cuda(state); }
void od_enc_opt_vtbl_init_x86(od_enc_ctx *enc) { od_enc_opt_vtbl_init_cuda(enc); }

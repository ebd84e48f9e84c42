/* oracle/ref_dering_vtbl.h -- TEST INFRASTRUCTURE ONLY.
 * The deringing function table the reference's own od_state would hold on this build: the SSE2 kernels
 * (src/x86/x86state.c:72-75) when the library is compiled with the x86 intrinsics (libdaala_ref_simd.so,
 * the CPU arm of bench.py), the plain C ones otherwise. */
#ifndef ORACLE_REF_DERING_VTBL_H
# define ORACLE_REF_DERING_VTBL_H
# include "dering.h"
# if defined(OD_X86ASM) && defined(OD_SSE2_INTRINSICS)
#  include "x86/x86int.h"
static const od_dering_opt_vtbl *oracle_dering_vtbl(void) {
  static od_dering_opt_vtbl v;
  static int ready;
  if (!ready) {
    OD_COPY(v.filter_dering_direction, OD_DERING_DIRECTION_SSE2, OD_DERINGSIZES);
    OD_COPY(v.filter_dering_orthogonal, OD_DERING_ORTHOGONAL_SSE2, OD_DERINGSIZES);
    ready = 1;
  }
  return &v;
}
# else
static const od_dering_opt_vtbl *oracle_dering_vtbl(void) {
  return &OD_DERING_VTBL_C;
}
# endif
#endif

/* oracle/ref_hooks_encode.c -- TEST INFRASTRUCTURE ONLY.
 * Compiles the reference's src/encode.c in place (see ref_hooks_pvq.c). */
#include "encode.c"

/* oracle/ref_hooks_mcenc.c -- TEST INFRASTRUCTURE ONLY.
 * Compiles the reference's src/mcenc.c in place (see ref_hooks_pvq.c). */
#include "mcenc.c"

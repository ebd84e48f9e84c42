/* oracle/dropin_encode.c -- TEST INFRASTRUCTURE ONLY: the frame encoder driver of the drop-in link test
 * (public API only; see dropin_main.c). */
#include <stdlib.h>
#include "daala/daalaenc.h"
#include "daala/daaladec.h"
#define ENCODE_FRAMES_NAME oracle_dropin_encode_frames
#include "encode_frames.inc"

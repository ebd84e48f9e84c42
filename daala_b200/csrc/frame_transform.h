// Internal aliases for the launch-parameter structs of the public C ABI.
#pragma once
#include "daala_b200.h"

typedef daala_b200_plane PlaneXform;
typedef daala_b200_frame FrameXformParams;

// Parameter blocks of the per-ray kernels (render.cu).
#pragma once
#include <stdint.h>
#include "../../include/rnb_b200.h"

namespace rnb {

constexpr int MAX_RAY_SAMPLES = 192;

using UpsampleParams = rnb_upsample_t;
using CompositeParams = rnb_composite_t;

}  // namespace rnb

// guiding_device.cuh -- device view of the guiding field (spatial kd-tree + per-cell vMF mixtures).
#pragma once
#include "device_math.cuh"

namespace pg {

struct GuideDevice {
    int enabled;
};

}  // namespace pg

// medium_device.cuh -- heterogeneous medium device routines (grid lookups, Woodcock tracking).
#pragma once
#include "device_scene.cuh"

namespace pg {
}  // namespace pg

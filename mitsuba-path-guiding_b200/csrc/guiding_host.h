// guiding_host.h -- host side of the guiding field (training schedule hooks).
#pragma once
#include <cuda_runtime.h>

#include "../../include/b200pg.h"
#include "host_scene.h"
#include "wavefront.cuh"

namespace pg {

struct GuidingHost {
    void init(const B200pgIntegratorParams &, const HostScene &, cudaStream_t) {}
    void configure(ShadeArgs &A) { A.G.enabled = 0; }
    void preprogression(int) {}
    void postprogression(int, void *) {}
    uint32_t numCells() const { return 0; }
};

}  // namespace pg

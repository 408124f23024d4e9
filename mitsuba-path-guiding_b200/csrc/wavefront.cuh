// wavefront.cuh -- wavefront path state, queues and kernel launch wrappers.
//
// Data layout in HBM (see DESIGN.md "Data layout"): path state lives in two ping-pong SoA
// buffers that are COMPACTED every bounce (a path's record moves to its new queue index), so
// that every stage reads and writes its state with fully coalesced 16-byte accesses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_scene.cuh"
#include "guiding_device.cuh"

namespace pg {

// flags word: bits 0-7 depth (rRec.depth), then:
enum : uint32_t {
    kFlagDead = 1u << 8,       // terminated but parked one bounce so its shadow ray can land
    kFlagPrevDelta = 1u << 9,  // last sampled lobe was a delta lobe (bRec.sampledType & EDelta)
    kFlagFirst = 1u << 10,     // rRec.type still includes EEmittedRadiance
    kFlagScattered = 1u << 11,
    kFlagPrevMedium = 1u << 12,  // last event was a medium scattering event
    kFlagNoTrace = 1u << 13,
    kFlagVertexClosed = 1u << 14,  // the last training vertex already has its L_k / distance
    kFlagLook = 1u << 15,          // volumetric: the emitter look-up along the sampled ray is still pending
    kDepthMask = 0xFFu
};

static constexpr int kShadeThreads = 128;
static constexpr int kVertShift = 16;  // flags bits 16-23: number of recorded training vertices

struct PathState {
    float4 *rayO;     // o.xyz, mint
    float4 *rayD;     // d.xyz, maxt
    float4 *thr;      // throughput rgb, eta
    float4 *rad;      // accumulated radiance rgb, pdf of the last sampled direction
    float4 *pos;      // samplePos.xy, rng state lo/hi (bits)
    uint32_t *flags;
    uint32_t *slot;   // index of the camera sample inside the batch
    int32_t *medium;  // current medium (-1 = none)
};

struct ShadowQueue {
    float4 *o;  // o.xyz, mint
    float4 *d;  // d.xyz, maxt
    float4 *c;  // contribution rgb, destination path index (bits)
    int32_t *medium;
    uint4 *aux;  // volumetric: forked rng state lo/hi, pixel, interactions left
};

// Rays handed from the traversal kernels to the cooperative one: queue indices, appended with one atomic per warp
struct TailList {
    uint32_t *list;
    uint32_t *count;
    uint32_t visits;  // node-visit budget of a ray in the first kernel; 0 = no budget, no cooperative kernel (small scenes)
    uint32_t visitsSmall;  // budget when the queue holds fewer than kTailSmallQueue rays (the thin late bounces)
    __host__ __device__ uint32_t budget(uint32_t n) const { return n < (512u << 10) ? visitsSmall : visits; }
};

struct IntegratorConfig {
    int maxDepth, rrDepth, strictNormals, hideEmitters, useNee, volumetric;
    float maxComponentValue;
    // guiding
    int guiding;
    float guidingProbability;
    int recordTraining;
    int guidedDistance;
};

// device counters: per bounce {paths in queue, shadow rays}, then global statistics
struct Counters {
    uint32_t queue[260];
    uint32_t shadow[260];
    uint32_t traceWork[260];   // dynamic work-fetch cursors, one per bounce
    uint32_t shadowWork[260];
    uint32_t lookWork[260];    // volumetric tracking stages (volpath.cu)
    uint32_t trackWork[260];
    uint32_t tailCount[260], tailWork[260];      // long rays handed to the cooperative kernel (k_trace_tail): list size, cursor
    uint32_t shTailCount[260], shTailWork[260];  // same for the shadow queues
    uint32_t partNeed[260];    // hit / miss partition of the shade queue (k_hit_partition): entries placed at the front / back
    uint32_t partRest[260];    // (volumetric path, k_event_partition: medium events / rest, and the surface events below)
    uint32_t partSurf[260];
    uint32_t misc[16];
    unsigned long long paths, normalRays, shadowRays, pathLen, nodesVisited, primsTested, trainSamples;
};

struct BatchDesc {
    uint32_t nPaths;
    uint32_t rowBegin, nRows;  // image band
    uint32_t firstSample, nSamples;
    uint32_t slotBase;           // first slot (splat record / training-vertex record index) of this batch
    const uint32_t *pixelList;   // optional explicit (pixel, sample) pairs
    const uint32_t *sampleList;
};

struct ShadeArgs {
    DeviceScene S;
    IntegratorConfig cfg;
    PathState cur, next;
    ShadowQueue shadow;
    const float4 *hits;
    Counters *C;
    float4 *film;
    float4 *splat;       // per slot, 2 x float4 = one 32-byte sector: {samplePos.xy, L.r, L.g} {L.b, 0, 0, 0} (written once)
    float *radianceOut;  // optional: per-slot radiance instead of film splats (b200pg_k_radiance)
    float4 *trkA, *trkB; // volumetric: per queued path, output of k_track_vol (volpath.cu)
    float4 *lookL;       // volumetric: per queued path, MIS-weighted emitter radiance found by k_look_vol
    GuideDevice G;
    int bounce;
    const uint32_t *perm;  // coherence sort / hit partition / event partition: queue position -> path index (nullptr = queue order)
    const uint32_t *permRest;  // volumetric event partition: the third class
};

// Coherence sort of the shade queue by guiding cell (north star, subsystem 2: "sorting by cell to restore coherence").
// k_trace bins every traced path by the cell of its hit point (bin 0 = nothing to shade: miss / parked) and takes its
// rank inside the bin with one warp-aggregated atomic; k_bin_scan turns the counts into offsets; k_bin_scatter writes
// the permutation the shade stage reads its paths through. Keys never leave the device; order inside a bin is arbitrary.
struct SortArgs {
    const uint4 *guideNodes;  // spatial tree of the guiding field
    uint32_t *binCount;       // 1 + cells counters (zeroed by k_bin_scan for the next bounce)
    uint32_t *binOffset;
    uint32_t *key, *rank;     // per queued path
    uint32_t *perm;
    const uint32_t *nCells;   // device-resident cell count of the field
};

}  // namespace pg

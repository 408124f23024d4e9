// integrator.cu -- host side of the wavefront integrator and the C-ABI of include/b200pg.h.
//
// Mirrors, for this path, what the reference's plugin interface does:
//   ProgressiveMonteCarloIntegrator::render / renderSamples (src/librender/progressiveintegrator.cpp:65-220):
//     numPasses = spp / samplesPerProgression, pre/postprogression hooks around each pass
//   BlockedRenderProcess + LocalWorker tiles (src/librender/renderproc.cpp, imageproc.cpp): replaced by
//     wavefront batches over the whole image band on one GPU
//   Film::put / develop (src/films/hdrfilm.cpp:391-546)
// There is no CPU fallback: without a CUDA device every entry point fails loudly.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200pg.h"
#include "devbuf.h"
#include "guiding_host.h"
#include "host_scene.h"
#include "wavefront.cuh"

namespace pg {

// launch wrappers implemented in kernels.cu
void launchGenerate(const DeviceScene &S, const BatchDesc &B, const PathState &P, Counters *C, cudaStream_t st);
void launchTrace(const DeviceScene &S, const PathState &P, float4 *hits, const uint32_t *nPtr, uint32_t *work, Counters *C,
                 bool count, const SortArgs *sort, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st);
void launchShadow(const DeviceScene &S, const ShadowQueue &Q, float4 *rad, const uint32_t *nPtr, uint32_t *work, Counters *C,
                  bool count, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st);
void launchShade(const ShadeArgs &A, cudaStream_t st);
void launchHitPartition(const float4 *hits, const uint32_t *flags, const uint32_t *nPtr, uint32_t *perm, uint32_t *cntNeed,
                        uint32_t *cntRest, cudaStream_t st);
void launchFlush(const ShadeArgs &A, cudaStream_t st);
void launchFilmAdd(float4 *film, const float4 *peer, uint32_t n, cudaStream_t st);
void launchFilmExportMerged(const float4 *film, const float4 *const *peers, int nPeers, float *out, uint32_t n, cudaStream_t st);
void launchFeatures(const DeviceScene &S, const PathState &P, const float4 *hits, const uint32_t *nPtr, float4 *feat, cudaStream_t st);
void launchFeatureColor(const FilmRecord &F, const float4 *splat, uint32_t n, float maxComponentValue, float4 *feat, cudaStream_t st);
// volpath.cu
void launchShadeVol(const ShadeArgs &A, cudaStream_t st);
void launchShadowVol(const DeviceScene &S, const ShadowQueue &Q, float4 *rad, const uint32_t *nPtr, uint32_t *work, Counters *C,
                     cudaStream_t st);
void launchGridLookup(const DeviceScene &S, int medium, const float *p, uint32_t n, float *out, cudaStream_t st);
void launchMediumTest(const DeviceScene &S, int medium, const float4 *rays, uint32_t n, float *outT, float *outTr, float *outWo,
                      float *outPdf, cudaStream_t st);
void launchFilmExport(const float4 *film, float *out, uint32_t n, int develop, cudaStream_t st);
void launchSplat(const FilmRecord &F, float4 *film, const float4 *splat, uint32_t n, float maxComponentValue, bool tile, cudaStream_t st);
void launchTraceRays(const DeviceScene &S, const float4 *rays, uint32_t n, float4 *hits, uint32_t *work, Counters *C, bool shadow,
                     bool count, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st);
void launchFilmSplat(const FilmRecord &F, float4 *film, const float2 *pos, const float3 *rgb, uint32_t n, float maxComponentValue,
                     cudaStream_t st);
void launchBsdfTest(const DeviceScene &S, int bsdfIndex, const float *wi, const float *wo, const float *u, uint32_t n, float *outEval,
                    float *outPdf, float *outWo, float *outWeight, float *outSpdf, uint32_t *outFlags, cudaStream_t st);

static thread_local std::string g_lastError;
static int fail(const std::string &msg, int code = -1) {
    g_lastError = msg;
    return code;
}
struct PathBuffers {
    DevBuf<float4> rayO, rayD, thr, rad, pos;
    DevBuf<uint32_t> flags, slot;
    DevBuf<int32_t> medium;
    void alloc(size_t n) {
        rayO.alloc(n); rayD.alloc(n); thr.alloc(n); rad.alloc(n); pos.alloc(n);
        flags.alloc(n); slot.alloc(n); medium.alloc(n);
    }
    PathState view() {
        PathState s;
        s.rayO = rayO.p; s.rayD = rayD.p; s.thr = thr.p; s.rad = rad.p; s.pos = pos.p;
        s.flags = flags.p; s.slot = slot.p; s.medium = medium.p;
        return s;
    }
};

struct Integrator {
    HostScene *scene = nullptr;
    B200pgIntegratorParams params;
    int device = 0;
    cudaStream_t stream = nullptr, copyStream = nullptr;
    cudaEvent_t evExport = nullptr;
    std::atomic<int> cancel{0};
    bool scenePinned = false;        // b200pg_scene_upload page-locks the scene's large host arrays on first use
    std::vector<void *> pinnedHost;

    // device scene
    DevBuf<float4> dNodes, dWideNodes, dPrimPlanes, dPrimRows, dRects;
    DevBuf<ShapeRecord> dShapes;
    DevBuf<MeshRecord> dMeshes;
    DevBuf<float> dPositions, dNormals, dTexcoords, dAreaCdf, dEmitterCdf, dDensity;
    DevBuf<uint32_t> dIndices, dPrimGlobal;
    DevBuf<PrimInfo> dPrimInfo;
    DevBuf<float4> dShadeTris;
    DevBuf<BsdfRecord> dBsdfs;
    DevBuf<EmitterRecord> dEmitters;
    DevBuf<MediumRecord> dMedia;
    DeviceScene S;

    // wavefront state: `lanes` independent sub-batches in flight, each on its own stream with its own queues and counters.
    // The persistent kernels of one lane leave the SMs one block at a time as their queue runs dry -- on an open scene the
    // last rays of a bounce are single dependent-load chains hundreds of node visits long (C4: bounces 2-5 take 260-470 us
    // each for < 4 % of the rays, profiles/r02_c4_launches.csv) -- and the next lane's kernels move into the freed SMs.
    // Film, splat records and the guiding buffers are shared (atomics / disjoint slot ranges).
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t evDone = nullptr, evShade = nullptr, evShadow = nullptr;
        cudaStream_t shadowStream = nullptr;
        PathBuffers bufA, bufB;
        DevBuf<float4> dHits, dShO, dShD, dShC;
        DevBuf<int32_t> dShMedium;
        DevBuf<uint4> dShAux;
        DevBuf<float4> dTrkA, dTrkB, dLookL;
        DevBuf<uint32_t> dSortKey, dSortRank, dSortPerm, dBinCount, dBinOffset;
        DevBuf<uint32_t> dTail, dShTail;  // long rays handed to the cooperative traversal kernel (closest / shadow queues)
        DevBuf<Counters> dCounters;
        size_t capacity = 0;
    };
    static constexpr int kMaxLanes = 8;
    Lane lane[kMaxLanes];
    bool overlapShadow = !(std::getenv("B200PG_OVERLAP_SHADOW") && std::atoi(std::getenv("B200PG_OVERLAP_SHADOW")) == 0);
    bool splatTile = true;  // k_splat accumulates a warp's 8 x 4 pixel tile in shared memory first (set_option "splat_tile")
    bool laneMajor = std::getenv("B200PG_LANE_MAJOR") && std::atoi(std::getenv("B200PG_LANE_MAJOR")) != 0;
    int lanes = std::getenv("B200PG_LANES") ? std::max(1, std::min(kMaxLanes, std::atoi(std::getenv("B200PG_LANES")))) : 2;
    DevBuf<float4> dSplat;
    DevBuf<float4> dFilm;
    DevBuf<float> dFilmOut, dFilmAsync;
    cudaEvent_t evFork = nullptr;
    const float4 *filmPeers[16] = {nullptr};  // peers' films mapped through CUDA IPC (b200pg_film_peers_connect)
    int nFilmPeers = 0;
    bool countTraversal = false;
    // coherence sort of the shade queue by guiding cell: bounces 1..sortBounces of a guided (sampling) progression
    // (default off: on C2 with ~800 cells the gathered state reads cost more than the coherent lobe loads save --
    // DESIGN.md, optimisation log 7; B200PG_SORT_BOUNCES / b200pg_set_option("sort_bounces") turn it on)
    int sortBounces = std::getenv("B200PG_SORT_BOUNCES") ? std::atoi(std::getenv("B200PG_SORT_BOUNCES")) : 0;
    // persistent speculative traversal with per-lane refill (kernels.cu: traceQueueSpeculative): bit 0 = closest-hit queues of
    // bounces >= 1, bit 1 = shadow queues, bit 2 = camera rays too (coherent: the batch kernel is as good there)
    // 8-ary quantised tree for large meshes (walked by the speculative kernels; the batch kernels keep the binary tree)
    // measured on C4 (profiles/r02_c4_summary.txt): node visits per ray 25.7 -> 9.5 and long-scoreboard stalls 9.7 -> 3.8 per issue,
    // but 1.65x the instructions -- the kernel turns issue-bound and ends up 5-15 % SLOWER than the binary tree. Off by default.
    bool useWide = std::getenv("B200PG_WIDE") && std::atoi(std::getenv("B200PG_WIDE")) != 0;
    // the wide tree only from this bounce on (late bounces hold few rays and are bound by the dependent-load chains of single
    // long rays, where three binary levels per visit pay; the bulk bounces are issue-bound on the wide node test)
    int wideFrom = std::getenv("B200PG_WIDE_FROM") ? std::atoi(std::getenv("B200PG_WIDE_FROM")) : 0;
    // node-visit budget of a ray before it is handed to the warp-cooperative kernel (kernels.cu: k_trace_tail); 0 = off. Default:
    // 96 when the BVH has >= 64 k nodes (a ray can only get that long in a deep tree), else off.
    int tailVisits = std::getenv("B200PG_TAIL_VISITS") ? std::atoi(std::getenv("B200PG_TAIL_VISITS")) : -1;
    uint32_t tailBudget() const {
        if (tailVisits >= 0) return (uint32_t)tailVisits;
        return scene->nodes.size() >= (64u << 10) ? 96u : 0u;
    }
    int tailVisitsSmall = std::getenv("B200PG_TAIL_VISITS_SMALL") ? std::atoi(std::getenv("B200PG_TAIL_VISITS_SMALL")) : -1;
    TailList tailList(uint32_t *list, uint32_t *count) const {
        const uint32_t v = tailBudget();
        return TailList{list, count, v, tailVisitsSmall >= 0 ? (uint32_t)tailVisitsSmall : v};
    }
    int traceSpec = std::getenv("B200PG_TRACE_SPEC") ? std::atoi(std::getenv("B200PG_TRACE_SPEC")) : 3;
    // hit / miss partition of the shade queue (kernels.cu: k_hit_partition) on bounces where, in the previous progression, fewer
    // than half of the queued paths still had a vertex to shade (open scenes: most bounce rays leave). 0 = never, 1 = adaptive
    // (default), 2 = every bounce >= 1. Affects the order in which paths are shaded, not what they compute.
    int partitionMode = std::getenv("B200PG_PARTITION") ? std::atoi(std::getenv("B200PG_PARTITION")) : 1;
    float shadeNeedRatio[260];   // per bounce, from lane 0's counters of the last progression
    uint32_t hQueue[260], hPartNeed[260];
    uint64_t partitionSkipUntil[260];  // a bounce the partition kernel itself found dense is left alone for a while
    bool partitionBounce(int b, bool sorted) const {
        if (sorted || params.volumetric || b < 1 || b >= 260) return false;
        return partitionMode == 2 ||
               (partitionMode == 1 && shadeNeedRatio[b] < 0.5f && stats.progressions_done >= partitionSkipUntil[b]);
    }
    // denoiser feature buffers (denoiser.cpp:138-144): 3 float4 per pixel, allocated by set_option("feature_buffers", 1)
    bool featureBuffers = false;
    DevBuf<float4> dFeat;

    B200pgStats stats;
    uint64_t peerPaths = 0, peerNormalRays = 0, peerShadowRays = 0, peerPathLen = 0;  // workers of multi-device renders
    cudaEvent_t ev[8];
    GuidingHost guide;

    // per-stage CUDA-event timing on the launching stream (drained after each progression)
    enum { kTimeTrace = 0, kTimeShade = 1, kTimeShadow = 2, kTimeFilm = 3, kTimeTrain = 4, kTimeKinds = 5 };
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<cudaEvent_t> evPool;
    std::vector<Span> spans;
    double stageSeconds[kTimeKinds] = {0, 0, 0, 0, 0};
    uint64_t stageLaunches[kTimeKinds] = {0, 0, 0, 0, 0};
    bool timing = true;
    cudaEvent_t getEvent() {
        if (evPool.empty()) {
            cudaEvent_t e;
            CUDA_OK(cudaEventCreate(&e));
            return e;
        }
        cudaEvent_t e = evPool.back();
        evPool.pop_back();
        return e;
    }
    // With several lanes in flight the spans of different streams overlap in wall time: the per-stage seconds are then
    // stream time, not exclusive GPU time (bench.py measures its per-kernel roofline figures with one lane).
    cudaEvent_t spanBegin(cudaStream_t st = nullptr) {
        if (!timing) return nullptr;
        cudaEvent_t a = getEvent();
        CUDA_OK(cudaEventRecord(a, st ? st : stream));
        return a;
    }
    void spanEnd(int kind, cudaEvent_t a, cudaStream_t st = nullptr) {
        if (!timing) return;
        cudaEvent_t b = getEvent();
        CUDA_OK(cudaEventRecord(b, st ? st : stream));
        spans.push_back(Span{kind, a, b});
    }
    void drainSpans() {  // requires a synchronised stream
        for (auto &sp : spans) {
            float ms = 0;
            CUDA_OK(cudaEventElapsedTime(&ms, sp.a, sp.b));
            stageSeconds[sp.kind] += ms * 1e-3;
            stageLaunches[sp.kind]++;
            evPool.push_back(sp.a);
            evPool.push_back(sp.b);
        }
        spans.clear();
        stats.seconds_trace = stageSeconds[kTimeTrace] + stageSeconds[kTimeShadow];
        stats.seconds_shade = stageSeconds[kTimeShade];
        stats.seconds_film = stageSeconds[kTimeFilm];
        stats.seconds_train = stageSeconds[kTimeTrain];
    }

    ~Integrator() {
        for (void *h : pinnedHost) cudaHostUnregister(h);
        if (stream) cudaStreamDestroy(stream);
        if (copyStream) cudaStreamDestroy(copyStream);
        if (evExport) cudaEventDestroy(evExport);
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        for (auto &sp : spans) {
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        for (auto &e : evPool) cudaEventDestroy(e);
        for (auto &L : lane) {
            if (L.stream) cudaStreamDestroy(L.stream);
            if (L.shadowStream) cudaStreamDestroy(L.shadowStream);
            if (L.evDone) cudaEventDestroy(L.evDone);
            if (L.evShade) cudaEventDestroy(L.evShade);
            if (L.evShadow) cudaEventDestroy(L.evShadow);
        }
        if (evFork) cudaEventDestroy(evFork);
        for (int r = 0; r < nFilmPeers; ++r) cudaIpcCloseMemHandle((void *)filmPeers[r]);
    }

    void init() {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw std::runtime_error("no CUDA device available (this library has no CPU path)");
        if (device < 0 || device >= count) throw std::runtime_error("invalid CUDA device index");
        CUDA_OK(cudaSetDevice(device));
        CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto &e2 : ev) {
            e2 = nullptr;
            CUDA_OK(cudaEventCreate(&e2));
        }
        std::memset(&stats, 0, sizeof(stats));
        for (float &r : shadeNeedRatio) r = 1.0f;
        for (uint64_t &u : partitionSkipUntil) u = 0;
        HostScene &H = *scene;
        dNodes.upload(reinterpret_cast<const float4 *>(H.nodes.data()), H.nodes.size() * 4, stream);
        dPrimPlanes.upload(reinterpret_cast<const float4 *>(H.primPlanes.data()), H.primPlanes.size() / 4, stream);
        dPrimRows.upload(reinterpret_cast<const float4 *>(H.primRows.data()), H.primRows.size() / 4, stream);
        if (!H.wideNodes.empty()) dWideNodes.upload(reinterpret_cast<const float4 *>(H.wideNodes.data()), H.wideNodes.size() * 6, stream);
        dRects.upload(reinterpret_cast<const float4 *>(H.rects.data()), H.rects.size() * 8, stream);
        dShapes.upload(H.shapeRecs, stream);
        dMeshes.upload(H.meshes, stream);
        dPositions.upload(H.positions, stream);
        dNormals.upload(H.normals, stream);
        dTexcoords.upload(H.texcoords, stream);
        dIndices.upload(H.indices, stream);
        dAreaCdf.upload(H.areaCdf, stream);
        dBsdfs.upload(H.bsdfRecs, stream);
        dEmitters.upload(H.emitterRecs, stream);
        dEmitterCdf.upload(H.emitterCdf, stream);
        dMedia.upload(H.mediumRecs, stream);
        dDensity.upload(H.densityPool, stream);
        dPrimGlobal.upload(H.primGlobalId, stream);
        dPrimInfo.upload(H.primInfo, stream);
        dShadeTris.upload(reinterpret_cast<const float4 *>(H.shadeTris.data()), H.shadeTris.size() / 4, stream);
        S.nodes = dNodes.p; S.primPlanes = dPrimPlanes.p; S.primRows = dPrimRows.p; S.rects = dRects.p;
        S.wideNodes = (H.wideNodes.empty() || !useWide) ? nullptr : dWideNodes.p;
        S.shapes = dShapes.p; S.meshes = dMeshes.p;
        S.positions = dPositions.p; S.normals = dNormals.p; S.texcoords = dTexcoords.p;
        S.indices = dIndices.p; S.areaCdf = dAreaCdf.p;
        S.bsdfs = dBsdfs.p; S.emitters = dEmitters.p; S.emitterCdf = dEmitterCdf.p;
        S.media = dMedia.p; S.density = dDensity.p; S.primGlobalId = dPrimGlobal.p;
        S.primInfo = dPrimInfo.p;
        S.shadeTris = dShadeTris.p;
        S.nEmitters = (uint32_t)H.emitters.size();
        S.nPrims = (uint32_t)H.primGlobalId.size();
        S.camera = H.camera;
        S.film = H.filmRec;
        S.seed = H.seed;
        CUDA_OK(cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming));
        int prioLeast = 0, prioGreatest = 0;
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest));
        int laneIndex = 0;
        for (auto &L : lane) {
            // earlier lanes get the SMs first (numerically lower = higher priority): with the lane-major enqueue order the lanes
            // run staggered instead of side by side
            const int prio = std::min(prioLeast, prioGreatest + laneIndex++);
            CUDA_OK(cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, prio));
            CUDA_OK(cudaStreamCreateWithPriority(&L.shadowStream, cudaStreamNonBlocking, prio));
            CUDA_OK(cudaEventCreateWithFlags(&L.evDone, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&L.evShade, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&L.evShadow, cudaEventDisableTiming));
            L.dCounters.alloc(1);
            CUDA_OK(cudaMemsetAsync(L.dCounters.p, 0, sizeof(Counters), stream));
        }
        dFilm.allocExact((size_t)H.film.width * H.film.height);
        CUDA_OK(cudaMemsetAsync(dFilm.p, 0, dFilm.n * sizeof(float4), stream));
        if (params.use_nee && H.emitters.empty()) params.use_nee = 0;  // nothing to sample
        guide.init(this->params, H, stream);
        CUDA_OK(cudaStreamSynchronize(stream));
    }

    void ensureLane(Lane &L, size_t n) {
        if (n <= L.capacity) return;
        L.bufA.alloc(n); L.bufB.alloc(n);
        L.dHits.alloc(n); L.dShO.alloc(n); L.dShD.alloc(n); L.dShC.alloc(n); L.dShMedium.alloc(n);
        if (params.volumetric) {
            L.dShAux.alloc(n); L.dTrkA.alloc(n); L.dTrkB.alloc(n); L.dLookL.alloc(n);
        }
        L.dSortKey.alloc(n); L.dSortRank.alloc(n); L.dSortPerm.alloc(n);
        L.dTail.alloc(n); L.dShTail.alloc(n);
        L.capacity = n;
    }

    IntegratorConfig config() const {
        IntegratorConfig c;
        c.maxDepth = params.max_depth;
        c.rrDepth = params.rr_depth;
        c.strictNormals = params.strict_normals;
        c.hideEmitters = params.hide_emitters;
        c.useNee = params.use_nee;
        c.volumetric = params.volumetric;
        c.maxComponentValue = params.max_component_value;
        c.guiding = 0;
        c.guidingProbability = params.guiding_probability;
        c.recordTraining = 0;
        c.guidedDistance = params.guided_distance;
        return c;
    }

    // Runs up to `lanes` wavefront batches to completion (all bounces), one lane each, their kernels interleaved bounce by
    // bounce on the lanes' streams. Everything is enqueued without host synchronisation unless maxDepth is infinite. The
    // batches' slot ranges [slotBase, slotBase + nPaths) must be disjoint (splat records, training-vertex records).
    // On return the main stream waits for every lane.
    void runBatches(const BatchDesc *batches, int nBatches, float *radianceOut) {
        size_t totalSlots = 0;
        for (int l = 0; l < nBatches; ++l) totalSlots = std::max<size_t>(totalSlots, (size_t)batches[l].slotBase + batches[l].nPaths);
        dSplat.alloc(2 * totalSlots);
        guide.ensureBatch(totalSlots);
        if (guide.active && guide.recording) {  // room for the training samples (renderProgression sizes it for a whole
            // progression beforehand; this covers stand-alone batches such as b200pg_k_radiance)
            const size_t want = std::min<size_t>(totalSlots * guide.maxVerts, (size_t)48 << 20);
            if (want > guide.sampleCapacity) {
                guide.dSRec.alloc(2 * want); guide.dSDist.alloc(want); guide.dSKey.alloc(want);
                guide.sampleCapacity = want;
            }
        }
        for (int l = 0; l < nBatches; ++l) ensureLane(lane[l], batches[l].nPaths);
        const bool single = nBatches == 1;  // one batch: run on the main stream itself (no fork / join events)
        if (!single) CUDA_OK(cudaEventRecord(evFork, stream));
        const size_t zeroBytes = offsetof(Counters, paths);
        ShadeArgs A[kMaxLanes];
        PathState cur[kMaxLanes], next[kMaxLanes];
        SortArgs sortArgs[kMaxLanes];
        cudaStream_t st[kMaxLanes];
        bool live[kMaxLanes];
        const bool sortOn = sortBounces > 0 && !params.volumetric && guide.active && guide.sampling && guide.trained;
        for (int l = 0; l < nBatches; ++l) {
            Lane &L = lane[l];
            st[l] = single ? stream : L.stream;
            live[l] = true;
            if (!single) CUDA_OK(cudaStreamWaitEvent(st[l], evFork, 0));
            CUDA_OK(cudaMemsetAsync(L.dCounters.p, 0, zeroBytes, st[l]));
            cur[l] = L.bufA.view();
            next[l] = L.bufB.view();
            launchGenerate(S, batches[l], cur[l], L.dCounters.p, st[l]);
            stats.kernel_launches++;
            ShadeArgs &a = A[l];
            a.S = S;
            a.cfg = config();
            guide.configure(a);
            a.shadow.o = L.dShO.p; a.shadow.d = L.dShD.p; a.shadow.c = L.dShC.p; a.shadow.medium = L.dShMedium.p; a.shadow.aux = L.dShAux.p;
            a.hits = L.dHits.p;
            a.C = L.dCounters.p;
            a.film = dFilm.p;
            a.radianceOut = radianceOut;
            a.splat = dSplat.p;
            a.trkA = L.dTrkA.p; a.trkB = L.dTrkB.p; a.lookL = L.dLookL.p;
            a.perm = nullptr;
            sortArgs[l] = SortArgs{};
            if (sortOn) {  // coherence sort (surface path, sampling from a trained field): bins = 1 + cells
                const size_t bins = 1 + (size_t)guide.numCells();
                if (bins > L.dBinCount.n) {
                    L.dBinCount.alloc(bins); L.dBinOffset.alloc(L.dBinCount.n);
                    CUDA_OK(cudaMemsetAsync(L.dBinCount.p, 0, L.dBinCount.n * sizeof(uint32_t), st[l]));
                }
                sortArgs[l].guideNodes = guide.dNodes.p;
                sortArgs[l].binCount = L.dBinCount.p; sortArgs[l].binOffset = L.dBinOffset.p;
                sortArgs[l].key = L.dSortKey.p; sortArgs[l].rank = L.dSortRank.p; sortArgs[l].perm = L.dSortPerm.p;
                sortArgs[l].nCells = guide.dCounts.p;
            }
        }
        const int maxBounces = params.max_depth > 0 ? std::min(params.max_depth + 1, 256) : 256;
        int lastBounce[kMaxLanes];
        for (int l = 0; l < nBatches; ++l) lastBounce[l] = maxBounces;
        // Enqueue order. Bounce-major (bounce b of every lane, then bounce b + 1) keeps symmetric lanes in lock step: their bulk
        // kernels compete and their latency-bound tails coincide. Lane-major (all bounces of lane 0, then lane 1, ...; finite
        // depth only -- nothing is polled) staggers them: while lane 0 is in its thin late bounces, lane 1's bulk fills the SMs.
        const bool byLane = laneMajor && params.max_depth > 0 && nBatches > 1;
        const int outer = byLane ? nBatches : maxBounces, inner = byLane ? maxBounces : nBatches;
        for (int oi = 0; oi < outer; ++oi) {
            if (cancel.load()) {
                for (int l = 0; l < nBatches; ++l)
                    if (live[l] && !(byLane && l < oi)) { lastBounce[l] = byLane ? 0 : oi; live[l] = false; }
                break;
            }
            bool any = false;
            for (int ii = 0; ii < inner; ++ii) {
                const int b = byLane ? ii : oi, l = byLane ? oi : ii;
                if (!live[l]) continue;
                any = true;
                Lane &L = lane[l];
                Counters *C = L.dCounters.p;
                cudaEvent_t t = spanBegin(st[l]);
                const bool sorted = sortOn && b >= 1 && b <= sortBounces;
                DeviceScene Sb = S;
                if (b < wideFrom) Sb.wideNodes = nullptr;
                launchTrace(Sb, cur[l], L.dHits.p, &C->queue[b], &C->traceWork[b], C, countTraversal, sorted ? &sortArgs[l] : nullptr,
                            (traceSpec & (b == 0 ? 4 : 1)) != 0, tailList(L.dTail.p, &C->tailCount[b]), &C->tailWork[b], st[l]);
                if (tailBudget() && !sorted && !Sb.wideNodes) stats.kernel_launches++;
                if (sorted) stats.kernel_launches += 2;
                spanEnd(kTimeTrace, t, st[l]);
                if (b == 0 && featureBuffers && !radianceOut) {
                    launchFeatures(S, cur[l], L.dHits.p, &C->queue[0], dFeat.p, st[l]);
                    stats.kernel_launches++;
                }
                ShadeArgs &a = A[l];
                a.perm = sorted ? L.dSortPerm.p : nullptr;
                a.permRest = nullptr;
                if (params.volumetric && partitionMode != 0) {  // event partition (volpath.cu), launched inside launchShadeVol
                    a.perm = L.dSortPerm.p;
                    a.permRest = L.dSortKey.p;
                    stats.kernel_launches++;
                }
                if (partitionBounce(b, sorted)) {
                    launchHitPartition(L.dHits.p, cur[l].flags, &C->queue[b], L.dSortPerm.p, &C->partNeed[b], &C->partRest[b], st[l]);
                    stats.kernel_launches++;
                    a.perm = L.dSortPerm.p;
                }
                a.cur = cur[l];
                a.next = next[l];
                a.bounce = b;
                // shadow rays of bounce b-1 add into the path records this shade stage reads: they must have landed
                if (overlapShadow && b > 0) CUDA_OK(cudaStreamWaitEvent(st[l], L.evShadow, 0));
                t = spanBegin(st[l]);
                if (params.volumetric) launchShadeVol(a, st[l]); else launchShade(a, st[l]);
                spanEnd(kTimeShade, t, st[l]);
                // The shadow stage of bounce b only touches the shadow queue and next.rad; the closest-hit stage of bounce
                // b + 1 reads rays and flags. With overlapShadow the two run concurrently (second stream): the shadow
                // kernel's blocks move into the SMs that the trace kernel's tail leaves idle.
                cudaStream_t ss = overlapShadow ? L.shadowStream : st[l];
                if (overlapShadow) {
                    CUDA_OK(cudaEventRecord(L.evShade, st[l]));
                    CUDA_OK(cudaStreamWaitEvent(ss, L.evShade, 0));
                }
                t = spanBegin(ss);
                if (params.volumetric)
                    launchShadowVol(S, a.shadow, next[l].rad, &C->shadow[b], &C->shadowWork[b], C, ss);
                else
                    launchShadow(Sb, a.shadow, next[l].rad, &C->shadow[b], &C->shadowWork[b], C, countTraversal, (traceSpec & 2) != 0,
                                 tailList(L.dShTail.p, &C->shTailCount[b]), &C->shTailWork[b], ss);
                if (!params.volumetric && (traceSpec & 2) && tailBudget() && !Sb.wideNodes) stats.kernel_launches++;
                spanEnd(kTimeShadow, t, ss);
                if (overlapShadow) CUDA_OK(cudaEventRecord(L.evShadow, ss));
                stats.kernel_launches += 3;
                std::swap(cur[l], next[l]);
                if (params.max_depth <= 0 && (b & 3) == 3) {  // infinite depth: poll the queue size
                    uint32_t nq = 0;
                    CUDA_OK(cudaMemcpyAsync(&nq, &C->queue[b + 1], sizeof(uint32_t), cudaMemcpyDeviceToHost, st[l]));
                    CUDA_OK(cudaStreamSynchronize(st[l]));
                    if (nq == 0) {
                        lastBounce[l] = b + 1;
                        live[l] = false;
                    }
                }
            }
            if (!any && !byLane) break;
        }
        for (int l = 0; l < nBatches; ++l) {
            Lane &L = lane[l];
            ShadeArgs &a = A[l];
            a.cur = cur[l];
            a.next = next[l];
            a.perm = nullptr;
            a.bounce = std::min(lastBounce[l], maxBounces);
            if (overlapShadow && a.bounce > 0) CUDA_OK(cudaStreamWaitEvent(st[l], L.evShadow, 0));
            launchFlush(a, st[l]);
            stats.kernel_launches++;
            if (!radianceOut && !cancel.load()) {  // every path of the batch has ended exactly once: rasterise them
                cudaEvent_t t = spanBegin(st[l]);
                const float4 *sp = dSplat.p + 2 * (size_t)batches[l].slotBase;
                launchSplat(S.film, dFilm.p, sp, batches[l].nPaths, params.max_component_value, splatTile, st[l]);
                stats.kernel_launches++;
                if (featureBuffers) {
                    launchFeatureColor(S.film, sp, batches[l].nPaths, params.max_component_value, dFeat.p, st[l]);
                    stats.kernel_launches++;
                }
                spanEnd(kTimeFilm, t, st[l]);
            }
            if (!single) {
                CUDA_OK(cudaEventRecord(L.evDone, st[l]));
                CUDA_OK(cudaStreamWaitEvent(stream, L.evDone, 0));
            }
        }
    }
    void runBatch(const BatchDesc &B, float *radianceOut) { runBatches(&B, 1, radianceOut); }

    void pullCounters() {
        struct Tail { unsigned long long paths, normalRays, shadowRays, pathLen, nodesVisited, primsTested, trainSamples; };
        static_assert(sizeof(Tail) == sizeof(Counters) - offsetof(Counters, paths), "Counters statistics tail");
        Tail h[kMaxLanes];
        for (int l = 0; l < kMaxLanes; ++l)
            CUDA_OK(cudaMemcpyAsync(&h[l], &lane[l].dCounters.p->paths, sizeof(Tail), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        uint64_t paths = 0, nr = 0, sr = 0, pl = 0, nv = 0, pt = 0;
        for (int l = 0; l < kMaxLanes; ++l) {
            paths += h[l].paths; nr += h[l].normalRays; sr += h[l].shadowRays; pl += h[l].pathLen;
            nv += h[l].nodesVisited; pt += h[l].primsTested;
        }
        stats.paths = paths + peerPaths;
        stats.normal_rays = nr + peerNormalRays;
        stats.shadow_rays = sr + peerShadowRays;
        stats.path_length_sum = pl + peerPathLen;
        stats.bvh_nodes_visited = nv;
        stats.prims_tested = pt;
        stats.train_samples = guide.samplesTrained;
    }

    // One progression over (rows, samples), split into batches of at most maxBatch paths.
    void renderProgression(int firstSample, int nSamples, int rowBegin, int rowEnd) {
        const int W = scene->film.width, H = scene->film.height;
        if (rowEnd <= 0 || rowEnd > H) rowEnd = H;
        if (rowBegin < 0) rowBegin = 0;
        if (rowBegin >= rowEnd || nSamples <= 0) return;
        size_t maxBatch = params.max_batch_paths > 0 ? (size_t)params.max_batch_paths : (size_t)4 << 20;
        const size_t rowPaths = (size_t)W;
        if (guide.active && guide.recording) {
            size_t want = std::min<size_t>((size_t)W * (rowEnd - rowBegin) * nSamples * guide.maxVerts, (size_t)48 << 20);
            if (want > guide.sampleCapacity) {
                guide.dSRec.alloc(2 * want); guide.dSDist.alloc(want); guide.dSKey.alloc(want);
                guide.sampleCapacity = want;
            }
        }
        // device time of the whole progression: CUDA events on the launching stream
        CUDA_OK(cudaEventRecord(ev[2], stream));
        // Split into batches: whole band x k samples where that fits, else row chunks per sample; `lanes` batches run
        // concurrently, so the batch size aims at total / lanes (not below 256 k paths: smaller kernels do not fill the GPU).
        const size_t bandPaths = rowPaths * (rowEnd - rowBegin);
        const size_t total = bandPaths * (size_t)nSamples;
        const size_t target = std::min(maxBatch, std::max<size_t>((total + lanes - 1) / lanes, (size_t)256 << 10));
        std::vector<BatchDesc> batches;
        if (bandPaths <= target) {
            const int spb = (int)std::max<size_t>(1, target / bandPaths);
            for (int s = 0; s < nSamples; s += spb) {
                BatchDesc B;
                std::memset(&B, 0, sizeof(B));
                B.rowBegin = rowBegin;
                B.nRows = rowEnd - rowBegin;
                B.firstSample = firstSample + s;
                B.nSamples = std::min(spb, nSamples - s);
                B.nPaths = (uint32_t)(bandPaths * B.nSamples);
                batches.push_back(B);
            }
        } else {
            int rowsPer = (int)std::max<size_t>(1, target / rowPaths);
            if (rowsPer >= 4) rowsPer &= ~3;  // k_generate's 8 x 4 pixel tiles
            for (int s = 0; s < nSamples; ++s)
                for (int r = rowBegin; r < rowEnd; r += rowsPer) {
                    BatchDesc B;
                    std::memset(&B, 0, sizeof(B));
                    B.rowBegin = r;
                    B.nRows = std::min(rowsPer, rowEnd - r);
                    B.firstSample = firstSample + s;
                    B.nSamples = 1;
                    B.nPaths = (uint32_t)(rowPaths * B.nRows);
                    batches.push_back(B);
                }
        }
        for (size_t i = 0; i < batches.size() && !cancel.load(); i += (size_t)lanes) {
            const int n = (int)std::min<size_t>((size_t)lanes, batches.size() - i);
            uint32_t base = 0;
            for (int l = 0; l < n; ++l) {  // disjoint slot ranges inside the group
                batches[i + l].slotBase = base;
                base += batches[i + l].nPaths;
            }
            runBatches(&batches[i], n, nullptr);
        }
        CUDA_OK(cudaEventRecord(ev[3], stream));
        uint32_t recorded = 0xFFFFFFFFu;
        if (guide.active && guide.recording)  // the training update needs the sample count: fetch it with this sync
            CUDA_OK(cudaMemcpyAsync(&recorded, guide.dSCount.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        if (partitionMode == 1 && !params.volumetric) {  // next progression's partition policy: who still had a vertex to shade
            CUDA_OK(cudaMemcpyAsync(hQueue, lane[0].dCounters.p->queue, sizeof(hQueue), cudaMemcpyDeviceToHost, stream));
            CUDA_OK(cudaMemcpyAsync(hPartNeed, lane[0].dCounters.p->partNeed, sizeof(hPartNeed), cudaMemcpyDeviceToHost, stream));
        }
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaGetLastError());
        if (partitionMode == 1 && !params.volumetric && !batches.empty()) {
            for (int b = 1; b + 1 < 260; ++b) {
                const bool ran = partitionBounce(b, false);
                // small queues are not worth a launch; with the partition on, the kernel's own count is exact, otherwise the
                // survivors of the bounce are the estimate (a path with a vertex to shade mostly lives on)
                if (hQueue[b] < (64u << 10)) shadeNeedRatio[b] = 1.0f;
                else shadeNeedRatio[b] = (float)(ran ? hPartNeed[b] : hQueue[b + 1]) / (float)hQueue[b];
                if (ran && shadeNeedRatio[b] >= 0.5f) partitionSkipUntil[b] = stats.progressions_done + 64;
            }
        }
        guide.pendingCount = recorded;
        drainSpans();
        float msTotal = 0;
        CUDA_OK(cudaEventElapsedTime(&msTotal, ev[2], ev[3]));
        stats.seconds_total += msTotal * 1e-3;
        stats.progressions_done++;
    }
};

struct SceneHandle {
    HostScene host;
};

}  // namespace pg

using namespace pg;

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

int b200pg_version(void) { return B200PG_VERSION; }
const char *b200pg_last_error(void) { return g_lastError.c_str(); }

void b200pg_integrator_params_default(B200pgIntegratorParams *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->max_depth = -1;               // integrator.cpp:203
    p->rr_depth = 5;                 // integrator.cpp:198
    p->samples_per_progression = 1;  // progressiveintegrator.cpp:297
    p->max_render_time = 0;
    p->max_component_value = std::numeric_limits<float>::infinity();
    p->use_nee = 1;                  // progressive_path.cpp:117
    p->guiding_probability = 0.5f;
    p->guide_max_components = 16;
    p->guide_max_cell_samples = 32768;
}

void *b200pg_scene_from_arrays(const B200pgSceneDesc *desc) {
    try {
        std::unique_ptr<SceneHandle> h(new SceneHandle());
        std::string err;
        if (!h->host.copyFrom(desc, err) || !h->host.compile(err)) {
            fail(err);
            return nullptr;
        }
        return h.release();
    } catch (const std::exception &e) {
        fail(e.what());
        return nullptr;
    }
}

void *b200pg_scene_load_xml(const char *path, const char *const *defines, char *errOut, size_t errlen) {
    try {
        std::unique_ptr<SceneHandle> h(new SceneHandle());
        std::string err;
        if (!loadSceneXml(path, defines, h->host, err) || !h->host.compile(err)) {
            fail(err);
            if (errOut && errlen) std::snprintf(errOut, errlen, "%s", err.c_str());
            return nullptr;
        }
        return h.release();
    } catch (const std::exception &e) {
        fail(e.what());
        if (errOut && errlen) std::snprintf(errOut, errlen, "%s", e.what());
        return nullptr;
    }
}

const B200pgSceneDesc *b200pg_scene_desc(void *scene) {
    if (!scene) return nullptr;
    return &((SceneHandle *)scene)->host.view;
}

int b200pg_scene_integrator_params(void *scene, B200pgIntegratorParams *out) {
    if (!scene || !out) return fail("null argument");
    *out = ((SceneHandle *)scene)->host.xmlParams;
    return 0;
}

void b200pg_scene_destroy(void *scene) { delete (SceneHandle *)scene; }

void *b200pg_integrator_create(void *scene, const B200pgIntegratorParams *params, int device) {
    if (!scene) {
        fail("null scene");
        return nullptr;
    }
    try {
        std::unique_ptr<Integrator> I(new Integrator());
        I->scene = &((SceneHandle *)scene)->host;
        if (params)
            I->params = *params;
        else
            I->params = I->scene->xmlParams;
        // MonteCarloIntegrator ctor checks (integrator.cpp:225-229)
        if (I->params.rr_depth <= 0) {
            fail("'rrDepth' must be set to a value greater than zero!");
            return nullptr;
        }
        if (I->params.max_depth <= 0 && I->params.max_depth != -1) {
            fail("'maxDepth' must be set to -1 (infinite) or a value greater than zero!");
            return nullptr;
        }
        if (I->params.max_depth > 250) {
            fail("maxDepth above 250 is not supported by the wavefront queues");
            return nullptr;
        }
        if (I->params.samples_per_progression <= 0) I->params.samples_per_progression = 1;
        I->device = device;
        I->init();
        return I.release();
    } catch (const std::exception &e) {
        fail(e.what());
        return nullptr;
    }
}

#define PG_TRY(I) \
    if (!(I)) return fail("null integrator"); \
    Integrator *self = (Integrator *)(I); \
    try { \
        CUDA_OK(cudaSetDevice(self->device));
#define PG_END \
    } catch (const std::exception &e) { return fail(e.what()); } \
    return 0;

int b200pg_progression_render(void *integ, int first_sample, int n_samples, int row_begin, int row_end) {
    PG_TRY(integ)
    self->cancel.store(0);  // a cancel request ends the call it interrupts, not every later one
    self->renderProgression(first_sample, n_samples, row_begin, row_end);
    self->pullCounters();
    PG_END
}

// ---- the progressive loop (ProgressiveMonteCarloIntegrator::renderSamples / renderTime, progressiveintegrator.cpp:65-168)
// One pass = preprogression (what this pass records / samples from) -> progression -> postprogression (refit the field from
// the pass' samples; the last training pass may discard the film). The SAME routine serves the sample budget, the time
// budget and every worker of a multi-device render, so the estimator does not depend on how the job was launched.
namespace pg {
struct PassPlan {
    int perPass, numPasses, trainPasses;
    bool timed;
    double limit;
    int rank, world;
};
static PassPlan makePlan(const Integrator &I, int rank, int world) {
    PassPlan P;
    P.perPass = std::max(1, I.params.samples_per_progression);
    const int passes = std::max(1, I.scene->sampleCount / P.perPass);
    P.world = std::max(1, world);
    P.rank = rank;
    P.numPasses = (passes + P.world - 1) / P.world;  // global passes: `world` sample blocks each, one per device
    P.timed = I.params.max_render_time > 0;
    P.limit = (double)I.params.max_render_time;
    // the sample budget trains at most as long as it renders; a time budget has no pass count to clamp against
    // How many GLOBAL passes train the field (a global pass = `world` sample blocks, one per device).
    //  * sample budget: ceil(training_progressions / world) -- the same number of training SAMPLES as a single-device run, so the
    //    split of the budget between training and rendering (and with guide_train_discard_film the film's sample count) does not
    //    depend on the device count;
    //  * time budget: training_progressions -- there is no such split to preserve, and the field needs its UPDATES: the spatial
    //    tree splits one level per update, and counting device passes here left an 8-GPU 4K job with two updates of its sixteen
    //    and a 4-cell field (guided == unguided, profiles/r02_equal_time_c5_4k_8gpu_before_schedule_fix.jsonl).
    const int T = I.guide.active ? std::max(0, I.params.training_progressions) : 0;
    P.trainPasses = P.timed ? T : std::min((T + P.world - 1) / P.world, P.numPasses);
    return P;
}
// global pass g of the plan on this integrator; `record` is the same on every rank
static void runPass(Integrator &I, const PassPlan &P, int g) {
    const bool record = g < P.trainPasses;
    I.guide.recording = I.guide.active && record;
    I.guide.sampling = true;
    I.renderProgression((g * P.world + P.rank) * P.perPass, P.perPass, 0, 0);
    if (I.guide.recording && !I.cancel.load()) {
        cudaEvent_t t = I.spanBegin();
        I.guide.trainLocal();  // with connected peers the statistics are summed over all devices inside the M-step kernel
        I.spanEnd(Integrator::kTimeTrain, t);
        CUDA_OK(cudaStreamSynchronize(I.stream));
        I.drainSpans();
        if (I.params.guide_train_discard_film && g + 1 >= P.trainPasses)
            CUDA_OK(cudaMemsetAsync(I.dFilm.p, 0, I.dFilm.n * sizeof(float4), I.stream));
    }
}

// Host-side rendezvous of the device workers of one render call: rank 0 publishes a verdict (continue / stop) that everybody
// obeys, so that all workers leave the loop -- and its in-kernel cross-GPU barriers -- together. A worker that fails aborts
// the rendezvous and thereby the job.
struct Rendezvous {
    std::mutex m;
    std::condition_variable cv;
    int world, arrived = 0, generation = 0;
    bool aborted = false, verdict = false;
    std::string error;
    explicit Rendezvous(int w) : world(w) {}
    // returns false when the job was aborted; *stop receives rank 0's verdict
    bool arrive(int rank, bool myStop, bool *stop) {
        std::unique_lock<std::mutex> lk(m);
        if (aborted) return false;
        if (rank == 0) verdict = myStop;
        const int gen = generation;
        if (++arrived == world) {
            arrived = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen || aborted; });
            if (aborted && generation == gen) return false;
        }
        *stop = verdict;
        return true;
    }
    void abort(const std::string &why) {
        std::lock_guard<std::mutex> lk(m);
        if (!aborted) error = why;
        aborted = true;
        cv.notify_all();
    }
};

static void renderLoop(Integrator &I, const PassPlan &P, Rendezvous *rv, std::atomic<int> &cancel) {
    CUDA_OK(cudaSetDevice(I.device));
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; P.timed || g < P.numPasses; ++g) {
        if (cancel.load()) I.cancel.store(1);
        runPass(I, P, g);
        bool stop = cancel.load() != 0;
        if (P.timed) stop = stop || std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= P.limit;
        if (rv && !rv->arrive(P.rank, stop, &stop)) return;
        if (stop) break;
    }
    I.pullCounters();
}
}  // namespace pg

int b200pg_render(void *integ, int device_count, const int *devices) {
    PG_TRY(integ)
    self->cancel.store(0);  // a cancelled render does not poison the next one (Integrator::cancel is per render job, integrator.h:84-87)
    if (device_count <= 1 || !devices) {
        if (devices && device_count == 1 && devices[0] != self->device)
            return fail("the integrator lives on device " + std::to_string(self->device) + ", not on the requested device");
        if (self->guide.commWorld > 1)
            return fail("peers are connected (b200pg_comm_connect): drive the progressions of every rank yourself, or render with a device list");
        renderLoop(*self, makePlan(*self, 0, 1), nullptr, self->cancel);
        return 0;
    }
    // ---- several devices inside this call (the reference registers all its workers inside one render(): mitsuba.cpp:278-327,
    // progressiveintegrator.cpp:84-103). Device devices[0] must be the integrator's own; the scene and the field are replicated,
    // device r renders sample block g * n + r of global pass g, the EM statistics are summed over NVLink peer memory inside
    // the M-step kernel, and the films are added into this integrator's film at the end.
    const int world = device_count;
    if (world > 16) return fail("at most 16 devices");
    if (devices[0] != self->device) return fail("devices[0] must be the device the integrator was created on");
    int available = 0;
    CUDA_OK(cudaGetDeviceCount(&available));
    for (int r = 0; r < world; ++r) {
        if (devices[r] < 0 || devices[r] >= available) return fail("invalid CUDA device index in the device list");
        for (int q = 0; q < r; ++q)
            if (devices[q] == devices[r]) return fail("a device is listed twice");
    }
    if (self->guide.commWorld > 1) return fail("peers are already connected (b200pg_comm_connect)");
    std::vector<std::unique_ptr<Integrator>> workers;  // ranks 1..world-1
    std::vector<Integrator *> ranks((size_t)world, nullptr);
    ranks[0] = self;
    for (int r = 1; r < world; ++r) {
        std::unique_ptr<Integrator> W(new Integrator());
        W->scene = self->scene;
        W->params = self->params;
        W->device = devices[r];
        W->traceSpec = self->traceSpec;
        W->sortBounces = self->sortBounces;
        W->timing = self->timing;
        W->init();
        if (self->guide.active) {  // a field trained by earlier calls stays the common start
            CUDA_OK(cudaSetDevice(self->device));
            const std::vector<uint32_t> words = self->guide.snapshot();
            CUDA_OK(cudaSetDevice(W->device));
            if (!W->guide.load(words.data(), words.size())) return fail("cannot replicate the guiding field");
        }
        ranks[(size_t)r] = W.get();
        workers.push_back(std::move(W));
    }
    for (int r = 0; r < world; ++r) {  // peer access in both directions (exchange blocks, films)
        CUDA_OK(cudaSetDevice(devices[r]));
        for (int q = 0; q < world; ++q) {
            if (q == r) continue;
            int can = 0;
            CUDA_OK(cudaDeviceCanAccessPeer(&can, devices[r], devices[q]));
            if (!can) return fail("devices " + std::to_string(devices[r]) + " and " + std::to_string(devices[q]) + " cannot access each other's memory");
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else CUDA_OK(e);
        }
    }
    if (self->guide.active) {
        std::vector<float *> blocks((size_t)world, nullptr);
        for (int r = 0; r < world; ++r) {
            CUDA_OK(cudaSetDevice(devices[r]));
            blocks[(size_t)r] = ranks[(size_t)r]->guide.commLocalBlock();
        }
        for (int r = 0; r < world; ++r) {
            CUDA_OK(cudaSetDevice(devices[r]));
            ranks[(size_t)r]->guide.commConnectPointers(r, world, blocks.data());
        }
    }
    Rendezvous rv(world);
    std::vector<std::thread> threads;
    for (int r = 1; r < world; ++r)
        threads.emplace_back([&, r] {
            try {
                renderLoop(*ranks[(size_t)r], makePlan(*ranks[(size_t)r], r, world), &rv, self->cancel);
            } catch (const std::exception &e) {
                rv.abort(std::string("device ") + std::to_string(devices[r]) + ": " + e.what());
            }
        });
    try {
        renderLoop(*self, makePlan(*self, 0, world), &rv, self->cancel);
    } catch (const std::exception &e) {
        rv.abort(std::string("device ") + std::to_string(devices[0]) + ": " + e.what());
    }
    for (auto &t : threads) t.join();
    CUDA_OK(cudaSetDevice(self->device));
    if (self->guide.active) {  // back to a single-device integrator; the (identical) field of rank 0 is the result
        for (int r = 0; r < world; ++r) {
            cudaSetDevice(devices[r]);
            ranks[(size_t)r]->guide.commDisconnect();
        }
        CUDA_OK(cudaSetDevice(self->device));
    }
    if (rv.aborted) return fail("multi-device render failed: " + rv.error);
    for (int r = 1; r < world; ++r) {  // film += the peers' films over NVLink, statistics summed
        Integrator &W = *ranks[(size_t)r];
        launchFilmAdd(self->dFilm.p, W.dFilm.p, (uint32_t)self->dFilm.n, self->stream);
        CUDA_OK(cudaStreamSynchronize(self->stream));
        self->stats.kernel_launches += 1 + W.stats.kernel_launches + W.guide.launches;
        self->peerPaths += W.stats.paths;
        self->peerNormalRays += W.stats.normal_rays;
        self->peerShadowRays += W.stats.shadow_rays;
        self->peerPathLen += W.stats.path_length_sum;
    }
    self->pullCounters();
    PG_END
}

int b200pg_cancel(void *integ) {
    if (!integ) return fail("null integrator");
    ((Integrator *)integ)->cancel.store(1);
    return 0;
}

int b200pg_film_clear(void *integ) {
    PG_TRY(integ)
    CUDA_OK(cudaMemsetAsync(self->dFilm.p, 0, self->dFilm.n * sizeof(float4), self->stream));
    if (self->dFeat.p) CUDA_OK(cudaMemsetAsync(self->dFeat.p, 0, self->dFeat.n * sizeof(float4), self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    PG_END
}

int b200pg_film_ipc_handle(void *integ, void *handle64) {
    PG_TRY(integ)
    if (!handle64) return fail("null argument");
    CUDA_OK(cudaStreamSynchronize(self->stream));
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, self->dFilm.p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &h, 64);
    PG_END
}

int b200pg_film_add_peers(void *integ, int rank, int world, const void *handles) {
    PG_TRY(integ)
    if (!handles || world < 1 || rank < 0 || rank >= world) return fail("invalid argument");
    for (int r = 0; r < world; ++r) {  // fixed peer order
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        launchFilmAdd(self->dFilm.p, (const float4 *)ptr, (uint32_t)self->dFilm.n, self->stream);
        cudaError_t e = cudaStreamSynchronize(self->stream);
        cudaIpcCloseMemHandle(ptr);
        CUDA_OK(e);
    }
    CUDA_OK(cudaGetLastError());
    PG_END
}

int b200pg_film_device_buffer(void *integ, void **dev_ptr, size_t *n_floats) {
    if (!integ) return fail("null integrator");
    Integrator *self = (Integrator *)integ;
    if (dev_ptr) *dev_ptr = self->dFilm.p;
    if (n_floats) *n_floats = self->dFilm.n * 4;
    return 0;
}

int b200pg_film_read(void *integ, float *rgbaw) {
    PG_TRY(integ)
    // expand (R,G,B,weight) -> (R,G,B,alpha,weight) on the device, then ONE device->host copy straight into the
    // caller's buffer (fast when that buffer is pinned). alpha == weight here: rRec.alpha stays 1 (integrator.h:218-225).
    self->dFilmOut.alloc(self->dFilm.n * 5);
    launchFilmExport(self->dFilm.p, self->dFilmOut.p, (uint32_t)self->dFilm.n, 0, self->stream);
    CUDA_OK(cudaMemcpyAsync(rgbaw, self->dFilmOut.p, self->dFilm.n * 5 * sizeof(float), cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    self->stats.kernel_launches++;
    PG_END
}

// Asynchronous variant for progressive previews: the 5-channel expansion is enqueued on the render stream (so it sees the
// film exactly as it is now), the device->host copy runs on a second stream and overlaps whatever is rendered next.
int b200pg_film_read_async(void *integ, float *rgbaw_pinned) {
    PG_TRY(integ)
    if (!self->copyStream) {
        CUDA_OK(cudaStreamCreateWithFlags(&self->copyStream, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&self->evExport, cudaEventDisableTiming));
    }
    CUDA_OK(cudaStreamSynchronize(self->copyStream));  // the previous read must have left the staging buffer
    // the asynchronous path stages through its OWN buffer: b200pg_film_read / _develop / _write re-use dFilmOut on the render
    // stream and must not tear a preview copy that is still in flight
    self->dFilmAsync.alloc(self->dFilm.n * 5);
    if (self->nFilmPeers > 0)
        launchFilmExportMerged(self->dFilm.p, self->filmPeers, self->nFilmPeers, self->dFilmAsync.p, (uint32_t)self->dFilm.n, self->stream);
    else
        launchFilmExport(self->dFilm.p, self->dFilmAsync.p, (uint32_t)self->dFilm.n, 0, self->stream);
    CUDA_OK(cudaEventRecord(self->evExport, self->stream));
    CUDA_OK(cudaStreamWaitEvent(self->copyStream, self->evExport, 0));
    CUDA_OK(cudaMemcpyAsync(rgbaw_pinned, self->dFilmAsync.p, self->dFilm.n * 5 * sizeof(float), cudaMemcpyDeviceToHost, self->copyStream));
    self->stats.kernel_launches++;
    PG_END
}
// Multi-GPU previews: maps the films of the other ranks (CUDA IPC handles from b200pg_film_ipc_handle, `world` x 64 bytes, own
// slot ignored). From then on b200pg_film_read_async on this integrator delivers own film + the peers' films, summed on the
// device over NVLink in fixed rank order -- ONE device->host copy per preview for the whole job instead of one per rank. The
// films are not modified. world = 1 (or handles = NULL) unmaps.
int b200pg_film_peers_connect(void *integ, int rank, int world, const void *handles) {
    PG_TRY(integ)
    CUDA_OK(cudaStreamSynchronize(self->stream));
    if (self->copyStream) CUDA_OK(cudaStreamSynchronize(self->copyStream));
    for (int r = 0; r < self->nFilmPeers; ++r) cudaIpcCloseMemHandle((void *)self->filmPeers[r]);
    self->nFilmPeers = 0;
    if (!handles || world <= 1) return 0;
    if (world > 16 || rank < 0 || rank >= world) return fail("invalid argument");
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        self->filmPeers[self->nFilmPeers++] = (const float4 *)ptr;
    }
    PG_END
}
int b200pg_film_read_wait(void *integ) {
    PG_TRY(integ)
    if (self->copyStream) CUDA_OK(cudaStreamSynchronize(self->copyStream));
    PG_END
}

int b200pg_film_develop(void *integ, float *rgb) {
    PG_TRY(integ)
    // weight normalisation RGB / weight (fmtconv.cpp:978-1005) on the device
    self->dFilmOut.alloc(self->dFilm.n * 5);
    launchFilmExport(self->dFilm.p, self->dFilmOut.p, (uint32_t)self->dFilm.n, 1, self->stream);
    CUDA_OK(cudaMemcpyAsync(rgb, self->dFilmOut.p, self->dFilm.n * 3 * sizeof(float), cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    self->stats.kernel_launches++;
    PG_END
}

// ---- image writers (Bitmap::write, src/libcore/bitmap.cpp: PFM, OpenEXR, RGBE) -----------------------------------------
static uint16_t floatToHalf(float f) {  // round to nearest even, IEEE binary16
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u));  // inf / nan
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                      // overflow -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                    // underflow -> 0
    int e = (int)(x >> 23) - 127 + 15;
    uint32_t m = x & 0x7FFFFFu;
    if (e <= 0) {  // subnormal half
        m |= 0x800000u;
        const int shift = 14 - e;
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}

static bool writePfm(const char *path, const float *rgb, int W, int H) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "PF\n%d %d\n-1.0\n", W, H);  // little endian, bottom-up scanlines (bitmap.cpp writePFM)
    for (int y = H - 1; y >= 0; --y) std::fwrite(rgb + (size_t)y * W * 3, sizeof(float), (size_t)W * 3, f);
    return std::fclose(f) == 0;
}

// Scanline OpenEXR, single part, no compression. Channels must be given in alphabetical order of their names (the
// format stores them sorted); channel c of pixel i is chans[c].base[i * chans[c].stride].
struct ExrChannel {
    std::string name;
    const float *base;
    size_t stride;
};
static bool writeExrChannels(const char *path, const std::vector<ExrChannel> &chans, int W, int H, bool half) {
    for (size_t c = 1; c < chans.size(); ++c)
        if (!(chans[c - 1].name < chans[c].name)) return false;
    FILE *f = std::fopen(path, "wb");
    if (!f) return false;
    std::vector<unsigned char> hdr;
    auto put = [&](const void *p, size_t n) { hdr.insert(hdr.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
    auto putStr = [&](const char *s) { put(s, std::strlen(s) + 1); };
    auto putI = [&](int32_t v) { put(&v, 4); };
    auto putF = [&](float v) { put(&v, 4); };
    auto attr = [&](const char *name, const char *type, int32_t size) { putStr(name); putStr(type); putI(size); };
    const uint32_t magic = 20000630u;
    put(&magic, 4);
    putI(2);  // version 2, scanline, single part
    int32_t chlist = 1;
    for (const auto &c : chans) chlist += (int32_t)c.name.size() + 1 + 16;
    attr("channels", "chlist", chlist);
    for (const auto &c : chans) {
        putStr(c.name.c_str());
        putI(half ? 1 : 2);  // HALF / FLOAT
        const unsigned char lin[4] = {0, 0, 0, 0};
        put(lin, 4);
        putI(1);
        putI(1);
    }
    hdr.push_back(0);
    attr("compression", "compression", 1);
    hdr.push_back(0);
    attr("dataWindow", "box2i", 16);
    putI(0); putI(0); putI(W - 1); putI(H - 1);
    attr("displayWindow", "box2i", 16);
    putI(0); putI(0); putI(W - 1); putI(H - 1);
    attr("lineOrder", "lineOrder", 1);
    hdr.push_back(0);
    attr("pixelAspectRatio", "float", 4);
    putF(1.0f);
    attr("screenWindowCenter", "v2f", 8);
    putF(0.0f); putF(0.0f);
    attr("screenWindowWidth", "float", 4);
    putF(1.0f);
    hdr.push_back(0);
    const size_t bpc = half ? 2 : 4, nc = chans.size(), lineBytes = (size_t)W * nc * bpc;
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    uint64_t offset = hdr.size() + (uint64_t)H * 8;
    for (int y = 0; y < H; ++y) {
        std::fwrite(&offset, 8, 1, f);
        offset += 8 + lineBytes;
    }
    std::vector<unsigned char> line(lineBytes);
    for (int y = 0; y < H; ++y) {
        const int32_t yy = y, sz = (int32_t)lineBytes;
        std::fwrite(&yy, 4, 1, f);
        std::fwrite(&sz, 4, 1, f);
        for (size_t c = 0; c < nc; ++c) {  // one plane per channel inside the scanline
            for (int x = 0; x < W; ++x) {
                const float v = chans[c].base[((size_t)y * W + x) * chans[c].stride];
                if (half) {
                    const uint16_t h = floatToHalf(v);
                    std::memcpy(&line[(c * W + x) * 2], &h, 2);
                } else {
                    std::memcpy(&line[(c * W + x) * 4], &v, 4);
                }
            }
        }
        std::fwrite(line.data(), 1, lineBytes, f);
    }
    return std::fclose(f) == 0;
}
static bool writeExr(const char *path, const float *rgb, int W, int H, bool half) {
    return writeExrChannels(path, {{"B", rgb + 2, 3}, {"G", rgb + 1, 3}, {"R", rgb, 3}}, W, H, half);
}

// Radiance RGBE, flat (uncompressed) scanlines, top to bottom (bitmap.cpp writeRGBE / rgbe.cpp)
static bool writeRgbe(const char *path, const float *rgb, int W, int H) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", H, W);
    std::vector<unsigned char> line((size_t)W * 4);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const float *p = rgb + ((size_t)y * W + x) * 3;
            const float v = std::max(p[0], std::max(p[1], p[2]));
            unsigned char *o = &line[(size_t)x * 4];
            if (!(v > 1e-32f)) {
                o[0] = o[1] = o[2] = o[3] = 0;
            } else {
                int e;
                const float sc = std::frexp(v, &e) * 256.0f / v;
                o[0] = (unsigned char)(p[0] * sc);
                o[1] = (unsigned char)(p[1] * sc);
                o[2] = (unsigned char)(p[2] * sc);
                o[3] = (unsigned char)(e + 128);
            }
        }
        std::fwrite(line.data(), 1, line.size(), f);
    }
    return std::fclose(f) == 0;
}

int b200pg_film_write(void *integ, const char *path) {
    if (!integ || !path) return fail("null argument");
    Integrator *self = (Integrator *)integ;
    const int W = self->scene->film.width, H = self->scene->film.height;
    std::vector<float> rgb((size_t)W * H * 3);
    int r = b200pg_film_develop(integ, rgb.data());
    if (r) return r;
    std::string p(path), ext;
    const size_t dot = p.find_last_of('.');
    if (dot != std::string::npos) ext = p.substr(dot);
    for (auto &c : ext) c = (char)std::tolower(c);
    bool ok;
    if (ext == ".pfm") ok = writePfm(path, rgb.data(), W, H);
    else if (ext == ".exr") ok = writeExr(path, rgb.data(), W, H, self->scene->film.component_format == 0);
    else if (ext == ".rgbe" || ext == ".hdr") ok = writeRgbe(path, rgb.data(), W, H);
    else return fail("unsupported image extension \"" + ext + "\" (hdrfilm writes .exr, .pfm or .rgbe)");
    if (!ok) return fail(std::string("cannot write ") + path);
    return 0;
}

// Denoiser feature buffers (denoiser.cpp:138-144): out = H*W*10 floats {color.rgb, albedo.rgb, normal.xyz, sample count}
int b200pg_features_read(void *integ, float *out) {
    PG_TRY(integ)
    if (!out) return fail("null argument");
    if (!self->dFeat.p) return fail("feature buffers are not enabled (b200pg_set_option(\"feature_buffers\", 1))");
    const size_t n = self->dFilm.n;
    std::vector<float4> h(n * 3);
    CUDA_OK(cudaMemcpyAsync(h.data(), self->dFeat.p, n * 3 * sizeof(float4), cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    for (size_t i = 0; i < n; ++i) {
        const float4 c = h[3 * i], a = h[3 * i + 1], nn = h[3 * i + 2];
        const float ic = c.w > 0 ? 1.0f / c.w : 0.0f, ia = a.w > 0 ? 1.0f / a.w : 0.0f;
        float *o = out + 10 * i;
        o[0] = c.x * ic; o[1] = c.y * ic; o[2] = c.z * ic;
        o[3] = a.x * ia; o[4] = a.y * ia; o[5] = a.z * ia;
        o[6] = nn.x * ia; o[7] = nn.y * ia; o[8] = nn.z * ia;
        o[9] = a.w;
    }
    PG_END
}

// Multi-channel OpenEXR (float32) with the layer names Denoiser::saveBuffers uses (denoiser.cpp:88-112): color, albedo,
// normal (normal.R/G/B = x/y/z)
int b200pg_features_write(void *integ, const char *path) {
    if (!integ || !path) return fail("null argument");
    Integrator *self = (Integrator *)integ;
    const int W = self->scene->film.width, H = self->scene->film.height;
    std::vector<float> f((size_t)W * H * 10);
    int r = b200pg_features_read(integ, f.data());
    if (r) return r;
    const float *b = f.data();
    const std::vector<ExrChannel> chans = {{"albedo.B", b + 5, 10}, {"albedo.G", b + 4, 10}, {"albedo.R", b + 3, 10},
                                           {"color.B", b + 2, 10},  {"color.G", b + 1, 10},  {"color.R", b + 0, 10},
                                           {"normal.B", b + 8, 10}, {"normal.G", b + 7, 10}, {"normal.R", b + 6, 10}};
    if (!writeExrChannels(path, chans, W, H, false)) return fail(std::string("cannot write ") + path);
    return 0;
}


int b200pg_stats(void *integ, B200pgStats *out) {
    if (!integ || !out) return fail("null argument");
    Integrator *self = (Integrator *)integ;
    *out = self->stats;
    out->kernel_launches += self->guide.launches;
    out->guide_cells = self->guide.numCells();
    return 0;
}

void b200pg_destroy(void *integ) { delete (Integrator *)integ; }

int b200pg_set_option(void *integ, const char *name, int value) {
    if (!integ || !name) return fail("null argument");
    Integrator *self = (Integrator *)integ;
    std::string n(name);
    if (n == "count_traversal") self->countTraversal = value != 0;
    else if (n == "timing") self->timing = value != 0;
    else if (n == "sort_bounces") self->sortBounces = value;
    else if (n == "trace_spec") self->traceSpec = value;
    else if (n == "wide_bvh") {
        self->useWide = value != 0;
        self->S.wideNodes = (self->scene->wideNodes.empty() || !self->useWide) ? nullptr : self->dWideNodes.p;
    }
    else if (n == "overlap_shadow") self->overlapShadow = value != 0;
    else if (n == "lanes") self->lanes = std::max(1, std::min((int)Integrator::kMaxLanes, value));
    else if (n == "partition") self->partitionMode = value;
    else if (n == "split_levels") self->guide.splitLevels = std::max(1, std::min(16, value));
    else if (n == "splat_tile") self->splatTile = value != 0;
    else if (n == "tail_visits") self->tailVisits = value;
    else if (n == "lane_major") self->laneMajor = value != 0;
    else if (n == "feature_buffers") {
        self->featureBuffers = value != 0;
        if (self->featureBuffers && !self->dFeat.p) {
            try {
                CUDA_OK(cudaSetDevice(self->device));
                self->dFeat.allocExact(self->dFilm.n * 3);
                CUDA_OK(cudaMemsetAsync(self->dFeat.p, 0, self->dFeat.n * sizeof(float4), self->stream));
                CUDA_OK(cudaStreamSynchronize(self->stream));
            } catch (const std::exception &e) { return fail(e.what()); }
        }
    }
    else return fail("unknown option " + n);
    return 0;
}

int b200pg_stage_times(void *integ, double *seconds5, uint64_t *launches5) {
    if (!integ) return fail("null integrator");
    Integrator *self = (Integrator *)integ;
    for (int i = 0; i < Integrator::kTimeKinds; ++i) {
        if (seconds5) seconds5[i] = self->stageSeconds[i];
        if (launches5) launches5[i] = self->stageLaunches[i];
    }
    return 0;
}

int b200pg_scene_upload(void *integ, size_t *bytes) {
    PG_TRY(integ)
    HostScene &H = *self->scene;
    size_t total = 0;
    if (!self->scenePinned) {  // page-lock the large host arrays once: repeated uploads then run at the PCIe rate
        auto pin = [&](const void *ptr, size_t nbytes) {
            if (nbytes < (1u << 20)) return;
            if (cudaHostRegister(const_cast<void *>(ptr), nbytes, cudaHostRegisterDefault) == cudaSuccess)
                self->pinnedHost.push_back(const_cast<void *>(ptr));
            else
                cudaGetLastError();  // already registered by another integrator of the same scene, or not pinnable: pageable copy
        };
        pin(H.nodes.data(), H.nodes.size() * sizeof(H.nodes[0]));
        pin(H.primPlanes.data(), H.primPlanes.size() * sizeof(float));
        pin(H.primRows.data(), H.primRows.size() * sizeof(float));
        pin(H.shadeTris.data(), H.shadeTris.size() * sizeof(float));
        pin(H.primInfo.data(), H.primInfo.size() * sizeof(H.primInfo[0]));
        pin(H.positions.data(), H.positions.size() * sizeof(float));
        pin(H.normals.data(), H.normals.size() * sizeof(float));
        pin(H.indices.data(), H.indices.size() * sizeof(uint32_t));
        pin(H.densityPool.data(), H.densityPool.size() * sizeof(float));
        self->scenePinned = true;
    }
    auto up = [&](auto &dst, const auto &src) {
        dst.upload(src, self->stream);
        total += src.size() * sizeof(src[0]);
    };
    self->dNodes.upload(reinterpret_cast<const float4 *>(H.nodes.data()), H.nodes.size() * 4, self->stream);
    self->dPrimPlanes.upload(reinterpret_cast<const float4 *>(H.primPlanes.data()), H.primPlanes.size() / 4, self->stream);
    self->dPrimRows.upload(reinterpret_cast<const float4 *>(H.primRows.data()), H.primRows.size() / 4, self->stream);
    if (!H.wideNodes.empty()) {
        self->dWideNodes.upload(reinterpret_cast<const float4 *>(H.wideNodes.data()), H.wideNodes.size() * 6, self->stream);
        total += H.wideNodes.size() * sizeof(WideNode);
    }
    self->dRects.upload(reinterpret_cast<const float4 *>(H.rects.data()), H.rects.size() * 8, self->stream);
    total += H.nodes.size() * 64 + H.primGlobalId.size() * 48 + H.rects.size() * 128;
    up(self->dShapes, H.shapeRecs);
    up(self->dPrimInfo, H.primInfo);
    self->dShadeTris.upload(reinterpret_cast<const float4 *>(H.shadeTris.data()), H.shadeTris.size() / 4, self->stream);
    total += H.shadeTris.size() * sizeof(float);
    up(self->dMeshes, H.meshes);
    up(self->dPositions, H.positions);
    up(self->dNormals, H.normals);
    up(self->dIndices, H.indices);
    up(self->dAreaCdf, H.areaCdf);
    up(self->dBsdfs, H.bsdfRecs);
    up(self->dEmitters, H.emitterRecs);
    up(self->dEmitterCdf, H.emitterCdf);
    up(self->dMedia, H.mediumRecs);
    up(self->dDensity, H.densityPool);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    if (bytes) *bytes = total;
    PG_END
}

// ---------------------------------------------------------------------------------------------
// per-kernel entry points
// ---------------------------------------------------------------------------------------------
int b200pg_k_trace_device(void *integ, const void *d_rays, size_t n, int shadow, void *d_hits, float *ms, uint64_t *counts) {
    PG_TRY(integ)
    Counters *C = self->lane[0].dCounters.p;
    CUDA_OK(cudaMemsetAsync(&C->misc[0], 0, 3 * sizeof(uint32_t), self->stream));  // work cursor, tail count, tail cursor
    self->lane[0].dTail.alloc(n);
    if (counts) {
        CUDA_OK(cudaMemsetAsync(&C->nodesVisited, 0, 2 * sizeof(unsigned long long), self->stream));
    }
    CUDA_OK(cudaEventRecord(self->ev[0], self->stream));
    launchTraceRays(self->S, (const float4 *)d_rays, (uint32_t)n, (float4 *)d_hits, &C->misc[0], C, shadow != 0, counts != nullptr,
                    (self->traceSpec & (shadow ? 2 : 1)) != 0, self->tailList(self->lane[0].dTail.p, &C->misc[1]), &C->misc[2], self->stream);
    CUDA_OK(cudaEventRecord(self->ev[1], self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    if (ms) CUDA_OK(cudaEventElapsedTime(ms, self->ev[0], self->ev[1]));
    if (counts) {
        unsigned long long h[2];
        CUDA_OK(cudaMemcpy(h, &C->nodesVisited, sizeof(h), cudaMemcpyDeviceToHost));
        counts[0] = h[0];
        counts[1] = h[1];
    }
    self->stats.kernel_launches++;
    PG_END
}

int b200pg_k_trace(void *integ, const float *rays, size_t n, int shadow, float *hits_tuv, uint32_t *hits_prim) {
    PG_TRY(integ)
    DevBuf<float4> dR, dH;
    dR.upload(reinterpret_cast<const float4 *>(rays), n * 2, self->stream);
    dH.alloc(n);
    int r = b200pg_k_trace_device(integ, dR.p, n, shadow, dH.p, nullptr, nullptr);
    if (r) return r;
    std::vector<float4> h(n);
    CUDA_OK(cudaMemcpy(h.data(), dH.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        if (hits_tuv) {
            hits_tuv[3 * i] = h[i].x;
            hits_tuv[3 * i + 1] = h[i].y;
            hits_tuv[3 * i + 2] = h[i].z;
        }
        uint32_t p;
        std::memcpy(&p, &h[i].w, 4);
        hits_prim[i] = p;
    }
    PG_END
}

int b200pg_k_bsdf(void *integ, int bsdf_index, const float *wi, const float *wo, const float *u, size_t n, float *out_eval,
                  float *out_pdf, float *out_wo, float *out_weight, float *out_spdf, uint32_t *out_flags) {
    PG_TRY(integ)
    if (bsdf_index < 0 || bsdf_index >= (int)self->scene->bsdfRecs.size()) return fail("bsdf index out of range");
    DevBuf<float> dWi, dWo, dU, dE, dP, dSo, dW, dSp;
    DevBuf<uint32_t> dF;
    dWi.upload(wi, 3 * n, self->stream);
    dWo.upload(wo, 3 * n, self->stream);
    dU.upload(u, 2 * n, self->stream);
    dE.alloc(3 * n); dP.alloc(n); dSo.alloc(3 * n); dW.alloc(3 * n); dSp.alloc(n); dF.alloc(n);
    launchBsdfTest(self->S, bsdf_index, dWi.p, dWo.p, dU.p, (uint32_t)n, dE.p, dP.p, dSo.p, dW.p, dSp.p, dF.p, self->stream);
    CUDA_OK(cudaMemcpyAsync(out_eval, dE.p, 3 * n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_pdf, dP.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_wo, dSo.p, 3 * n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_weight, dW.p, 3 * n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_spdf, dSp.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_flags, dF.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->stats.kernel_launches++;
    PG_END
}

int b200pg_k_radiance(void *integ, const uint32_t *pixel, const uint32_t *sample_index, size_t n, float *out_rgb) {
    PG_TRY(integ)
    DevBuf<uint32_t> dPix, dSmp;
    DevBuf<float> dOut;
    dPix.upload(pixel, n, self->stream);
    dSmp.upload(sample_index, n, self->stream);
    dOut.alloc(3 * n);
    CUDA_OK(cudaMemsetAsync(dOut.p, 0, 3 * n * 4, self->stream));
    BatchDesc B;
    std::memset(&B, 0, sizeof(B));
    B.nPaths = (uint32_t)n;
    B.pixelList = dPix.p;
    B.sampleList = dSmp.p;
    self->runBatch(B, dOut.p);
    self->guide.pendingCount = 0xFFFFFFFFu;  // a stand-alone batch may have recorded samples: the next update asks the device
    CUDA_OK(cudaMemcpyAsync(out_rgb, dOut.p, 3 * n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->drainSpans();
    self->pullCounters();
    PG_END
}

int b200pg_k_film_splat(void *integ, const float *pos, const float *rgb, size_t n) {
    PG_TRY(integ)
    DevBuf<float> dPos, dRgb;
    dPos.upload(pos, 2 * n, self->stream);
    dRgb.upload(rgb, 3 * n, self->stream);
    launchFilmSplat(self->S.film, self->dFilm.p, (const float2 *)dPos.p, (const float3 *)dRgb.p, (uint32_t)n,
                    self->params.max_component_value, self->stream);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->stats.kernel_launches++;
    PG_END
}

// ---------------------------------------------------------------------------------------------
// guiding: training hooks and per-kernel entry points
// ---------------------------------------------------------------------------------------------
int b200pg_guiding_mode(void *integ, int record, int sample) {
    if (!integ) return fail("null integrator");
    Integrator *self = (Integrator *)integ;
    if (!self->guide.active && (record || sample)) return fail("guiding is not enabled in the integrator parameters");
    self->guide.recording = record != 0;
    self->guide.sampling = sample != 0;
    return 0;
}

int b200pg_train_begin(void *integ, uint32_t *n_samples, uint32_t *n_cells) {
    PG_TRY(integ)
    if (!self->guide.active) return fail("guiding is not enabled in the integrator parameters");
    cudaEvent_t t = self->spanBegin();
    self->guide.begin();
    self->spanEnd(Integrator::kTimeTrain, t);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    self->drainSpans();
    if (n_samples) *n_samples = self->guide.nSamples;
    if (n_cells) *n_cells = self->guide.numCells();
    PG_END
}
int b200pg_train_accumulate(void *integ) {
    PG_TRY(integ)
    cudaEvent_t t = self->spanBegin();
    self->guide.accumulate();
    self->spanEnd(Integrator::kTimeTrain, t);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->drainSpans();
    PG_END
}
int b200pg_train_stats_buffer(void *integ, void **dev_ptr, size_t *n_floats) {
    if (!integ) return fail("null integrator");
    Integrator *self = (Integrator *)integ;
    if (dev_ptr) *dev_ptr = self->guide.dStats.p;
    if (n_floats) *n_floats = (size_t)self->guide.numCells() * self->guide.statsStride();
    return 0;
}
int b200pg_train_update(void *integ, int commit) {
    PG_TRY(integ)
    cudaEvent_t t = self->spanBegin();
    self->guide.update(commit != 0);
    self->spanEnd(Integrator::kTimeTrain, t);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->drainSpans();
    PG_END
}
int b200pg_train(void *integ, int n_iter, uint32_t *n_samples, uint32_t *n_cells) {
    PG_TRY(integ)
    if (!self->guide.active) return fail("guiding is not enabled in the integrator parameters");
    cudaEvent_t t = self->spanBegin();
    self->guide.train(n_iter > 0 ? n_iter : self->guide.emIterations);
    self->spanEnd(Integrator::kTimeTrain, t);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->drainSpans();
    if (n_samples) *n_samples = self->guide.nSamples;
    if (n_cells) *n_cells = self->guide.numCells();
    PG_END
}
int b200pg_comm_local_handle(void *integ, void *handle64) {
    PG_TRY(integ)
    if (!handle64) return fail("null argument");
    self->guide.commLocalHandle(handle64);
    PG_END
}
int b200pg_comm_connect(void *integ, int rank, int world, const void *handles) {
    PG_TRY(integ)
    if (!handles) return fail("null argument");
    self->guide.commConnect(rank, world, handles);
    PG_END
}
int b200pg_train_end(void *integ) {
    PG_TRY(integ)
    cudaEvent_t t = self->spanBegin();
    self->guide.end();
    self->spanEnd(Integrator::kTimeTrain, t);
    CUDA_OK(cudaStreamSynchronize(self->stream));
    self->drainSpans();
    PG_END
}

int b200pg_k_vmm_pdf_sample(void *integ, const float *pos, const float *dir, const float *u, size_t n, float *out_pdf,
                            float *out_dir, float *out_spdf, uint32_t *out_cell) {
    PG_TRY(integ)
    self->guide.query(pos, dir, u, n, out_pdf, out_dir, out_spdf, out_cell);
    PG_END
}
int b200pg_k_bin_samples(void *integ, const float *pos, size_t n, uint32_t *out_cell, uint32_t *out_perm, uint32_t *out_offsets,
                         uint32_t *n_cells) {
    PG_TRY(integ)
    self->guide.bin(pos, n, out_cell, out_perm, out_offsets, n_cells);
    PG_END
}
int b200pg_k_em_step(void *integ, const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist,
                     size_t n, int n_iter, float *stats_out) {
    PG_TRY(integ)
    GuidingHost &g = self->guide;
    g.beginExternal(pos, dir, weight, pdf, dist, n);
    if (n_iter <= 0) {
        g.accumulate();
        if (stats_out)
            CUDA_OK(cudaMemcpyAsync(stats_out, g.dStats.p, (size_t)g.numCells() * g.statsStride() * sizeof(float),
                                    cudaMemcpyDeviceToHost, self->stream));
        CUDA_OK(cudaStreamSynchronize(self->stream));
        CUDA_OK(cudaGetLastError());
    } else {
        for (int it = 0; it < n_iter; ++it) {
            g.accumulate();
            g.update(it == n_iter - 1);
        }
        if (stats_out)
            CUDA_OK(cudaMemcpyAsync(stats_out, g.dStats.p, (size_t)g.numCells() * g.statsStride() * sizeof(float),
                                    cudaMemcpyDeviceToHost, self->stream));
        g.end();
    }
    PG_END
}
int b200pg_k_em_exchange(void *integ, uint32_t n_cells, int n_iter, int mode, float *ms_per_iter) {
    PG_TRY(integ)
    if (!ms_per_iter || n_iter <= 0 || mode < 0 || mode > 3) return fail("invalid argument");
    if (!self->guide.active) return fail("guiding is not enabled in the integrator parameters");
    const int before = self->guide.commForceMode;
    self->guide.commForceMode = mode == 2 ? 0 : mode == 3 ? 1 : -1;
    try {
        *ms_per_iter = self->guide.exchangeBench(n_cells, n_iter, mode == 1);
    } catch (...) {
        self->guide.commForceMode = before;
        throw;
    }
    self->guide.commForceMode = before;
    PG_END
}
int b200pg_field_snapshot(void *integ, uint32_t *out, size_t *n_words) {
    PG_TRY(integ)
    std::vector<uint32_t> w = self->guide.snapshot();
    if (out && n_words && *n_words >= w.size()) std::memcpy(out, w.data(), w.size() * 4);
    if (n_words) *n_words = w.size();
    PG_END
}
int b200pg_field_load(void *integ, const uint32_t *in, size_t n_words) {
    PG_TRY(integ)
    if (!self->guide.load(in, n_words)) return fail("malformed guiding-field snapshot");
    self->guide.sampling = true;
    PG_END
}
int b200pg_k_grid_lookup(void *integ, int medium, const float *p, size_t n, float *out) {
    PG_TRY(integ)
    if (medium < 0 || medium >= (int)self->scene->media.size()) return fail("medium index out of range");
    DevBuf<float> dP, dO;
    dP.upload(p, 3 * n, self->stream);
    dO.alloc(n);
    launchGridLookup(self->S, medium, dP.p, (uint32_t)n, dO.p, self->stream);
    CUDA_OK(cudaMemcpyAsync(out, dO.p, n * sizeof(float), cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->stats.kernel_launches++;
    PG_END
}

int b200pg_k_medium_sample(void *integ, int medium, const float *rays, size_t n, float *out_t, float *out_tr, float *out_wo,
                           float *out_pdf) {
    PG_TRY(integ)
    if (medium < 0 || medium >= (int)self->scene->media.size()) return fail("medium index out of range");
    DevBuf<float4> dR;
    DevBuf<float> dT, dTr, dWo, dPdf;
    dR.upload(reinterpret_cast<const float4 *>(rays), 2 * n, self->stream);
    dT.alloc(n); dTr.alloc(n); dWo.alloc(3 * n); dPdf.alloc(n);
    launchMediumTest(self->S, medium, dR.p, (uint32_t)n, dT.p, dTr.p, dWo.p, dPdf.p, self->stream);
    CUDA_OK(cudaMemcpyAsync(out_t, dT.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_tr, dTr.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_wo, dWo.p, 3 * n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaMemcpyAsync(out_pdf, dPdf.p, n * 4, cudaMemcpyDeviceToHost, self->stream));
    CUDA_OK(cudaStreamSynchronize(self->stream));
    CUDA_OK(cudaGetLastError());
    self->stats.kernel_launches++;
    PG_END
}

}  // extern "C"

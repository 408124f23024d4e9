// kernels.cu -- the wavefront stages of the surface path as hand-written sm_100a CUDA kernels:
//   k_generate       camera-ray generation        (progressiveintegrator.cpp:246-271, perspective.cpp:271-298)
//   k_trace          closest-hit BVH traversal, warps take 32 rays at a time (coherent queues: camera rays)
//   k_trace_spec / k_shadow_spec   persistent speculative traversal with per-lane refill (bounce and shadow queues)
//   k_trace_tail     warp-cooperative traversal of the rays that outlive their node-visit budget in the two above
//   k_hit_partition  hit / miss partition of a sparse shade queue (open scenes)
//   k_shade          intersection fill, emitted radiance + MIS, Russian roulette, NEE, BSDF / guided sampling, queue
//                    compaction, training-vertex records; a finished path leaves a splat record (progressive_path.cpp:133-314)
//   k_shadow         batch any-hit kernel (B200PG_TRACE_SPEC without bit 1); unoccluded rays add their NEE term to the path record
//   k_splat          film accumulation of a finished batch in pixel order (imageblock.h:151-197)
//   k_film_*, k_features, k_feature_color, k_flush, test kernels (k_film_splat, k_trace_rays*, k_bsdf_test)
// All kernels are persistent (grid = k * #SMs) and read their work size from device counters so
// that a whole batch (all bounces) is enqueued without a host round trip.
#include <cooperative_groups.h>

#include "guiding_device.cuh"
#include "medium_device.cuh"
#include "wavefront.cuh"
#include "wavefront_device.cuh"

namespace pg {

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------
// Film: Gaussian-filtered splat (ImageBlock::put, include/mitsuba/render/imageblock.h:151-197).
// The arithmetic replays the reference's 32x32 tile + border coordinates (imageproc.cpp:27-78,
// rfilter.cpp:51) so that the discretised filter lookups land in the same table bins.
// Film texel = float4 (R, G, B, weight); alpha == weight on this path.
// ------------------------------------------------------------------------------------------
PG_DEV void filmSplat(const FilmRecord &F, float4 *film, float2 pos, float3 spec, float maxComponentValue) {
    float maxSpec = maxComp(spec);
    if (maxSpec > maxComponentValue) spec = spec * (maxComponentValue / maxSpec);  // progressiveintegrator.cpp:274-277
    // ImageBlock::put rejects NaN / negative samples (imageblock.h:154-158)
    if (!(isfinite(spec.x) && isfinite(spec.y) && isfinite(spec.z)) || spec.x < 0 || spec.y < 0 || spec.z < 0) return;
    const int border = (int)ceilf(F.radius - 0.5f);
    const int pxI = min(max((int)pos.x, 0), F.width - 1), pyI = min(max((int)pos.y, 0), F.height - 1);
    const int ox = (pxI >> 5) << 5, oy = (pyI >> 5) << 5;
    const int bw = min(32, F.width - ox) + 2 * border, bh = min(32, F.height - oy) + 2 * border;
    const float px = pos.x - 0.5f - (float)(ox - border), py = pos.y - 0.5f - (float)(oy - border);
    const int minx = max((int)ceilf(px - F.radius), 0), miny = max((int)ceilf(py - F.radius), 0);
    const int maxx = min((int)floorf(px + F.radius), bw - 1), maxy = min((int)floorf(py + F.radius), bh - 1);
    float wx[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int x = minx + i;
        wx[i] = x <= maxx ? F.values[min((int)fabsf((x - px) * F.scaleFactor), 31)] : 0.0f;
    }
    for (int y = miny; y <= maxy; ++y) {
        const float wy = F.values[min((int)fabsf((y - py) * F.scaleFactor), 31)];
        const int fy = y + oy - border;
        if (fy < 0 || fy >= F.height) continue;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int x = minx + i;
            const int fx = x + ox - border;
            if (x > maxx || fx < 0 || fx >= F.width) continue;
            const float w = wx[i] * wy;
            atomicAdd(film + (size_t)fy * F.width + fx, make_float4(w * spec.x, w * spec.y, w * spec.z, w));
        }
    }
}

__global__ void __launch_bounds__(256) k_film_splat(FilmRecord F, float4 *film, const float2 *pos, const float3 *rgb,
                                                    uint32_t n, float maxComponentValue) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *c = reinterpret_cast<const float *>(rgb) + 3 * (size_t)i;
        filmSplat(F, film, pos[i], f3(c[0], c[1], c[2]), maxComponentValue);
    }
}

// Film export: develop = 0 -> (R,G,B,alpha,weight) per pixel; develop = 1 -> RGB / weight (fmtconv.cpp:978-1005)
__global__ void __launch_bounds__(256) k_film_export(const float4 *__restrict__ film, float *__restrict__ out, uint32_t n, int develop) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 f = film[i];
        if (develop) {
            const float inv = f.w > 0 ? 1.0f / f.w : 0.0f;
            out[3 * (size_t)i + 0] = f.x * inv;
            out[3 * (size_t)i + 1] = f.y * inv;
            out[3 * (size_t)i + 2] = f.z * inv;
        } else {
            out[5 * (size_t)i + 0] = f.x;
            out[5 * (size_t)i + 1] = f.y;
            out[5 * (size_t)i + 2] = f.z;
            out[5 * (size_t)i + 3] = f.w;
            out[5 * (size_t)i + 4] = f.w;
        }
    }
}

// Merged preview of a multi-GPU job: (own film + the peers' films, read over NVLink) expanded to (R,G,B,alpha,weight) -- the
// films themselves are left untouched, so it can run after every progression while all ranks keep accumulating.
struct FilmPeers {
    const float4 *p[16];
    int n;
};
__global__ void __launch_bounds__(256) k_film_export_merged(const float4 *__restrict__ film, FilmPeers peers, float *__restrict__ out,
                                                            uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 f = film[i];
        for (int r = 0; r < peers.n; ++r) {  // fixed peer order
            const float4 g = ldStream(peers.p[r] + i);
            f.x += g.x; f.y += g.y; f.z += g.z; f.w += g.w;
        }
        out[5 * (size_t)i + 0] = f.x;
        out[5 * (size_t)i + 1] = f.y;
        out[5 * (size_t)i + 2] = f.z;
        out[5 * (size_t)i + 3] = f.w;
        out[5 * (size_t)i + 4] = f.w;
    }
}

// Multi-GPU film merge (every GPU holds a full-size film): film += a peer's film, read over NVLink
__global__ void __launch_bounds__(256) k_film_add(float4 *__restrict__ film, const float4 *__restrict__ peer, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 f = film[i];
        const float4 g = ldStream(peer + i);
        f.x += g.x; f.y += g.y; f.z += g.z; f.w += g.w;
        film[i] = f;
    }
}

// ------------------------------------------------------------------------------------------
// Camera rays
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_generate(DeviceScene S, BatchDesc B, PathState P, Counters *C) {
    const uint32_t W = S.film.width;
    const uint32_t nPix = B.nRows * W;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nPaths; i += gridDim.x * blockDim.x) {
        uint32_t pixel, sample;
        if (B.pixelList) {
            pixel = B.pixelList[i];
            sample = B.sampleList[i];
        } else {
            sample = B.firstSample + i / nPix;
            uint32_t j = i % nPix;
            // a warp covers an 8 x 4 pixel tile instead of 32 pixels of one row: its first hits are neighbours in two
            // dimensions (same material, same guiding cell), and queue compaction keeps that order for later bounces.
            // Samples are keyed by (pixel, sample index), so the order does not change any result.
            if ((W & 7u) == 0 && (B.nRows & 3u) == 0) {
                const uint32_t tile = j >> 5, within = j & 31u, tilesX = W >> 3;
                j = ((tile / tilesX) * 4 + (within >> 3)) * W + (tile % tilesX) * 8 + (within & 7u);
            }
            pixel = B.rowBegin * W + j;
        }
        Rng rng;
        rng.init(S.seed, pixel, sample);
        const float2 off = rng.next2D();
        const float2 samplePos = make_float2((float)(pixel % W) + off.x, (float)(pixel / W) + off.y);
        float3 o, d;
        float mint, maxt;
        sampleCameraRay(S.camera, samplePos, o, d, mint, maxt);
        P.rayO[i] = make_float4(o.x, o.y, o.z, mint);
        P.rayD[i] = make_float4(d.x, d.y, d.z, maxt);
        P.thr[i] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        P.rad[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        P.pos[i] = make_float4(samplePos.x, samplePos.y, __uint_as_float((uint32_t)rng.state),
                               __uint_as_float((uint32_t)(rng.state >> 32)));
        P.flags[i] = 1u | kFlagFirst;  // rRec.newQuery: depth = 1 (integrator.h:218-225)
        P.slot[i] = B.slotBase + i;
        P.medium[i] = S.camera.medium;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) C->queue[0] = B.nPaths;
}

// ------------------------------------------------------------------------------------------
// Traversal. One thread per ray, warps fetch 32 rays at a time from a device work counter
// (dynamic load balancing: rays of one queue differ a lot in traversal length).
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Long rays. On an open scene a few grazing rays need several hundred node visits (bvhsim, 10 M-triangle terrain: 99.9 %
// quantile 142, maximum 516 against a mean of 26): one dependent L2 / DRAM round trip per visit, ~0.5 us each, while the
// rest of the GPU idles -- per step of that config bounces 2-5 took 490-820 us each for < 10 % of the rays, ~3 ms of the 4.6 ms
// traversal time. The traversal kernels therefore give a ray a budget of node visits (TailList::visits); a ray that needs more (and whatever
// a warp still holds once the queue has run dry and kTailLanes or fewer lanes are busy) goes to a tail list, and
// k_trace_tail finishes it with a whole WARP: its pending subtrees sit on a shared-memory stack, every lane pops one entry
// per round, tests that node's two children (or that leaf's primitives) and pushes what the ray enters; the closest hit is
// a warp-wide minimum per round. A ray that misses everything -- the common case for the long ones -- has no ordering
// constraint at all, so the dependent chain shrinks from the number of visits to roughly the depth of the tree.
// Same node and primitive tests as traceRay(); a round prunes with the hit distance of the previous round, so a few more
// nodes are tested than sequentially. Exact-t ties between two primitives (shared edges) go to the lower primitive slot.
// ------------------------------------------------------------------------------------------
// The budget is a launch parameter (TailList::visits, integrator.cu: 96 for scenes whose BVH is deep enough to have such rays, 0 =
// off for the Cornell-size scenes, where the extra launches cost more than they can save; measured on C4: 64 / 96 / 160 visits ->
// 546 / 595 / 583 M paths/s, kTailLanes 0 / 4 / 12 -> 572 / 595 / 573).
#ifndef PG_TAIL_LANES
#define PG_TAIL_LANES 4
#endif
static constexpr int kTailLanes = PG_TAIL_LANES;

// all lanes of a converged warp; lanes with `pred` append rayIdx
PG_DEV void tailDefer(const TailList &T, bool pred, uint32_t rayIdx) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)laneId() == leader) base = atomicAdd(T.count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) T.list[base + (uint32_t)__popc(m & ((1u << laneId()) - 1u))] = rayIdx;
}

// traceRay() with the visit budget: false + aborted = true when the ray needs more than `budget` node visits
template <bool kCount>
PG_DEV void traceRayBudget(const DeviceScene &S, float3 o, float3 d, float mint, float maxt, Hit &hit, uint32_t budget, bool &aborted,
                           uint32_t *cntNodes, uint32_t *cntPrims) {
    const float3 idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[kTraceStack];
    int sp = 0, node = 0;
    uint32_t visits = 0;
    hit.prim = kMiss;
    hit.t = maxt;
    float tmax = maxt;
    aborted = false;
    while (true) {
        while (node >= 0) {
            if (kCount) (*cntNodes)++;
            if (++visits > budget) {
                aborted = true;
                return;
            }
            node = bvhNodeStep(S, node, o, idir, mint, tmax, stack, sp);
        }
        if (node == kDoneNode) break;
        bvhLeafStep<false, kCount>(S, node, o, d, mint, tmax, hit, cntPrims);
        node = sp ? stack[--sp] : kDoneNode;
        if (node == kDoneNode) break;
    }
}

// entries per warp. Up to kCoopWide entries a round takes up to 32 of them (and pushes at most 64: <= kCoopWide + 32 afterwards);
// above that it takes one -- a depth-first walk, which adds at most the tree depth (<= kTraceStack = 64) on top.
static constexpr int kCoopStack = 320, kCoopWide = 192;

// One ray, served by all 32 lanes (every lane holds the same ray; `hit` comes back identical in all lanes). cs = the warp's
// shared-memory stack.
template <bool kAnyHit, bool kCount>
PG_DEV void traceRayCooperative(const DeviceScene &S, int *cs, float3 o, float3 d, float mint, float tmax, Hit &best, uint32_t &cntNodes,
                                uint32_t &cntPrims) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = (int)laneId();
    const float3 idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    best.prim = kMiss;
    best.t = tmax;
    best.u = best.v = 0.0f;
    int top = 1;
    if (lane == 0) cs[0] = 0;  // the root
    __syncwarp();
    bool found = false;
    while (top > 0 && !found) {
        const int take = top > kCoopWide ? 1 : min(top, 32);
        const int entry = lane < take ? cs[top - 1 - lane] : kDoneNode;
        top -= take;
        __syncwarp();
        int c0 = kDoneNode, c1 = kDoneNode, nh = 0;
        Hit cand;
        cand.prim = kMiss;
        cand.t = kInf;
        cand.u = cand.v = 0.0f;
        if (entry >= 0) {
            if (kCount) cntNodes++;
            nh = bvhTestNode(S, entry, o, idir, mint, tmax, c0, c1);
        } else if (entry != kDoneNode) {
            float tm = tmax;
            if (bvhLeafStep<kAnyHit, kCount>(S, entry, o, d, mint, tm, cand, &cntPrims)) cand.t = 0.0f;  // any-hit: found
        }
        // push: the farther children first, the nearer ones on top (lane 0's on the very top)
        const unsigned mFar = __ballot_sync(kFull, nh == 2), mNear = __ballot_sync(kFull, nh >= 1);
        if (nh == 2) cs[top + __popc(mFar & ((1u << lane) - 1u))] = c1;
        if (nh >= 1) cs[top + __popc(mFar) + __popc(mNear >> lane) - 1] = c0;
        top += __popc(mFar) + __popc(mNear);
        // closest candidate of the round (non-negative floats order like their bit patterns)
        const unsigned mHit = __ballot_sync(kFull, cand.prim != kMiss);
        if (mHit) {
            if (kAnyHit) {
                best.prim = __shfl_sync(kFull, cand.prim, __ffs(mHit) - 1);
                found = true;
            } else {
                const uint32_t tb = cand.prim != kMiss ? __float_as_uint(cand.t) : 0x7F800000u;
                const uint32_t tmin = __reduce_min_sync(kFull, tb);
                const uint32_t pmin = __reduce_min_sync(kFull, tb == tmin ? cand.prim : 0xFFFFFFFFu);
                const unsigned mWin = __ballot_sync(kFull, tb == tmin && cand.prim == pmin);
                const int win = __ffs(mWin) - 1;
                const float tw = __uint_as_float(tmin);
                const float wu = __shfl_sync(kFull, cand.u, win), wv = __shfl_sync(kFull, cand.v, win);
                if (tw < tmax || (tw == tmax && pmin < best.prim)) {
                    best.prim = pmin;
                    best.t = tw;
                    best.u = wu;
                    best.v = wv;
                    tmax = tw;
                }
            }
        }
        __syncwarp();
    }
}

// The tail kernel: one warp per listed ray, from the root.
template <bool kAnyHit, bool kCount, typename Queue>
__global__ void __launch_bounds__(128) k_trace_tail(DeviceScene S, Queue Q, TailList T, uint32_t *work, Counters *C) {
    __shared__ int sStack[4][kCoopStack];
    int *cs = sStack[threadIdx.x >> 5];
    const uint32_t n = *T.count;
    uint32_t cn = 0, cp = 0;
    while (true) {
        uint32_t k = 0;
        if (laneId() == 0) k = atomicAdd(work, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n) break;
        const uint32_t i = T.list[k];
        float3 o, d;
        float mint, tmax;
        Q.refetch(i, o, d, mint, tmax);  // every lane reads the same ray (broadcast)
        Hit hit;
        traceRayCooperative<kAnyHit, kCount>(S, cs, o, d, mint, tmax, hit, cn, cp);
        if (laneId() == 0) Q.finish(i, hit);
    }
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}

template <bool kCount, bool kKey>
__global__ void __launch_bounds__(128) k_trace(DeviceScene S, const float4 *__restrict__ rayO, const float4 *__restrict__ rayD,
                                               const uint32_t *__restrict__ flags, float4 *__restrict__ hits,
                                               const uint32_t *nPtr, uint32_t *work, Counters *C, SortArgs Q, TailList T) {
    const uint32_t n = *nPtr;
    unsigned long long rays = 0;
    uint32_t cn = 0, cp = 0;
    while (true) {
        uint32_t base = 0;
        if (laneId() == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + laneId();
        uint32_t bin = 0xFFFFFFFFu;  // lanes past the end of the queue
        bool aborted = false;
        if (i < n) {
            Hit h;
            h.prim = kMiss;
            h.t = kInf;
            h.u = h.v = 0;
            bin = 0;
            if (!(flags[i] & (kFlagDead | kFlagNoTrace))) {
                const float4 ro = rayO[i], rd = rayD[i];
                const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
                const float mint = adaptiveMinT(o, ro.w, false);
                if (kKey) traceRay<false, kCount>(S, o, d, mint, rd.w, h, &cn, &cp);
                else traceRayBudget<kCount>(S, o, d, mint, rd.w, h, T.visits ? T.budget(n) : 0xFFFFFFFFu, aborted, &cn, &cp);
                if (h.prim == kMiss) h.t = kInf;
                else if (kKey) {  // guiding cell of the hit point (ordering only: the shade stage looks its cell up itself)
                    const float3 p = o + d * h.t;
                    uint32_t nd = 0;
                    while (true) {
                        const uint4 g = __ldg(Q.guideNodes + nd);
                        if (g.x == 3u) { bin = 1u + g.z; break; }
                        nd = comp(p, (int)g.x) < __uint_as_float(g.y) ? g.z : g.z + 1;
                    }
                }
                rays++;
            }
            if (!aborted) hits[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
        }
        if (!kKey) tailDefer(T, aborted, i);  // k_trace_tail writes the hit record of a deferred ray
        if (kKey) {  // rank inside the bin: one atomic per distinct bin of the warp
            const uint32_t peers = __match_any_sync(0xffffffffu, bin);
            const int leader = __ffs(peers) - 1;
            uint32_t first = 0;
            if (bin != 0xFFFFFFFFu && (int)laneId() == leader) first = atomicAdd(Q.binCount + bin, (uint32_t)__popc(peers));
            first = __shfl_sync(peers, first, leader);
            if (bin != 0xFFFFFFFFu) {
                Q.key[i] = bin;
                Q.rank[i] = first + (uint32_t)__popc(peers & ((1u << laneId()) - 1u));
            }
        }
    }
    warpAddU64(&C->normalRays, rays);
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}

// ------------------------------------------------------------------------------------------
// Persistent speculative traversal with per-lane refill (incoherent queues: bounces >= 1 and shadow rays).
//
// The batch-of-32 kernel above runs a warp until its slowest ray is done. On the 10 M-triangle mesh the secondary rays
// need 26 node visits on average but 80 at the 99th percentile, and a ray that has just been fetched needs ~23 visits to
// its first leaf while its neighbours need 2-3 to their next one: ncu shows 4.7-7.7 of 32 lanes per instruction there
// (profiles/r01: prof_trace_mesh). tools/bvhsim replays the same traversal on the host under different warp schedules
// (profiles/r02_bvhsim.txt): lock-step batches issue 93.7 node steps per ray slot on the diffuse-bounce set, refill +
// speculation 56.6. Hence:
//   * a lane that finishes its ray fetches the next one from the queue cursor as soon as fewer than kRefillBelow lanes are
//     busy (one atomic per refill of the warp);
//   * speculative descent: a lane that reaches a leaf postpones it and keeps descending (its interval is not shortened
//     yet, which is conservative) until every lane of the warp holds a leaf -- then all leaves are tested together.
// Results are identical to traceRay(): the same node / primitive tests, and a closest hit does not depend on visit order
// (ties: the lowest t wins, equal t keeps the first found -- the reference has no rule either, skdtree.cpp:112-142).
// ------------------------------------------------------------------------------------------
#ifndef PG_REFILL_BELOW
#define PG_REFILL_BELOW 24
#endif
#ifndef PG_TRACE_MIN_BLOCKS
#define PG_TRACE_MIN_BLOCKS 0  // resident CTAs per SM the speculative kernels are compiled for (0 = compiler's choice: 56 registers, 9 CTAs; A/B: 10 = 48 registers, 12 = 40)
#endif
#if PG_TRACE_MIN_BLOCKS > 0
#define PG_TRACE_BOUNDS __launch_bounds__(128, PG_TRACE_MIN_BLOCKS)
#else
#define PG_TRACE_BOUNDS __launch_bounds__(128)
#endif
static constexpr int kRefillBelow = PG_REFILL_BELOW;

// kWide: walk the 8-ary quantised tree (large meshes) instead of the binary one -- same leaves, same primitive tests.
template <bool kWide>
struct TraceStack;
template <>
struct TraceStack<false> {
    int e[kTraceStack];
    PG_DEV int step(const DeviceScene &S, int node, float3 o, float3 idir, float mint, float tmax, int &sp) {
#ifndef PG_TRACE_PREFETCH
#define PG_TRACE_PREFETCH 0  // measured (r2d): no gain on C4 (407.5 vs 407.7 M paths/s), C2 trace stage 1.33 -> 3.80 ms
#endif
        return bvhNodeStep<PG_TRACE_PREFETCH != 0>(S, node, o, idir, mint, tmax, e, sp);
    }
    PG_DEV int pop(int &sp, float) { return sp ? e[--sp] : kDoneNode; }
};
template <>
struct TraceStack<true> {
    WideEntry e[kWideStack];
    PG_DEV int step(const DeviceScene &S, int node, float3 o, float3 idir, float mint, float tmax, int &sp) {
        return wideNodeStep(S, node, o, idir, mint, tmax, e, sp);
    }
    PG_DEV int pop(int &sp, float tmax) { return widePop(e, sp, tmax); }
};

template <bool kAnyHit, bool kCount, bool kWide, typename Queue>
PG_DEV void traceQueueSpeculative(const DeviceScene &S, Queue Q, uint32_t n, uint32_t *work, uint32_t &cntNodes, uint32_t &cntPrims,
                                  unsigned long long &cntRays, TailList T) {
    constexpr unsigned kFull = 0xffffffffu;
    const bool kTail = !kWide && T.visits > 0;  // the cooperative kernel walks the binary tree
    const uint32_t budget = T.budget(n);
    TraceStack<kWide> stack;
    int sp = 0, node = kDoneNode, leaf = 0;
    uint32_t visits = 0;
    uint32_t rayIdx = 0xFFFFFFFFu;
    float3 o = f3(0.0f), d = f3(0.0f), idir = f3(0.0f);
    float mint = 0.0f, tmax = 0.0f;
    Hit hit;
    hit.prim = kMiss;
    hit.t = hit.u = hit.v = 0.0f;
    bool exhausted = false;  // warp-uniform: the queue cursor has passed the end
    while (true) {
        // ---- refill: lanes without a ray take consecutive queue entries
        unsigned idle = __ballot_sync(kFull, rayIdx == 0xFFFFFFFFu);
        while (!exhausted && __popc(idle) > 32 - kRefillBelow) {
            const uint32_t cnt = (uint32_t)__popc(idle);
            uint32_t base = 0;
            if (laneId() == 0) base = atomicAdd(work, cnt);
            base = __shfl_sync(kFull, base, 0);
            if (base + cnt >= n) exhausted = true;
            if (rayIdx == 0xFFFFFFFFu) {
                const uint32_t i = base + (uint32_t)__popc(idle & ((1u << laneId()) - 1u));
                if (i < n && Q.fetch(i, o, d, mint, tmax)) {
                    rayIdx = i;
                    idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    if (kWide) idir = wideClampIdir(idir);
                    node = 0;
                    sp = 0;
                    leaf = 0;
                    visits = 0;
                    hit.prim = kMiss;
                    hit.t = tmax;
                    hit.u = hit.v = 0.0f;
                    cntRays++;
                }
            }
            idle = __ballot_sync(kFull, rayIdx == 0xFFFFFFFFu);
        }
        if (idle == kFull) break;
        if (kTail) {
            // over budget, or the last few rays of a drained queue: hand them to the cooperative kernel
            const bool last = exhausted && __popc(~idle) <= kTailLanes;
            const bool defer = rayIdx != 0xFFFFFFFFu && (last || visits > budget);
            tailDefer(T, defer, rayIdx);
            if (defer) rayIdx = 0xFFFFFFFFu, node = kDoneNode, leaf = 0;
            if (last) break;
            if (__ballot_sync(kFull, rayIdx != 0xFFFFFFFFu) == 0u) continue;  // everything deferred: refill
        }
        // ---- speculative descent: until every lane with a ray holds a leaf (or has nothing left to visit)
        while (__any_sync(kFull, node >= 0 && leaf == 0)) {
            if (node >= 0) {
                if (kCount) cntNodes++;
                visits++;
                node = stack.step(S, node, o, idir, mint, tmax, sp);
                if (node < 0 && node != kDoneNode && leaf == 0) {  // first leaf: postpone it, continue with the next subtree
                    leaf = node;
#if PG_TRACE_PREFETCH
                    prefetchRef(S, leaf);  // its primitive records are tested only after the speculative descent
#endif
                    node = stack.pop(sp, tmax);
                }
            }
        }
        // ---- leaves: the postponed one, then any leaf the lane is standing on
        while (leaf != 0) {
            const bool found = bvhLeafStep<kAnyHit, kCount>(S, leaf, o, d, mint, tmax, hit, &cntPrims);
            leaf = 0;
            if (kAnyHit && found) {
                node = kDoneNode;
            } else if (node < 0 && node != kDoneNode) {
                leaf = node;
                node = stack.pop(sp, tmax);
            }
        }
        // ---- finished rays hand in their result and become idle
        if (rayIdx != 0xFFFFFFFFu && node == kDoneNode) {
            Q.finish(rayIdx, hit);
            rayIdx = 0xFFFFFFFFu;
        }
    }
}

struct ClosestQueue {  // the wavefront's ray queue (path state) -> hit records
    const float4 *rayO, *rayD;
    const uint32_t *flags;
    float4 *hits;
    PG_DEV bool fetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {
        if (flags[i] & (kFlagDead | kFlagNoTrace)) {
            hits[i] = make_float4(kInf, 0.0f, 0.0f, __uint_as_float(kMiss));
            return false;
        }
        const float4 ro = rayO[i], rd = rayD[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, false);
        tmax = rd.w;
        return true;
    }
    PG_DEV void finish(uint32_t i, const Hit &h) const {
        hits[i] = h.prim == kMiss ? make_float4(kInf, 0.0f, 0.0f, __uint_as_float(kMiss))
                                  : make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    }
    PG_DEV void refetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {  // a ray that passed fetch() before
        const float4 ro = rayO[i], rd = rayD[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, false);
        tmax = rd.w;
    }
};
struct ShadowRayQueue {  // shadow queue -> adds the contribution of unoccluded rays to the path record
    ShadowQueue Q;
    float4 *rad;
    PG_DEV bool fetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {
        const float4 ro = Q.o[i], rd = Q.d[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, true);
        tmax = rd.w;
        if (!(tmax > mint)) {  // empty interval: unoccluded (as k_shadow)
            Hit none;
            none.prim = kMiss;
            finish(i, none);
            return false;
        }
        return true;
    }
    PG_DEV void refetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {
        const float4 ro = Q.o[i], rd = Q.d[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, true);
        tmax = rd.w;
    }
    PG_DEV void finish(uint32_t i, const Hit &h) const {
        if (h.prim != kMiss) return;
        const float4 c = Q.c[i];
        const uint32_t dst = __float_as_uint(c.w);
        float4 r = rad[dst];  // exactly one shadow ray per path and bounce: no race
        r.x += c.x;
        r.y += c.y;
        r.z += c.z;
        rad[dst] = r;
    }
};

struct RayListQueue {  // stand-alone ray queries (b200pg_k_trace*): rays as {o, mint}{d, maxt}, hits with global primitive ids
    const float4 *rays;
    float4 *hits;
    const uint32_t *primGlobalId;
    bool shadow;
    PG_DEV bool fetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {
        const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, shadow);
        tmax = rd.w;
        if (!(tmax > mint)) {
            hits[i] = make_float4(kInf, 0.0f, 0.0f, __uint_as_float(kMiss));
            return false;
        }
        return true;
    }
    PG_DEV void refetch(uint32_t i, float3 &o, float3 &d, float &mint, float &tmax) const {
        const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        mint = adaptiveMinT(o, ro.w, shadow);
        tmax = rd.w;
    }
    PG_DEV void finish(uint32_t i, const Hit &h) const {
        // any-hit mode reports the primitive only (as k_trace_rays: t, u, v of an occluder are whatever was found first)
        hits[i] = h.prim == kMiss ? make_float4(kInf, 0.0f, 0.0f, __uint_as_float(kMiss))
                                  : make_float4(h.t, h.u, h.v, __uint_as_float(primGlobalId[h.prim]));
    }
};
template <bool kShadow, bool kCount, bool kWide>
__global__ void __launch_bounds__(128) k_trace_rays_spec(DeviceScene S, RayListQueue Q, uint32_t n, uint32_t *work, Counters *C,
                                                         TailList T) {
    unsigned long long rays = 0;
    uint32_t cn = 0, cp = 0;
    traceQueueSpeculative<kShadow, kCount, kWide>(S, Q, n, work, cn, cp, rays, T);
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}
template <bool kCount, bool kWide>
__global__ void PG_TRACE_BOUNDS k_trace_spec(DeviceScene S, ClosestQueue Q, const uint32_t *nPtr, uint32_t *work, Counters *C,
                                                    TailList T) {
    unsigned long long rays = 0;
    uint32_t cn = 0, cp = 0;
    traceQueueSpeculative<false, kCount, kWide>(S, Q, *nPtr, work, cn, cp, rays, T);
    warpAddU64(&C->normalRays, rays);
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}
template <bool kCount, bool kWide>
__global__ void PG_TRACE_BOUNDS k_shadow_spec(DeviceScene S, ShadowRayQueue Q, const uint32_t *nPtr, uint32_t *work, Counters *C,
                                                     TailList T) {
    unsigned long long rays = 0;
    uint32_t cn = 0, cp = 0;
    const uint32_t n = *nPtr;
    traceQueueSpeculative<true, kCount, kWide>(S, Q, n, work, cn, cp, rays, T);
    // every queued shadow ray counts (the batch kernel counts empty intervals as well)
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&C->shadowRays, (unsigned long long)n);
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}

// Coherence sort, step 2: exclusive scan of the 1 + cells bin counters (single block; the counters are re-zeroed for the
// next bounce) and step 3: permutation.
__global__ void __launch_bounds__(1024) k_bin_scan(SortArgs Q) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sCarry;
    const uint32_t nBins = 1u + *Q.nCells;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nBins; base += 1024u) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t v = b < nBins ? Q.binCount[b] : 0u;
        if (b < nBins) Q.binCount[b] = 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)laneId() >= o) x += y;
        }
        if (laneId() == 31) sWarp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = sWarp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if ((int)laneId() >= o) w += y;
            }
            sWarp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warpBase = (threadIdx.x >> 5) ? sWarp[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t carry = sCarry;
        if (b < nBins) Q.binOffset[b] = carry + warpBase + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) sCarry = carry + warpBase + x;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_bin_scatter(SortArgs Q, const uint32_t *nPtr) {
    const uint32_t n = *nPtr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        Q.perm[__ldg(Q.binOffset + Q.key[i]) + Q.rank[i]] = i;
}

// Hit / miss partition of a traced queue (open scenes): perm = [entries with a vertex to shade ..., the rest (missed rays, parked
// records) in reverse]. On the 10 M-triangle terrain 90-95 % of the bounce-1 rays leave the scene; shaded in queue order, 4 of 5 warps
// still hold one or two live vertices and walk the whole shading code for them (ncu, guided C4 step: k_shade 2.8 ms on bounce 1 for
// 0.3 M live vertices, 6 lanes per instruction). Through the permutation the live vertices fill whole warps and the misses form
// warps that only close their path. Order inside each class follows the queue (block-wise), so the gathers stay nearly coalesced.
__global__ void __launch_bounds__(256) k_hit_partition(const float4 *__restrict__ hits, const uint32_t *__restrict__ flags,
                                                       const uint32_t *nPtr, uint32_t *__restrict__ perm, uint32_t *cntNeed,
                                                       uint32_t *cntRest) {
    __shared__ uint32_t sWarp[2][8];
    __shared__ uint32_t sBase[2];
    const uint32_t n = *nPtr;
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        const uint32_t i = base + threadIdx.x;
        bool need = false, rest = false;
        if (i < n) {
            need = __float_as_uint(ldStream(&hits[i].w)) != kMiss && !(ldStream(flags + i) & kFlagDead);
            rest = !need;
        }
        const unsigned mN = __ballot_sync(0xffffffffu, need), mR = __ballot_sync(0xffffffffu, rest);
        if (laneId() == 0) {
            sWarp[0][warp] = __popc(mN);
            sWarp[1][warp] = __popc(mR);
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            uint32_t total = 0;
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = sWarp[threadIdx.x][w];
                sWarp[threadIdx.x][w] = total;
                total += c;
            }
            sBase[threadIdx.x] = total ? atomicAdd(threadIdx.x ? cntRest : cntNeed, total) : 0u;
        }
        __syncthreads();
        const unsigned lt = (1u << laneId()) - 1u;
        if (need) perm[sBase[0] + sWarp[0][warp] + __popc(mN & lt)] = i;
        if (rest) perm[n - 1u - (sBase[1] + sWarp[1][warp] + __popc(mR & lt))] = i;
        __syncthreads();
    }
}

// Shadow rays: any-hit; an unoccluded ray adds its contribution to the path record it belongs to.
template <bool kCount>
__global__ void __launch_bounds__(128) k_shadow(DeviceScene S, ShadowQueue Q, float4 *__restrict__ rad, const uint32_t *nPtr,
                                                uint32_t *work, Counters *C) {
    const uint32_t n = *nPtr;
    unsigned long long rays = 0;
    uint32_t cn = 0, cp = 0;
    while (true) {
        uint32_t base = 0;
        if (laneId() == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + laneId();
        if (i < n) {
            const float4 ro = Q.o[i], rd = Q.d[i];
            const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
            const float mint = adaptiveMinT(o, ro.w, true);
            Hit h;
            rays++;
            bool occluded = rd.w > mint && traceRay<true, kCount>(S, o, d, mint, rd.w, h, &cn, &cp);
            if (!occluded) {
                const float4 c = Q.c[i];
                const uint32_t dst = __float_as_uint(c.w);
                float4 r = rad[dst];  // exactly one shadow ray per path and bounce: no race
                r.x += c.x;
                r.y += c.y;
                r.z += c.z;
                rad[dst] = r;
            }
        }
    }
    warpAddU64(&C->shadowRays, rays);
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}

// Standalone ray queries for the parity tests / traversal benchmark (b200pg_k_trace*).
template <bool kShadow, bool kCount>
__global__ void __launch_bounds__(128) k_trace_rays(DeviceScene S, const float4 *__restrict__ rays, uint32_t n,
                                                    float4 *__restrict__ hits, uint32_t *work, Counters *C) {
    uint32_t cn = 0, cp = 0;
    while (true) {
        uint32_t base = 0;
        if (laneId() == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + laneId();
        if (i < n) {
            const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
            const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
            const float mint = adaptiveMinT(o, ro.w, kShadow);
            Hit h;
            h.prim = kMiss;
            h.t = kInf;
            h.u = h.v = 0;
            if (rd.w > mint) traceRay<kShadow, kCount>(S, o, d, mint, rd.w, h, &cn, &cp);
            if (h.prim == kMiss) { h.t = kInf; h.u = h.v = 0; }
            const uint32_t gid = h.prim == kMiss ? kMiss : S.primGlobalId[h.prim];
            hits[i] = make_float4(h.t, h.u, h.v, __uint_as_float(gid));
        }
    }
    if (kCount) {
        warpAddU64(&C->nodesVisited, cn);
        warpAddU64(&C->primsTested, cp);
    }
}

// ------------------------------------------------------------------------------------------
// Shade stage (surface path tracer, ProgressiveMIPathTracer::Li)
// ------------------------------------------------------------------------------------------

#ifndef PG_SHADE_BLOCKS
#define PG_SHADE_BLOCKS 8
#endif
__global__ void __launch_bounds__(kShadeThreads, PG_SHADE_BLOCKS) k_shade(ShadeArgs A) {
    const DeviceScene &S = A.S;
    const IntegratorConfig &cfg = A.cfg;
    const uint32_t n = A.C->queue[A.bounce];
    uint32_t *nextCount = &A.C->queue[A.bounce + 1];
    uint32_t *shadowCount = &A.C->shadow[A.bounce];
    unsigned long long donePaths = 0, doneLen = 0;
    __shared__ uint32_t sAppend[2 * 3 * (kShadeThreads / 32 + 1)];
    uint32_t appendParity = 0;

    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t q = base + threadIdx.x;
        const bool valid = q < n;
        const uint32_t i = (valid && A.perm) ? A.perm[q] : q;

        bool alive = false;      // continues into the next queue
        bool wantShadow = false;
        bool terminate = false;
        float4 ro, rd, thr4, rad4, pos4;
        uint32_t fl = 0, slot = 0;
        int medium = -1;
        float3 L = f3(0.0f), thr = f3(1.0f);
        float eta = 1.0f;
        Rng rng;
        rng.state = 0;
        rng.inc = 1;
        float3 newO = f3(0.0f), newD = f3(0.0f);
        float newPdf = 0.0f;
        float3 shO = f3(0.0f), shD = f3(0.0f), shC = f3(0.0f);
        float shMaxT = 0.0f;
        uint32_t depth = 0, vcount = 0;

        if (valid) {
            ro = ldStream(A.cur.rayO + i);
            rd = ldStream(A.cur.rayD + i);
            thr4 = ldStream(A.cur.thr + i);
            rad4 = ldStream(A.cur.rad + i);
            pos4 = ldStream(A.cur.pos + i);
            fl = ldStream(A.cur.flags + i);
            slot = ldStream(A.cur.slot + i);
            medium = ldStream(A.cur.medium + i);
            const float4 h4 = ldStream(A.hits + i);
            L = f3(rad4.x, rad4.y, rad4.z);
            thr = f3(thr4.x, thr4.y, thr4.z);
            eta = thr4.w;
            depth = fl & kDepthMask;
            const float2 samplePos = make_float2(pos4.x, pos4.y);
            const uint32_t pixel = (uint32_t)samplePos.y * (uint32_t)S.film.width + (uint32_t)samplePos.x;
            rng.state = ((uint64_t)__float_as_uint(pos4.w) << 32) | (uint64_t)__float_as_uint(pos4.z);
            rng.inc = ((uint64_t)pixel << 1) | 1ULL;
            const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
            Hit h;
            h.t = h4.x;
            h.u = h4.y;
            h.v = h4.z;
            h.prim = __float_as_uint(h4.w);
            vcount = (fl >> kVertShift) & 0xFFu;
            if (A.G.record && vcount > 0 && !(fl & kFlagVertexClosed)) {
                // close the previous training vertex: radiance gathered so far (its NEE has landed by now)
                // and the distance to this hit
                guideVertexClose(A.G, slot, vcount - 1, thr, h.prim == kMiss ? 0.0f : h.t, L);
                fl |= kFlagVertexClosed;
            }

            if (fl & kFlagDead) {
                terminate = true;
            } else if (h.prim == kMiss) {
                terminate = true;  // no environment emitter on this path (progressive_path.cpp:150-159)
            } else {
                Intersection its;
                fillIntersection(S, o, d, h, its);
                const BsdfRecord &bsdf = S.bsdfs[its.bsdf];

                // ---- emitted radiance: directly visible (first hit) or reached by BSDF sampling (MIS)
                if (its.emitter >= 0) {
                    const float3 Le = dot(its.sh.n, -d) <= 0 ? f3(0.0f) : ld3(S.emitters[its.emitter].radiance);  // area.cpp:104-109
                    if (fl & kFlagFirst) {
                        if (!cfg.hideEmitters) L += thr * Le;  // progressive_path.cpp:167-169
                    } else {
                        const float lumPdf = (cfg.useNee && !(fl & kFlagPrevDelta))
                                                 ? pdfEmitterDirect(S, its.emitter, d, its.sh.n, its.t) : 0.0f;
                        const float weight = cfg.useNee ? miWeight(rad4.w, lumPdf) : 1.0f;
                        L += thr * Le * weight;  // progressive_path.cpp:276-284
                    }
                }
                // ---- Russian roulette closes the previous loop iteration (progressive_path.cpp:296-306)
                if (!(fl & kFlagFirst)) {
                    if (depth++ >= (uint32_t)cfg.rrDepth) {
                        const float q = fminf(maxComp(thr) * eta * eta, 0.95f);
                        if (rng.next1D() >= q)
                            terminate = true;
                        else
                            thr = thr / q;
                    }
                }
                // ---- loop condition and depth / strict-normal stops (:149, :175-184)
                if (!terminate && !((int)depth <= cfg.maxDepth || cfg.maxDepth < 0)) terminate = true;
                if (!terminate && (((int)depth >= cfg.maxDepth && cfg.maxDepth > 0) ||
                                   (cfg.strictNormals && dot(d, its.geoN) * its.wi.z >= 0)))
                    terminate = true;

                if (!terminate) {
                    // ---- direct illumination sampling (:191-219)
                    const uint32_t btype = bsdf.typeFlags;
                    // guided vertex: smooth BSDF only -- delta lobes are never guided
                    const bool guided = A.G.enabled && (btype & kSmooth);
                    const uint32_t gcell = guided ? guideLookup(A.G, its.p) : 0u;
                    // NEE (:191-219). At a guided vertex the MIS weight needs the mixture pdf of the light direction; it is
                    // evaluated together with the pdf of the sampled direction in ONE pass over the cell's lobes below.
                    bool neePending = false;
                    float neeBsdfPdf = 0.0f, neeLightPdf = 0.0f;
                    float3 neeContrib = f3(0.0f);
                    if (cfg.useNee && (btype & kSmooth)) {
                        const float3 refN = (btype & (kTransmission | kBackSide)) == 0 ? its.sh.n : f3(0.0f);  // records.inl:160-164
                        DirectSample dRec;
                        const float2 u = rng.next2D();
                        const float3 value = sampleEmitterDirect(S, its.p, refN, u, dRec);
                        if (!isZero(value)) {
                            const float3 woL = its.sh.toLocal(dRec.d);
                            const float3 bsdfVal = bsdfEval(bsdf, its.wi, woL);
                            if (!isZero(bsdfVal) && (!cfg.strictNormals || dot(its.geoN, dRec.d) * woL.z > 0)) {
                                neeBsdfPdf = bsdfPdf(bsdf, its.wi, woL);
                                neeLightPdf = dRec.pdf;
                                neeContrib = thr * value * bsdfVal;
                                shO = its.p;
                                shD = dRec.d;
                                shMaxT = dRec.dist * (1 - kShadowEpsilon);  // scene.cpp:883-884
                                neePending = true;
                            }
                        }
                    }
                    // ---- BSDF sampling (:226-238)
                    float bPdf, bEta;
                    uint32_t sampledType;
                    float3 woL, wo, bsdfWeight;
                    if (guided) {
                        // one-sample MIS between the guiding mixture (probability alpha) and the BSDF
                        float u0 = rng.next1D();
                        const float2 u12 = rng.next2D();
                        float3 fcos;
                        float pb;
                        bool ok = true;
                        if (u0 < A.G.alpha) {
                            u0 /= A.G.alpha;
                            wo = guideSample(A.G, gcell, u0, u12.x, u12.y);
                            woL = its.sh.toLocal(wo);
                            fcos = bsdfEval(bsdf, its.wi, woL);
                            pb = bsdfPdf(bsdf, its.wi, woL);
                            bEta = 1.0f;
                            sampledType = kGlossyReflection;
                        } else {
                            const float3 w = bsdfSample(bsdf, its.wi, u12, woL, pb, bEta, sampledType);
                            ok = !isZero(w);
                            fcos = w * pb;
                            wo = its.sh.toWorld(woL);
                        }
                        // one pass over the cell's lobes for both directions (a single code path keeps the warp converged;
                        // an unused direction is evaluated on a dummy and discarded)
                        float gNee = 0.0f, gWo = 0.0f;
                        if (neePending || ok) guidePdf2(A.G, gcell, neePending ? shD : wo, ok ? wo : shD, gNee, gWo);
                        if (neePending) neeBsdfPdf = A.G.alpha * gNee + (1 - A.G.alpha) * neeBsdfPdf;
                        bPdf = ok ? A.G.alpha * gWo + (1 - A.G.alpha) * pb : 0.0f;
                        bsdfWeight = (ok && !isZero(fcos) && bPdf > 0) ? fcos / bPdf : f3(0.0f);
                    } else {
                        bsdfWeight = bsdfSample(bsdf, its.wi, rng.next2D(), woL, bPdf, bEta, sampledType);
                        wo = its.sh.toWorld(woL);
                    }
                    if (neePending) {
                        shC = neeContrib * miWeight(neeLightPdf, neeBsdfPdf);
                        wantShadow = true;
                    }
                    if (isZero(bsdfWeight)) {
                        terminate = true;
                    } else {
                        if (cfg.strictNormals && dot(its.geoN, wo) * woL.z <= 0) {
                            terminate = true;
                        } else {
                            thr *= bsdfWeight;  // (:271-272; a miss of the new ray terminates next bounce)
                            eta *= bEta;
                            newO = its.p;
                            newD = wo;
                            newPdf = bPdf;
                            fl &= ~(kFlagFirst | kFlagPrevDelta);
                            if (sampledType & kDelta) fl |= kFlagPrevDelta;
                            if (sampledType != kNull) fl |= kFlagScattered;
                            alive = true;
                            if (A.G.record && (btype & kSmooth) && (int)vcount < A.G.maxVerts) {
                                guideVertexOpen(A.G, slot, vcount, its.p, bPdf, wo, gcell);
                                vcount++;
                                fl &= ~kFlagVertexClosed;
                            }
                        }
                    }
                }
            }
            if (terminate && wantShadow) {
                // park the record for one bounce so that the shadow ray has somewhere to land
                alive = true;
                fl |= kFlagDead;
                terminate = false;
            }
        }

        // ---- compaction into the next queue / shadow queue (one atomic per block each)
        const AppendResult ap = blockAppend3(nextCount, alive, shadowCount, wantShadow, A.G.record ? A.G.sCount : nullptr,
                                             (A.G.record && valid && terminate) ? vcount : 0u, sAppend, appendParity);
        appendParity ^= 1u;
        const uint32_t j = ap.idxA, sidx = ap.idxB;
        if (alive) {
            stStream(A.next.rayO + j, make_float4(newO.x, newO.y, newO.z, kEpsilon));
            stStream(A.next.rayD + j, make_float4(newD.x, newD.y, newD.z, kInf));
            stStream(A.next.thr + j, make_float4(thr.x, thr.y, thr.z, eta));
            stStream(A.next.rad + j, make_float4(L.x, L.y, L.z, newPdf));
            stStream(A.next.pos + j, make_float4(pos4.x, pos4.y, __uint_as_float((uint32_t)rng.state),
                                                 __uint_as_float((uint32_t)(rng.state >> 32))));
            stStream(A.next.flags + j, (fl & ~(kDepthMask | (0xFFu << kVertShift))) | (depth & kDepthMask) | (vcount << kVertShift));
            stStream(A.next.slot + j, slot);
            stStream(A.next.medium + j, medium);
        }
        if (wantShadow) {
            stStream(A.shadow.o + sidx, make_float4(shO.x, shO.y, shO.z, kEpsilon));
            stStream(A.shadow.d + sidx, make_float4(shD.x, shD.y, shD.z, shMaxT));
            stStream(A.shadow.c + sidx, make_float4(shC.x, shC.y, shC.z, __uint_as_float(j)));
        }
        if (valid && terminate) {
            donePaths++;
            doneLen += depth;
            finishPath(A, slot, pos4, L);
        }
        if (A.G.record) emitTrainingSamples(A.G, (valid && terminate) ? vcount : 0u, ap.inclC, ap.warpTotalC, ap.baseC, slot, L);
    }
    warpAddU64(&A.C->paths, donePaths);
    warpAddU64(&A.C->pathLen, doneLen);
}

// Film accumulation for a finished batch: one thread per camera sample, in slot (= pixel) order.
//
// A sample's Gaussian footprint covers up to 5 x 5 texels, so a direct scatter issues 25 float4 reductions per sample -- 105 M
// of them per C2 step, 0.29 ms, bound by the L2's atomic throughput. k_generate lays 32 consecutive slots out as an 8 x 4 pixel
// tile, so the footprints of one warp's samples overlap heavily: 800 contributions land on 12 x 8 texels. The warp therefore
// accumulates them in a shared-memory tile first -- 25 steps; in step (dx, dy) every lane adds to the texel at that offset from
// ITS pixel, which is a different texel for every lane, so plain read-modify-writes need no atomics -- and issues three
// reductions per lane for the 96 texels afterwards: 8 x fewer atomics. The weights are those of filmSplat (same table bins,
// same 32 x 32-tile + border coordinates, same rejection rules); only the order of the float additions differs. Warps whose 32
// slots are not such a tile (pixel lists, odd image sizes, the end of a batch) and filters wider than the tile's 2-texel apron take
// the scatter path.
#ifndef PG_SPLAT_TILE
#define PG_SPLAT_TILE 1
#endif
__global__ void __launch_bounds__(256) k_splat(FilmRecord F, float4 *film, const float4 *__restrict__ splat, uint32_t n,
                                               float maxComponentValue, int useTile) {
    __shared__ float4 sTile[8][96];
    float4 *tile = sTile[threadIdx.x >> 5];
    const uint32_t lane = laneId();
    const int border = (int)ceilf(F.radius - 0.5f);
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        const bool valid = i < n;
        float4 a = make_float4(0, 0, 0, 0), b = a;
        if (valid) {
            a = splat[2 * (size_t)i];
            b = splat[2 * (size_t)i + 1];
        }
        const float2 pos = make_float2(a.x, a.y);
        float3 spec = f3(a.z, a.w, b.x);
        const int pxI = min(max((int)pos.x, 0), F.width - 1), pyI = min(max((int)pos.y, 0), F.height - 1);
        const int x0 = __shfl_sync(0xffffffffu, pxI, 0), y0 = __shfl_sync(0xffffffffu, pyI, 0);
        const bool inTile = valid && pxI == x0 + (int)(lane & 7u) && pyI == y0 + (int)(lane >> 3);
        if (!PG_SPLAT_TILE || !useTile || !(F.radius <= 2.0f) || __ballot_sync(0xffffffffu, inTile) != 0xffffffffu) {
            if (valid) filmSplat(F, film, pos, spec, maxComponentValue);
            continue;
        }
        // ---- the sample's value and per-axis weights, exactly as filmSplat forms them
        const float maxSpec = maxComp(spec);
        if (maxSpec > maxComponentValue) spec = spec * (maxComponentValue / maxSpec);
        const bool ok = isfinite(spec.x) && isfinite(spec.y) && isfinite(spec.z) && spec.x >= 0 && spec.y >= 0 && spec.z >= 0;
        const int ox = (pxI >> 5) << 5, oy = (pyI >> 5) << 5;
        const int bw = min(32, F.width - ox) + 2 * border, bh = min(32, F.height - oy) + 2 * border;
        const float px = pos.x - 0.5f - (float)(ox - border), py = pos.y - 0.5f - (float)(oy - border);
        const int minx = max((int)ceilf(px - F.radius), 0), miny = max((int)ceilf(py - F.radius), 0);
        const int maxx = min((int)floorf(px + F.radius), bw - 1), maxy = min((int)floorf(py + F.radius), bh - 1);
        float wx[5], wy[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const int x = pxI + d - 2 - ox + border, y = pyI + d - 2 - oy + border;  // block-local texel coordinates
            wx[d] = (ok && x >= minx && x <= maxx) ? F.values[min((int)fabsf((x - px) * F.scaleFactor), 31)] : 0.0f;
            wy[d] = (ok && y >= miny && y <= maxy) ? F.values[min((int)fabsf((y - py) * F.scaleFactor), 31)] : 0.0f;
        }
        tile[lane] = tile[lane + 32] = tile[lane + 64] = make_float4(0, 0, 0, 0);
        __syncwarp();
        const int tx = (int)(lane & 7u), ty = (int)(lane >> 3);
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float w = wx[dx] * wy[dy];
                float4 *t = tile + (ty + dy) * 12 + tx + dx;  // a different texel for every lane
                float4 v = *t;
                v.x += w * spec.x; v.y += w * spec.y; v.z += w * spec.z; v.w += w;
                *t = v;
                __syncwarp();
            }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int t = (int)lane + 32 * k;
            const int fx = x0 - 2 + t % 12, fy = y0 - 2 + t / 12;
            const float4 v = tile[t];
            if (fx >= 0 && fx < F.width && fy >= 0 && fy < F.height && v.w != 0.0f) atomicAdd(film + (size_t)fy * F.width + fx, v);
        }
        __syncwarp();
    }
}

// Denoiser feature buffers (src/librender/denoiser.cpp:138-144, Denoiser::add keeps per-pixel running means of the sample
// colour, albedo and normal; here: per-pixel sums + counts, divided on read). feat = 3 float4 per pixel:
// {colour sum, #colour samples} {albedo sum, #samples} {normal sum, 0}. The reference never fills albedo / normal (nothing
// calls its denoiser), so the definitions are this repo's: first intersection of the camera ray; albedo =
// diffuse reflectance (diffuse, roughplastic), specular reflectance (roughconductor), 1 (dielectric, null); normal =
// shading normal; 0 when the ray leaves the scene. k_features runs once per batch after the first trace.
__global__ void __launch_bounds__(256) k_features(DeviceScene S, PathState P, const float4 *__restrict__ hits, const uint32_t *nPtr,
                                                  float4 *feat) {
    const uint32_t n = *nPtr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 ro = P.rayO[i], rd = P.rayD[i], pos4 = P.pos[i], h4 = hits[i];
        const int px = min(max((int)pos4.x, 0), S.film.width - 1), py = min(max((int)pos4.y, 0), S.film.height - 1);
        const size_t pixel = (size_t)py * S.film.width + px;
        float3 albedo = f3(0.0f), normal = f3(0.0f);
        Hit h;
        h.t = h4.x;
        h.u = h4.y;
        h.v = h4.z;
        h.prim = __float_as_uint(h4.w);
        if (h.prim != kMiss) {
            Intersection its;
            fillIntersection(S, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), h, its);
            const BsdfRecord &b = S.bsdfs[its.bsdf];
            albedo = (b.type == B200PG_BSDF_DIFFUSE || b.type == B200PG_BSDF_ROUGHPLASTIC) ? ld3(b.reflectance)
                     : (b.type == B200PG_BSDF_ROUGHCONDUCTOR ? ld3(b.specRefl) : f3(1.0f));
            normal = its.sh.n;
        }
        atomicAdd(feat + 3 * pixel + 1, make_float4(albedo.x, albedo.y, albedo.z, 1.0f));
        atomicAdd(feat + 3 * pixel + 2, make_float4(normal.x, normal.y, normal.z, 0.0f));
    }
}
__global__ void __launch_bounds__(256) k_feature_color(FilmRecord F, const float4 *__restrict__ splat, uint32_t n, float maxComponentValue,
                                                       float4 *feat) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = splat[2 * (size_t)i], b = splat[2 * (size_t)i + 1];
        float3 spec = f3(a.z, a.w, b.x);
        const float maxSpec = maxComp(spec);
        if (maxSpec > maxComponentValue) spec = spec * (maxComponentValue / maxSpec);
        if (!(isfinite(spec.x) && isfinite(spec.y) && isfinite(spec.z)) || spec.x < 0 || spec.y < 0 || spec.z < 0) continue;
        const int px = min(max((int)a.x, 0), F.width - 1), py = min(max((int)a.y, 0), F.height - 1);
        atomicAdd(feat + 3 * ((size_t)py * F.width + px), make_float4(spec.x, spec.y, spec.z, 1.0f));
    }
}

// Flush whatever is still queued after the last bounce (only parked/dead records can remain).
__global__ void __launch_bounds__(256) k_flush(ShadeArgs A) {
    const uint32_t n = A.C->queue[A.bounce];
    unsigned long long donePaths = 0, doneLen = 0;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + laneId();
        if (i < n) {
            const float4 rad4 = A.cur.rad[i], pos4 = A.cur.pos[i];
            const uint32_t slot = A.cur.slot[i];
            donePaths++;
            doneLen += A.cur.flags[i] & kDepthMask;
            finishPath(A, slot, pos4, f3(rad4.x, rad4.y, rad4.z));
        }
    }
    warpAddU64(&A.C->paths, donePaths);
    warpAddU64(&A.C->pathLen, doneLen);
}

// ------------------------------------------------------------------------------------------
// BSDF test kernel (b200pg_k_bsdf)
// ------------------------------------------------------------------------------------------
__global__ void k_bsdf_test(DeviceScene S, int bsdfIndex, const float *wi, const float *wo, const float *u, uint32_t n,
                            float *outEval, float *outPdf, float *outWo, float *outWeight, float *outSpdf, uint32_t *outFlags) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const BsdfRecord &b = S.bsdfs[bsdfIndex];
        const float3 vi = ld3(wi + 3 * (size_t)i), vo = ld3(wo + 3 * (size_t)i);
        const float3 e = bsdfEval(b, vi, vo);
        outEval[3 * (size_t)i] = e.x; outEval[3 * (size_t)i + 1] = e.y; outEval[3 * (size_t)i + 2] = e.z;
        outPdf[i] = bsdfPdf(b, vi, vo);
        float3 so;
        float spdf, eta;
        uint32_t st;
        float3 w = bsdfSample(b, vi, make_float2(u[2 * (size_t)i], u[2 * (size_t)i + 1]), so, spdf, eta, st);
        if (isZero(w)) { so = f3(0.0f); spdf = 0; }
        outWo[3 * (size_t)i] = so.x; outWo[3 * (size_t)i + 1] = so.y; outWo[3 * (size_t)i + 2] = so.z;
        outWeight[3 * (size_t)i] = w.x; outWeight[3 * (size_t)i + 1] = w.y; outWeight[3 * (size_t)i + 2] = w.z;
        outSpdf[i] = spdf;
        outFlags[i] = st;
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers (called from integrator.cpp, compiled by the host compiler)
// ------------------------------------------------------------------------------------------
static int g_numSMs = 0;
static int numSMs() {
    if (!g_numSMs) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_numSMs, cudaDevAttrMultiProcessorCount, dev);
        if (g_numSMs <= 0) g_numSMs = 148;
    }
    return g_numSMs;
}
template <typename K>
static int persistentGrid(K kernel, int block) {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, block, 0);
    if (perSM <= 0) perSM = 1;
    return numSMs() * perSM;  // grid = multiple of the SM count, one resident wave
}

void launchGenerate(const DeviceScene &S, const BatchDesc &B, const PathState &P, Counters *C, cudaStream_t st) {
    static int grid = persistentGrid(k_generate, 256);
    k_generate<<<grid, 256, 0, st>>>(S, B, P, C);
}
void launchTrace(const DeviceScene &S, const PathState &P, float4 *hits, const uint32_t *nPtr, uint32_t *work, Counters *C,
                 bool count, const SortArgs *sort, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st) {
    static int grid = persistentGrid(k_trace<false, false>, 128);
    static int gridKey = persistentGrid(k_trace<false, true>, 128);
    static int gridSpec = persistentGrid(k_trace_spec<false, false>, 128);
    static int gridWide = persistentGrid(k_trace_spec<false, true>, 128);
    static int gridTail = persistentGrid(k_trace_tail<false, false, ClosestQueue>, 128);
    const SortArgs none = {};
    const ClosestQueue Q = {P.rayO, P.rayD, P.flags, hits};
    bool tail = T.visits > 0;
    if (sort) {
        tail = false;
        if (count)
            k_trace<true, true><<<gridKey, 128, 0, st>>>(S, P.rayO, P.rayD, P.flags, hits, nPtr, work, C, *sort, T);
        else
            k_trace<false, true><<<gridKey, 128, 0, st>>>(S, P.rayO, P.rayD, P.flags, hits, nPtr, work, C, *sort, T);
        k_bin_scan<<<1, 1024, 0, st>>>(*sort);
        k_bin_scatter<<<numSMs() * 4, 256, 0, st>>>(*sort, nPtr);
    } else if (speculative) {
        if (S.wideNodes) {
            tail = false;
            if (count) k_trace_spec<true, true><<<gridWide, 128, 0, st>>>(S, Q, nPtr, work, C, T);
            else k_trace_spec<false, true><<<gridWide, 128, 0, st>>>(S, Q, nPtr, work, C, T);
        } else if (count)
            k_trace_spec<true, false><<<gridSpec, 128, 0, st>>>(S, Q, nPtr, work, C, T);
        else
            k_trace_spec<false, false><<<gridSpec, 128, 0, st>>>(S, Q, nPtr, work, C, T);
    } else if (count)
        k_trace<true, false><<<grid, 128, 0, st>>>(S, P.rayO, P.rayD, P.flags, hits, nPtr, work, C, none, T);
    else
        k_trace<false, false><<<grid, 128, 0, st>>>(S, P.rayO, P.rayD, P.flags, hits, nPtr, work, C, none, T);
    if (tail) {  // the long rays the kernel above handed over
        if (count) k_trace_tail<false, true, ClosestQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
        else k_trace_tail<false, false, ClosestQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
    }
}
void launchHitPartition(const float4 *hits, const uint32_t *flags, const uint32_t *nPtr, uint32_t *perm, uint32_t *cntNeed,
                        uint32_t *cntRest, cudaStream_t st) {
    k_hit_partition<<<numSMs() * 4, 256, 0, st>>>(hits, flags, nPtr, perm, cntNeed, cntRest);
}
void launchShadow(const DeviceScene &S, const ShadowQueue &Q, float4 *rad, const uint32_t *nPtr, uint32_t *work, Counters *C,
                  bool count, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st) {
    static int grid = persistentGrid(k_shadow<false>, 128);
    static int gridSpec = persistentGrid(k_shadow_spec<false, false>, 128);
    static int gridWide = persistentGrid(k_shadow_spec<false, true>, 128);
    static int gridTail = persistentGrid(k_trace_tail<true, false, ShadowRayQueue>, 128);
    if (speculative) {
        const ShadowRayQueue R = {Q, rad};
        if (S.wideNodes) {
            if (count) k_shadow_spec<true, true><<<gridWide, 128, 0, st>>>(S, R, nPtr, work, C, T);
            else k_shadow_spec<false, true><<<gridWide, 128, 0, st>>>(S, R, nPtr, work, C, T);
        } else {
            if (count) k_shadow_spec<true, false><<<gridSpec, 128, 0, st>>>(S, R, nPtr, work, C, T);
            else k_shadow_spec<false, false><<<gridSpec, 128, 0, st>>>(S, R, nPtr, work, C, T);
            if (T.visits > 0) {
                if (count) k_trace_tail<true, true, ShadowRayQueue><<<gridTail, 128, 0, st>>>(S, R, T, tailWork, C);
                else k_trace_tail<true, false, ShadowRayQueue><<<gridTail, 128, 0, st>>>(S, R, T, tailWork, C);
            }
        }
    } else if (count)
        k_shadow<true><<<grid, 128, 0, st>>>(S, Q, rad, nPtr, work, C);
    else
        k_shadow<false><<<grid, 128, 0, st>>>(S, Q, rad, nPtr, work, C);
}
void launchShade(const ShadeArgs &A, cudaStream_t st) {
    static int grid = persistentGrid(k_shade, kShadeThreads);
    k_shade<<<grid, kShadeThreads, 0, st>>>(A);
}
void launchSplat(const FilmRecord &F, float4 *film, const float4 *splat, uint32_t n, float maxComponentValue, bool tile, cudaStream_t st) {
    static int grid = persistentGrid(k_splat, 256);
    k_splat<<<grid, 256, 0, st>>>(F, film, splat, n, maxComponentValue, tile ? 1 : 0);
}
void launchFilmExport(const float4 *film, float *out, uint32_t n, int develop, cudaStream_t st) {
    k_film_export<<<numSMs() * 4, 256, 0, st>>>(film, out, n, develop);
}
void launchFeatures(const DeviceScene &S, const PathState &P, const float4 *hits, const uint32_t *nPtr, float4 *feat, cudaStream_t st) {
    k_features<<<numSMs() * 4, 256, 0, st>>>(S, P, hits, nPtr, feat);
}
void launchFeatureColor(const FilmRecord &F, const float4 *splat, uint32_t n, float maxComponentValue, float4 *feat, cudaStream_t st) {
    k_feature_color<<<numSMs() * 4, 256, 0, st>>>(F, splat, n, maxComponentValue, feat);
}
void launchFilmExportMerged(const float4 *film, const float4 *const *peers, int nPeers, float *out, uint32_t n, cudaStream_t st) {
    FilmPeers P;
    P.n = nPeers;
    for (int r = 0; r < 16; ++r) P.p[r] = r < nPeers ? peers[r] : nullptr;
    k_film_export_merged<<<numSMs() * 4, 256, 0, st>>>(film, P, out, n);
}
void launchFilmAdd(float4 *film, const float4 *peer, uint32_t n, cudaStream_t st) {
    k_film_add<<<numSMs() * 4, 256, 0, st>>>(film, peer, n);
}
void launchFlush(const ShadeArgs &A, cudaStream_t st) {
    static int grid = persistentGrid(k_flush, 256);
    k_flush<<<grid, 256, 0, st>>>(A);
}
void launchTraceRays(const DeviceScene &S, const float4 *rays, uint32_t n, float4 *hits, uint32_t *work, Counters *C, bool shadow,
                     bool count, bool speculative, TailList T, uint32_t *tailWork, cudaStream_t st) {
    static int grid = persistentGrid(k_trace_rays<false, false>, 128);
    static int gridSpec = persistentGrid(k_trace_rays_spec<false, false, false>, 128);
    static int gridWide = persistentGrid(k_trace_rays_spec<false, false, true>, 128);
    static int gridTail = persistentGrid(k_trace_tail<false, false, RayListQueue>, 128);
    if (speculative) {
        const RayListQueue Q = {rays, hits, S.primGlobalId, shadow};
        if (S.wideNodes) {
            if (shadow) {
                if (count) k_trace_rays_spec<true, true, true><<<gridWide, 128, 0, st>>>(S, Q, n, work, C, T);
                else k_trace_rays_spec<true, false, true><<<gridWide, 128, 0, st>>>(S, Q, n, work, C, T);
            } else {
                if (count) k_trace_rays_spec<false, true, true><<<gridWide, 128, 0, st>>>(S, Q, n, work, C, T);
                else k_trace_rays_spec<false, false, true><<<gridWide, 128, 0, st>>>(S, Q, n, work, C, T);
            }
        } else {
            if (shadow) {
                if (count) k_trace_rays_spec<true, true, false><<<gridSpec, 128, 0, st>>>(S, Q, n, work, C, T);
                else k_trace_rays_spec<true, false, false><<<gridSpec, 128, 0, st>>>(S, Q, n, work, C, T);
            } else {
                if (count) k_trace_rays_spec<false, true, false><<<gridSpec, 128, 0, st>>>(S, Q, n, work, C, T);
                else k_trace_rays_spec<false, false, false><<<gridSpec, 128, 0, st>>>(S, Q, n, work, C, T);
            }
            if (T.visits > 0) {
                if (shadow) {
                    if (count) k_trace_tail<true, true, RayListQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
                    else k_trace_tail<true, false, RayListQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
                } else {
                    if (count) k_trace_tail<false, true, RayListQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
                    else k_trace_tail<false, false, RayListQueue><<<gridTail, 128, 0, st>>>(S, Q, T, tailWork, C);
                }
            }
        }
    } else if (shadow) {
        if (count) k_trace_rays<true, true><<<grid, 128, 0, st>>>(S, rays, n, hits, work, C);
        else k_trace_rays<true, false><<<grid, 128, 0, st>>>(S, rays, n, hits, work, C);
    } else {
        if (count) k_trace_rays<false, true><<<grid, 128, 0, st>>>(S, rays, n, hits, work, C);
        else k_trace_rays<false, false><<<grid, 128, 0, st>>>(S, rays, n, hits, work, C);
    }
}
void launchFilmSplat(const FilmRecord &F, float4 *film, const float2 *pos, const float3 *rgb, uint32_t n, float maxComponentValue,
                     cudaStream_t st) {
    int grid = numSMs() * 4;
    k_film_splat<<<grid, 256, 0, st>>>(F, film, pos, rgb, n, maxComponentValue);
}
void launchBsdfTest(const DeviceScene &S, int bsdfIndex, const float *wi, const float *wo, const float *u, uint32_t n, float *outEval,
                    float *outPdf, float *outWo, float *outWeight, float *outSpdf, uint32_t *outFlags, cudaStream_t st) {
    k_bsdf_test<<<numSMs(), 128, 0, st>>>(S, bsdfIndex, wi, wo, u, n, outEval, outPdf, outWo, outWeight, outSpdf, outFlags);
}

}  // namespace pg

// volpath.cu -- shade and connection stages of the volumetric path tracer as sm_100a CUDA kernels:
//   k_shade_vol   one loop iteration of ProgressiveVolumetricPathTracer::Li
//                 (src/integrators/path/progressive_volpath.cpp:98-374) per queued path:
//                 [emitter look-up along the ray sampled one iteration earlier (:401-460) + Russian roulette],
//                 free-flight distance sampling in the current medium (heterogeneous.cpp:589-663),
//                 then either a medium scattering event (phase function, hg.cpp / isotropic.cpp) or a
//                 surface event (BSDF; index-matched "null" boundaries pass straight through, :318-328)
//   k_shadow_vol  attenuated emitter connections: Scene::evalTransmittance (src/librender/scene.cpp:662-722)
//                 through any number of index-matched boundaries, Woodcock-tracking transmittance
//                 (heterogeneous.cpp:546-587)
//   k_look_vol    the emitter look-up along the ray sampled one iteration earlier (:401-460), as its own stage
//   k_track_vol   Russian roulette + free-flight distance sampling of every queued path, as its own stage
// The Woodcock loops run a geometrically distributed number of steps per path, and only for paths inside a medium:
// inside the shade kernel they ran with 4.7 of 32 lanes (profiles/r01_v3_vol_summary.txt). The three tracking stages
// (k_look_vol, k_track_vol, k_shadow_vol) are therefore persistent warps with PER-LANE refill: every lane owns one
// job, advances it by one tracking step per iteration, and fetches the next job from the device cursor as soon as its
// own is finished, so the step body always runs with (nearly) all lanes. k_shade_vol only handles the events.
// The closest-hit query of every newly sampled ray is k_trace's (kernels.cu); follow-up segments behind
// index-matched boundaries are traced inline, by the lane that owns the job.
//
// RNG: transmittance estimates of one connection draw from a stream forked off the path's stream
// (Rng::fork) so that the main stream advances by a fixed amount per connection (DESIGN.md "RNG").
#include "medium_device.cuh"
#include "wavefront.cuh"
#include "wavefront_device.cuh"

namespace pg {

struct BoundaryHit {  // what the connection loops need to know about a surface hit
    bool valid, isNull, transition;
    int emitter, interior, exterior;
    float3 geoN, shN;
    float t;
};

PG_DEV BoundaryHit boundaryOf(const DeviceScene &S, const Hit &h, const Intersection &its) {
    BoundaryHit b;
    b.valid = h.prim != kMiss;
    b.t = b.valid ? h.t : kInf;
    b.isNull = false;
    b.transition = false;
    b.emitter = -1;
    b.interior = b.exterior = -1;
    b.geoN = b.shN = f3(0.0f);
    if (b.valid) {
        const ShapeRecord sr = S.shapes[its.shape];
        b.isNull = (S.bsdfs[its.bsdf].typeFlags & kNull) != 0;
        b.emitter = its.emitter;
        b.interior = sr.interiorMedium;
        b.exterior = sr.exteriorMedium;
        b.transition = sr.interiorMedium >= 0 || sr.exteriorMedium >= 0;
        b.geoN = its.geoN;
        b.shN = its.sh.n;
    }
    return b;
}

// ---- persistent warps with per-lane refill -----------------------------------------------------------------------
// Job interface: bool fetch(uint32_t i)  -> set up job i; false when it completed right away (output already written)
//                bool tracking           -> the job's next unit of work is a tracking step
//                void trackStep()        -> one tentative collision (cheap, uniform); clears `tracking` when the walk is over
//                bool logicStep()        -> everything between two walks (boundary logic, inline traversal: expensive,
//                                           divergent); true when the job is finished
//                void finish()           -> write the job's output
// Schedule: tracking steps run in bursts while at least kMinTrack lanes are walking; lanes whose walk is over wait for
// the burst to end, then all of them run their logic step together, finished lanes are refilled, and the next burst
// starts. The expensive divergent code is thus executed once per burst instead of once per tracking step.
static constexpr int kMinTrack = 12;
template <typename Job>
PG_DEV void runWithRefill(Job &job, uint32_t n, uint32_t *work) {
    const unsigned ltMask = (1u << laneId()) - 1u;
    bool active = false, exhausted = false;
    while (true) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && !exhausted) {
            const int leader = __ffs(idle) - 1, cnt = __popc(idle);
            uint32_t base = 0;
            if ((int)laneId() == leader) base = atomicAdd(work, (uint32_t)cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!active) {
                const uint32_t i = base + __popc(idle & ltMask);
                if (i < n) active = job.fetch(i);
            }
            exhausted = base + (uint32_t)cnt >= n;  // warp-uniform
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (exhausted) break;
            continue;
        }
        while (true) {  // tracking burst
            const bool walking = active && job.tracking;
            const int nWalk = __popc(__ballot_sync(0xffffffffu, walking));
            if (nWalk == 0) break;
            if (nWalk < kMinTrack) {
                const int nWait = __popc(__ballot_sync(0xffffffffu, active && !job.tracking));
                const int nIdle = __popc(__ballot_sync(0xffffffffu, !active));
                if (nWait > 0 || (nIdle > 0 && !exhausted)) break;
            }
            if (walking) job.trackStep();
        }
        if (active && !job.tracking && job.logicStep()) {
            job.finish();
            active = false;
        }
    }
}

// The two ratio-free tracking trials of HeterogeneousMedium::evalTransmittance (heterogeneous.cpp:546-587), one
// tentative collision per step() so that it can be interleaved with other lanes' work.
struct TrialTracker {
    float t, mint, maxt, result;
    int trial;
    // false: no walk needed -- the segment misses the density box (transmittance 1, no random numbers consumed), or the
    // medium integrates deterministically (method = simpson: the whole quadrature runs here, value() holds the result)
    PG_DEV bool begin(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt) {
        result = 2.0f;  // value() == 1
        if (M.method == B200PG_MEDIUM_SIMPSON) {
            result = 2.0f * expf(-mediumIntegrateDensity(M, density, o, d, rmint, rmaxt));
            return false;
        }
        float a, b;
        if (!mediumClip(M, o, d, a, b)) return false;
        mint = fmaxf(a, rmint);
        maxt = fminf(b, rmaxt);
        t = mint;
        result = 0.0f;
        trial = 0;
        return true;
    }
    // true when both trials are over
    PG_DEV bool step(const MediumRecord &M, const float *density, float3 o, float3 d, Rng &rng) {
        bool over;
        t -= logf(1 - rng.next1D()) * M.invMaxDensity;
        if (t >= maxt) {
            result += 1;
            over = true;
        } else {
            const float dens = gridLookup(M, density, o + d * t) * M.scale;
            over = dens * M.invMaxDensity > rng.next1D();
        }
        if (over) {
            t = mint;
            return ++trial == 2;
        }
        return false;
    }
    PG_DEV float value() const { return result * 0.5f; }
};

PG_DEV Rng loadRng(float4 pos4, uint32_t width, uint32_t &pixel) {
    Rng rng;
    pixel = (uint32_t)pos4.y * width + (uint32_t)pos4.x;
    rng.state = ((uint64_t)__float_as_uint(pos4.w) << 32) | (uint64_t)__float_as_uint(pos4.z);
    rng.inc = ((uint64_t)pixel << 1) | 1ULL;
    return rng;
}

// ---- k_look_vol: rayIntersectAndLookForEmitter (progressive_volpath.cpp:401-460) + the MIS weight of its result
struct LookJob {
    const ShadeArgs &A;
    unsigned long long rays = 0;
    uint32_t idx;
    float3 o, d, thr;
    float prevPdf, transmittance;
    uint32_t fl;
    int medium, interactions, maxInteractions;
    BoundaryHit cur;
    Rng fr;
    TrialTracker trk;
    bool tracking, zero;
    PG_DEV LookJob(const ShadeArgs &a) : A(a) {}

    PG_DEV void beginSegment() {
        tracking = false;
        if (medium >= 0) {
            tracking = trk.begin(A.S.media[medium], A.S.density, o, d, 0.0f, cur.t);
            if (!tracking) transmittance *= trk.value();
        }
    }
    PG_DEV bool fetch(uint32_t i) {
        idx = i;
        fl = A.cur.flags[i];
        if ((fl & kFlagDead) || !(fl & kFlagLook)) {
            A.lookL[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            return false;
        }
        const float4 ro = A.cur.rayO[i], rd = A.cur.rayD[i], thr4 = A.cur.thr[i], h4 = A.hits[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        thr = f3(thr4.x, thr4.y, thr4.z);
        prevPdf = A.cur.rad[i].w;
        medium = A.cur.medium[i];
        uint32_t pixel;
        Rng rng = loadRng(A.cur.pos[i], (uint32_t)A.S.film.width, pixel);
        fr = rng.fork();
        Hit h;
        h.t = h4.x;
        h.u = h4.y;
        h.v = h4.z;
        h.prim = __float_as_uint(h4.w);
        Intersection its;
        if (h.prim != kMiss) fillIntersection(A.S, o, d, h, its);
        cur = boundaryOf(A.S, h, its);
        // the look-up ran with maxDepth - depth - 1, depth = the value before the loop's depth++
        maxInteractions = A.cfg.maxDepth - (int)(fl & kDepthMask) - 1;
        interactions = 0;
        transmittance = 1.0f;
        zero = false;
        beginSegment();
        return true;
    }
    PG_DEV void trackStep() {
        if (trk.step(A.S.media[medium], A.S.density, o, d, fr)) {
            transmittance *= trk.value();
            tracking = false;
        }
    }
    PG_DEV bool logicStep() {
        const DeviceScene &S = A.S;
        // the segment's transmittance is in: boundary logic of the reference loop
        if (cur.valid && (interactions == maxInteractions || !cur.isNull || cur.emitter >= 0)) return true;
        if (!cur.valid) return true;
        if (transmittance == 0) {
            zero = true;
            return true;
        }
        if (cur.transition) medium = dot(d, cur.geoN) > 0 ? cur.exterior : cur.interior;
        o = o + d * cur.t;
        if (++interactions > 100) {
            zero = true;
            return true;
        }
        Hit h;
        const float mint = adaptiveMinT(o, kEpsilon, false);
        uint32_t cn = 0, cp = 0;
        traceRay<false, false>(S, o, d, mint, kInf, h, &cn, &cp);
        rays++;
        Intersection its;
        if (h.prim != kMiss) fillIntersection(S, o, d, h, its);
        cur = boundaryOf(S, h, its);
        beginSegment();
        return false;
    }
    PG_DEV void finish() {
        const DeviceScene &S = A.S;
        float3 contrib = f3(0.0f);
        if (!zero && cur.valid && cur.emitter >= 0) {
            const float3 Le = dot(cur.shN, -d) <= 0 ? f3(0.0f) : ld3(S.emitters[cur.emitter].radiance);  // area.cpp:104-109
            const float3 value = Le * transmittance;
            const bool prevMedium = (fl & kFlagPrevMedium) != 0;
            if (!isZero(value) && (!prevMedium || fminf(value.x, fminf(value.y, value.z)) > 0.0f)) {  // (:288-303, :330-345)
                // cur.t is the length of the LAST segment, as in the reference (dRec.setQuery(ray, *its) after the ray origin was
                // moved to the last index-matched boundary, progressive_volpath.cpp:437-452 / records.inl:170-178) -- not the
                // distance from the path vertex. Kept for parity; DESIGN.md section 2 ("volumetric Li") has the consequence.
                const float emitterPdf = (A.cfg.useNee && !(fl & kFlagPrevDelta)) ? pdfEmitterDirect(S, cur.emitter, d, cur.shN, cur.t) : 0.0f;
                const float weight = A.cfg.useNee ? miWeight(prevPdf, emitterPdf) : 1.0f;
                contrib = thr * value * weight;
            }
        }
        A.lookL[idx] = make_float4(contrib.x, contrib.y, contrib.z, 0.0f);
    }
};

__global__ void __launch_bounds__(128) k_look_vol(ShadeArgs A) {
    LookJob job(A);
    runWithRefill(job, A.C->queue[A.bounce], &A.C->lookWork[A.bounce]);
    warpAddU64(&A.C->normalRays, job.rays);
}

// ---- k_track_vol: the part of the loop between the emitter look-up and the event: Russian roulette of the previous
// iteration (:350-360), loop condition, free-flight distance (:118-121). Output per queued path:
//   trkA = {distance of the medium event (inf: none), rng state lo, hi (bits), flags: 1 = path ends here}
//   trkB = {throughput after roulette and (guided) tracking weights, density at the event}
struct TrackJob {
    const ShadeArgs &A;
    uint32_t idx;
    float3 o, d, thr, albedo;
    float t, maxt, dens, albedoAvg;
    int medium;
    bool guidedDist, event, tracking;
    Rng rng;
    PG_DEV TrackJob(const ShadeArgs &a) : A(a) {}

    PG_DEV void write(float tEvent, uint32_t bits) {
        A.trkA[idx] = make_float4(tEvent, __uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32)),
                                  __uint_as_float(bits));
        A.trkB[idx] = make_float4(thr.x, thr.y, thr.z, dens);
    }
    PG_DEV bool fetch(uint32_t i) {
        idx = i;
        const uint32_t fl = A.cur.flags[i];
        const float4 thr4 = A.cur.thr[i];
        thr = f3(thr4.x, thr4.y, thr4.z);
        dens = 0.0f;
        uint32_t pixel;
        rng = loadRng(A.cur.pos[i], (uint32_t)A.S.film.width, pixel);
        if (fl & kFlagDead) {
            write(kInf, 1u);
            return false;
        }
        uint32_t depth = fl & kDepthMask;
        bool terminate = false;
        if (fl & kFlagLook) {
            rng.nextU32();  // the look-up's stream was forked off here (Rng::fork consumes two words)
            rng.nextU32();
            if (depth++ >= (uint32_t)A.cfg.rrDepth) {
                const float q = fminf(maxComp(thr) * thr4.w * thr4.w, 0.95f);
                if (rng.next1D() >= q)
                    terminate = true;
                else
                    thr = thr / q;
            }
        }
        if (!terminate && !((int)depth <= A.cfg.maxDepth || A.cfg.maxDepth < 0)) terminate = true;
        if (terminate) {
            write(kInf, 1u);
            return false;
        }
        medium = A.cur.medium[i];
        if (medium < 0) {
            write(kInf, 0u);
            return false;
        }
        const float4 ro = A.cur.rayO[i], rd = A.cur.rayD[i], h4 = A.hits[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        const MediumRecord &M = A.S.media[medium];
        if (M.method == B200PG_MEDIUM_SIMPSON) {  // deterministic quadrature: one unit of work, done right here
            MediumSample mRec;
            const bool ev = mediumSampleDistanceSimpson(M, A.S.density, o, d, 0.0f, __float_as_uint(h4.w) != kMiss ? h4.x : kInf, mRec, rng);
            // the throughput factor is applied here (k_shade_vol skips it for this method, bit 2):
            // event: sigma_s * transmittance / pdfSuccess, no event: transmittance / pdfFailure  (:125-126, :231-232)
            if (ev)
                thr = thr * ((mRec.sigmaS * mRec.transmittance) / mRec.pdfSuccess);
            else
                thr = thr * (mRec.transmittance / mRec.pdfFailure);
            write(ev ? mRec.t : kInf, 2u);
            return false;
        }
        float a, b;
        if (!mediumClip(M, o, d, a, b)) {
            write(kInf, 0u);
            return false;
        }
        t = fmaxf(a, 0.0f);
        maxt = fminf(b, __float_as_uint(h4.w) != kMiss ? h4.x : kInf);
        guidedDist = A.G.enabled && A.cfg.guidedDistance;  // (Woodcock only)
        albedo = ld3(M.albedo);
        albedoAvg = (albedo.x + albedo.y + albedo.z) * (1.0f / 3.0f);
        event = false;
        tracking = true;
        return true;
    }
    PG_DEV void trackStep() { tracking = !walk(); }
    PG_DEV bool logicStep() { return true; }
    // one tentative collision: heterogeneous.cpp:589-663 (Woodcock), or its guided variant (medium_device.cuh)
    PG_DEV bool walk() {
        const MediumRecord &M = A.S.media[medium];
        t -= logf(1 - rng.next1D()) * M.invMaxDensity;
        if (t >= maxt) return true;
        const float3 p = o + d * t;
        if (!guidedDist) {
            dens = gridLookup(M, A.S.density, p) * M.scale;
            if (dens * M.invMaxDensity > rng.next1D()) {
                event = true;
                return true;
            }
            return false;
        }
        const float a = fminf(gridLookup(M, A.S.density, p) * M.scale * M.invMaxDensity, 1.0f);
        const float u = rng.next1D();
        if (!(a > 0)) return false;
        const float g = 4 * kPi * guidePdf(A.G, guideLookup(A.G, p), d);
        const float num = a * albedoAvg, den = num + (1 - a) * g;
        const float pGuided = den > 0 ? num / den : a;
        const float pReal = 0.5f * a + 0.5f * pGuided;
        if (u < pReal) {
            thr = thr * (albedo * (a / pReal));
            event = true;
            return true;
        }
        thr = thr * ((1 - a) / (1 - pReal));
        return false;
    }
    PG_DEV void finish() { write(event ? t : kInf, 0u); }
};

__global__ void __launch_bounds__(128) k_track_vol(ShadeArgs A) {
    TrackJob job(A);
    runWithRefill(job, A.C->queue[A.bounce], &A.C->trackWork[A.bounce]);
}

// Event partition of the queue in front of k_shade_vol: a path arrives either with a medium scattering event (phase function,
// guided direction at a medium vertex), with a surface event (BSDF code) or with nothing left to shade (parked, ended by Russian
// roulette, left the scene). In queue order every warp holds all three and walks both long branches with half of its lanes.
// perm = [medium events ... | ... surface events], rest = the third class; k_shade_vol maps its queue position through them so that
// (all but two) warps hold one class. The classification mirrors the branch conditions of k_shade_vol.
__global__ void __launch_bounds__(256) k_event_partition(ShadeArgs A, uint32_t *__restrict__ perm, uint32_t *__restrict__ rest) {
    __shared__ uint32_t sWarp[3][8];
    __shared__ uint32_t sBase[3];
    const uint32_t n = A.C->queue[A.bounce];
    uint32_t *cnt[3] = {&A.C->partNeed[A.bounce], &A.C->partSurf[A.bounce], &A.C->partRest[A.bounce]};
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        const uint32_t i = base + threadIdx.x;
        int cls = -1;
        if (i < n) {
            const uint32_t fl = A.cur.flags[i];
            const float4 tA = A.trkA[i];
            if ((fl & kFlagDead) || (__float_as_uint(tA.w) & 1u)) cls = 2;
            else if (tA.x < kInf) cls = 0;
            else cls = __float_as_uint(A.hits[i].w) != kMiss ? 1 : 2;
        }
        unsigned m[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            m[c] = __ballot_sync(0xffffffffu, cls == c);
            if (laneId() == 0) sWarp[c][warp] = __popc(m[c]);
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            uint32_t total = 0;
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = sWarp[threadIdx.x][w];
                sWarp[threadIdx.x][w] = total;
                total += c;
            }
            sBase[threadIdx.x] = total ? atomicAdd(cnt[threadIdx.x], total) : 0u;
        }
        __syncthreads();
        if (cls >= 0) {
            const uint32_t r = sBase[cls] + sWarp[cls][warp] + __popc(m[cls] & ((1u << laneId()) - 1u));
            if (cls == 0) perm[r] = i;
            else if (cls == 1) perm[n - 1u - r] = i;
            else rest[r] = i;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kShadeThreads, 6) k_shade_vol(ShadeArgs A) {
    const DeviceScene &S = A.S;
    const IntegratorConfig &cfg = A.cfg;
    const uint32_t n = A.C->queue[A.bounce];
    const uint32_t nMed = A.perm ? A.C->partNeed[A.bounce] : 0u, nSurf = A.perm ? A.C->partSurf[A.bounce] : 0u;
    uint32_t *nextCount = &A.C->queue[A.bounce + 1];
    uint32_t *shadowCount = &A.C->shadow[A.bounce];
    unsigned long long donePaths = 0, doneLen = 0;
    __shared__ uint32_t sAppend[2 * 3 * (kShadeThreads / 32 + 1)];
    uint32_t appendParity = 0;

    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t q = base + threadIdx.x;
        const bool valid = q < n;
        uint32_t i = q;
        if (valid && A.perm)  // event partition: medium events, then surface events, then the paths with nothing to shade
            i = q < nMed ? A.perm[q] : (q < nMed + nSurf ? A.perm[n - 1u - (q - nMed)] : A.permRest[q - nMed - nSurf]);

        bool alive = false, wantShadow = false, terminate = false;
        float4 pos4 = make_float4(0, 0, 0, 0);
        uint32_t fl = 0, slot = 0, pixel = 0;
        int medium = -1;
        float3 L = f3(0.0f), thr = f3(1.0f);
        float eta = 1.0f;
        Rng rng;
        rng.state = 0;
        rng.inc = 1;
        float3 newO = f3(0.0f), newD = f3(0.0f);
        float newPdf = 0.0f, newMint = kEpsilon;
        float3 shO = f3(0.0f), shD = f3(0.0f), shC = f3(0.0f);
        float shMaxT = 0.0f;
        int shMedium = -1, shInteractions = 0;
        uint32_t shOnSurface = 0;
        uint64_t shRng = 0;
        uint32_t depth = 0, vcount = 0;
        float neeDirPdf = 0.0f, neeLightPdf = 0.0f;

        if (valid) {
            // streamed-once state: evict-first, as in k_shade
            const float4 ro = ldStream(A.cur.rayO + i), rd = ldStream(A.cur.rayD + i), thr4 = ldStream(A.cur.thr + i),
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
            pixel = (uint32_t)pos4.y * (uint32_t)S.film.width + (uint32_t)pos4.x;
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
                // close the previous training vertex: radiance gathered so far (its NEE has landed by now) and the
                // distance to the first surface along the sampled ray
                guideVertexClose(A.G, slot, vcount - 1, thr, h.prim == kMiss ? 0.0f : h.t, L);
                fl |= kFlagVertexClosed;
            }

            if (fl & kFlagDead) {
                terminate = true;
            } else {
                Intersection its;
                const bool hitValid = h.prim != kMiss;
                if (hitValid) fillIntersection(S, o, d, h, its);
                const int maxDepth = cfg.maxDepth;

                // ---- results of the two tracking stages for this path: attenuated emitter radiance found along the ray
                // (k_look_vol, already MIS-weighted), Russian roulette / loop condition / free-flight distance (k_track_vol)
                const float4 tA = A.trkA[i], tB = A.trkB[i];
                if (fl & kFlagLook) {
                    const float4 lk = A.lookL[i];
                    L += f3(lk.x, lk.y, lk.z);
                    fl &= ~(kFlagFirst | kFlagLook);  // rRec.type &= ~EEmittedRadiance
                    depth++;
                    fl |= kFlagScattered;
                }
                if (__float_as_uint(tA.w) & 1u) terminate = true;  // Russian roulette or the loop condition ended the path
                thr = f3(tB.x, tB.y, tB.z);
                rng.state = ((uint64_t)__float_as_uint(tA.z) << 32) | (uint64_t)__float_as_uint(tA.y);

                if (!terminate) {
                    MediumSample mRec;
                    const bool mediumEvent = tA.x < kInf;
                    const bool thrDone = (__float_as_uint(tA.w) & 2u) != 0;  // simpson: k_track_vol applied the medium factor
                    const bool guidedDist = (A.G.enabled && cfg.guidedDistance && medium >= 0) || thrDone;
                    if (mediumEvent) {
                        mRec.t = tA.x;
                        mRec.p = o + d * tA.x;
                        const float densityAtT = tB.w;
                        mRec.sigmaS = ld3(S.media[medium].albedo) * densityAtT;
                        float tr = densityAtT != 0.0f ? 1.0f / densityAtT : 0.0f;
                        if (!isfinite(tr)) tr = 0.0f;
                        mRec.transmittance = tr;
                    }
                    if (mediumEvent) {
                        const MediumRecord &M = S.media[medium];
                        if ((int)depth >= maxDepth && maxDepth != -1) {
                            terminate = true;  // (:128-129)
                        } else {
                            if (!guidedDist) thr *= mRec.sigmaS * mRec.transmittance;  // pdfSuccess == 1 for Woodcock tracking
                            const bool guided = A.G.enabled != 0;
                            const uint32_t gcell = guided ? guideLookup(A.G, mRec.p) : 0u;
                            if (cfg.useNee) {  // (:137-168)
                                DirectSample dRec;
                                const float2 u = rng.next2D();
                                const float3 value = sampleEmitterDirect(S, mRec.p, f3(0.0f), u, dRec);
                                Rng fr = rng.fork();
                                if (!isZero(value)) {
                                    const float phaseVal = phaseEval(M, -d, dRec.d);
                                    if (phaseVal != 0) {
                                        // the MIS weight needs the direction pdf (phase pdf == phase value; at a guided vertex
                                        // mixed with the mixture pdf, evaluated below together with the sampled direction)
                                        neeDirPdf = phaseVal;
                                        neeLightPdf = dRec.pdf;
                                        shC = thr * value * phaseVal;
                                        shO = mRec.p;
                                        shD = dRec.d;
                                        shMaxT = dRec.dist;
                                        shMedium = medium;
                                        shInteractions = maxDepth - (int)depth - 1;
                                        shOnSurface = 0;
                                        shRng = fr.state;
                                        wantShadow = true;
                                    }
                                }
                            }
                            float phasePdf;
                            float3 wo;
                            bool dirOk = true;
                            if (guided) {  // one-sample MIS between the mixture and the phase function
                                const float u0 = rng.next1D();
                                const float2 u12 = rng.next2D();
                                float pp;
                                if (u0 < A.G.alpha) {
                                    wo = guideSample(A.G, gcell, u0 / A.G.alpha, u12.x, u12.y);
                                    pp = phaseEval(M, -d, wo);
                                } else {
                                    wo = phaseSample(M, -d, u12, pp);
                                }
                                float gNee = 0.0f, gWo = 0.0f;  // one pass over the cell's lobes for both directions
                                guidePdf2(A.G, gcell, wantShadow ? shD : wo, wo, gNee, gWo);
                                if (wantShadow) neeDirPdf = A.G.alpha * gNee + (1 - A.G.alpha) * neeDirPdf;
                                phasePdf = A.G.alpha * gWo + (1 - A.G.alpha) * pp;
                                dirOk = pp > 0 && phasePdf > 0;
                                if (dirOk) thr = thr * (pp / phasePdf);
                            } else {
                                wo = phaseSample(M, -d, rng.next2D(), phasePdf);  // phase weight == 1
                            }
                            if (wantShadow) shC = shC * miWeight(neeLightPdf, neeDirPdf);
                            if (!dirOk) {
                                terminate = true;
                            } else {
                            if (A.G.record && (int)vcount < A.G.maxVerts) {
                                guideVertexOpen(A.G, slot, vcount, mRec.p, phasePdf, wo, gcell);
                                vcount++;
                                fl &= ~kFlagVertexClosed;
                            }
                            newO = mRec.p;
                            newD = wo;
                            newMint = 0.0f;
                            newPdf = phasePdf;
                            fl &= ~kFlagPrevDelta;
                            fl |= kFlagPrevMedium | kFlagLook;
                            alive = true;
                            }
                        }
                    } else if (!hitValid) {
                        terminate = true;  // no environment emitter on this path (:183-195)
                    } else {
                        const BsdfRecord &bsdf = S.bsdfs[its.bsdf];
                        const ShapeRecord sr = S.shapes[its.shape];
                        const bool transition = sr.interiorMedium >= 0 || sr.exteriorMedium >= 0;
                        // ---- emitted radiance of a directly visible / null-boundary-visible emitter (:199-201)
                        if (its.emitter >= 0 && (fl & kFlagFirst) && (!cfg.hideEmitters || (fl & kFlagScattered))) {
                            const float3 Le = dot(its.sh.n, -d) <= 0 ? f3(0.0f) : ld3(S.emitters[its.emitter].radiance);
                            L += thr * Le;
                        }
                        if (((int)depth >= maxDepth && maxDepth != -1) || (cfg.strictNormals && dot(d, its.geoN) * its.wi.z >= 0)) {
                            terminate = true;  // (:207-216)
                        } else {
                            const uint32_t btype = bsdf.typeFlags;
                            const bool guided = A.G.enabled && (btype & kSmooth);
                            const uint32_t gcell = guided ? guideLookup(A.G, its.p) : 0u;
                            if (cfg.useNee && (btype & kSmooth)) {  // (:226-262)
                                const float3 refN = (btype & (kTransmission | kBackSide)) == 0 ? its.sh.n : f3(0.0f);
                                DirectSample dRec;
                                const float2 u = rng.next2D();
                                const float3 value = sampleEmitterDirect(S, its.p, refN, u, dRec);
                                Rng fr = rng.fork();
                                if (!isZero(value)) {
                                    const float3 woL = its.sh.toLocal(dRec.d);
                                    const float3 bsdfVal = bsdfEval(bsdf, its.wi, woL);
                                    if (!isZero(bsdfVal) && (!cfg.strictNormals || dot(its.geoN, dRec.d) * woL.z > 0)) {
                                        neeDirPdf = bsdfPdf(bsdf, its.wi, woL);
                                        neeLightPdf = dRec.pdf;
                                        shC = thr * value * bsdfVal;
                                        shO = its.p;
                                        shD = dRec.d;
                                        shMaxT = dRec.dist;
                                        shMedium = transition ? (dot(dRec.d, its.geoN) > 0 ? sr.exteriorMedium : sr.interiorMedium) : medium;
                                        shInteractions = maxDepth - (int)depth - 1;
                                        shOnSurface = 1;
                                        shRng = fr.state;
                                        wantShadow = true;
                                    }
                                }
                            }
                            // ---- BSDF sampling (:268-283)
                            float bPdf, bEta;
                            uint32_t sampledType;
                            float3 woL, wo, bsdfWeight;
                            if (guided) {
                                const float u0 = rng.next1D();
                                const float2 u12 = rng.next2D();
                                float3 fcos;
                                float pb;
                                bool ok = true;
                                if (u0 < A.G.alpha) {
                                    wo = guideSample(A.G, gcell, u0 / A.G.alpha, u12.x, u12.y);
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
                                float gNee = 0.0f, gWo = 0.0f;
                                if (wantShadow || ok) guidePdf2(A.G, gcell, wantShadow ? shD : wo, ok ? wo : shD, gNee, gWo);
                                if (wantShadow) neeDirPdf = A.G.alpha * gNee + (1 - A.G.alpha) * neeDirPdf;
                                bPdf = ok ? A.G.alpha * gWo + (1 - A.G.alpha) * pb : 0.0f;
                                bsdfWeight = (ok && !isZero(fcos) && bPdf > 0) ? fcos / bPdf : f3(0.0f);
                            } else {
                                bsdfWeight = bsdfSample(bsdf, its.wi, rng.next2D(), woL, bPdf, bEta, sampledType);
                                wo = its.sh.toWorld(woL);
                            }
                            if (wantShadow) shC = shC * miWeight(neeLightPdf, neeDirPdf);
                            if (isZero(bsdfWeight) || (cfg.strictNormals && dot(its.geoN, wo) * woL.z <= 0)) {
                                terminate = true;
                            } else {
                                thr *= bsdfWeight;
                                eta *= bEta;
                                newO = its.p;
                                newD = wo;
                                newMint = kEpsilon;
                                newPdf = bPdf;
                                if (A.G.record && (btype & kSmooth) && (int)vcount < A.G.maxVerts) {
                                    guideVertexOpen(A.G, slot, vcount, its.p, bPdf, wo, gcell);
                                    vcount++;
                                    fl &= ~kFlagVertexClosed;
                                }
                                if (transition) medium = dot(wo, its.geoN) > 0 ? sr.exteriorMedium : sr.interiorMedium;  // (:314-315)
                                fl &= ~(kFlagPrevDelta | kFlagPrevMedium);
                                if (sampledType == kNull) {
                                    // index-matched boundary (:318-328): no emitter look-up, no Russian roulette
                                    if (fl & kFlagScattered) fl &= ~kFlagFirst; else fl |= kFlagFirst;
                                    depth++;  // MTS_IGNORE_NULLBSDF_INTERSECTIONS is ON by default (MitsubaBuildOptions.cmake:143-145)
                                } else {
                                    if (sampledType & kDelta) fl |= kFlagPrevDelta;
                                    fl |= kFlagLook;
                                }
                                alive = true;
                            }
                        }
                    }
                }
            }
            if (terminate && wantShadow) {  // park the record for one bounce so that the connection has somewhere to land
                alive = true;
                fl |= kFlagDead;
                terminate = false;
            }
        }

        const AppendResult ap = blockAppend3(nextCount, alive, shadowCount, wantShadow, A.G.record ? A.G.sCount : nullptr,
                                             (A.G.record && valid && terminate) ? vcount : 0u, sAppend, appendParity);
        appendParity ^= 1u;
        const uint32_t j = ap.idxA, sidx = ap.idxB;
        if (alive) {
            stStream(A.next.rayO + j, make_float4(newO.x, newO.y, newO.z, newMint));
            stStream(A.next.rayD + j, make_float4(newD.x, newD.y, newD.z, kInf));
            stStream(A.next.thr + j, make_float4(thr.x, thr.y, thr.z, eta));
            stStream(A.next.rad + j, make_float4(L.x, L.y, L.z, newPdf));
            stStream(A.next.pos + j, make_float4(pos4.x, pos4.y, __uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32))));
            stStream(A.next.flags + j, (fl & ~(kDepthMask | (0xFFu << kVertShift))) | (depth & kDepthMask) | (vcount << kVertShift));
            stStream(A.next.slot + j, slot);
            stStream(A.next.medium + j, medium);
        }
        if (wantShadow) {
            A.shadow.o[sidx] = make_float4(shO.x, shO.y, shO.z, shOnSurface ? kEpsilon : 0.0f);
            A.shadow.d[sidx] = make_float4(shD.x, shD.y, shD.z, shMaxT);
            A.shadow.c[sidx] = make_float4(shC.x, shC.y, shC.z, __uint_as_float(j));
            A.shadow.medium[sidx] = shMedium;
            A.shadow.aux[sidx] = make_uint4((uint32_t)shRng, (uint32_t)(shRng >> 32), pixel, (uint32_t)shInteractions);
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

// Scene::evalTransmittance (scene.cpp:662-722) for the queued emitter connections; an unoccluded
// connection adds contribution * transmittance to the path record it belongs to.
struct ShadowJob {
    const DeviceScene &S;
    const ShadowQueue &Q;
    float4 *rad;
    unsigned long long rays = 0;
    uint32_t idx;
    float3 o, d;
    float remaining, maxt, mintRaw, transmittance, segT;
    int medium, interactions, maxInteractions;
    bool surface, tracking;
    ShapeRecord sr;
    uint32_t primSlot, primIdx;
    Rng rng;
    TrialTracker trk;
    PG_DEV ShadowJob(const DeviceScene &s, const ShadowQueue &q, float4 *r) : S(s), Q(q), rad(r) {}

    // one segment: closest hit with the shadow-style epsilon (ShapeKDTree::rayIntersect(ray, t, shape, n, uv),
    // skdtree.cpp:144-205), occluder test, then the segment's transmittance trials. true = connection decided.
    PG_DEV bool beginSegment() {
        if (!(remaining > 0)) return true;
        Hit h;
        h.prim = kMiss;
        const float mint = adaptiveMinT(o, mintRaw, true);
        uint32_t cn = 0, cp = 0;
        rays++;
        surface = false;
        if (maxt > mint) surface = traceRay<false, false>(S, o, d, mint, maxt, h, &cn, &cp);
        segT = surface ? h.t : kInf;
        if (surface) {
            const PrimInfo pi = S.primInfo[h.prim];
            sr = S.shapes[pi.shape];
            primSlot = h.prim;
            primIdx = pi.prim;
            if (interactions == maxInteractions || !(S.bsdfs[sr.bsdf].typeFlags & kNull)) {
                transmittance = 0.0f;  // occluder
                return true;
            }
        }
        tracking = false;
        if (medium >= 0) {
            tracking = trk.begin(S.media[medium], S.density, o, d, 0.0f, fminf(segT, remaining));
            if (!tracking) transmittance *= trk.value();
        }
        return false;
    }
    PG_DEV bool fetch(uint32_t i) {
        idx = i;
        const float4 ro = Q.o[i], rd = Q.d[i];
        const uint4 aux = Q.aux[i];
        o = f3(ro.x, ro.y, ro.z);
        d = f3(rd.x, rd.y, rd.z);
        medium = Q.medium[i];
        maxInteractions = (int)aux.w;
        rng.state = ((uint64_t)aux.y << 32) | (uint64_t)aux.x;
        rng.inc = ((uint64_t)aux.z << 1) | 1ULL;
        remaining = rd.w;
        mintRaw = ro.w;
        maxt = remaining * (1.0f - kShadowEpsilon);  // the far end lies on the emitter's surface
        transmittance = 1.0f;
        interactions = 0;
        tracking = false;
        if (beginSegment()) {
            finish();
            return false;
        }
        return true;
    }
    PG_DEV void trackStep() {
        if (trk.step(S.media[medium], S.density, o, d, rng)) {
            transmittance *= trk.value();
            tracking = false;
        }
    }
    PG_DEV bool logicStep() {
        if (!surface || transmittance == 0) return true;
        if (sr.interiorMedium >= 0 || sr.exteriorMedium >= 0) {
            const float3 nrm = windingNormal(S, primSlot, sr, primIdx);
            const int expected = dot(-d, nrm) > 0 ? sr.exteriorMedium : sr.interiorMedium;
            if (medium != expected) {  // medium inconsistency (scene.cpp:703-707)
                transmittance = 0.0f;
                return true;
            }
            medium = dot(d, nrm) > 0 ? sr.exteriorMedium : sr.interiorMedium;
        }
        if (++interactions > 100) return true;
        o = o + d * segT;
        remaining -= segT;
        maxt = remaining * (1.0f - kShadowEpsilon);
        mintRaw = kEpsilon;
        return beginSegment();
    }
    PG_DEV void finish() {
        if (transmittance != 0) {
            const float4 c = Q.c[idx];
            const uint32_t dst = __float_as_uint(c.w);
            float4 r = rad[dst];  // at most one connection per path and bounce: no race
            r.x += c.x * transmittance;
            r.y += c.y * transmittance;
            r.z += c.z * transmittance;
            rad[dst] = r;
        }
    }
};

__global__ void __launch_bounds__(128) k_shadow_vol(DeviceScene S, ShadowQueue Q, float4 *__restrict__ rad, const uint32_t *nPtr,
                                                    uint32_t *work, Counters *C) {
    ShadowJob job(S, Q, rad);
    runWithRefill(job, *nPtr, work);
    warpAddU64(&C->shadowRays, job.rays);
}

// b200pg_k_grid_lookup: GridDataSource::lookupFloat on a batch of points
__global__ void __launch_bounds__(256) k_grid_lookup(DeviceScene S, int medium, const float *__restrict__ p, uint32_t n,
                                                     float *__restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = gridLookup(S.media[medium], S.density, ld3(p + 3 * (size_t)i));
}

// b200pg_k_medium_sample: sampleDistance, evalTransmittance and one phase-function sample per ray
__global__ void __launch_bounds__(128) k_medium_test(DeviceScene S, int medium, const float4 *__restrict__ rays, uint32_t n,
                                                     float *__restrict__ outT, float *__restrict__ outTr, float *__restrict__ outWo,
                                                     float *__restrict__ outPdf) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
        const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
        const MediumRecord &M = S.media[medium];
        Rng rng;
        rng.init(S.seed, i, 0);
        MediumSample mRec;
        const bool ok = mediumSampleDistance(M, S.density, o, d, ro.w, rd.w, mRec, rng);
        outT[i] = ok ? mRec.t : kInf;
        outTr[i] = mediumTransmittance(M, S.density, o, d, ro.w, rd.w, rng);
        float pdf;
        const float3 wo = phaseSample(M, -d, rng.next2D(), pdf);
        outWo[3 * (size_t)i] = wo.x;
        outWo[3 * (size_t)i + 1] = wo.y;
        outWo[3 * (size_t)i + 2] = wo.z;
        outPdf[i] = pdf;
    }
}

template <typename K>
static int volGrid(K kernel, int block) {
    int dev = 0, sms = 0, perSM = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, block, 0);
    if (sms <= 0) sms = 148;
    if (perSM <= 0) perSM = 1;
    return sms * perSM;
}

void launchShadeVol(const ShadeArgs &A, cudaStream_t st) {
    static int gridLook = volGrid(k_look_vol, 128), gridTrack = volGrid(k_track_vol, 128), grid = volGrid(k_shade_vol, kShadeThreads);
    k_look_vol<<<gridLook, 128, 0, st>>>(A);
    k_track_vol<<<gridTrack, 128, 0, st>>>(A);
    if (A.perm) k_event_partition<<<gridLook, 256, 0, st>>>(A, const_cast<uint32_t *>(A.perm), const_cast<uint32_t *>(A.permRest));
    k_shade_vol<<<grid, kShadeThreads, 0, st>>>(A);
}
void launchShadowVol(const DeviceScene &S, const ShadowQueue &Q, float4 *rad, const uint32_t *nPtr, uint32_t *work, Counters *C,
                     cudaStream_t st) {
    static int grid = volGrid(k_shadow_vol, 128);
    k_shadow_vol<<<grid, 128, 0, st>>>(S, Q, rad, nPtr, work, C);
}
void launchGridLookup(const DeviceScene &S, int medium, const float *p, uint32_t n, float *out, cudaStream_t st) {
    static int grid = volGrid(k_grid_lookup, 256);
    k_grid_lookup<<<grid, 256, 0, st>>>(S, medium, p, n, out);
}

void launchMediumTest(const DeviceScene &S, int medium, const float4 *rays, uint32_t n, float *outT, float *outTr, float *outWo,
                      float *outPdf, cudaStream_t st) {
    static int grid = volGrid(k_medium_test, 128);
    k_medium_test<<<grid, 128, 0, st>>>(S, medium, rays, n, outT, outTr, outWo, outPdf);
}

}  // namespace pg

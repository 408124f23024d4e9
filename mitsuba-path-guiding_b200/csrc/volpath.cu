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
// The closest-hit query of every newly sampled ray is k_trace's (kernels.cu); follow-up segments behind
// index-matched boundaries are traced inline, by the thread that owns the path.
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

// rayIntersectAndLookForEmitter (progressive_volpath.cpp:401-460). `first` is the closest hit of (o, d)
// found by k_trace. Returns the attenuated emitted radiance and the direct-sampling record fields that
// pdfEmitterDirect needs.
PG_DEV float3 lookForEmitter(const DeviceScene &S, Rng &fr, int medium, int maxInteractions, float3 o, float3 d, BoundaryHit cur,
                             int &emitter, float3 &emN, float &emDist, unsigned long long &rays) {
    float transmittance = 1.0f;
    int interactions = 0;
    emitter = -1;
    while (true) {
        if (medium >= 0) transmittance *= mediumTransmittance(S.media[medium], S.density, o, d, 0.0f, cur.t, fr);
        if (cur.valid && (interactions == maxInteractions || !cur.isNull || cur.emitter >= 0)) break;
        if (!cur.valid) break;
        if (transmittance == 0) return f3(0.0f);
        if (cur.transition) medium = dot(d, cur.geoN) > 0 ? cur.exterior : cur.interior;
        o = o + d * cur.t;
        if (++interactions > 100) return f3(0.0f);
        Hit h;
        const float mint = adaptiveMinT(o, kEpsilon, false);
        uint32_t cn = 0, cp = 0;
        traceRay<false, false>(S, o, d, mint, kInf, h, &cn, &cp);
        rays++;
        Intersection its;
        if (h.prim != kMiss) fillIntersection(S, o, d, h, its);
        cur = boundaryOf(S, h, its);
    }
    if (cur.valid && cur.emitter >= 0) {
        emitter = cur.emitter;
        emN = cur.shN;
        emDist = cur.t;
        const float3 Le = dot(cur.shN, -d) <= 0 ? f3(0.0f) : ld3(S.emitters[cur.emitter].radiance);  // area.cpp:104-109
        return Le * transmittance;
    }
    return f3(0.0f);
}

__global__ void __launch_bounds__(kShadeThreads) k_shade_vol(ShadeArgs A) {
    const DeviceScene &S = A.S;
    const IntegratorConfig &cfg = A.cfg;
    const uint32_t n = A.C->queue[A.bounce];
    uint32_t *nextCount = &A.C->queue[A.bounce + 1];
    uint32_t *shadowCount = &A.C->shadow[A.bounce];
    unsigned long long donePaths = 0, doneLen = 0, extraRays = 0;
    __shared__ uint32_t sAppend[2 * 3 * (kShadeThreads / 32 + 1)];
    uint32_t appendParity = 0;

    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < n;

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

        if (valid) {
            const float4 ro = A.cur.rayO[i], rd = A.cur.rayD[i], thr4 = A.cur.thr[i], rad4 = A.cur.rad[i];
            pos4 = A.cur.pos[i];
            fl = A.cur.flags[i];
            slot = A.cur.slot[i];
            medium = A.cur.medium[i];
            const float4 h4 = A.hits[i];
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

                // ---- second half of the previous loop iteration: emitter look-up along the sampled ray with
                // MIS (:288-303 phase, :330-345 BSDF), then Russian roulette (:350-360)
                if (fl & kFlagLook) {
                    Rng fr = rng.fork();
                    int emitter;
                    float3 emN = f3(0.0f);
                    float emDist = 0.0f;
                    // depth here is the value before the loop's depth++ (the look-up ran with maxDepth - depth - 1)
                    const float3 value = lookForEmitter(S, fr, medium, maxDepth - (int)depth - 1, o, d, boundaryOf(S, h, its), emitter,
                                                        emN, emDist, extraRays);
                    const bool prevMedium = (fl & kFlagPrevMedium) != 0;
                    if (!isZero(value) && (!prevMedium || fminf(value.x, fminf(value.y, value.z)) > 0.0f)) {
                        const float emitterPdf = (cfg.useNee && !(fl & kFlagPrevDelta)) ? pdfEmitterDirect(S, emitter, d, emN, emDist) : 0.0f;
                        const float weight = cfg.useNee ? miWeight(rad4.w, emitterPdf) : 1.0f;
                        L += thr * value * weight;
                    }
                    fl &= ~(kFlagFirst | kFlagLook);  // rRec.type &= ~EEmittedRadiance
                    if (depth++ >= (uint32_t)cfg.rrDepth) {
                        const float q = fminf(maxComp(thr) * eta * eta, 0.95f);
                        if (rng.next1D() >= q)
                            terminate = true;
                        else
                            thr = thr / q;
                    }
                    fl |= kFlagScattered;
                }
                if (!terminate && !((int)depth <= maxDepth || maxDepth < 0)) terminate = true;

                if (!terminate) {
                    // ---- free-flight distance in the current medium (:118-121)
                    MediumSample mRec;
                    bool mediumEvent = false;
                    const bool guidedDist = A.G.enabled && cfg.guidedDistance && medium >= 0;
                    float3 distWeight = f3(1.0f);
                    if (guidedDist)
                        mediumEvent = mediumSampleDistanceGuided(S.media[medium], S.density, A.G, o, d, 0.0f, hitValid ? h.t : kInf, mRec,
                                                                 distWeight, rng);
                    else if (medium >= 0)
                        mediumEvent = mediumSampleDistance(S.media[medium], S.density, o, d, 0.0f, hitValid ? h.t : kInf, mRec, rng);
                    if (guidedDist) thr *= distWeight;  // real-collision and null-collision weights alike
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
                                        float dirPdf = phaseVal;  // phase pdf == phase value
                                        if (guided) dirPdf = A.G.alpha * guidePdf(A.G, gcell, dRec.d) + (1 - A.G.alpha) * phaseVal;
                                        const float weight = miWeight(dRec.pdf, dirPdf);
                                        shC = thr * value * phaseVal * weight;
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
                                phasePdf = A.G.alpha * guidePdf(A.G, gcell, wo) + (1 - A.G.alpha) * pp;
                                dirOk = pp > 0 && phasePdf > 0;
                                if (dirOk) thr = thr * (pp / phasePdf);
                            } else {
                                wo = phaseSample(M, -d, rng.next2D(), phasePdf);  // phase weight == 1
                            }
                            if (!dirOk) {
                                terminate = true;
                            } else {
                            if (A.G.record && (int)vcount < A.G.maxVerts) {
                                guideVertexOpen(A.G, slot, vcount, mRec.p, phasePdf, wo);
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
                                        float bPdf = bsdfPdf(bsdf, its.wi, woL);
                                        if (guided) bPdf = A.G.alpha * guidePdf(A.G, gcell, dRec.d) + (1 - A.G.alpha) * bPdf;
                                        const float weight = miWeight(dRec.pdf, bPdf);
                                        shC = thr * value * bsdfVal * weight;
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
                                bPdf = ok ? A.G.alpha * guidePdf(A.G, gcell, wo) + (1 - A.G.alpha) * pb : 0.0f;
                                bsdfWeight = (ok && !isZero(fcos) && bPdf > 0) ? fcos / bPdf : f3(0.0f);
                            } else {
                                bsdfWeight = bsdfSample(bsdf, its.wi, rng.next2D(), woL, bPdf, bEta, sampledType);
                                wo = its.sh.toWorld(woL);
                            }
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
                                    guideVertexOpen(A.G, slot, vcount, its.p, bPdf, wo);
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
            A.next.rayO[j] = make_float4(newO.x, newO.y, newO.z, newMint);
            A.next.rayD[j] = make_float4(newD.x, newD.y, newD.z, kInf);
            A.next.thr[j] = make_float4(thr.x, thr.y, thr.z, eta);
            A.next.rad[j] = make_float4(L.x, L.y, L.z, newPdf);
            A.next.pos[j] = make_float4(pos4.x, pos4.y, __uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32)));
            A.next.flags[j] = (fl & ~(kDepthMask | (0xFFu << kVertShift))) | (depth & kDepthMask) | (vcount << kVertShift);
            A.next.slot[j] = slot;
            A.next.medium[j] = medium;
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
    warpAddU64(&A.C->normalRays, extraRays);
}

// Scene::evalTransmittance (scene.cpp:662-722) for the queued emitter connections; an unoccluded
// connection adds contribution * transmittance to the path record it belongs to.
__global__ void __launch_bounds__(128) k_shadow_vol(DeviceScene S, ShadowQueue Q, float4 *__restrict__ rad, const uint32_t *nPtr,
                                                    uint32_t *work, Counters *C) {
    const uint32_t n = *nPtr;
    unsigned long long rays = 0;
    while (true) {
        uint32_t base = 0;
        if (laneId() == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + laneId();
        if (i >= n) continue;
        const float4 ro = Q.o[i], rd = Q.d[i];
        const uint4 aux = Q.aux[i];
        float3 o = f3(ro.x, ro.y, ro.z);
        const float3 d = f3(rd.x, rd.y, rd.z);
        int medium = Q.medium[i];
        const int maxInteractions = (int)aux.w;
        Rng rng;
        rng.state = ((uint64_t)aux.y << 32) | (uint64_t)aux.x;
        rng.inc = ((uint64_t)aux.z << 1) | 1ULL;
        float remaining = rd.w;
        const float lengthFactor = 1.0f - kShadowEpsilon;  // the far end lies on the emitter's surface
        float mintRaw = ro.w, maxt = remaining * lengthFactor;
        float transmittance = 1.0f;
        int interactions = 0;
        while (remaining > 0) {
            // ShapeKDTree::rayIntersect(ray, t, shape, n, uv) (skdtree.cpp:144-205): closest hit, shadow-style epsilon
            Hit h;
            h.prim = kMiss;
            const float mint = adaptiveMinT(o, mintRaw, true);
            uint32_t cn = 0, cp = 0;
            rays++;
            bool surface = false;
            if (maxt > mint) surface = traceRay<false, false>(S, o, d, mint, maxt, h, &cn, &cp);
            const float t = surface ? h.t : kInf;
            ShapeRecord sr;
            uint32_t primIdx = 0;
            if (surface) {
                const PrimInfo pi = S.primInfo[h.prim];
                sr = S.shapes[pi.shape];
                primIdx = pi.prim;
                if (interactions == maxInteractions || !(S.bsdfs[sr.bsdf].typeFlags & kNull)) {
                    transmittance = 0.0f;  // occluder
                    break;
                }
            }
            if (medium >= 0) transmittance *= mediumTransmittance(S.media[medium], S.density, o, d, 0.0f, fminf(t, remaining), rng);
            if (!surface || transmittance == 0) break;
            if (sr.interiorMedium >= 0 || sr.exteriorMedium >= 0) {
                const float3 nrm = windingNormal(S, h.prim, sr, primIdx);
                const int expected = dot(-d, nrm) > 0 ? sr.exteriorMedium : sr.interiorMedium;
                if (medium != expected) {  // medium inconsistency (scene.cpp:703-707)
                    transmittance = 0.0f;
                    break;
                }
                medium = dot(d, nrm) > 0 ? sr.exteriorMedium : sr.interiorMedium;
            }
            if (++interactions > 100) break;
            o = o + d * t;
            remaining -= t;
            maxt = remaining * lengthFactor;
            mintRaw = kEpsilon;
        }
        if (transmittance != 0) {
            const float4 c = Q.c[i];
            const uint32_t dst = __float_as_uint(c.w);
            float4 r = rad[dst];  // at most one connection per path and bounce: no race
            r.x += c.x * transmittance;
            r.y += c.y * transmittance;
            r.z += c.z * transmittance;
            rad[dst] = r;
        }
    }
    warpAddU64(&C->shadowRays, rays);
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
    static int grid = volGrid(k_shade_vol, kShadeThreads);
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

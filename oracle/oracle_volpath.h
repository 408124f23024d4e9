// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_volpath.h: ProgressiveVolumetricPathTracer::Li (src/integrators/path/progressive_volpath.cpp:98-374),
// rayIntersectAndLookForEmitter (:401-460), Scene::evalTransmittance (src/librender/scene.cpp:662-722) and
// Scene::sampleAttenuatedEmitterDirect (scene.cpp:897-939).
//
// RNG note: the reference feeds ONE sampler to distance sampling, direction sampling and the
// stochastic transmittance estimates. Here the transmittance estimates of a connection (attenuated NEE,
// look-for-emitter) draw from a stream FORKED off the main one at the start of the connection, so that the main
// stream advances by a fixed amount per connection; the CUDA path does the same. Distributions are unchanged.
#pragma once

namespace orc {

// ShapeKDTree::rayIntersect(ray, t, shape, n, uv) (skdtree.cpp:144-205): closest hit, shadow-style epsilon,
// geometric normal from the triangle winding (not flipped towards the shading normal).
static bool rayIntersectT(const Scene &scene, const Ray &ray, Float &t, int &shapeIdx, Vec3 &n, Stats *st) {
    IntersectionCache cache;
    Float mint, maxt;
    t = std::numeric_limits<Float>::infinity();
    if (st) st->shadowRays++;
    if (scene.kd.aabb.rayIntersect(ray, mint, maxt)) {
        Float rayMinT = ray.mint;
        if (rayMinT == Epsilon) rayMinT *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        if (maxt > mint) {
            if (scene.rayIntersectHavran<false>(ray, mint, maxt, t, cache, st ? &st->trav : nullptr)) {
                shapeIdx = (int)cache.shapeIndex;
                const Shape &s = scene.shapes[shapeIdx];
                if (s.type == B200PG_SHAPE_TRIMESH) {
                    const Vec3 &p0 = s.positions[s.indices[3 * cache.primIndex]], &p1 = s.positions[s.indices[3 * cache.primIndex + 1]],
                               &p2 = s.positions[s.indices[3 * cache.primIndex + 2]];
                    n = normalize(cross(p1 - p0, p2 - p0));
                } else {
                    n = s.frame.n;
                }
                return true;
            }
        }
    }
    return false;
}

static inline int targetMedium(const Shape &s, const Vec3 &geoN, const Vec3 &d) {  // records.inl:81-86
    return dot(d, geoN) > 0 ? s.exteriorMedium : s.interiorMedium;
}

// Scene::evalTransmittance, scene.cpp:662-722
static Float evalTransmittance(const Scene &scene, const Vec3 &p1, bool p1OnSurface, const Vec3 &p2, bool p2OnSurface, int medium,
                               int &interactions, Rng &rng, Stats *st) {
    Vec3 d = p2 - p1;
    Float remaining = length(d);
    d /= remaining;
    Float lengthFactor = p2OnSurface ? (1 - ShadowEpsilon) : 1;
    Ray ray(p1, d, p1OnSurface ? Epsilon : 0, remaining * lengthFactor);
    Float transmittance = 1.0f;
    int maxInteractions = interactions;
    interactions = 0;
    while (remaining > 0) {
        Float t;
        int shapeIdx = -1;
        Vec3 n;
        bool surface = rayIntersectT(scene, ray, t, shapeIdx, n, st);
        if (surface && (interactions == maxInteractions || !(scene.bsdfOf(scene.shapes[shapeIdx]).typeFlags() & ENull)))
            return 0.0f;  // occluder
        if (medium >= 0)
            transmittance *= scene.media[medium].evalTransmittance(ray.o, ray.d, 0, std::min(t, remaining), rng);
        if (!surface || transmittance == 0) break;
        // null BSDF eval with typeMask = ENull, EDiscrete: 1 (null.cpp:52-58)
        const Shape &s = scene.shapes[shapeIdx];
        if (s.isMediumTransition()) {
            if (medium != targetMedium(s, n, -d)) return 0.0f;  // medium inconsistency (scene.cpp:703-707)
            medium = targetMedium(s, n, d);
        }
        if (++interactions > 100) break;
        ray.o = ray(t);
        remaining -= t;
        ray.maxt = remaining * lengthFactor;
        ray.mint = Epsilon;
    }
    return transmittance;
}

// REFERENCE QUIRK, reproduced on purpose (parity): when the sampled ray reaches an emitter THROUGH index-matched boundaries,
// dRec.setQuery(ray, *its) (records.inl:170-178) takes dist = its.t of the LAST segment -- the ray origin has been moved to
// the last boundary (progressive_volpath.cpp:437) -- not the distance from the path vertex. pdfEmitterDirect converts the
// area density with dist^2 (shape.cpp:117-126), so the emitter pdf that enters the MIS weight is too small, the
// BSDF/phase-sampled contribution is over-weighted, and the estimator is biased bright wherever a null boundary lies
// between a vertex and the light (the furnace test measures +7..10 %; without next-event estimation, or with the total
// distance, it returns the exact value: tests/test_oracle_transport.py). g_lookupTotalDistance = 1 (test hook
// orc_debug_flag) switches to the total distance to demonstrate this; the default is the reference's behaviour.
static int g_lookupTotalDistance = 0;

// progressive_volpath.cpp:401-460
static void rayIntersectAndLookForEmitter(const Scene &scene, Rng &rng, int medium, int maxInteractions, Ray ray, Intersection &_its,
                                          Scene::DirectSample &dRec, Vec3 &value, Stats &st) {
    Intersection its2, *its = &_its;
    Float transmittance = 1.0f;
    bool surface = false;
    int interactions = 0;
    Float walked = 0.0f;  // length of the segments before the last one
    while (true) {
        surface = scene.rayIntersect(ray, *its, &st);
        if (medium >= 0) transmittance *= scene.media[medium].evalTransmittance(ray.o, ray.d, 0, its->t, rng);
        if (surface && (interactions == maxInteractions || !(scene.bsdfOf(scene.shapes[its->shape]).typeFlags() & ENull) ||
                        scene.shapes[its->shape].emitter >= 0))
            break;
        if (!surface) break;
        if (transmittance == 0) return;
        const Shape &s = scene.shapes[its->shape];
        if (s.isMediumTransition()) medium = targetMedium(s, its->geoN, ray.d);
        walked += its->t;
        ray.o = ray(its->t);
        ray.mint = Epsilon;
        its = &its2;
        if (++interactions > 100) return;
    }
    if (surface && scene.shapes[its->shape].emitter >= 0) {
        dRec.p = its->p;  // dRec.setQuery(ray, *its)
        dRec.n = its->shFrame.n;
        dRec.emitter = scene.shapes[its->shape].emitter;
        dRec.d = ray.d;
        dRec.dist = g_lookupTotalDistance ? walked + its->t : its->t;  // see REFERENCE QUIRK above
        value = scene.emitterEval(*its, -ray.d) * transmittance;
    }
}

// Guided variant (G != nullptr): this repo's design on top of the reference loop --
//   * direction sampling at medium vertices and at surface vertices with a smooth BSDF: one-sample MIS between the phase
//     function / BSDF and the cell's vMF mixture (probability alpha for the mixture), exactly as in Li_path;
//   * P.guided_distance: guided collision probabilities in the free-flight sampling (Medium::sampleDistanceGuided);
//   * every such vertex is recorded as a training sample.
static Vec3 Li_volpath(const Scene &scene, const B200pgIntegratorParams &P, const Ray &r, Rng &rng, Stats &st,
                       const GuideCtx *G = nullptr) {
    const GuideField *field = G ? G->field : nullptr;
    const bool record = G && G->rec;
    const Float alpha = G ? G->alpha : 0.0f;
    GuideVertex verts[64];
    int nVerts = 0;
    Intersection its;
    MediumSample mRec;
    Ray ray(r);
    Vec3 Li(0.0f);
    Float eta = 1.0f;
    int depth = 1;
    int medium = scene.camera.medium;
    bool emittedRadiance = true;  // rRec.type & EEmittedRadiance
    const int maxDepth = P.max_depth;

    scene.rayIntersect(ray, its, &st);
    Vec3 throughput(1.0f);
    bool scattered = false;

    while (depth <= maxDepth || maxDepth < 0) {
        bool mediumEvent = false;
        const bool guidedDist = field && P.guided_distance && medium >= 0;
        Vec3 distWeight(1.0f);
        if (guidedDist) {
            const Vec3 rd = ray.d;
            mediumEvent = scene.media[medium].sampleDistanceGuided(ray.o, ray.d, 0, its.t, mRec, distWeight, rng,
                                                                   [&](const Vec3 &p) { return field->pdf(field->lookup(p), rd); });
        } else if (medium >= 0) {
            mediumEvent = scene.media[medium].sampleDistance(ray.o, ray.d, 0, its.t, mRec, rng);
        }
        if (mediumEvent) {
            const Medium &med = scene.media[medium];
            if (depth >= maxDepth && maxDepth != -1) break;
            if (guidedDist)
                throughput *= distWeight;
            else
                throughput *= mRec.sigmaS * mRec.transmittance / mRec.pdfSuccess;
            const bool guided = field != nullptr;
            const uint32_t gcell = guided ? field->lookup(mRec.p) : 0u;

            Scene::DirectSample dRec;
            dRec.ref = mRec.p;
            dRec.refN = Vec3(0.0f);
            if (P.use_nee) {
                int interactions = maxDepth - depth - 1;
                Vec3 value = scene.sampleEmitterDirectNoVis(dRec, rng.next2D());
                Rng fr = rng.fork();
                if (dRec.pdf != 0) {
                    value *= evalTransmittance(scene, dRec.ref, false, dRec.p, true, medium, interactions, fr, &st);
                } else {
                    value = Vec3(0.0f);
                }
                if (!value.isZero()) {
                    Float phaseVal = med.phaseEval(-ray.d, dRec.d);
                    if (phaseVal != 0) {
                        Float phasePdf = phaseVal;  // phase->pdf == eval for hg / isotropic
                        if (guided) phasePdf = alpha * field->pdf(gcell, dRec.d) + (1 - alpha) * phasePdf;
                        const Float weight = miWeight(dRec.pdf, phasePdf);
                        Li += throughput * value * phaseVal * weight;
                    }
                }
            }
            const Vec3 LafterNee = Li;
            Float phasePdf;
            Vec3 wo;
            if (guided) {  // one-sample MIS between the mixture and the phase function
                Float u0 = rng.next1D();
                Vec2 u12 = rng.next2D();
                Float pp;
                if (u0 < alpha) {
                    wo = field->sample(gcell, u0 / alpha, u12.x, u12.y);
                    pp = med.phaseEval(-ray.d, wo);
                } else {
                    wo = med.phaseSample(-ray.d, u12, pp);
                }
                phasePdf = alpha * field->pdf(gcell, wo) + (1 - alpha) * pp;
                if (!(pp > 0) || !(phasePdf > 0)) break;
                throughput *= pp / phasePdf;
            } else {
                wo = med.phaseSample(-ray.d, rng.next2D(), phasePdf);  // phaseWeight == 1
            }
            if (record && nVerts < 64) {
                GuideVertex &gv = verts[nVerts++];
                gv.pos = mRec.p;
                gv.dir = wo;
                gv.pdf = phasePdf;
                gv.T = throughput;
                gv.Lk = LafterNee;
                gv.dist = 0.0f;
            }
            ray = Ray(mRec.p, wo, 0.0f);
            Vec3 value(0.0f);
            Rng fr = rng.fork();
            rayIntersectAndLookForEmitter(scene, fr, medium, maxDepth - depth - 1, ray, its, dRec, value, st);
            if (record && nVerts > 0 && verts[nVerts - 1].dist == 0.0f && its.isValid()) verts[nVerts - 1].dist = its.t;
            if (!value.isZero() && std::min(value.x, std::min(value.y, value.z)) > 0.f) {
                const Float emitterPdf = P.use_nee ? scene.pdfEmitterDirect(dRec) : 0.0f;
                const Float weight = P.use_nee ? miWeight(phasePdf, emitterPdf) : 1.0f;
                Li += throughput * value * weight;
            }
            emittedRadiance = false;
        } else {
            if (guidedDist)
                throughput *= distWeight;
            else if (medium >= 0)
                throughput *= mRec.transmittance / mRec.pdfFailure;
            if (!its.isValid()) break;
            const Shape &shape = scene.shapes[its.shape];
            const Bsdf &bsdf = scene.bsdfOf(shape);
            if (shape.emitter >= 0 && emittedRadiance && (!P.hide_emitters || scattered)) Li += throughput * scene.emitterEval(its, -ray.d);
            if (depth >= maxDepth && maxDepth != -1) break;
            Float wiDotGeoN = -dot(its.geoN, ray.d), wiDotShN = Frame::cosTheta(its.wi);
            if (wiDotGeoN * wiDotShN < 0 && P.strict_normals) break;

            Scene::DirectSample dRec;
            dRec.ref = its.p;
            dRec.refN = Vec3(0.0f);
            const unsigned btype = bsdf.typeFlags();
            if ((btype & (ETransmission | EBackSide)) == 0) dRec.refN = its.shFrame.n;
            const bool guided = field && (btype & ESmooth);
            const uint32_t gcell = guided ? field->lookup(its.p) : 0u;
            if (P.use_nee && (btype & ESmooth)) {
                int interactions = maxDepth - depth - 1;
                Vec3 value = scene.sampleEmitterDirectNoVis(dRec, rng.next2D());
                Rng fr = rng.fork();
                if (dRec.pdf != 0) {
                    int med = medium;
                    if (shape.isMediumTransition()) med = targetMedium(shape, its.geoN, dRec.d);
                    value *= evalTransmittance(scene, its.p, true, dRec.p, true, med, interactions, fr, &st);
                } else {
                    value = Vec3(0.0f);
                }
                if (!value.isZero()) {
                    Vec3 woL = its.toLocal(dRec.d);
                    const Vec3 bsdfVal = bsdf.eval(its.wi, woL);
                    Float woDotGeoN = dot(its.geoN, dRec.d);
                    if (!bsdfVal.isZero() && (!P.strict_normals || woDotGeoN * Frame::cosTheta(woL) > 0)) {
                        Float bsdfPdf = bsdf.pdf(its.wi, woL);
                        if (guided) bsdfPdf = alpha * field->pdf(gcell, dRec.d) + (1 - alpha) * bsdfPdf;
                        const Float weight = miWeight(dRec.pdf, bsdfPdf);
                        Li += throughput * (value * bsdfVal * weight);
                    }
                }
            }
            const Vec3 LafterNee = Li;
            Float bsdfPdf, bEta;
            unsigned sampledType;
            Vec3 woLocal, bsdfWeight, wo;
            if (guided) {
                Float u0 = rng.next1D();
                Vec2 u12 = rng.next2D();
                Vec3 fcos;
                Float pb;
                if (u0 < alpha) {
                    wo = field->sample(gcell, u0 / alpha, u12.x, u12.y);
                    woLocal = its.toLocal(wo);
                    fcos = bsdf.eval(its.wi, woLocal);
                    pb = bsdf.pdf(its.wi, woLocal);
                    bEta = 1.0f;
                    sampledType = EGlossyReflection;
                } else {
                    Vec3 w = bsdf.sample(its.wi, u12, woLocal, pb, bEta, sampledType);
                    if (w.isZero()) break;
                    fcos = w * pb;
                    wo = its.toWorld(woLocal);
                }
                bsdfPdf = alpha * field->pdf(gcell, wo) + (1 - alpha) * pb;
                if (fcos.isZero() || !(bsdfPdf > 0)) break;
                bsdfWeight = fcos / bsdfPdf;
            } else {
                bsdfWeight = bsdf.sample(its.wi, rng.next2D(), woLocal, bsdfPdf, bEta, sampledType);
                if (bsdfWeight.isZero()) break;
                wo = its.toWorld(woLocal);
            }
            Float woDotGeoN = dot(its.geoN, wo);
            if (woDotGeoN * Frame::cosTheta(woLocal) <= 0 && P.strict_normals) break;
            ray = Ray(its.p, wo);
            throughput *= bsdfWeight;
            eta *= bEta;
            if (record && (btype & ESmooth) && nVerts < 64) {
                GuideVertex &gv = verts[nVerts++];
                gv.pos = its.p;
                gv.dir = wo;
                gv.pdf = bsdfPdf;
                gv.T = throughput;
                gv.Lk = LafterNee;
                gv.dist = 0.0f;
            }
            if (shape.isMediumTransition()) medium = targetMedium(shape, its.geoN, ray.d);
            if (sampledType == ENull) {  // index-matched boundary (:318-328)
                emittedRadiance = !scattered;
                scene.rayIntersect(ray, its, &st);
                depth++;  // MTS_IGNORE_NULLBSDF_INTERSECTIONS is defined by default (MitsubaBuildOptions.cmake:141-145)
                continue;
            }
            Vec3 value(0.0f);
            Rng fr = rng.fork();
            rayIntersectAndLookForEmitter(scene, fr, medium, maxDepth - depth - 1, ray, its, dRec, value, st);
            if (record && (btype & ESmooth) && nVerts > 0 && its.isValid()) verts[nVerts - 1].dist = its.t;
            if (!value.isZero()) {
                const Float emitterPdf = (P.use_nee && !(sampledType & EDelta)) ? scene.pdfEmitterDirect(dRec) : 0;
                const Float weight = P.use_nee ? miWeight(bsdfPdf, emitterPdf) : 1.0f;
                Li += throughput * value * weight;
            }
            emittedRadiance = false;
        }
        if (depth++ >= P.rr_depth) {
            Float q = std::min(throughput.maxc() * eta * eta, 0.95f);
            if (rng.next1D() >= q) break;
            throughput /= q;
        }
        scattered = true;
    }
    st.paths++;
    st.pathLen += depth;
    if (record) {  // same training-sample definition as Li_path
        for (int v = 0; v < nVerts; ++v) {
            const GuideVertex &gv = verts[v];
            Vec3 dL = Li - gv.Lk, est(0.0f);
            for (int c = 0; c < 3; ++c) est[c] = gv.T[c] > 0 ? dL[c] / gv.T[c] : 0.0f;
            Float w = est.average() / gv.pdf;
            if (!std::isfinite(w) || w < 0) w = 0.0f;
            G->rec->push(gv.pos, gv.dir, w, gv.pdf, gv.dist);
        }
    }
    return Li;
}

}  // namespace orc

// TEST INFRASTRUCTURE ONLY -- CPU oracle (restatement of the reference's algorithm).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may build, link or execute anything under oracle/. The product path never does.
//
// oracle_pt.cpp: scene assembly, Havran kd traversal, intersection records, area-light NEE,
// perspective camera, Gaussian-splat film and ProgressiveMIPathTracer::Li, restated from the
// reference files cited at each function (paths relative to the reference root).
//
// PARITY PINNING: the restatement is pinned SAMPLE BY SAMPLE to the reference's own code -- its libraries and the plugins
// of this path compiled from /root/reference into oracle/_ref by oracle/Makefile.ref and driven through
// oracle/ref_harness/ref_harness.cpp with a replay sampler (tests/test_upstream.py + tests/golden/upstream.npz,
// tests/test_ref_pin.py, tests/test_ref_xml_semantics.py, tests/test_ref_meshes.py; DESIGN.md "Reference build status") --
// and, as before, through the properties the reference's own tests check (test_chisquare: sample<->pdf<->eval consistency,
// tests/test_oracle_bsdf.py; test_kd / test_dgeom known answers, tests/test_oracle_kd_film.py) and closed forms
// (tests/test_oracle_transport.py). The guiding part (oracle_guiding.h) has no reference counterpart in the snapshot:
// parity unpinned there, and it says so.
#include <omp.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include "oracle_bsdf.h"
#include "oracle_medium.h"
#include "oracle_scene.h"
#include "oracle_guiding.h"

namespace orc {

struct Emitter {
    Vec3 radiance;
    Float samplingWeight;
    int shape;
};

struct Camera {  // perspective.cpp:126-180, 271-298
    Mat4 toWorld;
    Mat4 sampleToCamera;
    Float nearClip, farClip;
    Vec2 invResolution;
    int medium;
};

struct Film {
    int width, height;
    Float radius;
    Float values[32];  // MTS_FILTER_RESOLUTION + 1
    Float scaleFactor;
    int borderSize;
    void configure(Float stddev) {  // gaussian.cpp:25-51, rfilter.cpp:38-56
        radius = 4 * stddev;
        Float alpha = -1.0f / (2.0f * stddev * stddev);
        Float sum = 0.0f;
        for (int i = 0; i < 31; ++i) {
            Float x = (radius * i) / 31;
            Float value = std::max(0.0f, std::exp(alpha * x * x) - std::exp(alpha * radius * radius));
            values[i] = value;
            sum += value;
        }
        values[31] = 0.0f;
        scaleFactor = 31 / radius;
        borderSize = (int)std::ceil(radius - 0.5f);
        sum *= 2 * radius / 31;
        Float normalization = 1.0f / sum;
        for (int i = 0; i < 31; ++i) values[i] *= normalization;
    }
    Float evalDiscretized(Float x) const { return values[std::min((int)std::abs(x * scaleFactor), 31)]; }
};

// ImageBlock (imageblock.h:131-197): RGB, alpha, weight; bordered.
struct ImageBlock {
    int ox, oy, w, h, border;
    std::vector<Float> data;  // (h+2b)*(w+2b)*5
    void init(int ox_, int oy_, int w_, int h_, int border_) {
        ox = ox_; oy = oy_; w = w_; h = h_; border = border_;
        data.assign((size_t)(w + 2 * border) * (h + 2 * border) * 5, 0.0f);
    }
    bool put(const Film &f, const Vec2 &_pos, const Float *value) {
        for (int i = 0; i < 5; ++i)
            if (!std::isfinite(value[i]) || value[i] < 0) return false;
        const int sx = w + 2 * border, sy = h + 2 * border;
        const Float px = _pos.x - 0.5f - (ox - border), py = _pos.y - 0.5f - (oy - border);
        const int minx = std::max((int)std::ceil(px - f.radius), 0), miny = std::max((int)std::ceil(py - f.radius), 0);
        const int maxx = std::min((int)std::floor(px + f.radius), sx - 1),
                  maxy = std::min((int)std::floor(py + f.radius), sy - 1);
        Float wx[16], wy[16];
        for (int x = minx, idx = 0; x <= maxx; ++x) wx[idx++] = f.evalDiscretized(x - px);
        for (int y = miny, idx = 0; y <= maxy; ++y) wy[idx++] = f.evalDiscretized(y - py);
        for (int y = miny, yr = 0; y <= maxy; ++y, ++yr) {
            const Float weightY = wy[yr];
            Float *dest = data.data() + ((size_t)y * sx + minx) * 5;
            for (int x = minx, xr = 0; x <= maxx; ++x, ++xr) {
                const Float weight = wx[xr] * weightY;
                for (int k = 0; k < 5; ++k) *dest++ += weight * value[k];
            }
        }
        return true;
    }
};

struct Stats {
    uint64_t paths = 0, normalRays = 0, shadowRays = 0, pathLen = 0;
    TraversalCounters trav;
};

struct Scene {
    std::vector<Shape> shapes;
    std::vector<Bsdf> bsdfs;
    std::vector<Emitter> emitters;
    std::vector<Medium> media;
    std::vector<Float> emitterCdf;
    std::vector<TriAccel> triAccel;  // one per global primitive
    std::vector<AABB> primBoxes;
    KDTree kd;
    Camera camera;
    Film film;
    int sampleCount;
    uint64_t seed;
    int defaultDiffuseHalf = -1, defaultBlack = -1, defaultNull = -1;

    // ---------------- primitive intersection (skdtree.h:248-336) ----------------
    inline bool rectIntersect(const Shape &s, const Ray &_ray, Float mint, Float maxt, Float &t, Float *temp) const {
        // rectangle.cpp:125-148
        Vec3 o = s.worldToObject.pointAffine(_ray.o), d = s.worldToObject.vector(_ray.d);
        Float hit = -o.z / d.z;
        if (!(hit >= mint && hit <= maxt)) return false;
        Vec3 local = o + d * hit;
        if (std::abs(local.x) <= 1 && std::abs(local.y) <= 1) {
            t = hit;
            if (temp) {
                temp[0] = local.x;
                temp[1] = local.y;
            }
            return true;
        }
        return false;
    }

    inline bool intersectPrim(const Ray &ray, uint32_t idx, Float mint, Float maxt, Float &t, IntersectionCache &cache) const {
        const TriAccel &ta = triAccel[idx];
        if (ta.k != KNoTriangleFlag) {
            Float tempU, tempV, tempT;
            if (ta.rayIntersect(ray, mint, maxt, tempU, tempV, tempT)) {
                t = tempT;
                cache.shapeIndex = ta.shapeIndex;
                cache.primIndex = ta.primIndex;
                cache.u = tempU;
                cache.v = tempV;
                return true;
            }
        } else {
            Float temp[2];
            if (rectIntersect(shapes[ta.shapeIndex], ray, mint, maxt, t, temp)) {
                cache.shapeIndex = ta.shapeIndex;
                cache.primIndex = KNoTriangleFlag;
                cache.u = temp[0];
                cache.v = temp[1];
                return true;
            }
        }
        return false;
    }

    // ---------------- Havran traversal (sahkdtree3.h:178-308) ----------------
    template <bool shadowRay>
    bool rayIntersectHavran(const Ray &ray, Float mint, Float maxt, Float &t, IntersectionCache &cache,
                            TraversalCounters *cnt) const {
        struct KDStackEntryHavran {
            const KDNode *node;
            Float t;
            uint32_t prev;
            Vec3 p;
        } stack[48];
        uint32_t mailbox[8];
        std::memset(mailbox, 0xFF, sizeof(mailbox));

        uint32_t enPt = 0;
        stack[enPt].t = mint;
        stack[enPt].p = ray(mint);
        uint32_t exPt = 1;
        stack[exPt].t = maxt;
        stack[exPt].p = ray(maxt);
        stack[exPt].node = nullptr;

        bool foundIntersection = false;
        const KDNode *currNode = kd.nodes.data();
        while (currNode != nullptr) {
            while (!currNode->isLeaf()) {
                if (cnt) cnt->nodes++;
                const Float splitVal = currNode->split;
                const int axis = currNode->axis();
                const KDNode *left = currNode + currNode->leftOffset();
                const KDNode *farChild;
                if (stack[enPt].p[axis] <= splitVal) {
                    if (stack[exPt].p[axis] <= splitVal) {
                        currNode = left;
                        continue;
                    }
                    if (stack[enPt].p[axis] == splitVal) {
                        currNode = left + 1;
                        continue;
                    }
                    currNode = left;
                    farChild = currNode + 1;
                } else {
                    if (splitVal < stack[exPt].p[axis]) {
                        currNode = left + 1;
                        continue;
                    }
                    farChild = left;
                    currNode = farChild + 1;
                }
                Float distToSplit = (splitVal - ray.o[axis]) * ray.dRcp[axis];
                const uint32_t tmp = exPt++;
                if (exPt == enPt) ++exPt;
                stack[exPt].prev = tmp;
                stack[exPt].t = distToSplit;
                stack[exPt].node = farChild;
                stack[exPt].p = ray(distToSplit);
                stack[exPt].p[axis] = splitVal;
            }
            if (cnt) cnt->nodes++;
            if (cnt) cnt->leaves++;
            for (uint32_t entry = currNode->primStart(), last = currNode->primEnd(); entry != last; entry++) {
                const uint32_t primIdx = kd.indices[entry];
                if (cnt) cnt->indices++;
                if (mailbox[primIdx & 7] == primIdx) continue;
                if (cnt) cnt->prims++;
                bool result = intersectPrim(ray, primIdx, mint, maxt, t, cache);
                if (result) {
                    if (shadowRay) return true;
                    maxt = t;
                    foundIntersection = true;
                }
                mailbox[primIdx & 7] = primIdx;
            }
            if (stack[exPt].t > maxt) break;
            enPt = exPt;
            currNode = stack[exPt].node;
            exPt = stack[enPt].prev;
        }
        return foundIntersection;
    }

    // skdtree.h:343-428 + rectangle.cpp:155-168
    void fillIntersectionRecord(const Ray &ray, const IntersectionCache &cache, Intersection &its) const {
        const Shape &shape = shapes[cache.shapeIndex];
        its.shape = (int)cache.shapeIndex;
        its.primIndex = cache.primIndex;
        if (shape.type == B200PG_SHAPE_TRIMESH) {
            const Vec3 b(1 - cache.u - cache.v, cache.u, cache.v);
            const uint32_t idx0 = shape.indices[3 * cache.primIndex], idx1 = shape.indices[3 * cache.primIndex + 1],
                           idx2 = shape.indices[3 * cache.primIndex + 2];
            const Vec3 &p0 = shape.positions[idx0], &p1 = shape.positions[idx1], &p2 = shape.positions[idx2];
            its.p = p0 * b.x + p1 * b.y + p2 * b.z;
            Vec3 side1(p1 - p0), side2(p2 - p0);
            Vec3 faceNormal(cross(side1, side2));
            Float len = length(faceNormal);
            if (!faceNormal.isZero()) faceNormal /= len;
            // skdtree.h:374-381: the per-triangle UV tangent when the mesh has one (TriMesh::configure computes them for every
            // mesh with texture coordinates, trimesh.cpp:383-385), the first edge otherwise
            its.dpdu = shape.uvTangents.empty() ? side1 : shape.uvTangents[cache.primIndex];
            if (!shape.normals.empty()) {
                const Vec3 &n0 = shape.normals[idx0], &n1 = shape.normals[idx1], &n2 = shape.normals[idx2];
                its.shFrame.n = normalize(n0 * b.x + n1 * b.y + n2 * b.z);
                if (dot(faceNormal, its.shFrame.n) < 0) faceNormal = -faceNormal;
            } else {
                its.shFrame.n = faceNormal;
            }
            its.geoN = faceNormal;
            if (!shape.texcoords.empty()) {
                const Vec2 &t0 = shape.texcoords[idx0], &t1 = shape.texcoords[idx1], &t2 = shape.texcoords[idx2];
                its.uv = Vec2(t0.x * b.x + t1.x * b.y + t2.x * b.z, t0.y * b.x + t1.y * b.y + t2.y * b.z);
            } else {
                its.uv = Vec2(b.y, b.z);
            }
        } else {
            its.geoN = shape.frame.n;
            its.shFrame.n = shape.frame.n;
            its.dpdu = shape.dpdu;
            its.uv = Vec2(0.5f * (cache.u + 1), 0.5f * (cache.v + 1));
            its.p = ray(its.t);
        }
        computeShadingFrame(its.shFrame.n, its.dpdu, its.shFrame);
        its.wi = its.toLocal(-ray.d);
    }

    // skdtree.cpp:112-142
    bool rayIntersect(const Ray &ray, Intersection &its, Stats *st, IntersectionCache *outCache = nullptr) const {
        IntersectionCache cache;
        its.t = std::numeric_limits<Float>::infinity();
        Float mint, maxt;
        if (st) st->normalRays++;
        if (kd.aabb.rayIntersect(ray, mint, maxt)) {
            Float rayMinT = ray.mint;
            if (rayMinT == Epsilon)
                rayMinT *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), Epsilon);
            if (rayMinT > mint) mint = rayMinT;
            if (ray.maxt < maxt) maxt = ray.maxt;
            if (maxt > mint) {
                if (rayIntersectHavran<false>(ray, mint, maxt, its.t, cache, st ? &st->trav : nullptr)) {
                    fillIntersectionRecord(ray, cache, its);
                    if (outCache) *outCache = cache;
                    return true;
                }
            }
        }
        return false;
    }

    // skdtree.cpp:207-226 (note: no inner max with Epsilon here)
    bool rayIntersectShadow(const Ray &ray, Stats *st) const {
        Float mint, maxt, t = std::numeric_limits<Float>::infinity();
        IntersectionCache cache;
        if (st) st->shadowRays++;
        if (kd.aabb.rayIntersect(ray, mint, maxt)) {
            Float rayMinT = ray.mint;
            if (rayMinT == Epsilon) rayMinT *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
            if (rayMinT > mint) mint = rayMinT;
            if (ray.maxt < maxt) maxt = ray.maxt;
            if (maxt > mint)
                if (rayIntersectHavran<true>(ray, mint, maxt, t, cache, st ? &st->trav : nullptr)) return true;
        }
        return false;
    }

    const Bsdf &bsdfOf(const Shape &s) const { return bsdfs[s.bsdf]; }

    // ---------------- emitters ----------------
    // Shape::samplePosition: rectangle.cpp:210-216, trimesh.cpp:412-423 + triangle.cpp:24-59
    void samplePosition(const Shape &s, Vec2 sample, Vec3 &p, Vec3 &n, Float &pdf) const {
        if (s.type == B200PG_SHAPE_RECTANGLE) {
            p = s.objectToWorld.point(Vec3(sample.x * 2 - 1, sample.y * 2 - 1, 0));
            n = s.frame.n;
            pdf = s.invSurfaceArea;
        } else {
            Float dummy;
            size_t index = cdfSampleReuse(s.areaCdf, sample.y, dummy);
            const Vec3 &p0 = s.positions[s.indices[3 * index]], &p1 = s.positions[s.indices[3 * index + 1]],
                       &p2 = s.positions[s.indices[3 * index + 2]];
            Vec2 bary = squareToUniformTriangle(sample);
            Vec3 sideA = p1 - p0, sideB = p2 - p0;
            p = p0 + (sideA * bary.x) + (sideB * bary.y);
            if (!s.normals.empty()) {
                const Vec3 &n0 = s.normals[s.indices[3 * index]], &n1 = s.normals[s.indices[3 * index + 1]],
                           &n2 = s.normals[s.indices[3 * index + 2]];
                n = normalize(n0 * (1.0f - bary.x - bary.y) + n1 * bary.x + n2 * bary.y);
            } else {
                n = normalize(cross(sideA, sideB));
            }
            pdf = s.invSurfaceArea;
        }
    }

    struct DirectSample {
        Vec3 ref, refN, p, n, d;
        Float dist, pdf;
        int emitter;
    };

    // Scene::sampleEmitterDirect without the visibility test (scene.cpp:871-895) +
    // AreaLight::sampleDirect (area.cpp:158-173) + Shape::sampleDirect (shape.cpp:102-116)
    Vec3 sampleEmitterDirectNoVis(DirectSample &dRec, Vec2 sample) const {
        Float emPdf;
        size_t index = cdfSampleReuse(emitterCdf, sample.x, emPdf);
        const Emitter &em = emitters[index];
        samplePosition(shapes[em.shape], sample, dRec.p, dRec.n, dRec.pdf);
        dRec.d = dRec.p - dRec.ref;
        Float distSquared = dot(dRec.d, dRec.d);
        dRec.dist = std::sqrt(distSquared);
        dRec.d /= dRec.dist;
        Float dp = absDot(dRec.d, dRec.n);
        dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
        Vec3 value;
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) {
            value = em.radiance / dRec.pdf;
        } else {
            dRec.pdf = 0.0f;
            return Vec3(0.0f);
        }
        dRec.emitter = (int)index;
        dRec.pdf *= emPdf;
        value /= emPdf;
        return value;
    }

    // Scene::pdfEmitterDirect (scene.cpp:992-995) + area.cpp:175-183 + shape.cpp:117-126
    Float pdfEmitterDirect(const DirectSample &dRec) const {
        const Emitter &em = emitters[dRec.emitter];
        Float discrete = emitterCdf[dRec.emitter + 1] - emitterCdf[dRec.emitter];
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0)
            return shapes[em.shape].invSurfaceArea * (dRec.dist * dRec.dist) / absDot(dRec.d, dRec.n) * discrete;
        return 0.0f;
    }

    // AreaLight::eval (area.cpp:104-109)
    Vec3 emitterEval(const Intersection &its, const Vec3 &d) const {
        const Emitter &em = emitters[shapes[its.shape].emitter];
        if (dot(its.shFrame.n, d) <= 0) return Vec3(0.0f);
        return em.radiance;
    }

    // perspective.cpp:271-298
    Ray sampleRay(const Vec2 &pixelSample) const {
        Vec3 nearP = camera.sampleToCamera.point(
            Vec3(pixelSample.x * camera.invResolution.x, pixelSample.y * camera.invResolution.y, 0.0f));
        Vec3 d = normalize(nearP);
        Float invZ = 1.0f / d.z;
        return Ray(camera.toWorld.pointAffine(Vec3(0.0f)), camera.toWorld.vector(d), camera.nearClip * invZ,
                   camera.farClip * invZ);
    }
};

static inline Float miWeight(Float pdfA, Float pdfB) {  // progressive_path.cpp:316-320
    pdfA *= pdfA;
    pdfB *= pdfB;
    return pdfA / (pdfA + pdfB);
}

// Guiding context of one Li call: the field to sample from (may be null) and the training-sample sink (may be
// null). The guided branches below are this repo's design (oracle_guiding.h), everything else is the reference's.
struct GuideCtx {
    const GuideField *field = nullptr;
    GuideSamples *rec = nullptr;
    Float alpha = 0.5f;  // one-sample MIS selection probability of the guiding distribution
};
struct GuideVertex {
    Vec3 pos, dir, T, Lk;
    Float pdf, dist;
};

// ProgressiveMIPathTracer::Li, progressive_path.cpp:133-314 (no environment emitter, no subsurface)
static Vec3 Li_path(const Scene &scene, const B200pgIntegratorParams &P, const Ray &r, Rng &rng, Stats &st,
                    const GuideCtx *G = nullptr) {
    Intersection its;
    Ray ray(r);
    Vec3 Li(0.0f);
    bool scattered = false;
    int depth = 1;
    bool emittedRadiance = true;  // rRec.type & EEmittedRadiance
    const GuideField *field = G ? G->field : nullptr;
    const bool record = G && G->rec;
    const Float alpha = G ? G->alpha : 0.0f;
    GuideVertex verts[64];
    int nVerts = 0;

    scene.rayIntersect(ray, its, &st);
    ray.mint = Epsilon;

    Vec3 throughput(1.0f);
    Float eta = 1.0f;
    const int maxDepth = P.max_depth;

    while (depth <= maxDepth || maxDepth < 0) {
        if (!its.isValid()) break;
        const Shape &shape = scene.shapes[its.shape];
        const Bsdf &bsdf = scene.bsdfOf(shape);

        if (shape.emitter >= 0 && emittedRadiance && (!P.hide_emitters || scattered))
            Li += throughput * scene.emitterEval(its, -ray.d);

        if ((depth >= maxDepth && maxDepth > 0) ||
            (P.strict_normals && dot(ray.d, its.geoN) * Frame::cosTheta(its.wi) >= 0))
            break;

        /* Direct illumination sampling */
        Scene::DirectSample dRec;
        dRec.ref = its.p;
        dRec.refN = Vec3(0.0f);
        const unsigned btype = bsdf.typeFlags();
        if ((btype & (ETransmission | EBackSide)) == 0) dRec.refN = its.shFrame.n;  // records.inl:160-164
        // guided vertex: smooth BSDF (delta lobes are never guided, cf. getDeltaSamplingRate bsdf.h:374-382)
        const bool guided = field && (btype & ESmooth);
        const uint32_t gcell = guided ? field->lookup(its.p) : 0u;

        if (P.use_nee && (btype & ESmooth)) {
            Vec3 value = scene.sampleEmitterDirectNoVis(dRec, rng.next2D());
            if (dRec.pdf != 0) {
                // visibility test of Scene::sampleEmitterDirect (scene.cpp:882-886)
                Ray shadow(dRec.ref, dRec.d, Epsilon, dRec.dist * (1 - ShadowEpsilon));
                if (scene.rayIntersectShadow(shadow, &st)) value = Vec3(0.0f);
            } else {
                value = Vec3(0.0f);
            }
            if (!value.isZero()) {
                Vec3 wo = its.toLocal(dRec.d);
                const Vec3 bsdfVal = bsdf.eval(its.wi, wo);
                if (!bsdfVal.isZero() && (!P.strict_normals || dot(its.geoN, dRec.d) * Frame::cosTheta(wo) > 0)) {
                    Float bsdfPdf = bsdf.pdf(its.wi, wo);
                    if (guided)  // the direction-sampling technique is the one-sample-MIS mixture
                        bsdfPdf = alpha * field->pdf(gcell, dRec.d) + (1 - alpha) * bsdfPdf;
                    Float weight = miWeight(dRec.pdf, bsdfPdf);
                    Li += throughput * value * bsdfVal * weight;
                }
            }
        }
        const Vec3 LafterNee = Li;

        /* BSDF sampling (one-sample MIS between the BSDF and the guiding mixture at guided vertices) */
        Float bsdfPdf, bEta;
        unsigned sampledType;
        Vec3 woLocal, bsdfWeight, wo;
        if (guided) {
            Float u0 = rng.next1D();
            Vec2 u12 = rng.next2D();
            Vec3 fcos;
            Float pb;
            if (u0 < alpha) {
                u0 /= alpha;
                wo = field->sample(gcell, u0, u12.x, u12.y);
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
        scattered |= sampledType != ENull;

        Float woDotGeoN = dot(its.geoN, wo);
        if (P.strict_normals && woDotGeoN * Frame::cosTheta(woLocal) <= 0) break;

        bool hitEmitter = false;
        Vec3 value;
        ray = Ray(its.p, wo);
        const bool recordVertex = record && (btype & ESmooth) && nVerts < 64;
        if (recordVertex) {
            GuideVertex &gv = verts[nVerts++];
            gv.pos = its.p;
            gv.dir = wo;
            gv.pdf = bsdfPdf;
            gv.T = throughput * bsdfWeight;
            gv.Lk = LafterNee;
            gv.dist = 0.0f;
        }
        const bool hitSomething = scene.rayIntersect(ray, its, &st);
        if (recordVertex) verts[nVerts - 1].dist = hitSomething ? its.t : 0.0f;
        if (hitSomething) {
            if (scene.shapes[its.shape].emitter >= 0) {
                value = scene.emitterEval(its, -ray.d);
                // dRec.setQuery(ray, its), records.inl:170-178
                dRec.p = its.p;
                dRec.n = its.shFrame.n;
                dRec.emitter = scene.shapes[its.shape].emitter;
                dRec.d = ray.d;
                dRec.dist = its.t;
                hitEmitter = true;
            }
        } else {
            break;  // no environment emitter
        }

        throughput *= bsdfWeight;
        eta *= bEta;

        if (hitEmitter) {
            const Float lumPdf = (P.use_nee && !(sampledType & EDelta)) ? scene.pdfEmitterDirect(dRec) : 0;
            const Float weight = P.use_nee ? miWeight(bsdfPdf, lumPdf) : 1.0f;
            Li += throughput * value * weight;
        }

        emittedRadiance = false;  // rRec.type = ERadianceNoEmission

        if (depth++ >= P.rr_depth) {
            Float q = std::min(throughput.maxc() * eta * eta, 0.95f);
            if (rng.next1D() >= q) break;
            throughput /= q;
        }
    }
    st.paths++;
    st.pathLen += depth;
    if (record) {
        // incident-radiance estimate along each sampled direction: everything the path gathered after the
        // vertex, divided by the throughput right after the vertex; sample weight = avg_rgb(estimate) / pdf
        for (int v = 0; v < nVerts; ++v) {
            const GuideVertex &gv = verts[v];
            Vec3 d = Li - gv.Lk, est(0.0f);
            for (int c = 0; c < 3; ++c) est[c] = gv.T[c] > 0 ? d[c] / gv.T[c] : 0.0f;
            Float w = est.average() / gv.pdf;
            if (!std::isfinite(w) || w < 0) w = 0.0f;
            G->rec->push(gv.pos, gv.dir, w, gv.pdf, gv.dist);
        }
    }
    return Li;
}

}  // namespace orc

#include "oracle_volpath.h"

namespace orc {

static Vec3 Li(const Scene &scene, const B200pgIntegratorParams &P, const Ray &r, Rng &rng, Stats &st,
               const GuideCtx *G = nullptr) {
    if (P.volumetric) return Li_volpath(scene, P, r, rng, st, G);
    return Li_path(scene, P, r, rng, st, G);
}

// ---------------------------------------------------------------------------
// Scene assembly from the flat description
// ---------------------------------------------------------------------------
static Scene *buildScene(const B200pgSceneDesc *desc) {
    std::unique_ptr<Scene> sc(new Scene());
    sc->sampleCount = desc->sample_count;
    sc->seed = desc->seed;
    for (int i = 0; i < desc->n_bsdfs; ++i) {
        Bsdf b;
        b.d = desc->bsdfs[i];
        b.configure();
        sc->bsdfs.push_back(b);
    }
    auto addDefault = [&](int type, float refl) {
        Bsdf b;
        std::memset(&b.d, 0, sizeof(b.d));
        b.d.type = type;
        for (int c = 0; c < 3; ++c) b.d.reflectance[c] = refl;
        b.d.int_ior = 1.5046f;
        b.d.ext_ior = 1.000277f;
        b.configure();
        sc->bsdfs.push_back(b);
        return (int)sc->bsdfs.size() - 1;
    };
    for (int i = 0; i < desc->n_media; ++i) sc->media.push_back(Medium::fromDesc(desc->media[i]));

    uint32_t primCount = 0;
    for (int i = 0; i < desc->n_shapes; ++i) {
        const B200pgShape &d = desc->shapes[i];
        Shape s;
        s.type = d.type;
        s.bsdf = d.bsdf;
        s.emitter = d.emitter;
        s.interiorMedium = d.interior_medium;
        s.exteriorMedium = d.exterior_medium;
        s.primOffset = primCount;
        if (s.bsdf < 0) {  // Shape::configure, shape.cpp:48-70
            if (s.emitter >= 0) {
                if (sc->defaultBlack < 0) sc->defaultBlack = addDefault(B200PG_BSDF_DIFFUSE, 0.0f);
                s.bsdf = sc->defaultBlack;
            } else if (!s.isMediumTransition()) {
                if (sc->defaultDiffuseHalf < 0) sc->defaultDiffuseHalf = addDefault(B200PG_BSDF_DIFFUSE, 0.5f);
                s.bsdf = sc->defaultDiffuseHalf;
            } else {
                if (sc->defaultNull < 0) sc->defaultNull = addDefault(B200PG_BSDF_NULL, 0.0f);
                s.bsdf = sc->defaultNull;
            }
        }
        if (d.type == B200PG_SHAPE_RECTANGLE) {
            s.objectToWorld = Mat4::fromArray(d.to_world);
            if (!s.objectToWorld.invert(s.worldToObject)) return nullptr;
            // rectangle.cpp:98-107
            s.dpdu = s.objectToWorld.vector(Vec3(2, 0, 0));
            s.dpdv = s.objectToWorld.vector(Vec3(0, 2, 0));
            Vec3 normal = normalize(s.worldToObject.normalFromInverse(Vec3(0, 0, 1)));
            s.frame = Frame(normalize(s.dpdu), normalize(s.dpdv), normal);
            s.invSurfaceArea = 1.0f / (length(s.dpdu) * length(s.dpdv));
            TriAccel ta;
            std::memset(&ta, 0, sizeof(ta));
            ta.shapeIndex = i;
            ta.k = KNoTriangleFlag;
            sc->triAccel.push_back(ta);
            AABB box;  // rectangle.cpp:109-116
            box.expandBy(s.objectToWorld.point(Vec3(-1, -1, 0)));
            box.expandBy(s.objectToWorld.point(Vec3(1, -1, 0)));
            box.expandBy(s.objectToWorld.point(Vec3(1, 1, 0)));
            box.expandBy(s.objectToWorld.point(Vec3(-1, 1, 0)));
            sc->primBoxes.push_back(box);
            primCount += 1;
        } else {
            s.positions.resize(d.n_vertices);
            for (uint32_t v = 0; v < d.n_vertices; ++v)
                s.positions[v] = Vec3(d.positions[3 * v], d.positions[3 * v + 1], d.positions[3 * v + 2]);
            if (d.normals) {
                s.normals.resize(d.n_vertices);
                for (uint32_t v = 0; v < d.n_vertices; ++v)
                    s.normals[v] = Vec3(d.normals[3 * v], d.normals[3 * v + 1], d.normals[3 * v + 2]);
            }
            if (d.texcoords) {
                s.texcoords.resize(d.n_vertices);
                for (uint32_t v = 0; v < d.n_vertices; ++v) s.texcoords[v] = Vec2(d.texcoords[2 * v], d.texcoords[2 * v + 1]);
            }
            s.indices.assign(d.indices, d.indices + 3 * (size_t)d.n_triangles);
            s.areaCdf.assign(1, 0.0f);  // trimesh.cpp:395-402
            for (uint32_t t = 0; t < d.n_triangles; ++t) {
                const Vec3 &p0 = s.positions[s.indices[3 * t]], &p1 = s.positions[s.indices[3 * t + 1]],
                           &p2 = s.positions[s.indices[3 * t + 2]];
                TriAccel ta;
                std::memset(&ta, 0, sizeof(ta));
                ta.load(p0, p1, p2);  // degenerate -> k = 3, never hit (skdtree.cpp:80-96)
                ta.shapeIndex = i;
                ta.primIndex = t;
                sc->triAccel.push_back(ta);
                AABB box;
                box.expandBy(p0);
                box.expandBy(p1);
                box.expandBy(p2);
                sc->primBoxes.push_back(box);
                Float area = 0.5f * length(cross(p1 - p0, p2 - p0));  // triangle.cpp surfaceArea
                s.areaCdf.push_back(s.areaCdf.back() + area);
            }
            Float total = cdfNormalize(s.areaCdf);
            s.invSurfaceArea = 1.0f / total;
            if (!s.texcoords.empty()) {  // TriMesh::computeUVTangents, trimesh.cpp:683-735
                s.uvTangents.assign(d.n_triangles, Vec3(0.0f));
                for (uint32_t t = 0; t < d.n_triangles; ++t) {
                    const uint32_t i0 = s.indices[3 * t], i1 = s.indices[3 * t + 1], i2 = s.indices[3 * t + 2];
                    const Vec3 dP1 = s.positions[i1] - s.positions[i0], dP2 = s.positions[i2] - s.positions[i0];
                    const Vec2 dUV1(s.texcoords[i1].x - s.texcoords[i0].x, s.texcoords[i1].y - s.texcoords[i0].y);
                    const Vec2 dUV2(s.texcoords[i2].x - s.texcoords[i0].x, s.texcoords[i2].y - s.texcoords[i0].y);
                    const Vec3 n = cross(dP1, dP2);
                    const Float len = length(n);
                    if (len == 0) continue;  // degenerate triangle: the zero tangent stays (never hit)
                    const Float determinant = dUV1.x * dUV2.y - dUV1.y * dUV2.x;
                    if (determinant == 0) {  // degenerate parameterisation: any tangent perpendicular to the face normal
                        Vec3 a, b;
                        coordinateSystem(n / len, a, b);
                        s.uvTangents[t] = a;
                    } else {
                        const Float invDet = 1.0f / determinant;
                        s.uvTangents[t] = (dP1 * dUV2.y - dP2 * dUV1.y) * invDet;
                    }
                }
            }
            primCount += d.n_triangles;
        }
        sc->shapes.push_back(std::move(s));
    }
    sc->emitterCdf.assign(1, 0.0f);
    for (int i = 0; i < desc->n_emitters; ++i) {
        Emitter e;
        e.radiance = Vec3(desc->emitters[i].radiance[0], desc->emitters[i].radiance[1], desc->emitters[i].radiance[2]);
        e.samplingWeight = desc->emitters[i].sampling_weight;
        e.shape = desc->emitters[i].shape;
        sc->emitters.push_back(e);
        sc->emitterCdf.push_back(sc->emitterCdf.back() + e.samplingWeight);  // scene.cpp:419-423
    }
    if (desc->n_emitters > 0) cdfNormalize(sc->emitterCdf);

    // clipping is on by default (m_clip, gkdtree.h:739): triangles through Triangle::getClippedAABB, other shapes through
    // Shape::getClippedAABB = bounding box clipped to the cell (shape.cpp)
    Scene *scp = sc.get();
    sc->kd.build(sc->primBoxes, [scp](uint32_t prim, const AABB &box) {
        const TriAccel &ta = scp->triAccel[prim];
        if (ta.k == KNoTriangleFlag) {
            AABB r = scp->primBoxes[prim];
            for (int j = 0; j < 3; ++j) {
                r.min[j] = std::max(r.min[j], box.min[j]);
                r.max[j] = std::min(r.max[j], box.max[j]);
            }
            return r;
        }
        const Shape &s = scp->shapes[ta.shapeIndex];
        return clippedTriangleAABB(s.positions[s.indices[3 * ta.primIndex]], s.positions[s.indices[3 * ta.primIndex + 1]],
                                   s.positions[s.indices[3 * ta.primIndex + 2]], box);
    });

    // camera, perspective.cpp:126-155 (no crop window)
    const B200pgSensor &sd = desc->sensor;
    sc->film.width = desc->film.width;
    sc->film.height = desc->film.height;
    sc->film.configure(desc->film.filter_stddev);
    Float aspect = (Float)desc->film.width / (Float)desc->film.height;
    Float xfov = sd.fov;
    int axis = sd.fov_axis;
    if (axis == 3) axis = aspect > 1 ? 1 : 0;  // smaller
    if (axis == 4) axis = aspect > 1 ? 0 : 1;  // larger
    if (axis == 1) {  // setYFov, sensor.cpp
        xfov = 2.0f * std::atan(std::tan(0.5f * sd.fov * PI_F / 180.0f) * aspect) * 180.0f / PI_F;
    } else if (axis == 2) {  // setDiagonalFov
        Float diagonal = 2 * std::tan(0.5f * sd.fov * PI_F / 180.0f);
        Float width = diagonal / std::sqrt(1.0f + 1.0f / (aspect * aspect));
        xfov = 2.0f * std::atan(width * 0.5f) * 180.0f / PI_F;
    }
    Float recip = 1.0f / (sd.far_clip - sd.near_clip);
    Float cot = 1.0f / std::tan((xfov / 2.0f) * PI_F / 180.0f);
    Mat4 persp = Mat4::identity();  // transform.cpp:99-123
    persp.m[0][0] = cot; persp.m[1][1] = cot;
    persp.m[2][2] = sd.far_clip * recip; persp.m[2][3] = -sd.near_clip * sd.far_clip * recip;
    persp.m[3][2] = 1; persp.m[3][3] = 0;
    Mat4 tr = Mat4::identity();
    tr.m[0][3] = -1.0f; tr.m[1][3] = -1.0f / aspect;
    Mat4 scl = Mat4::identity();
    scl.m[0][0] = -0.5f; scl.m[1][1] = -0.5f * aspect;
    Mat4 cameraToSample = scl * tr * persp;
    if (!cameraToSample.invert(sc->camera.sampleToCamera)) return nullptr;
    sc->camera.toWorld = Mat4::fromArray(sd.to_world);
    sc->camera.nearClip = sd.near_clip;
    sc->camera.farClip = sd.far_clip;
    sc->camera.invResolution = Vec2(1.0f / desc->film.width, 1.0f / desc->film.height);
    sc->camera.medium = sd.medium;
    return sc.release();
}

}  // namespace orc

// ---------------------------------------------------------------------------
// C API (ctypes)
// ---------------------------------------------------------------------------
using namespace orc;

extern "C" {

void *orc_scene_create(const B200pgSceneDesc *desc) { return buildScene(desc); }
void orc_scene_destroy(void *s) { delete (Scene *)s; }

// Depth-first dump of the kd-tree (test hook, same layout as the reference harness' ref_kd_dump): per node
// {axis or -1 for a leaf, split or primitive count, depth}
int orc_kd_dump(void *s, float *out, int max_nodes) {
    Scene *sc = (Scene *)s;
    std::vector<std::pair<const KDNode *, int>> stack;
    stack.push_back(std::make_pair(sc->kd.nodes.data(), 0));
    int n = 0;
    while (!stack.empty() && n < max_nodes) {
        const KDNode *node = stack.back().first;
        int depth = stack.back().second;
        stack.pop_back();
        if (node->isLeaf()) {
            out[3 * n] = -1; out[3 * n + 1] = (float)(node->primEnd() - node->primStart()); out[3 * n + 2] = (float)depth;
        } else {
            const KDNode *left = node + node->leftOffset();
            out[3 * n] = (float)node->axis(); out[3 * n + 1] = node->split; out[3 * n + 2] = (float)depth;
            stack.push_back(std::make_pair(left + 1, depth + 1));
            stack.push_back(std::make_pair(left, depth + 1));
        }
        ++n;
    }
    return n;
}

int orc_kd_info(void *s, uint64_t *out /* nodes, indices, prims */) {
    Scene *sc = (Scene *)s;
    out[0] = sc->kd.nodes.size();
    out[1] = sc->kd.indices.size();
    out[2] = sc->triAccel.size();
    return 0;
}

// rays n*8 (o, mint, d, maxt); out tuv n*3, prim n (global prim id), counters[4] summed (may be NULL)
int orc_trace(void *s, const float *rays, size_t n, int shadow, float *tuv, uint32_t *prim, uint64_t *counters,
              int nthreads) {
    Scene *sc = (Scene *)s;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4096) reduction(+ : c0, c1, c2, c3)
    for (long long i = 0; i < (long long)n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[4], r[5], r[6]), r[3], r[7]);
        Stats st;
        if (shadow) {
            bool hit = sc->rayIntersectShadow(ray, &st);
            prim[i] = hit ? 0u : 0xFFFFFFFFu;
            if (tuv) tuv[3 * i] = tuv[3 * i + 1] = tuv[3 * i + 2] = 0;
        } else {
            Intersection its;
            IntersectionCache cache;
            if (sc->rayIntersect(ray, its, &st, &cache)) {
                tuv[3 * i] = its.t;
                tuv[3 * i + 1] = cache.u;
                tuv[3 * i + 2] = cache.v;
                prim[i] = sc->shapes[cache.shapeIndex].primOffset +
                          (cache.primIndex == KNoTriangleFlag ? 0u : cache.primIndex);
            } else {
                tuv[3 * i] = std::numeric_limits<float>::infinity();
                tuv[3 * i + 1] = tuv[3 * i + 2] = 0;
                prim[i] = 0xFFFFFFFFu;
            }
        }
        c0 += st.trav.nodes;
        c1 += st.trav.indices;
        c2 += st.trav.prims;
        c3 += st.trav.leaves;
    }
    if (counters) {  // 4 words: nodes visited (inner + leaf), index entries read, primitive tests, leaf visits
        counters[0] = c0;
        counters[1] = c1;
        counters[2] = c2;
        counters[3] = c3;
    }
    return 0;
}

// Brute-force closest hit over all primitives (cross-check for the kd-tree; test_kd-style)
int orc_trace_bruteforce(void *s, const float *rays, size_t n, float *tuv, uint32_t *prim) {
    Scene *sc = (Scene *)s;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[4], r[5], r[6]), r[3], r[7]);
        Float mint = ray.mint, maxt = ray.maxt;
        if (mint == Epsilon)
            mint *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), Epsilon);
        Float t = std::numeric_limits<Float>::infinity();
        IntersectionCache cache{}, best{};
        bool found = false;
        for (uint32_t p = 0; p < sc->triAccel.size(); ++p) {
            Float tt;
            if (sc->intersectPrim(ray, p, mint, maxt, tt, cache)) {
                if (tt < t || !found) {
                    t = tt;
                    maxt = tt;
                    best = cache;
                    found = true;
                }
            }
        }
        if (found) {
            tuv[3 * i] = t;
            tuv[3 * i + 1] = best.u;
            tuv[3 * i + 2] = best.v;
            prim[i] = sc->shapes[best.shapeIndex].primOffset + (best.primIndex == KNoTriangleFlag ? 0u : best.primIndex);
        } else {
            tuv[3 * i] = std::numeric_limits<float>::infinity();
            tuv[3 * i + 1] = tuv[3 * i + 2] = 0;
            prim[i] = 0xFFFFFFFFu;
        }
    }
    return 0;
}

int orc_camera_rays(void *s, const float *pos, size_t n, float *rays) {
    Scene *sc = (Scene *)s;
    for (size_t i = 0; i < n; ++i) {
        Ray r = sc->sampleRay(Vec2(pos[2 * i], pos[2 * i + 1]));
        float *o = rays + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.mint;
        o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.maxt;
    }
    return 0;
}

// Next-event estimation on its own, the way src/tests/test_chisquare.cpp test03_EmitterDirect drives an emitter
// (EmitterAdapter: sampleDirect for the samples, pdfDirect for the density). Reference point ref with normal refN
// (refN = 0: no facing test, records.inl:160-164).
//   u != NULL: out_d[i], out_dist[i], out_pdf[i], out_value[i] = Scene::sampleEmitterDirect(u[i]) WITHOUT the visibility test
//   u == NULL: out_pdf[i] = Scene::pdfEmitterDirect for the given direction d[i]: the ray (ref, d) is traced, and if the
//              first surface it meets is an emitter the query record is filled from the hit (records.inl:170-178), as
//              the path tracer does for BSDF-sampled rays (progressive_path.cpp:243-262); 0 otherwise.
int orc_emitter_direct(void *s, const float *ref, const float *refN, const float *u, float *d, size_t n, float *out_dist,
                       float *out_pdf, float *out_value) {
    Scene *sc = (Scene *)s;
    for (size_t i = 0; i < n; ++i) {
        Scene::DirectSample dRec;
        dRec.ref = Vec3(ref[0], ref[1], ref[2]);
        dRec.refN = Vec3(refN[0], refN[1], refN[2]);
        if (u) {
            Vec3 value = sc->sampleEmitterDirectNoVis(dRec, Vec2(u[2 * i], u[2 * i + 1]));
            d[3 * i] = dRec.d.x; d[3 * i + 1] = dRec.d.y; d[3 * i + 2] = dRec.d.z;
            if (out_dist) out_dist[i] = dRec.dist;
            if (out_pdf) out_pdf[i] = dRec.pdf;
            if (out_value) { out_value[3 * i] = value.x; out_value[3 * i + 1] = value.y; out_value[3 * i + 2] = value.z; }
        } else {
            Ray ray(dRec.ref, Vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
            Intersection its;
            Stats st;
            Float pdf = 0.0f;
            if (sc->rayIntersect(ray, its, &st) && its.isValid() && sc->shapes[its.shape].emitter >= 0) {
                dRec.p = its.p;
                dRec.n = its.shFrame.n;
                dRec.emitter = sc->shapes[its.shape].emitter;
                dRec.d = ray.d;
                dRec.dist = its.t;
                pdf = sc->pdfEmitterDirect(dRec);
                if (out_dist) out_dist[i] = its.t;
            } else if (out_dist) {
                out_dist[i] = std::numeric_limits<float>::infinity();
            }
            out_pdf[i] = pdf;
        }
    }
    return 0;
}

// Test hook: 1 = the emitter look-up of the volumetric path uses the total distance from the path vertex (see the REFERENCE
// QUIRK note in oracle_volpath.h); 0 (default) = the reference's behaviour. Returns the previous value.
int orc_debug_lookup_total_distance(int on) {
    const int before = g_lookupTotalDistance;
    g_lookupTotalDistance = on ? 1 : 0;
    return before;
}

// Full intersection records (ShapeKDTree::rayIntersect + fillIntersectionRecord, skdtree.h:343-428), the quantities the
// reference's src/tests/test_dgeom.cpp asserts: out[18 * i] = {t, p.xyz, uv.xy, geoFrame.n, shFrame.n, shFrame.s, dpdu};
// t = inf for a miss.
int orc_intersect(void *s, const float *rays, size_t n, float *out) {
    Scene *sc = (Scene *)s;
    for (size_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[4], r[5], r[6]), r[3], r[7]);
        Intersection its;
        Stats st;
        float *o = out + 18 * i;
        for (int k = 0; k < 18; ++k) o[k] = 0.0f;
        if (!sc->rayIntersect(ray, its, &st) || !its.isValid()) {
            o[0] = std::numeric_limits<float>::infinity();
            continue;
        }
        o[0] = its.t;
        o[1] = its.p.x; o[2] = its.p.y; o[3] = its.p.z;
        o[4] = its.uv.x; o[5] = its.uv.y;
        o[6] = its.geoN.x; o[7] = its.geoN.y; o[8] = its.geoN.z;
        o[9] = its.shFrame.n.x; o[10] = its.shFrame.n.y; o[11] = its.shFrame.n.z;
        o[12] = its.shFrame.s.x; o[13] = its.shFrame.s.y; o[14] = its.shFrame.s.z;
        o[15] = its.dpdu.x; o[16] = its.dpdu.y; o[17] = its.dpdu.z;
    }
    return 0;
}

// MicrofacetDistribution on its own (src/tests/test_microfacet.cpp drives the class directly): for a fixed incident
// direction wi, out_m[i] = sampleVisible(wi, u[i]) when u != NULL; for every normal m[i] (the sampled ones, or the given
// ones when u == NULL): pdfVisible(wi, m), D(m) = eval(m) and smithG1(wi, m).
int orc_microfacet(int type, float alphaU, float alphaV, const float *wi, const float *u, float *m, size_t n, float *out_pdf,
                   float *out_D, float *out_G1) {
    MicrofacetDistribution distr(type, alphaU, alphaV);
    const Vec3 vi(wi[0], wi[1], wi[2]);
    for (size_t i = 0; i < n; ++i) {
        Vec3 mm;
        if (u) {
            mm = distr.sampleVisible(vi, Vec2(u[2 * i], u[2 * i + 1]));
            m[3 * i] = mm.x; m[3 * i + 1] = mm.y; m[3 * i + 2] = mm.z;
        } else {
            mm = Vec3(m[3 * i], m[3 * i + 1], m[3 * i + 2]);
        }
        if (out_pdf) out_pdf[i] = distr.pdfVisible(vi, mm);
        if (out_D) out_D[i] = distr.eval(mm);
        if (out_G1) out_G1[i] = distr.smithG1(vi, mm);
    }
    return 0;
}

int orc_bsdf(void *s, int bsdfIndex, const float *wi, const float *wo, const float *u, size_t n, float *out_eval,
             float *out_pdf, float *out_wo, float *out_weight, float *out_spdf, uint32_t *out_flags) {
    Scene *sc = (Scene *)s;
    if (bsdfIndex < 0 || bsdfIndex >= (int)sc->bsdfs.size()) return -1;
    const Bsdf &b = sc->bsdfs[bsdfIndex];
    for (size_t i = 0; i < n; ++i) {
        Vec3 vi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), vo(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        Vec3 e = b.eval(vi, vo);
        out_eval[3 * i] = e.x; out_eval[3 * i + 1] = e.y; out_eval[3 * i + 2] = e.z;
        out_pdf[i] = b.pdf(vi, vo);
        Vec3 so(0.0f);
        Float spdf = 0, eta;
        unsigned st = 0;
        Vec3 w = b.sample(vi, Vec2(u[2 * i], u[2 * i + 1]), so, spdf, eta, st);
        if (w.isZero()) { so = Vec3(0.0f); spdf = 0; }
        out_wo[3 * i] = so.x; out_wo[3 * i + 1] = so.y; out_wo[3 * i + 2] = so.z;
        out_weight[3 * i] = w.x; out_weight[3 * i + 1] = w.y; out_weight[3 * i + 2] = w.z;
        out_spdf[i] = spdf;
        out_flags[i] = st;
    }
    return 0;
}

// field (may be NULL): guiding field to sample from; sink (may be NULL): receives the training samples of all
// n paths in index order (deterministic).
int orc_radiance(void *s, const B200pgIntegratorParams *P, const uint32_t *pixel, const uint32_t *sample, size_t n,
                 float *out_rgb, void *field, void *sink) {
    Scene *sc = (Scene *)s;
    const int nChunks = (int)((n + 255) / 256);
    std::vector<GuideSamples> chunkSamples(sink ? nChunks : 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int ch = 0; ch < nChunks; ++ch)
    for (long long i = (long long)ch * 256; i < std::min<long long>((long long)n, (long long)(ch + 1) * 256); ++i) {
        GuideCtx G;
        G.field = (const GuideField *)field;
        G.rec = sink ? &chunkSamples[ch] : nullptr;
        G.alpha = P->guiding_probability;
        Rng rng;
        rng.init(sc->seed, pixel[i], sample[i]);
        int px = pixel[i] % sc->film.width, py = pixel[i] / sc->film.width;
        Vec2 off = rng.next2D();
        Vec2 samplePos(px + off.x, py + off.y);
        Ray ray = sc->sampleRay(samplePos);
        Stats st;
        Vec3 L = Li(*sc, *P, ray, rng, st, (field || sink) ? &G : nullptr);
        out_rgb[3 * i] = L.x; out_rgb[3 * i + 1] = L.y; out_rgb[3 * i + 2] = L.z;
    }
    if (sink) {
        GuideSamples *out = (GuideSamples *)sink;
        for (auto &c : chunkSamples)
            for (size_t j = 0; j < c.size(); ++j) out->push(c.pos[j], c.dir[j], c.weight[j], c.pdf[j], c.dist[j]);
    }
    return 0;
}

// Denoiser feature buffers (src/librender/denoiser.cpp:138-144, Denoiser::add): per-pixel RUNNING MEANS of the sample colour,
// the albedo and the normal over the samples of the pixel (box: every sample counts for the pixel its position falls in).
// The reference never fills Sample::albedo / normal (nothing in the tree calls the denoiser), so the definitions are
// this repo's (DESIGN.md "Feature buffers"): first intersection of the camera ray; albedo = diffuse reflectance
// (diffuse, roughplastic), specular reflectance (roughconductor), 1 (dielectric, null); normal = shading normal
// (world space); both 0 when the ray leaves the scene. out: H*W*10 = {color.rgb, albedo.rgb, normal.xyz, sample count}.
static Vec3 featureAlbedo(const Bsdf &b) {
    switch (b.d.type) {
        case B200PG_BSDF_DIFFUSE:
        case B200PG_BSDF_ROUGHPLASTIC: return b.R();
        case B200PG_BSDF_ROUGHCONDUCTOR: return b.SR();
        default: return Vec3(1.0f);
    }
}
int orc_features(void *s, const B200pgIntegratorParams *P, int first_sample, int n_samples, float *out) {
    Scene *sc = (Scene *)s;
    const int W = sc->film.width, H = sc->film.height;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long pix = 0; pix < (long long)W * H; ++pix) {
        float *o = out + 10 * pix;
        const int x = (int)(pix % W), y = (int)(pix / W);
        for (int j = 0; j < n_samples; ++j) {
            Rng rng;
            rng.init(sc->seed, (uint32_t)pix, (uint32_t)(first_sample + j));
            Vec2 off = rng.next2D();
            Vec2 samplePos(x + off.x, y + off.y);
            Ray ray = sc->sampleRay(samplePos);
            Stats st;
            Intersection its;
            Vec3 albedo(0.0f), normal(0.0f);
            if (sc->rayIntersect(ray, its, &st) && its.isValid()) {
                albedo = featureAlbedo(sc->bsdfOf(sc->shapes[its.shape]));
                normal = its.shFrame.n;
            }
            Vec3 L = Li(*sc, *P, ray, rng, st, nullptr);
            float maxSpec = L.maxc();
            if (maxSpec > P->max_component_value) L *= P->max_component_value / maxSpec;
            o[9] += 1.0f;
            const float a = 1.0f / o[9];  // Denoiser::add
            const float v[9] = {L.x, L.y, L.z, albedo.x, albedo.y, albedo.z, normal.x, normal.y, normal.z};
            for (int k = 0; k < 9; ++k) o[k] = (1.0f - a) * o[k] + a * v[k];
        }
    }
    return 0;
}

// GridDataSource::lookupFloat on a batch of points (gridvolume.cpp:337-388)
int orc_grid_lookup(void *s, int medium, const float *p, size_t n, float *out) {
    Scene *sc = (Scene *)s;
    if (medium < 0 || medium >= (int)sc->media.size()) return -1;
    for (size_t i = 0; i < n; ++i) out[i] = sc->media[medium].lookup(Vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    return 0;
}

// sampleDistance, evalTransmittance and one phase sample per ray; ray i draws from the stream (seed, pixel = i, sample = 0)
int orc_medium_sample(void *s, int medium, const float *rays, size_t n, float *out_t, float *out_tr, float *out_wo, float *out_pdf) {
    Scene *sc = (Scene *)s;
    if (medium < 0 || medium >= (int)sc->media.size()) return -1;
    const Medium &M = sc->media[medium];
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        const float *r = rays + 8 * i;
        Vec3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        Rng rng;
        rng.init(sc->seed, (uint32_t)i, 0);
        MediumSample mRec;
        bool ok = M.sampleDistance(o, d, r[3], r[7], mRec, rng);
        out_t[i] = ok ? mRec.t : std::numeric_limits<Float>::infinity();
        out_tr[i] = M.evalTransmittance(o, d, r[3], r[7], rng);
        Float pdf;
        Vec3 wo = M.phaseSample(-d, rng.next2D(), pdf);
        out_wo[3 * i] = wo.x; out_wo[3 * i + 1] = wo.y; out_wo[3 * i + 2] = wo.z;
        out_pdf[i] = pdf;
    }
    return 0;
}

// phase function eval for (wi, wo) pairs and sample for (wi, u) (hg.cpp:74-110, isotropic.cpp:62-78)
int orc_phase(void *s, int medium, const float *wi, const float *wo, const float *u, size_t n, float *out_eval, float *out_wo,
              float *out_pdf) {
    Scene *sc = (Scene *)s;
    if (medium < 0 || medium >= (int)sc->media.size()) return -1;
    const Medium &M = sc->media[medium];
    for (size_t i = 0; i < n; ++i) {
        Vec3 vi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), vo(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        out_eval[i] = M.phaseEval(vi, vo);
        Float pdf;
        Vec3 w = M.phaseSample(vi, Vec2(u[2 * i], u[2 * i + 1]), pdf);
        out_wo[3 * i] = w.x; out_wo[3 * i + 1] = w.y; out_wo[3 * i + 2] = w.z;
        out_pdf[i] = pdf;
    }
    return 0;
}

// Triangle::getClippedAABB on one triangle (tri = 9 floats, box = min xyz, max xyz; out = min xyz, max xyz, inverted when
// the triangle is clipped away): the known-answer vectors of src/tests/test_kd.cpp:34-83
int orc_clipped_aabb(const float *tri, const float *box, float *out) {
    AABB b;
    b.min = Vec3(box[0], box[1], box[2]);
    b.max = Vec3(box[3], box[4], box[5]);
    AABB r = clippedTriangleAABB(Vec3(tri[0], tri[1], tri[2]), Vec3(tri[3], tri[4], tri[5]), Vec3(tri[6], tri[7], tri[8]), b);
    for (int j = 0; j < 3; ++j) {
        out[j] = r.min[j];
        out[3 + j] = r.max[j];
    }
    return 0;
}

int orc_film_splat(void *s, const float *pos, const float *rgb, size_t n, float *film /* H*W*5 */) {
    Scene *sc = (Scene *)s;
    ImageBlock blk;
    blk.init(0, 0, sc->film.width, sc->film.height, sc->film.borderSize);
    for (size_t i = 0; i < n; ++i) {
        Float v[5] = {rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 1.0f, 1.0f};
        blk.put(sc->film, Vec2(pos[2 * i], pos[2 * i + 1]), v);
    }
    const int b = blk.border, sx = blk.w + 2 * b;
    for (int y = 0; y < blk.h; ++y)
        for (int x = 0; x < blk.w; ++x)
            for (int k = 0; k < 5; ++k)
                film[((size_t)y * blk.w + x) * 5 + k] += blk.data[((size_t)(y + b) * sx + (x + b)) * 5 + k];
    return 0;
}

// One or more progressions over the whole image (progressiveintegrator.cpp:65-114, 222-282):
// 32x32 tiles handed to threads (imageproc.cpp:27-78), per-tile ImageBlock merged into the film
// under a mutex (renderproc.cpp:141-148). film: H*W*5 accumulators (+=). stats: 7 u64
// (paths, normal rays, shadow rays, path length sum, kd nodes, kd indices, prim tests).
int orc_render(void *s, const B200pgIntegratorParams *P, int first_sample, int n_samples, int row_begin, int row_end,
               float *film, int nthreads, uint64_t *stats, double *seconds, void *field, void *sink) {
    Scene *sc = (Scene *)s;
    const int W = sc->film.width, H = sc->film.height;
    if (row_end <= 0 || row_end > H) row_end = H;
    if (row_begin < 0) row_begin = 0;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const int TS = 32;
    const int tx = (W + TS - 1) / TS, ty0 = row_begin / TS, ty1 = (row_end + TS - 1) / TS;
    const int ntiles = tx * (ty1 - ty0);
    std::mutex filmMutex;
    Stats total;
    std::vector<GuideSamples> tileSamples(sink ? ntiles : 0);
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nthreads)
    {
        Stats st;
        ImageBlock blk;
#pragma omp for schedule(dynamic, 1)
        for (int tile = 0; tile < ntiles; ++tile) {
            int bx = (tile % tx) * TS, by = (ty0 + tile / tx) * TS;
            int y0 = std::max(by, row_begin), y1 = std::min(std::min(by + TS, H), row_end);
            int bw = std::min(TS, W - bx), bh = y1 - y0;
            if (bh <= 0) continue;
            blk.init(bx, y0, bw, bh, sc->film.borderSize);
            for (int y = y0; y < y1; ++y)
                for (int x = bx; x < bx + bw; ++x) {
                    uint32_t pixIdx = (uint32_t)y * W + x;
                    for (int j = 0; j < n_samples; ++j) {
                        Rng rng;
                        rng.init(sc->seed, pixIdx, (uint32_t)(first_sample + j));
                        Vec2 off = rng.next2D();
                        Vec2 samplePos(x + off.x, y + off.y);
                        Ray ray = sc->sampleRay(samplePos);
                        GuideCtx G;
                        G.field = (const GuideField *)field;
                        G.rec = sink ? &tileSamples[tile] : nullptr;
                        G.alpha = P->guiding_probability;
                        Vec3 spec = Li(*sc, *P, ray, rng, st, (field || sink) ? &G : nullptr);
                        float maxSpec = spec.maxc();  // progressiveintegrator.cpp:274-277
                        if (maxSpec > P->max_component_value) spec *= P->max_component_value / maxSpec;
                        Float v[5] = {spec.x, spec.y, spec.z, 1.0f, 1.0f};
                        blk.put(sc->film, samplePos, v);
                    }
                }
            {
                std::lock_guard<std::mutex> lock(filmMutex);
                const int b = blk.border, sx = blk.w + 2 * b, sy = blk.h + 2 * b;
                for (int yy = 0; yy < sy; ++yy) {
                    int fy = blk.oy - b + yy;
                    if (fy < 0 || fy >= H) continue;
                    for (int xx = 0; xx < sx; ++xx) {
                        int fx = blk.ox - b + xx;
                        if (fx < 0 || fx >= W) continue;
                        for (int k = 0; k < 5; ++k) film[((size_t)fy * W + fx) * 5 + k] += blk.data[((size_t)yy * sx + xx) * 5 + k];
                    }
                }
            }
        }
#pragma omp critical
        {
            total.paths += st.paths;
            total.normalRays += st.normalRays;
            total.shadowRays += st.shadowRays;
            total.pathLen += st.pathLen;
            total.trav.nodes += st.trav.nodes;
            total.trav.indices += st.trav.indices;
            total.trav.prims += st.trav.prims;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (sink) {
        // concatenate the per-tile sample lists in tile order (deterministic); parallel copy so that the merge does
        // not serialise the CPU baseline
        GuideSamples *out = (GuideSamples *)sink;
        std::vector<size_t> ofs(tileSamples.size() + 1, out->size());
        for (size_t t = 0; t < tileSamples.size(); ++t) ofs[t + 1] = ofs[t] + tileSamples[t].size();
        const size_t tot = ofs.back();
        out->pos.resize(tot); out->dir.resize(tot); out->weight.resize(tot); out->pdf.resize(tot); out->dist.resize(tot);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
        for (long long t = 0; t < (long long)tileSamples.size(); ++t) {
            const GuideSamples &c = tileSamples[t];
            std::copy(c.pos.begin(), c.pos.end(), out->pos.begin() + ofs[t]);
            std::copy(c.dir.begin(), c.dir.end(), out->dir.begin() + ofs[t]);
            std::copy(c.weight.begin(), c.weight.end(), out->weight.begin() + ofs[t]);
            std::copy(c.pdf.begin(), c.pdf.end(), out->pdf.begin() + ofs[t]);
            std::copy(c.dist.begin(), c.dist.end(), out->dist.begin() + ofs[t]);
        }
    }
    if (stats) {
        stats[0] = total.paths;
        stats[1] = total.normalRays;
        stats[2] = total.shadowRays;
        stats[3] = total.pathLen;
        stats[4] = total.trav.nodes;
        stats[5] = total.trav.indices;
        stats[6] = total.trav.prims;
    }
    return 0;
}

// Rough-transmittance reduction (rtrans.h:81-388): reads a MTS_TRANSMITTANCE .dat (or this repo's
// packed copy, same float order) and reduces it exactly like RoughPlastic::configure
// (roughplastic.cpp:290-307): external: setEta(eta), setAlpha(alpha); internal: setEta(1/eta).
// raw = full table as in the file after the header: for i<2*etaN, j<alphaN: thetaN trans + 1 diff.
int orc_rtrans_reduce(const float *raw, int etaN, int alphaN, int thetaN, float etaMin, float etaMax, float alphaMin,
                      float alphaMax, float eta, float alpha, float *ext_trans /* thetaN */, float *ext_diff,
                      float *int_diff) {
    std::vector<Float> trans((size_t)2 * etaN * alphaN * thetaN), diff((size_t)2 * etaN * alphaN);
    const float *ptr = raw;
    size_t fdr = 0, de = 0;
    for (int i = 0; i < 2 * etaN; ++i)
        for (int j = 0; j < alphaN; ++j) {
            for (int k = 0; k < thetaN; ++k) trans[de++] = *ptr++;
            diff[fdr++] = *ptr++;
        }
    auto setEta = [&](Float e, std::vector<Float> &outTrans, std::vector<Float> &outDiff) {
        const Float *tr = trans.data(), *df = diff.data();
        if (e < 1) {
            tr += (size_t)etaN * alphaN * thetaN;
            df += (size_t)etaN * alphaN;
            e = 1.0f / e;
        }
        if (e < etaMin) e = etaMin;
        Float warpedEta = std::pow((e - etaMin) / (etaMax - etaMin), 0.25f);
        outTrans.resize((size_t)alphaN * thetaN);
        outDiff.resize(alphaN);
        Float dAlpha = 1.0f / (alphaN - 1), dTheta = 1.0f / (thetaN - 1);
        size_t size3[3] = {(size_t)thetaN, (size_t)alphaN, (size_t)etaN};
        size_t size2[2] = {(size_t)alphaN, (size_t)etaN};
        for (int i = 0; i < alphaN; ++i) {
            for (int j = 0; j < thetaN; ++j) {
                Float p[3] = {j * dTheta, i * dAlpha, warpedEta};
                outTrans[(size_t)i * thetaN + j] = evalCubicInterpND<3>(p, tr, size3);
            }
            Float p2[2] = {i * dAlpha, warpedEta};
            outDiff[i] = evalCubicInterpND<2>(p2, df, size2);
        }
    };
    std::vector<Float> eT, eD, iT, iD;
    setEta(eta, eT, eD);
    setEta(1.0f / eta, iT, iD);
    Float warpedAlpha = std::pow((alpha - alphaMin) / (alphaMax - alphaMin), 0.25f);
    Float dTheta = 1.0f / (thetaN - 1);
    size_t size2[2] = {(size_t)thetaN, (size_t)alphaN};
    for (int i = 0; i < thetaN; ++i) {
        Float p[2] = {i * dTheta, warpedAlpha};
        ext_trans[i] = evalCubicInterpND<2>(p, eT.data(), size2);
    }
    *ext_diff = evalCubicInterp1D(warpedAlpha, eD.data(), alphaN, 0.0f, 1.0f);
    // internal: only eta fixed -> evalDiffuse(alpha) = clamp(evalCubicInterp1D(warpedAlpha, diffTrans))
    Float r = evalCubicInterp1D(warpedAlpha, iD.data(), alphaN, 0.0f, 1.0f);
    *int_diff = std::min(1.0f, std::max(0.0f, r));
    return 0;
}

int orc_num_threads(void) { return omp_get_max_threads(); }

// ---- guiding field (this repo's own algorithm; see oracle_guiding.h)
void *orc_field_create(int K, const float *bmin, const float *bmax) {
    GuideField *f = new GuideField();
    f->init(K, bmin, bmax);
    return f;
}
void orc_field_destroy(void *f) { delete (GuideField *)f; }
size_t orc_field_snapshot(void *f, uint32_t *out, size_t cap) {
    std::vector<uint32_t> w = ((GuideField *)f)->snapshot();
    if (out && cap >= w.size()) std::memcpy(out, w.data(), w.size() * 4);
    return w.size();
}
int orc_field_load(void *f, const uint32_t *w, size_t n) { return ((GuideField *)f)->load(w, n) ? 0 : -1; }
int orc_field_info(void *f, uint32_t *out /* nodes, cells, K */) {
    GuideField *F = (GuideField *)f;
    out[0] = (uint32_t)F->nodes.size();
    out[1] = F->numCells();
    out[2] = (uint32_t)F->K;
    return 0;
}
// u: n*3 (lobe selection, u1, u2). out_pdf: pdf of `dir`; out_dir / out_spdf: sampled direction and its pdf.
int orc_vmm_pdf_sample(void *f, const float *pos, const float *dir, const float *u, size_t n, float *out_pdf, float *out_dir,
                       float *out_spdf, uint32_t *out_cell) {
    GuideField *F = (GuideField *)f;
    for (size_t i = 0; i < n; ++i) {
        uint32_t c = F->lookup(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
        out_cell[i] = c;
        out_pdf[i] = F->pdf(c, Vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        Vec3 d = F->sample(c, u[3 * i], u[3 * i + 1], u[3 * i + 2]);
        out_dir[3 * i] = d.x; out_dir[3 * i + 1] = d.y; out_dir[3 * i + 2] = d.z;
        out_spdf[i] = F->pdf(c, d);
    }
    return 0;
}
int orc_bin_samples(void *f, const float *pos, size_t n, uint32_t *out_cell, uint32_t *out_perm, uint32_t *out_offsets) {
    GuideField *F = (GuideField *)f;
    std::vector<Vec3> p(n);
    for (size_t i = 0; i < n; ++i) p[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    std::vector<uint32_t> cell, perm, offsets;
    guideBin(*F, p.data(), n, cell, perm, offsets);
    std::memcpy(out_cell, cell.data(), n * 4);
    std::memcpy(out_perm, perm.data(), n * 4);
    std::memcpy(out_offsets, offsets.data(), offsets.size() * 4);
    return 0;
}
static void fillSamples(GuideSamples &S, const float *pos, const float *dir, const float *weight, const float *pdf,
                        const float *dist, size_t n) {
    for (size_t i = 0; i < n; ++i)
        S.push(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), Vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), weight[i],
               pdf[i], dist[i]);
}
// E-step only: stats_out has numCells * (4K + 8) floats (double accumulation, rounded once)
int orc_estep(void *f, const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist, size_t n,
              float *stats_out) {
    GuideField *F = (GuideField *)f;
    GuideSamples S;
    fillSamples(S, pos, dir, weight, pdf, dist, n);
    std::vector<uint32_t> cell, perm, offsets;
    guideBin(*F, S.pos.data(), n, cell, perm, offsets);
    std::vector<double> st;
    guideEStep(*F, S, perm, offsets, st);
    for (size_t i = 0; i < st.size(); ++i) stats_out[i] = (float)st[i];
    return 0;
}
int orc_train(void *f, const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist, size_t n,
              int nIter, float maxCellSamples) {
    GuideField *F = (GuideField *)f;
    GuideSamples S;
    fillSamples(S, pos, dir, weight, pdf, dist, n);
    guideTrain(*F, S, nIter, maxCellSamples);
    return 0;
}
// training update straight from a sample sink (no copies through the caller): what the CPU baseline times
int orc_train_sink(void *f, void *sink, int nIter, float maxCellSamples) {
    guideTrain(*(GuideField *)f, *(GuideSamples *)sink, nIter, maxCellSamples);
    return 0;
}
// the same update with several split levels (oracle_guiding.h: guideTrain, splitLevels)
int orc_train_levels(void *f, const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist, size_t n,
                     int nIter, float maxCellSamples, int splitLevels) {
    GuideField *F = (GuideField *)f;
    GuideSamples S;
    fillSamples(S, pos, dir, weight, pdf, dist, n);
    guideTrain(*F, S, nIter, maxCellSamples, splitLevels);
    return 0;
}
void *orc_samples_create(void) { return new GuideSamples(); }
void orc_samples_destroy(void *h) { delete (GuideSamples *)h; }
size_t orc_samples_size(void *h) { return ((GuideSamples *)h)->size(); }
void orc_samples_clear(void *h) { *((GuideSamples *)h) = GuideSamples(); }
int orc_samples_get(void *h, float *pos, float *dir, float *weight, float *pdf, float *dist) {
    GuideSamples *S = (GuideSamples *)h;
    for (size_t i = 0; i < S->size(); ++i) {
        pos[3 * i] = S->pos[i].x; pos[3 * i + 1] = S->pos[i].y; pos[3 * i + 2] = S->pos[i].z;
        dir[3 * i] = S->dir[i].x; dir[3 * i + 1] = S->dir[i].y; dir[3 * i + 2] = S->dir[i].z;
        weight[i] = S->weight[i]; pdf[i] = S->pdf[i]; dist[i] = S->dist[i];
    }
    return 0;
}

}  // extern "C"

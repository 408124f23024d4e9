// device_scene.cuh -- device view of the compiled scene, BVH traversal, primitive tests,
// intersection records and area-light sampling.
//
// What each routine has to agree with in the reference (paths relative to the reference root):
//   traceRay         closest / any hit as ShapeKDTree::rayIntersect (src/librender/skdtree.cpp:112-142, 207-226):
//                    adaptive ray epsilon, [mint, maxt] interval, TriAccel test (include/mitsuba/render/triaccel.h:96-158),
//                    rectangle test (src/shapes/rectangle.cpp:125-148). The accelerator is this repo's own BVH2
//                    (the north star allows a rebuilt BVH); hit results do not depend on it.
//   fillIntersection ShapeKDTree::fillIntersectionRecord<true> (include/mitsuba/render/skdtree.h:343-428),
//                    Rectangle::fillIntersectionRecord (rectangle.cpp:155-168), computeShadingFrame (util.cpp:605-610)
//   sampleEmitter    Scene::sampleEmitterDirect (src/librender/scene.cpp:871-895), AreaLight::sampleDirect
//                    (src/emitters/area.cpp:158-173), Shape::sampleDirect (src/librender/shape.cpp:102-116)
#pragma once
#include "device_bsdf.cuh"
#include "pg_types.h"

namespace pg {

struct DeviceScene {
    const float4 *nodes;
    const float4 *wideNodes;  // 6 x float4 per node (pg_types.h: WideNode), nullptr when the scene has no wide tree
    const float4 *primPlanes;  // per primitive slot: plane row of the affine map (16 B)
    const float4 *primRows;    // per primitive slot: the two (u, v) rows (32 B)
    const float4 *rects;
    const ShapeRecord *shapes;
    const MeshRecord *meshes;
    const float *positions;
    const float *normals;
    const float *texcoords;
    const uint32_t *indices;
    const float *areaCdf;
    const BsdfRecord *bsdfs;
    const EmitterRecord *emitters;
    const float *emitterCdf;
    const MediumRecord *media;
    const float *density;
    const uint32_t *primGlobalId;
    const PrimInfo *primInfo;
    const float4 *shadeTris;  // 6 x float4 per primitive slot (pg_types.h: ShadeTri)
    uint32_t nEmitters;
    uint32_t nPrims;
    CameraRecord camera;
    FilmRecord film;
    uint64_t seed;
};

struct Hit {
    float t, u, v;
    uint32_t prim;  // index into DeviceScene::prims (BVH order), kMiss = none
};

// Ray interval exactly as ShapeKDTree::rayIntersect sets it up (skdtree.cpp:119-133 / 212-222):
// an epsilon of exactly `Epsilon` is scaled by the largest origin coordinate.
PG_DEV float adaptiveMinT(float3 o, float mint, bool shadow) {
    if (mint == kEpsilon) {
        float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
        if (!shadow) m = fmaxf(m, kEpsilon);
        mint *= m;
    }
    return mint;
}

// Packed FP32 pairs (sm_100a: FADD2 / FMUL2 process two floats per instruction). The slab distances of one box axis,
// (min - o) * idir and (max - o) * idir, are exactly the two lanes of one pair, so the twelve subtract + twelve multiply of a
// node visit become six + six instructions with bit-identical results.
#ifndef PG_PACKED_SLABS
#define PG_PACKED_SLABS 1
#endif
PG_DEV float2 slabPair(float lo, float hi, float o, float idir) {
#if PG_PACKED_SLABS
    float2 v = make_float2(lo, hi), o2 = make_float2(o, o), i2 = make_float2(idir, idir);
    unsigned long long d, e;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long &>(v)), "l"(reinterpret_cast<unsigned long long &>(o2)));
    asm("mul.f32x2 %0, %1, %2;" : "=l"(e) : "l"(d), "l"(reinterpret_cast<unsigned long long &>(i2)));
    return reinterpret_cast<float2 &>(e);
#else
    return make_float2((lo - o) * idir, (hi - o) * idir);
#endif
}

// Box tests of one inner node (4 x LDG.128): returns how many children the ray interval [mint, tmax] overlaps; c0 = the
// nearer (or only) hit child, c1 = the farther one.
static constexpr int kDoneNode = (int)0x80000000;  // ~kDoneNode is not a valid leaf code
PG_DEV int bvhTestNode(const DeviceScene &S, int node, float3 o, float3 idir, float mint, float tmax, int &c0, int &c1) {
    const F8 lo = ldg256(S.nodes + 4 * node), hi = ldg256(S.nodes + 4 * node + 2);  // 64-byte node = 2 x LDG.256
    const float4 n0 = lo.a, n1 = lo.b, n2 = hi.a, n3 = hi.b;
    // slabs; fminf/fmaxf drop NaNs from 0*inf
    const float2 x0 = slabPair(n0.x, n0.y, o.x, idir.x), y0 = slabPair(n0.z, n0.w, o.y, idir.y), z0 = slabPair(n2.x, n2.y, o.z, idir.z);
    const float2 x1 = slabPair(n1.x, n1.y, o.x, idir.x), y1 = slabPair(n1.z, n1.w, o.y, idir.y), z1 = slabPair(n2.z, n2.w, o.z, idir.z);
    const float t0n = fmaxf(fmaxf(fminf(x0.x, x0.y), fminf(y0.x, y0.y)), fmaxf(fminf(z0.x, z0.y), mint));
    const float t0f = fminf(fminf(fmaxf(x0.x, x0.y), fmaxf(y0.x, y0.y)), fminf(fmaxf(z0.x, z0.y), tmax));
    const float t1n = fmaxf(fmaxf(fminf(x1.x, x1.y), fminf(y1.x, y1.y)), fmaxf(fminf(z1.x, z1.y), mint));
    const float t1f = fminf(fminf(fmaxf(x1.x, x1.y), fmaxf(y1.x, y1.y)), fminf(fmaxf(z1.x, z1.y), tmax));
    // conservative far bound (flat boxes, rounding): 1 + 2*gamma(3)
    const bool h0 = t0n <= t0f * 1.0000004f;
    const bool h1 = t1n <= t1f * 1.0000004f;
    c0 = __float_as_int(n3.x);
    c1 = __float_as_int(n3.y);
    if (h0 && h1) {
        if (t1n < t0n) {
            const int tmp = c0; c0 = c1; c1 = tmp;
        }
        return 2;
    }
    if (h1) c0 = c1;
    return (h0 | h1) ? 1 : 0;
}

// One inner-node step of the BVH2 descent: continues with the nearer hit child and pushes the farther one.
// kDoneNode when the stack runs empty.
static constexpr int kTraceStack = 64;  // the builder bounds the tree depth (host_scene.cpp: median splits below depth 36)
// Deferred subtrees are prefetched into L2 when they are pushed (kPrefetch): on a mesh whose node + primitive arrays exceed the
// L2 every pop is otherwise a cold DRAM round trip on the critical path of a latency-bound warp.
PG_DEV void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
PG_DEV void prefetchRef(const DeviceScene &S, int ref) {
    if (ref >= 0) {
        prefetchL2(S.nodes + 4 * (size_t)ref);       // 64-byte node = two sectors of one line
        prefetchL2(S.nodes + 4 * (size_t)ref + 2);
    } else {
        const uint32_t code = (uint32_t)(~ref);
        prefetchL2(S.primPlanes + (code >> kLeafShift));  // <= 4 planes x 16 B, 32-byte aligned: one line
    }
}
template <bool kPrefetch = false>
PG_DEV int bvhNodeStep(const DeviceScene &S, int node, float3 o, float3 idir, float mint, float tmax, int *stack, int &sp) {
    int c0, c1;
    const int nh = bvhTestNode(S, node, o, idir, mint, tmax, c0, c1);
    if (nh == 2) {
        if (kPrefetch) prefetchRef(S, c1);
        stack[sp++] = c1;
        return c0;
    }
    if (nh == 1) return c0;
    return sp ? stack[--sp] : kDoneNode;
}

// ---- wide (8-ary, quantised) tree -------------------------------------------------------------------------------------
// Stack entries carry the entry distance of the subtree, so a popped entry whose distance lies behind the current hit is
// dropped without touching memory.
static constexpr int kWideStack = 7 * kWideMaxDepth + 16;
struct WideEntry {
    int ref;
    float tn;
};
// The wide test forms (origin - o) * idir - 2^23 * 2^e * idir: with an infinite reciprocal (a zero direction component) that is
// inf - inf. A reciprocal clamped to 1e25 decides every slab exactly as infinity does (any offset of the ray origin from a
// plane beyond 1e-20 maps to a distance outside every interval) and keeps the arithmetic finite.
PG_DEV float3 wideClampIdir(float3 idir) {
    return f3(copysignf(fminf(fabsf(idir.x), 1e25f), idir.x), copysignf(fminf(fabsf(idir.y), 1e25f), idir.y),
              copysignf(fminf(fabsf(idir.z), 1e25f), idir.z));
}
PG_DEV int widePop(const WideEntry *stack, int &sp, float tmax) {
    while (sp) {
        const WideEntry e = stack[--sp];
        if (e.tn <= tmax) return e.ref;
    }
    return kDoneNode;
}
// ray distance of the quantised plane `byte` of word w: fma(2^23 + q, s, b) with s = 2^e * idir, b = (origin - o) * idir - 2^23 s.
// __byte_perm builds the float 2^23 + q (0x4B0000qq) in one instruction: no integer -> float conversion on the slow pipe.
template <int kByte>
PG_DEV float widePlane(uint32_t w, float s, float b) {
    return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + kByte)), s, b);
}
// One node visit: tests the (up to) 8 children, continues with the nearest hit child and pushes the others; kDoneNode when
// nothing is left. Same conservative far-bound factor as the binary test.
PG_DEV int wideNodeStep(const DeviceScene &S, int node, float3 o, float3 idir, float mint, float tmax, WideEntry *stack, int &sp) {
    const float4 *N = S.wideNodes + 6 * (size_t)node;
    const float4 h = __ldg(N);
    const int4 r0 = __ldg(reinterpret_cast<const int4 *>(N + 1)), r1 = __ldg(reinterpret_cast<const int4 *>(N + 2));
    const uint4 qa = __ldg(reinterpret_cast<const uint4 *>(N + 3)), qb = __ldg(reinterpret_cast<const uint4 *>(N + 4)),
                qc = __ldg(reinterpret_cast<const uint4 *>(N + 5));
    const uint32_t meta = __float_as_uint(h.w);
    const float sx = __uint_as_float((meta & 0xFFu) << 23) * idir.x, sy = __uint_as_float(((meta >> 8) & 0xFFu) << 23) * idir.y,
                sz = __uint_as_float(((meta >> 16) & 0xFFu) << 23) * idir.z;
    const float bx = fmaf(-8388608.0f, sx, (h.x - o.x) * idir.x), by = fmaf(-8388608.0f, sy, (h.y - o.y) * idir.y),
                bz = fmaf(-8388608.0f, sz, (h.z - o.z) * idir.z);
    int best = kDoneNode;
    float bestT = kInf;
#define PG_WIDE_CHILD(REF, LOX, LOY, LOZ, HIX, HIY, HIZ, B)                                                         \
    {                                                                                                                \
        const float ax = widePlane<B>(LOX, sx, bx), cx = widePlane<B>(HIX, sx, bx);                                  \
        const float ay = widePlane<B>(LOY, sy, by), cy = widePlane<B>(HIY, sy, by);                                  \
        const float az = widePlane<B>(LOZ, sz, bz), cz = widePlane<B>(HIZ, sz, bz);                                  \
        const float tn = fmaxf(fmaxf(fminf(ax, cx), fminf(ay, cy)), fmaxf(fminf(az, cz), mint));                     \
        const float tf = fminf(fminf(fmaxf(ax, cx), fmaxf(ay, cy)), fminf(fmaxf(az, cz), tmax));                     \
        const bool hit = (REF) != kWideEmpty && tn <= tf * 1.0000004f;                                               \
        const bool better = hit && tn < bestT;                                                                       \
        /* branch-free: the nearer of (this child, best so far) stays in registers, the other one goes to the stack */ \
        WideEntry out;                                                                                               \
        out.ref = better ? best : (REF);                                                                             \
        out.tn = better ? bestT : tn;                                                                                \
        if (hit && out.ref != kDoneNode) stack[sp++] = out;                                                          \
        best = better ? (REF) : best;                                                                                \
        bestT = better ? tn : bestT;                                                                                 \
    }
    PG_WIDE_CHILD(r0.x, qa.x, qa.z, qb.x, qb.z, qc.x, qc.z, 0)
    PG_WIDE_CHILD(r0.y, qa.x, qa.z, qb.x, qb.z, qc.x, qc.z, 1)
    PG_WIDE_CHILD(r0.z, qa.x, qa.z, qb.x, qb.z, qc.x, qc.z, 2)
    PG_WIDE_CHILD(r0.w, qa.x, qa.z, qb.x, qb.z, qc.x, qc.z, 3)
    PG_WIDE_CHILD(r1.x, qa.y, qa.w, qb.y, qb.w, qc.y, qc.w, 0)
    PG_WIDE_CHILD(r1.y, qa.y, qa.w, qb.y, qb.w, qc.y, qc.w, 1)
    PG_WIDE_CHILD(r1.z, qa.y, qa.w, qb.y, qb.w, qc.y, qc.w, 2)
    PG_WIDE_CHILD(r1.w, qa.y, qa.w, qb.y, qb.w, qc.y, qc.w, 3)
#undef PG_WIDE_CHILD
    if (best != kDoneNode) return best;
    return widePop(stack, sp, tmax);
}

// All primitives of one leaf (<= 4): the branch-free affine test that serves rectangles and triangles alike.
// Returns true on the first accepted hit in any-hit mode.
// The record is split (pg_types.h): the plane row {M_z, w_z} of all primitives in `planes` (16 B each; a leaf starts at an even
// index, so two neighbouring planes are one aligned 256-bit load), the two (u, v) rows in `rows` (32 B, fetched only for a
// primitive whose plane distance falls inside the ray interval). Same arithmetic as one 48-byte record.
template <bool kAnyHit, bool kCount>
PG_DEV bool bvhLeafStep(const DeviceScene &S, int node, float3 o, float3 d, float mint, float &tmax, Hit &hit, uint32_t *cntPrims) {
    const uint32_t code = (uint32_t)(~node);
    const uint32_t first = code >> kLeafShift, count = code & 7u, rectMask = (code >> 3) & 15u;
    F8 pl[2];
    pl[0] = ldg256(S.primPlanes + first);
    pl[1] = pl[0];
    if (count > 2) pl[1] = ldg256(S.primPlanes + first + 2);
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        if (i >= count) break;
        const float4 r2 = (i & 1u) ? pl[i >> 1].b : pl[i >> 1].a;
        if (kCount) (*cntPrims)++;
        // local = M o + w, local direction = M d (rectangle.cpp:125-133 for rectangles)
        const float loz = r2.x * o.x + r2.y * o.y + r2.z * o.z + r2.w;
        const float ldz = r2.x * d.x + r2.y * d.y + r2.z * d.z;
        const float t = -loz / ldz;
        if (t >= mint && t <= tmax) {
            const F8 uv = ldg256(S.primRows + 2 * (size_t)(first + i));
            const float4 r0 = uv.a, r1 = uv.b;
            const float lox = r0.x * o.x + r0.y * o.y + r0.z * o.z + r0.w;
            const float loy = r1.x * o.x + r1.y * o.y + r1.z * o.z + r1.w;
            const float ldx = r0.x * d.x + r0.y * d.y + r0.z * d.z;
            const float ldy = r1.x * d.x + r1.y * d.y + r1.z * d.z;
            const float u = lox + ldx * t, v = loy + ldy * t;
            const bool isRect = (rectMask >> i) & 1u;
            const bool ok = isRect ? (fabsf(u) <= 1 && fabsf(v) <= 1) : (u >= 0 && v >= 0 && u + v <= 1.0f);
            if (ok) {
                hit.prim = first + i;
                if (kAnyHit) return true;
                tmax = t;
                hit.t = t;
                hit.u = u;
                hit.v = v;
            }
        }
    }
    return false;
}

template <bool kAnyHit, bool kCount>
PG_DEV bool traceRay(const DeviceScene &S, float3 o, float3 d, float mint, float maxt, Hit &hit, uint32_t *cntNodes,
                     uint32_t *cntPrims) {
    // "while-while" traversal: every lane first descends to its next leaf, then the warp processes
    // leaves together, so that the (uniform, branch-free) primitive test runs with many lanes active.
    const float3 idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[kTraceStack];
    int sp = 0;
    int node = 0;
    hit.prim = kMiss;
    hit.t = maxt;
    float tmax = maxt;
    while (true) {
        while (node >= 0) {
            if (kCount) (*cntNodes)++;
            node = bvhNodeStep(S, node, o, idir, mint, tmax, stack, sp);
        }
        if (node == kDoneNode) break;
        if (bvhLeafStep<kAnyHit, kCount>(S, node, o, d, mint, tmax, hit, cntPrims)) return true;
        node = sp ? stack[--sp] : kDoneNode;
        if (node == kDoneNode) break;
    }
    return hit.prim != kMiss;
}

struct Intersection {
    float3 p, geoN;
    Frame sh;
    float3 wi;
    float2 uv;
    float t;
    int shape, bsdf, emitter;
    uint32_t primIndex;
};

PG_DEV void fillIntersection(const DeviceScene &S, float3 o, float3 d, const Hit &h, Intersection &its) {
    const float4 *T = S.shadeTris + 6 * (size_t)h.prim;  // 96-byte record = 3 x LDG.256
    const F8 t01 = ldg256(T), t23 = ldg256(T + 2), t45 = ldg256(T + 4);
    const float4 r0 = t01.a, r1 = t01.b, r2 = t23.a, r3 = t23.b, r4 = t45.a;
    const uint32_t flags = __float_as_uint(r1.w);
    its.shape = (int)__float_as_uint(r0.w);
    its.bsdf = (int)__float_as_uint(r3.w);
    its.emitter = (int)(flags >> 8) - 1;
    its.primIndex = __float_as_uint(r2.w);
    its.t = h.t;
    float3 dpdu, shN;
    if (flags & 1u) {  // triangle (skdtree.h:343-428)
        const float3 b = f3(1 - h.u - h.v, h.u, h.v);
        const float3 p0 = f3(r0.x, r0.y, r0.z), p1 = f3(r1.x, r1.y, r1.z), p2 = f3(r2.x, r2.y, r2.z);
        its.p = p0 * b.x + p1 * b.y + p2 * b.z;
        float3 side1 = p1 - p0, side2 = p2 - p0;
        float3 faceNormal = cross(side1, side2);
        float len = length(faceNormal);
        if (!isZero(faceNormal)) faceNormal = faceNormal / len;
        // the UV tangent of meshes with texture coordinates (trimesh.cpp:683-735), else the first edge (skdtree.h:374-381)
        dpdu = (flags & 4u) ? side1 * r4.w + side2 * t45.b.w : side1;
        if (flags & 2u) {
            const float4 r5 = t45.b;
            const float3 n0 = f3(r3.x, r3.y, r3.z), n1 = f3(r4.x, r4.y, r4.z), n2 = f3(r5.x, r5.y, r5.z);
            shN = normalize(n0 * b.x + n1 * b.y + n2 * b.z);
            if (dot(faceNormal, shN) < 0) faceNormal = -faceNormal;
        } else {
            shN = faceNormal;
        }
        its.geoN = faceNormal;
        its.uv = make_float2(b.y, b.z);
    } else {  // rectangle (rectangle.cpp:155-168)
        const uint32_t rect = S.shapes[its.shape].meshOffset;
        const float4 q3 = __ldg(S.rects + 8 * rect + 3), q4 = __ldg(S.rects + 8 * rect + 4);
        shN = f3(q3.x, q3.y, q3.z);
        its.geoN = shN;
        dpdu = f3(q4.x, q4.y, q4.z);
        its.uv = make_float2(0.5f * (h.u + 1), 0.5f * (h.v + 1));
        its.p = o + d * h.t;
    }
    its.sh = shadingFrame(shN, dpdu);
    its.wi = its.sh.toLocal(-d);
}

struct DirectSample {
    float3 d, n;
    float dist, pdf;
    int emitter;
};

// Shape::samplePosition for rectangles (rectangle.cpp:210-216) and meshes (trimesh.cpp:412-423, triangle.cpp:24-59)
PG_DEV void samplePosition(const DeviceScene &S, const ShapeRecord &sr, float2 sample, float3 &p, float3 &n, float &pdf) {
    if (sr.type == B200PG_SHAPE_RECTANGLE) {
        const float4 *r = S.rects + 8 * sr.meshOffset;
        const float4 r3 = __ldg(r + 3), m0 = __ldg(r + 5), m1 = __ldg(r + 6), m2 = __ldg(r + 7);
        float lx = sample.x * 2 - 1, ly = sample.y * 2 - 1;
        p = f3(m0.x * lx + m0.y * ly + m0.w, m1.x * lx + m1.y * ly + m1.w, m2.x * lx + m2.y * ly + m2.w);
        n = f3(r3.x, r3.y, r3.z);
        pdf = r3.w;
    } else {
        const MeshRecord mr = S.meshes[sr.meshOffset];
        const float *cdf = S.areaCdf + mr.cdfOffset;
        uint32_t index = cdfSample(cdf, mr.nTriangles + 1, sample.y);
        sample.y = (sample.y - cdf[index]) / (cdf[index + 1] - cdf[index]);  // sampleReuse, pmf.h:163-168
        const uint32_t *idx = S.indices + 3 * ((size_t)mr.indexOffset + index);
        const uint32_t i0 = idx[0] + mr.vertexOffset, i1 = idx[1] + mr.vertexOffset, i2 = idx[2] + mr.vertexOffset;
        const float3 p0 = ld3(S.positions + 3 * (size_t)i0), p1 = ld3(S.positions + 3 * (size_t)i1), p2 = ld3(S.positions + 3 * (size_t)i2);
        float2 bary = squareToUniformTriangle(sample);
        float3 sideA = p1 - p0, sideB = p2 - p0;
        p = p0 + (sideA * bary.x) + (sideB * bary.y);
        if (mr.hasNormals) {
            const float3 n0 = ld3(S.normals + 3 * (size_t)i0), n1 = ld3(S.normals + 3 * (size_t)i1), n2 = ld3(S.normals + 3 * (size_t)i2);
            n = normalize(n0 * (1.0f - bary.x - bary.y) + n1 * bary.x + n2 * bary.y);
        } else {
            n = normalize(cross(sideA, sideB));
        }
        pdf = mr.invSurfaceArea;
    }
}

// Emitter sampling without the visibility test (the shadow ray goes to the shadow queue).
// Returns radiance / pdf (zero when rejected).
PG_DEV float3 sampleEmitterDirect(const DeviceScene &S, float3 ref, float3 refN, float2 sample, DirectSample &dRec) {
    uint32_t index = cdfSample(S.emitterCdf, S.nEmitters + 1, sample.x);
    float emPdf = S.emitterCdf[index + 1] - S.emitterCdf[index];
    sample.x = (sample.x - S.emitterCdf[index]) / (S.emitterCdf[index + 1] - S.emitterCdf[index]);
    const EmitterRecord em = S.emitters[index];
    const ShapeRecord sr = S.shapes[em.shape];
    float3 p;
    samplePosition(S, sr, sample, p, dRec.n, dRec.pdf);
    dRec.d = p - ref;
    float distSquared = dot(dRec.d, dRec.d);
    dRec.dist = sqrtf(distSquared);
    dRec.d = dRec.d / dRec.dist;
    float dp = fabsf(dot(dRec.d, dRec.n));
    dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    dRec.emitter = (int)index;
    if (dot(dRec.d, refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) {
        float3 value = ld3(em.radiance) / dRec.pdf;
        dRec.pdf *= emPdf;
        return value / emPdf;
    }
    dRec.pdf = 0.0f;
    return f3(0.0f);
}

// Scene::pdfEmitterDirect after DirectSamplingRecord::setQuery (records.inl:170-178; scene.cpp:992-995;
// area.cpp:175-183; shape.cpp:117-126). The refN test of AreaLight::pdfDirect always passes for a direction
// that was just sampled from the BSDF on the same side as refN, so it is not carried along.
PG_DEV float pdfEmitterDirect(const DeviceScene &S, int emitter, float3 d, float3 n, float dist) {
    if (!(dot(d, n) < 0)) return 0.0f;
    const EmitterRecord em = S.emitters[emitter];
    const ShapeRecord sr = S.shapes[em.shape];
    float invArea = sr.type == B200PG_SHAPE_RECTANGLE ? __ldg(S.rects + 8 * sr.meshOffset + 3).w
                                                       : S.meshes[sr.meshOffset].invSurfaceArea;
    float discrete = S.emitterCdf[emitter + 1] - S.emitterCdf[emitter];
    return invArea * (dist * dist) / fabsf(dot(d, n)) * discrete;
}

// PerspectiveCameraImpl::sampleRayDifferential without differentials (perspective.cpp:271-298)
PG_DEV void sampleCameraRay(const CameraRecord &C, float2 pixelSample, float3 &o, float3 &d, float &mint, float &maxt) {
    const float *m = C.sampleToCamera;
    float sx = pixelSample.x * C.invResX, sy = pixelSample.y * C.invResY;
    float x = m[0] * sx + m[1] * sy + m[3];
    float y = m[4] * sx + m[5] * sy + m[7];
    float z = m[8] * sx + m[9] * sy + m[11];
    float w = m[12] * sx + m[13] * sy + m[15];
    float3 nearP = f3(x, y, z);
    if (w != 1.0f) nearP = nearP / w;
    float3 dl = normalize(nearP);
    float invZ = 1.0f / dl.z;
    mint = C.nearClip * invZ;
    maxt = C.farClip * invZ;
    const float *t = C.toWorld;
    o = f3(t[3], t[7], t[11]);
    d = f3(t[0] * dl.x + t[1] * dl.y + t[2] * dl.z, t[4] * dl.x + t[5] * dl.y + t[6] * dl.z,
           t[8] * dl.x + t[9] * dl.y + t[10] * dl.z);
}

}  // namespace pg

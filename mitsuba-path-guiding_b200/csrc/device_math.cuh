// device_math.cuh -- float3 helpers, frames, the counter-based RNG and sampling warps for the
// sm_100a kernels. Reference formulas are cited where a routine has to agree with Mitsuba's
// arithmetic (paths relative to the reference root).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pg {

#define PG_DEV __device__ __forceinline__

// Streamed-once data (wavefront state, queues, training records): evict-first loads / stores keep it from displacing the
// small hot working set (guiding lobes, tree nodes, scene records) in L1 and L2.
#ifndef PG_STREAM
#define PG_STREAM 1
#endif
template <typename T>
PG_DEV T ldStream(const T *p) {
#if PG_STREAM
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T>
PG_DEV void stStream(T *p, T v) {
#if PG_STREAM
    __stcs(p, v);
#else
    *p = v;
#endif
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256, PTX ld/st.global.v8.f32). A scattered record costs the L1TEX data pipe
// one wavefront per (instruction, cache line) pair -- a divergent warp's LDG.128 is ~32 of them -- so a 32-byte-aligned 64-byte
// BVH node read as two 256-bit loads instead of four 128-bit ones halves the pipe time of a node visit (the traversal and the
// shade stage are bound by that pipe, profiles/r02_*). Addresses must be 32-byte aligned.
struct __align__(32) F8 {
    float4 a, b;
};
#ifndef PG_LD256
#define PG_LD256 1  // 0 = two 128-bit loads (A/B runs)
#endif
PG_DEV F8 ldg256(const float4 *p) {  // read-only data (scene, guiding field)
    F8 r;
#if PG_LD256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
        : "l"(p));
#else
    r.a = __ldg(p);
    r.b = __ldg(p + 1);
#endif
    return r;
}
PG_DEV F8 ldStream256(const float4 *p) {  // streamed-once data (evict-first)
    F8 r;
#if PG_LD256
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p)
                 : "memory");
#else
    r.a = ldStream(p);
    r.b = ldStream(p + 1);
#endif
    return r;
}
PG_DEV void stStream256(float4 *p, float4 a, float4 b) {
#if PG_LD256
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                 "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
#else
    stStream(p, a);
    stStream(p + 1, b);
#endif
}

static constexpr float kEpsilon = 1e-4f;        // include/mitsuba/core/constants.h:28
static constexpr float kShadowEpsilon = 1e-3f;  // constants.h:29
static constexpr float kPi = 3.14159265358979323846f;
static constexpr float kInvPi = 0.31830988618379067154f;
static constexpr float kInf = __builtin_huge_valf();

PG_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PG_DEV float3 f3(float v) { return make_float3(v, v, v); }
PG_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PG_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PG_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PG_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PG_DEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
PG_DEV float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PG_DEV float3 operator/(float3 a, float s) {
    float r = 1.0f / s;  // TVector3::operator/ multiplies by the reciprocal (vector.h)
    return f3(a.x * r, a.y * r, a.z * r);
}
PG_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PG_DEV void operator+=(float3 &a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
PG_DEV void operator*=(float3 &a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
PG_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PG_DEV float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PG_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
PG_DEV float3 normalize(float3 a) { return a / length(a); }
PG_DEV bool isZero(float3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
PG_DEV float maxComp(float3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
PG_DEV float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
PG_DEV float safeSqrt(float v) { return sqrtf(fmaxf(0.0f, v)); }  // math.h:260-267
PG_DEV float signum(float v) { return v < 0 ? -1.0f : (v > 0 ? 1.0f : 0.0f); }
PG_DEV float3 ld3(const float *p) { return f3(p[0], p[1], p[2]); }

struct Frame {
    float3 s, t, n;
    PG_DEV float3 toLocal(float3 v) const { return f3(dot(v, s), dot(v, t), dot(v, n)); }
    PG_DEV float3 toWorld(float3 v) const { return s * v.x + t * v.y + n * v.z; }
};
// coordinateSystem, src/libcore/util.cpp:594-603
PG_DEV void coordinateSystem(float3 a, float3 &b, float3 &c) {
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = f3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = f3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}
// computeShadingFrame, util.cpp:605-610
PG_DEV Frame shadingFrame(float3 n, float3 dpdu) {
    Frame f;
    f.n = n;
    f.s = normalize(dpdu - n * dot(n, dpdu));
    f.t = cross(n, f.s);
    return f;
}
PG_DEV Frame frameFromNormal(float3 n) {
    Frame f;
    f.n = n;
    coordinateSystem(n, f.s, f.t);
    return f;
}

// ---------------------------------------------------------------------------------------
// RNG: PCG32 (XSH-RR), stream = pixel index, seeded per (seed, sample index). Specified in
// DESIGN.md; the CPU oracle implements the same specification independently. The reference's
// per-pixel SFMT samplers (independent.cpp:52-104) are not replicated; the CONSUMPTION ORDER
// (next2D = two next1D, NEE 2D -> BSDF 2D -> RR 1D) follows progressive_path.cpp.
// ---------------------------------------------------------------------------------------
PG_DEV uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state, inc;
    PG_DEV uint32_t nextU32() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    PG_DEV void init(uint64_t seed, uint32_t pixel, uint32_t sample) {
        inc = ((uint64_t)pixel << 1) | 1ULL;
        state = 0;
        nextU32();
        state += mix64(seed + (uint64_t)sample * 0x9E3779B97F4A7C15ULL);
        nextU32();
    }
    PG_DEV float next1D() { return (float)(nextU32() >> 8) * (1.0f / 16777216.0f); }
    PG_DEV float2 next2D() {
        float a = next1D();
        float b = next1D();
        return make_float2(a, b);
    }
    // independent child stream (used for transmittance estimates so that the main stream's
    // consumption does not depend on how many medium interactions a connection crosses)
    PG_DEV Rng fork() {
        Rng r;
        uint64_t a = nextU32();
        uint64_t b = nextU32();
        r.inc = inc;
        r.state = mix64((a << 32) | b);
        return r;
    }
};

// ---------------------------------------------------------------------------------------
// warps, src/libcore/warp.cpp
// ---------------------------------------------------------------------------------------
PG_DEV float2 squareToUniformDiskConcentric(float2 sample) {  // warp.cpp:79-100
    float r1 = 2.0f * sample.x - 1.0f;
    float r2 = 2.0f * sample.y - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (kPi / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (kPi / 2.0f) - (r1 / r2) * (kPi / 4.0f);
    }
    float s, c;
    sincosf(phi, &s, &c);
    return make_float2(r * c, r * s);
}
PG_DEV float3 squareToCosineHemisphere(float2 sample) {  // warp.cpp:41-49
    float2 p = squareToUniformDiskConcentric(sample);
    float z = safeSqrt(1.0f - p.x * p.x - p.y * p.y);
    if (z == 0) z = 1e-10f;
    return f3(p.x, p.y, z);
}
PG_DEV float2 squareToUniformTriangle(float2 sample) {  // warp.cpp:74-77
    float a = safeSqrt(1.0f - sample.x);
    return make_float2(1 - a, a * sample.y);
}
PG_DEV float3 squareToUniformSphere(float2 sample) {  // warp.cpp:24-31
    float z = 1.0f - 2.0f * sample.y;
    float r = safeSqrt(1.0f - z * z);
    float s, c;
    sincosf(2.0f * kPi * sample.x, &s, &c);
    return f3(r * c, r * s, z);
}

// power heuristic, progressive_path.cpp:316-320
PG_DEV float miWeight(float pdfA, float pdfB) {
    pdfA *= pdfA;
    pdfB *= pdfB;
    return pdfA / (pdfA + pdfB);
}

// DiscreteDistribution::sample over a normalised cdf of n+1 entries (pmf.h:124-136):
// lower_bound, minus one, clamped, then skip zero-probability entries.
PG_DEV uint32_t cdfSample(const float *__restrict__ cdf, uint32_t nEntries /* = n+1 */, float v) {
    uint32_t lo = 0, hi = nEntries;  // first index with cdf[idx] >= v
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cdf[mid] < v) lo = mid + 1; else hi = mid;
    }
    int idx = (int)lo - 1;
    if (idx < 0) idx = 0;
    if (idx > (int)nEntries - 2) idx = (int)nEntries - 2;
    while (cdf[idx + 1] - cdf[idx] == 0 && idx < (int)nEntries - 1) ++idx;
    return (uint32_t)idx;
}

}  // namespace pg

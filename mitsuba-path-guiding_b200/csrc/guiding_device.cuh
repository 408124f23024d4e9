// guiding_device.cuh -- device view of the guiding field and the per-vertex query routines.
//
// The guided integrator and its guiding library (Intel Open PGL) are NOT part of the reference
// snapshot: this is the repo's own algorithm, specified in DESIGN.md "Guiding field"
// and stated a second time on the CPU in oracle/oracle_guiding.h, which the parity tests compare against.
//
//   field   = spatial kd-tree over the scene box; leaves ("cells") hold a mixture of K von Mises-Fisher lobes
//   query   = tree walk (16-byte nodes) -> cell; pdf / sample over the cell's K lobes (3 x float4 each)
//   use     = one-sample MIS between the BSDF and the mixture at every vertex with a smooth BSDF
#pragma once
#include "device_math.cuh"

namespace pg {

static constexpr int kGuideMaxK = 32;
static constexpr float kGuidePriorWeight = 0.01f;
static constexpr float kGuidePriorMeanCos = 0.8f;
static constexpr float kGuideKappaMin = 0.01f, kGuideKappaMax = 5000.0f;
static constexpr float kGuideDecay = 0.25f;
static constexpr float kGuideInitKappa = 5.0f;

struct GuideDevice {
    const uint4 *nodes;    // {axis (3 = leaf), split bits, left | cell, 0}
    const float4 *lobes;   // cells * K * 3: {pi, mu.xyz} {kappa, norm, exp(-2 kappa), 0} {S, R.xyz}
    int K;
    int enabled;           // sample from the field in the shade stage
    int record;            // record training vertices
    float alpha;           // selection probability of the guiding distribution
    // training-vertex records, indexed [slot * maxVerts + v]
    float4 *vPos;          // position, pdf of the sampled direction
    float4 *vDir;          // sampled direction, distance to the next vertex
    float4 *vThr;          // throughput right after the vertex
    float4 *vL;            // radiance gathered up to and including the vertex' NEE
    int maxVerts;
    // training samples (output of finished paths)
    float4 *sPos;          // position, weight
    float4 *sDir;          // direction, pdf
    float *sDist;
    uint32_t *sCount;
    uint32_t sCapacity;
};

PG_DEV uint32_t guideLookup(const GuideDevice &G, float3 p) {
    uint32_t n = 0;
    while (true) {
        const uint4 nd = __ldg(G.nodes + n);
        if (nd.x == 3u) return nd.z;
        n = comp(p, (int)nd.x) < __uint_as_float(nd.y) ? nd.z : nd.z + 1;
    }
}

PG_DEV float guidePdf(const GuideDevice &G, uint32_t cell, float3 w) {
    const float4 *L = G.lobes + (size_t)cell * G.K * 3;
    float s = 0;
    for (int k = 0; k < G.K; ++k) {
        const float4 a = __ldg(L + 3 * k), b = __ldg(L + 3 * k + 1);
        const float c = a.y * w.x + a.z * w.y + a.w * w.z;
        s += a.x * b.y * expf(b.x * (c - 1.0f));
    }
    return s;
}

PG_DEV float3 guideSample(const GuideDevice &G, uint32_t cell, float u0, float u1, float u2) {
    const float4 *L = G.lobes + (size_t)cell * G.K * 3;
    int k = 0;
    float4 a = __ldg(L);
    while (k < G.K - 1 && u0 >= a.x) {
        u0 -= a.x;
        ++k;
        a = __ldg(L + 3 * k);
    }
    const float4 b = __ldg(L + 3 * k + 1);
    float cosT = 1.0f + logf(u1 + (1.0f - u1) * b.z) / b.x;
    cosT = fminf(1.0f, fmaxf(-1.0f, cosT));
    const float sinT = safeSqrt(1.0f - cosT * cosT);
    float sp, cp;
    sincosf(2.0f * kPi * u2, &sp, &cp);
    const float3 mu = f3(a.y, a.z, a.w);
    float3 s, t;
    coordinateSystem(mu, s, t);
    return s * (sinT * cp) + t * (sinT * sp) + mu * cosT;
}

}  // namespace pg

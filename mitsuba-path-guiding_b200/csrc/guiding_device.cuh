// guiding_device.cuh -- device view of the guiding field and the per-vertex query routines.
//
// The guided integrator and its guiding library (Intel Open PGL) are NOT part of the reference
// snapshot: this is the repo's own algorithm, specified in DESIGN.md "Guiding field"
// and stated a second time on the CPU in oracle/oracle_guiding.h, which the parity tests compare against.
//
//   field   = spatial kd-tree over the scene box; leaves ("cells") hold a mixture of K von Mises-Fisher lobes
//   query   = tree walk (16-byte nodes) -> cell; pdf / sample over the cell's K lobes (2 x float4 each, 32 B;
//             the running EM statistics live in a separate array so that queries do not drag them through L1/L2)
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
    const float4 *lobes;   // cells * K * 2: {pi, mu.xyz} {kappa, norm, exp(-2 kappa), 0}
    const float4 *lobeStats;  // cells * K: running statistics {S, R.xyz} (training only)
    int K;
    int enabled;           // sample from the field in the shade stage
    int record;            // record training vertices
    float alpha;           // selection probability of the guiding distribution
    // training-vertex records, 64 B = 4 x float4 at [(slot * maxVerts + v) * 4]: two whole 32-byte sectors, each written
    // by ONE shade invocation (HBM3e is ECC-protected: a partially written sector costs a read-modify-write):
    //   sector A (when the vertex is sampled): {position, pdf of the sampled direction} {sampled direction, 0}
    //   sector B (one bounce later, "closing"): {throughput right after the vertex, distance to the next hit}
    //                                           {radiance gathered up to and including the vertex' NEE, 0}
    float4 *vRec;
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

// pi_k * norm_k * exp(kappa_k (mu_k . w - 1)). The exponential is the hardware ex2 path (__expf: MUFU.EX2 after a
// multiply by log2 e): its relative error is ~2^-22 plus |x| * 2^-24 from the argument scaling, i.e. < 2e-6 for every
// term that contributes to the sum (x > -30) -- inside the 1e-5 parity bar against the oracle's std::exp.
PG_DEV float guideLobeTerm(float4 a, float4 b, float3 w) {
    const float c = a.y * w.x + a.z * w.y + a.w * w.z;
    return a.x * b.y * __expf(b.x * (c - 1.0f));
}

PG_DEV float guidePdf(const GuideDevice &G, uint32_t cell, float3 w) {
    const float4 *L = G.lobes + (size_t)cell * G.K * 2;
    float s = 0;
#pragma unroll 4
    for (int k = 0; k < G.K; ++k) {
        const float4 a = __ldg(L + 2 * k), b = __ldg(L + 2 * k + 1);
        s += guideLobeTerm(a, b, w);
    }
    return s;
}

// pdf of two directions in one pass over the cell's lobes (the NEE direction and the sampled direction of a
// guided vertex): every lobe is fetched once. Same summation order as guidePdf.
PG_DEV void guidePdf2(const GuideDevice &G, uint32_t cell, float3 w1, float3 w2, float &p1, float &p2) {
    const float4 *L = G.lobes + (size_t)cell * G.K * 2;
    float s1 = 0, s2 = 0;
#ifndef PG_PDF2_UNROLL
#define PG_PDF2_UNROLL 2  // 2 beats 4 and 8 at 64 registers (fewer spills; the loads are L1/L2 hits)
#endif
    constexpr int kUnroll = PG_PDF2_UNROLL;
#pragma unroll kUnroll
    for (int k = 0; k < G.K; ++k) {
        const float4 a = __ldg(L + 2 * k), b = __ldg(L + 2 * k + 1);
        s1 += guideLobeTerm(a, b, w1);
        s2 += guideLobeTerm(a, b, w2);
    }
    p1 = s1;
    p2 = s2;
}

PG_DEV float3 guideSample(const GuideDevice &G, uint32_t cell, float u0, float u1, float u2) {
    const float4 *L = G.lobes + (size_t)cell * G.K * 2;
    int k = 0;
    float4 a = __ldg(L);
    while (k < G.K - 1 && u0 >= a.x) {
        u0 -= a.x;
        ++k;
        a = __ldg(L + 2 * k);
    }
    const float4 b = __ldg(L + 2 * k + 1);
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

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
    float4 *sRec;          // 32 B per sample = one 256-bit store / load: {position, weight} {direction, pdf}. One record, because the
                           // training update gathers samples in random order and every gathered 16-byte piece costs a 64-byte
                           // DRAM fetch (ncu: k_gather_partition moved 2.85 GB for 0.9 GB of samples with two separate arrays)
    float *sDist;
    uint32_t *sKey;        // guiding cell of the sample's vertex, as the shade stage looked it up (the binning key: no second tree walk)
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
        const F8 lb = ldg256(L + 2 * k);  // 32-byte lobe = one LDG.256
        s += guideLobeTerm(lb.a, lb.b, w);
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
        const F8 lb = ldg256(L + 2 * k);
        s1 += guideLobeTerm(lb.a, lb.b, w1);
        s2 += guideLobeTerm(lb.a, lb.b, w2);
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

// ------------------------------------------------------------------------------------------
// Warp-cooperative queries (north star, subsystem 2: "warp-cooperative component evaluation").
//
// A per-thread loop over the K lobes of the lane's own cell issues 2K scattered 16-byte loads per lane, and every such
// warp instruction touches up to 32 different cache lines: the L1TEX data pipe takes about one line per cycle, and ncu showed
// it 72-77 % busy on the guided bounces of k_shade with ~60 % of its wavefronts coming from these loads
// (profiles/r02_shade_coop_summary.txt, r02_shade_final_summary.txt: l1tex__data_pipe_lsu_wavefronts; L1-wavefront-bound, not DRAM-bound).
// Here the warp serves its lanes one path at a time: lane k fetches lobe k of THAT path's cell, so a load instruction covers
// 16 lanes x 32 B = 4 lines; K <= 16 uses the two half-warps for the two directions of a pdf query (or for two paths of a
// lobe selection), K <= 32 one lobe per lane. Idle lanes of a divergent warp (finished, parked or unguided paths) help.
// The sum over the lobes is a butterfly instead of the oracle's left-to-right loop: same terms, different rounding of the
// sum (<= K/2 ulp), inside the 1e-5 parity bar. All functions must be called by all 32 lanes of a converged warp;
// `sq` = kCoopFloat4PerWarp float4 of shared memory owned by the warp.
// ------------------------------------------------------------------------------------------
static constexpr unsigned kFullWarp = 0xffffffffu;

PG_DEV float4 coopLoad(const float4 *p, bool has) { return has ? __ldg(p) : make_float4(0.0f, 0.0f, 0.0f, 0.0f); }

// N values per lane, groups of 2N lanes: lane l of a group returns the group's sum of value (l >> 1). A butterfly that halves the
// number of live values at every level: 2N - 1 shuffles for N sums instead of N * log2(2N) -- shuffles go through the same
// LSU data pipe as the loads this scheme is meant to relieve.
template <int N>
PG_DEV float transposeReduce(float (&v)[N], unsigned laneInGroup) {
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool up = (laneInGroup & (unsigned)(2 * h)) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(kFullWarp, send, 2 * h);
        }
    }
    return v[0] + __shfl_xor_sync(kFullWarp, v[0], 1);
}

// pdf of two directions per lane with `want`: p1 = pdf(w1), p2 = pdf(w2) in that lane's cell.
// The queries are compacted by rank into the staging area and served eight at a time: eight independent 256-bit lobe loads in
// flight per lane, then one transposed butterfly for all eight.
static constexpr int kCoopFloat4PerWarp = 64 + 16;  // 32 queries x 2 float4 + 64 result floats
PG_DEV void guidePdf2Coop(const GuideDevice &G, float4 *sq, bool want, uint32_t cell, float3 w1, float3 w2, float &p1, float &p2) {
    const unsigned mask = __ballot_sync(kFullWarp, want);
    if (!mask) return;
    const unsigned lane = threadIdx.x & 31u;
    const int count = __popc(mask), rank = __popc(mask & ((1u << lane) - 1u));
    float *res = reinterpret_cast<float *>(sq + 64);
    if (want) {
        sq[2 * rank] = make_float4(w1.x, w1.y, w1.z, __uint_as_float(cell));
        sq[2 * rank + 1] = make_float4(w2.x, w2.y, w2.z, __uint_as_float(cell));
    }
    __syncwarp();
    if (G.K <= 16) {  // low half-warp: w1, high half-warp: w2; lane & 15 = lobe
        const unsigned half = lane >> 4, k = lane & 15u;
        const bool has = (int)k < G.K;
        for (int c0 = 0; c0 < count; c0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 q = sq[2 * min(c0 + i, count - 1) + half];
                v[i] = 0.0f;
                if (has) {
                    const F8 lb = ldg256(G.lobes + ((size_t)__float_as_uint(q.w) * G.K + k) * 2);
                    v[i] = guideLobeTerm(lb.a, lb.b, f3(q.x, q.y, q.z));
                }
            }
            const float sum = transposeReduce<8>(v, k);
            if (!(k & 1u)) res[2 * (c0 + (k >> 1)) + half] = sum;
        }
    } else {  // lane = lobe, both directions per lane
        const bool has = (int)lane < G.K;
        for (int c0 = 0; c0 < count; c0 += 8) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = min(c0 + i, count - 1);
                const float4 q1 = sq[2 * e], q2 = sq[2 * e + 1];
                v[i] = v[i + 8] = 0.0f;
                if (has) {
                    const F8 lb = ldg256(G.lobes + ((size_t)__float_as_uint(q1.w) * G.K + lane) * 2);
                    v[i] = guideLobeTerm(lb.a, lb.b, f3(q1.x, q1.y, q1.z));
                    v[i + 8] = guideLobeTerm(lb.a, lb.b, f3(q2.x, q2.y, q2.z));
                }
            }
            const float sum = transposeReduce<16>(v, lane);
            if (!(lane & 1u)) res[2 * (c0 + ((lane >> 1) & 7u)) + (lane >> 4)] = sum;
        }
    }
    __syncwarp();
    if (want) {
        p1 = res[2 * rank];
        p2 = res[2 * rank + 1];
    }
    __syncwarp();
}

// Lobe selection by the weights pi_k (the loop of guideSample): returns, for lanes with `want`, the first k whose
// inclusive weight prefix exceeds u0 (K - 1 if none does).
PG_DEV int guideSelectCoop(const GuideDevice &G, float4 *sq, bool want, uint32_t cell, float u0) {
    unsigned todo = __ballot_sync(kFullWarp, want);
    int sel = 0;
    if (!todo) return sel;
    const unsigned lane = threadIdx.x & 31u;
    float2 *sp = reinterpret_cast<float2 *>(sq);
    sp[lane] = make_float2(u0, __uint_as_float(cell));
    __syncwarp();
    if (G.K <= 16) {  // two paths per pass, one per half-warp
        const unsigned half = lane >> 4, k = lane & 15u;
        const bool has = (int)k < G.K;
        while (todo) {
            const int j0 = __ffs(todo) - 1;
            todo &= todo - 1;
            const int j1 = todo ? __ffs(todo) - 1 : j0;
            todo &= todo - 1;
            const float2 q = sp[half ? j1 : j0];
            float c = has ? __ldg(&G.lobes[((size_t)__float_as_uint(q.y) * G.K + k) * 2].x) : 0.0f;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const float v = __shfl_up_sync(kFullWarp, c, o, 16);
                if ((int)k >= o) c += v;
            }
            const unsigned m = __ballot_sync(kFullWarp, has && q.x < c);
            const unsigned m0 = m & 0xFFFFu, m1 = m >> 16;
            if ((int)lane == j0) sel = m0 ? __ffs(m0) - 1 : G.K - 1;
            if ((int)lane == j1) sel = m1 ? __ffs(m1) - 1 : G.K - 1;
        }
    } else {
        const bool has = (int)lane < G.K;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const float2 q = sp[j];
            float c = has ? __ldg(&G.lobes[((size_t)__float_as_uint(q.y) * G.K + lane) * 2].x) : 0.0f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(kFullWarp, c, o);
                if ((int)lane >= o) c += v;
            }
            const unsigned m = __ballot_sync(kFullWarp, has && q.x < c);
            if ((int)lane == j) sel = m ? __ffs(m) - 1 : G.K - 1;
        }
    }
    __syncwarp();
    return sel;
}

// direction from lobe k of the cell (the second half of guideSample)
PG_DEV float3 guideSampleLobe(const GuideDevice &G, uint32_t cell, int k, float u1, float u2) {
    const F8 lb = ldg256(G.lobes + ((size_t)cell * G.K + k) * 2);
    const float4 a = lb.a, b = lb.b;
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

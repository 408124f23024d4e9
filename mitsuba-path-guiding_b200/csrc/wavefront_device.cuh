// wavefront_device.cuh -- device helpers shared by the shade kernels (kernels.cu, volpath.cu):
// queue appends, counter reductions, path termination and training-sample emission.
#pragma once
#include "guiding_device.cuh"
#include "wavefront.cuh"

namespace pg {

PG_DEV uint32_t laneId() { return threadIdx.x & 31u; }

// Warp-aggregated append: returns the destination index for lanes with `pred`, one atomic per warp.
PG_DEV uint32_t warpAppend(uint32_t *counter, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    uint32_t base = 0;
    if (laneId() == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << laneId()) - 1u));
}

// Block-aggregated append for the two queues and the training-sample buffer at once: ONE global atomic per block
// and counter instead of one per warp (the per-warp version serialised ~130k same-address atomics per bounce at L2
// and was 35% of k_shade's stall samples, profiles/r01_v1_summary.txt). Must be called by all threads of the block.
// `smem` holds 2 * 3 * (warps + 1) words and is double-buffered by `parity` (toggled by the caller every loop
// iteration), which saves the trailing barrier. countC = number of items this thread appends to counter C;
// inclC = inclusive prefix of countC inside the warp, warpTotalC = the warp's sum, baseC = index of the warp's first item.
struct AppendResult {
    uint32_t idxA, idxB, baseC, inclC, warpTotalC;
};
PG_DEV AppendResult blockAppend3(uint32_t *counterA, bool predA, uint32_t *counterB, bool predB, uint32_t *counterC, uint32_t countC,
                                 uint32_t *smem, uint32_t parity) {
    const uint32_t warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const unsigned maskA = __ballot_sync(0xffffffffu, predA), maskB = __ballot_sync(0xffffffffu, predB);
    uint32_t incl = countC;
    if (counterC) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)laneId() >= o) incl += v;
        }
    }
    const uint32_t warpTotal = counterC ? __shfl_sync(0xffffffffu, incl, 31) : 0u;
    uint32_t *cntA = smem + parity * 3 * (nWarps + 1), *cntB = cntA + nWarps + 1, *cntC = cntB + nWarps + 1;
    if (laneId() == 0) {
        cntA[warp] = __popc(maskA);
        cntB[warp] = __popc(maskB);
        cntC[warp] = warpTotal;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t *cnt = threadIdx.x == 0 ? cntA : (threadIdx.x == 1 ? cntB : cntC);
        uint32_t *counter = threadIdx.x == 0 ? counterA : (threadIdx.x == 1 ? counterB : counterC);
        uint32_t total = 0;
        for (uint32_t w = 0; w < nWarps; ++w) {
            const uint32_t c = cnt[w];
            cnt[w] = total;  // exclusive prefix
            total += c;
        }
        cnt[nWarps] = (total && counter) ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    AppendResult r;
    r.idxA = cntA[nWarps] + cntA[warp] + __popc(maskA & ((1u << laneId()) - 1u));
    r.idxB = cntB[nWarps] + cntB[warp] + __popc(maskB & ((1u << laneId()) - 1u));
    r.baseC = cntC[nWarps] + cntC[warp];
    r.inclC = incl;
    r.warpTotalC = warpTotal;
    return r;
}

PG_DEV void warpAddU64(unsigned long long *counter, unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (laneId() == 0 && v) atomicAdd(counter, v);
}

// Training samples of a finished path: for every recorded vertex the incident-radiance estimate along the
// sampled direction is everything the path gathered after the vertex divided by the throughput right after it;
// sample weight = avg_rgb(estimate) / pdf. Called by whole warps; the output range [base, base + total) of the warp was
// reserved by blockAppend3 (nMine = this lane's vertex count if its path just finished, incl = its inclusive prefix).
PG_DEV void emitTrainingSamples(const GuideDevice &G, uint32_t nMine, uint32_t incl, uint32_t total, uint32_t base, uint32_t slot,
                                    float3 Lfinal) {
    if (total == 0) return;
    // The warp emits its `total` samples cooperatively, 32 at a time: item t belongs to the first lane whose inclusive
    // prefix exceeds t (binary search over the prefix with shuffles), so the vertex loads/stores run with all lanes
    // busy instead of a serial per-path loop with a handful of finished lanes.
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + laneId();
        uint32_t owner = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t v = __shfl_sync(0xffffffffu, incl, (owner + o - 1) & 31);
            if (v <= t) owner += o;
        }
        owner &= 31u;
        const uint32_t oIncl = __shfl_sync(0xffffffffu, incl, owner), oN = __shfl_sync(0xffffffffu, nMine, owner);
        const uint32_t oSlot = __shfl_sync(0xffffffffu, slot, owner);
        const float Lx = __shfl_sync(0xffffffffu, Lfinal.x, owner), Ly = __shfl_sync(0xffffffffu, Lfinal.y, owner),
                    Lz = __shfl_sync(0xffffffffu, Lfinal.z, owner);
        const uint32_t dst = base + t;
        if (t >= total || dst >= G.sCapacity) continue;
        const uint32_t v = t - (oIncl - oN);
        const size_t vi = (size_t)oSlot * G.maxVerts + v;
        const float4 *rec = G.vRec + 4 * vi;
        const F8 ra = ldStream256(rec), rb = ldStream256(rec + 2);  // the two 32-byte sectors of the vertex record
        const float4 p = ra.a, d4 = ra.b, T = rb.a, Lk = rb.b;
        const float4 d = make_float4(d4.x, d4.y, d4.z, T.w);  // direction, distance to the next hit
        const float3 diff = f3(Lx, Ly, Lz) - f3(Lk.x, Lk.y, Lk.z);
        const float ex = T.x > 0 ? diff.x / T.x : 0.0f, ey = T.y > 0 ? diff.y / T.y : 0.0f, ez = T.z > 0 ? diff.z / T.z : 0.0f;
        float w = ((ex + ey + ez) * (1.0f / 3.0f)) / p.w;
        if (!isfinite(w) || w < 0) w = 0.0f;
        stStream256(G.sRec + 2 * (size_t)dst, make_float4(p.x, p.y, p.z, w), make_float4(d.x, d.y, d.z, p.w));
        stStream(G.sDist + dst, d.w);
        stStream(G.sKey + dst, __float_as_uint(d4.w));  // the vertex' guiding cell rides in the record's spare word
    }
}

PG_DEV void guideVertexOpen(const GuideDevice &G, uint32_t slot, uint32_t v, float3 p, float pdf, float3 wo, uint32_t cell) {
    float4 *rec = G.vRec + 4 * ((size_t)slot * G.maxVerts + v);
    stStream256(rec, make_float4(p.x, p.y, p.z, pdf), make_float4(wo.x, wo.y, wo.z, __uint_as_float(cell)));
}
// thr = the throughput the path record carried into this bounce (before Russian roulette rescales it) = the
// throughput right after the vertex; L = radiance gathered so far (the vertex' NEE has landed by now)
PG_DEV void guideVertexClose(const GuideDevice &G, uint32_t slot, uint32_t v, float3 thr, float dist, float3 L) {
    float4 *rec = G.vRec + 4 * ((size_t)slot * G.maxVerts + v);
    stStream256(rec + 2, make_float4(thr.x, thr.y, thr.z, dist), make_float4(L.x, L.y, L.z, 0.0f));
}

// A terminated path leaves its final sample value in the splat buffer (indexed by the path's slot, i.e.
// in pixel order); k_splat rasterises the whole batch afterwards with all lanes busy.
PG_DEV void finishPath(const ShadeArgs &A, uint32_t slot, float4 pos4, float3 L) {
    if (A.radianceOut) {
        A.radianceOut[3 * (size_t)slot + 0] = L.x;
        A.radianceOut[3 * (size_t)slot + 1] = L.y;
        A.radianceOut[3 * (size_t)slot + 2] = L.z;
    } else {
        // one whole 32-byte sector per sample, one 256-bit store
        stStream256(A.splat + 2 * (size_t)slot, make_float4(pos4.x, pos4.y, L.x, L.y), make_float4(L.z, 0.0f, 0.0f, 0.0f));
    }
}

}  // namespace pg

// guiding.cu -- training of the guiding field (north-star subsystem 3) and the standalone query kernel:
//   k_guide_cells      sample position -> cell index (kd-tree walk)
//   k_radix_*          hand-written stable LSD radix sort (8-bit digits) of (cell, sample index) pairs = binning;
//                      bit-exact against the oracle's stable counting sort (oracle_guiding.h: guideBin)
//   k_estep            weighted-EM E-step, one warp per chunk of one cell's samples: samples streamed through shared memory
//                      by bulk async copies (cp.async.bulk + mbarrier ring), lobes in shared memory, lane = sample,
//                      per-lane sufficient statistics in registers, butterfly reduction per chunk
//   k_gather_partition gather into sorted order + per-chunk partition (usable weights first) + per-chunk position
//                      moments (split statistics), once per training update
//   k_reduce_partials  per-cell sum of the chunk partials in a fixed order (deterministic)
//   k_mstep            M-step with decayed running statistics and MAP priors, one warp per cell
//   k_guide_query      pdf / sample of the field at arbitrary points (b200pg_k_vmm_pdf_sample)
// This is the repo's own algorithm (no guiding code exists in the reference snapshot); the CPU
// statement it is tested against is oracle/oracle_guiding.h.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "guiding_device.cuh"
#include "guiding_host.h"

namespace pg {

static constexpr int kSortThreads = 256;
static constexpr int kSortRounds = 8;
static constexpr int kSortTile = kSortThreads * kSortRounds;
static constexpr int kChunk = 2048;  // samples per E-step work item (one warp)

__device__ __forceinline__ uint32_t lane() { return threadIdx.x & 31u; }

__global__ void __launch_bounds__(256) k_guide_cells(GuideDevice G, const float4 *__restrict__ sRec, uint32_t n,
                                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = sRec[2 * (size_t)i];
        keys[i] = guideLookup(G, f3(p.x, p.y, p.z));
        vals[i] = i;
    }
}

// binning keys straight from the recorded samples (see GuidingHost::keysValid): the tree walk already happened in the shade stage
__global__ void __launch_bounds__(256) k_keys_from_samples(const uint32_t *__restrict__ sKey, uint32_t n, uint32_t *__restrict__ keys,
                                                           uint32_t *__restrict__ vals) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        keys[i] = sKey[i];
        vals[i] = i;
    }
}

// ---- radix sort pass: per-block digit histogram, laid out digit-major (hist[digit * nBlocks + block])
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift,
                                                             uint32_t *__restrict__ hist, uint32_t nBlocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
    for (int r = 0; r < kSortRounds; ++r) {
        const uint32_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over all (digit, block) counters in three small kernels: per-segment scan + segment totals,
// scan of the totals (one block), add-back. Integer sums: the result does not depend on the schedule.
static constexpr int kScanSeg = 2048;  // entries per block (256 threads x 8)
__global__ void __launch_bounds__(256) k_scan_segments(uint32_t *__restrict__ data, uint32_t m, uint32_t *__restrict__ totals) {
    __shared__ uint32_t warpSums[8];
    const uint32_t base = blockIdx.x * kScanSeg + threadIdx.x * 8;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = base + i < m ? data[base + i] : 0u;
        sum += v[i];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane() >= o) incl += t;
    }
    if (lane() == 31) warpSums[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t warpBase = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) warpBase += warpSums[w];
    uint32_t run = warpBase + incl - sum;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (base + i < m) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 255) totals[blockIdx.x] = run;
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t *__restrict__ totals, uint32_t n) {
    // n <= a few thousand: one block, chunked sequential carry
    __shared__ uint32_t carry;
    __shared__ uint32_t warpSums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? totals[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane() >= o) incl += t;
        }
        if (lane() == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t warpBase = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) warpBase += warpSums[w];
        const uint32_t c = carry;
        if (i < n) totals[i] = c + warpBase + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warpBase + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t *__restrict__ data, uint32_t m, const uint32_t *__restrict__ totals) {
    const uint32_t add = totals[blockIdx.x];
    const uint32_t base = blockIdx.x * kScanSeg + threadIdx.x * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (base + i < m) data[base + i] += add;
}

// stable scatter: keys are visited in index order (round by round, warp by warp), so equal digits keep their order
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const uint32_t *__restrict__ keysIn, const uint32_t *__restrict__ valsIn,
                                                                uint32_t *__restrict__ keysOut, uint32_t *__restrict__ valsOut,
                                                                uint32_t n, int shift, const uint32_t *__restrict__ hist,
                                                                uint32_t nBlocks) {
    __shared__ uint32_t offs[256];
    offs[threadIdx.x] = hist[threadIdx.x * nBlocks + blockIdx.x];
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
    const uint32_t warp = threadIdx.x >> 5;
    for (int r = 0; r < kSortRounds; ++r) {
        const uint32_t i = base + r * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keysIn[i] : 0u;
        const uint32_t d = (key >> shift) & 255u;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        unsigned peers = 0;
        if (valid) peers = __match_any_sync(act, d);
        const uint32_t rankInWarp = __popc(peers & ((1u << lane()) - 1u));
        const bool leader = valid && rankInWarp == 0;
        uint32_t pos = 0;
        for (uint32_t w = 0; w < kSortThreads / 32; ++w) {
            if (warp == w) {
                if (valid) pos = offs[d] + rankInWarp;
                __syncwarp();
                if (leader) offs[d] += __popc(peers);
            }
            __syncthreads();
        }
        if (valid) {
            keysOut[pos] = key;
            valsOut[pos] = valsIn[i];
        }
    }
}

// first index of every cell that occurs in the sorted key array (no atomics: one writer per cell)
__global__ void __launch_bounds__(256) k_cell_starts(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ start) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        if (i == 0 || keys[i - 1] != k) start[k] = i;
    }
}

// ---- block-wide helpers for the single-block bookkeeping kernels (1024 threads) ---------------------------------
__device__ __forceinline__ uint32_t blockExclusiveScan1024(uint32_t v, uint32_t *smem33, uint32_t &total) {
    // smem33: 33 words. Returns the exclusive prefix of v over the block's 1024 threads; total = block sum.
    const uint32_t ln = lane(), wp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)ln >= o) incl += t;
    }
    __syncthreads();
    if (ln == 31) smem33[wp] = incl;
    __syncthreads();
    if (wp == 0) {
        uint32_t w = smem33[ln], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)ln >= o) wi += t;
        }
        smem33[ln] = wi - w;
        if (ln == 31) smem33[32] = wi;
    }
    __syncthreads();
    total = smem33[32];
    return smem33[wp] + incl - v;
}

// Work list of one training update, built on the device (no host round trip between binning and the EM iterations):
//   offsets[c]  = first sorted sample of cell c (suffix minimum of the cell starts; empty cells inherit the next start)
//   workOfs[c]  = first work item of cell c, items = chunks of <= kChunk consecutive samples of ONE cell
//   counts[2]   = number of work items
// One block of 1024 threads (each walks ceil(nCells / 1024) consecutive cells); nCells <= kCommMaxCells = 65536.
__global__ void __launch_bounds__(1024) k_build_work(const uint32_t *__restrict__ cellStart, uint32_t nCells, uint32_t n, uint32_t chunk,
                                                     uint32_t *__restrict__ offsets, uint32_t *__restrict__ workOfs,
                                                     uint32_t *__restrict__ counts) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry;
    const uint32_t per = (nCells + 1023) / 1024;  // consecutive cells per thread
    const uint32_t c0 = threadIdx.x * per, c1 = min(c0 + per, nCells);
    // ---- suffix minimum: thread-local, then across threads (reverse inclusive scan with min)
    uint32_t local = 0xFFFFFFFFu;
    for (uint32_t c = c1; c-- > c0;) local = min(local, cellStart[c]);
    // reverse exclusive "min-scan" over threads via shared memory (1024 values): simple log-step
    __shared__ uint32_t mins[1024];
    mins[threadIdx.x] = local;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint32_t v = mins[threadIdx.x];
        if (threadIdx.x + o < 1024) v = min(v, mins[threadIdx.x + o]);
        __syncthreads();
        mins[threadIdx.x] = v;
        __syncthreads();
    }
    uint32_t after = threadIdx.x + 1 < 1024 ? mins[threadIdx.x + 1] : 0xFFFFFFFFu;  // min over all later threads
    after = min(after, n);
    for (uint32_t c = c1; c-- > c0;) {
        after = min(after, cellStart[c]);
        offsets[c] = after;
    }
    if (threadIdx.x == 0) {
        offsets[nCells] = n;
        carry = 0;
    }
    __syncthreads();
    // ---- chunks per cell -> exclusive scan
    uint32_t mine = 0;
    for (uint32_t c = c0; c < c1; ++c) mine += (offsets[c + 1] - offsets[c] + chunk - 1) / chunk;
    uint32_t total;
    uint32_t run = blockExclusiveScan1024(mine, sm, total);
    for (uint32_t c = c0; c < c1; ++c) {
        workOfs[c] = run;
        run += (offsets[c + 1] - offsets[c] + chunk - 1) / chunk;
    }
    if (threadIdx.x == 0) {
        workOfs[nCells] = total;
        counts[2] = total;
    }
}

__global__ void __launch_bounds__(256) k_fill_work(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ workOfs, uint32_t nCells,
                                                   uint32_t chunk, const uint32_t *__restrict__ counts, uint4 *__restrict__ work) {
    const uint32_t nWork = counts[2];
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nWork; w += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = nCells;  // last cell with workOfs[c] <= w
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (workOfs[mid] <= w) lo = mid; else hi = mid;
        }
        const uint32_t b = offsets[lo] + (w - workOfs[lo]) * chunk;
        work[w] = make_uint4(lo, b, min(b + chunk, offsets[lo + 1]), 0u);
    }
}

// Per-cell sums of the chunks' cell statistics (count, first / second position moments) into the cell slots of the stats buffer, for
// a further split level: one warp per cell, lanes 0..7 = the eight slots, chunks in index order, double accumulators.
__global__ void __launch_bounds__(256) k_cell_moments(const float *__restrict__ partials, const uint32_t *__restrict__ workOfs,
                                                      uint32_t nCells, int K, int stride, float *__restrict__ stats) {
    const uint32_t warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    const int e = (int)lane();
    for (uint32_t c = warpGlobal; c < nCells; c += nWarps) {
        if (e >= 8) continue;
        double acc = 0.0;
        if (e != 1)  // slot 1 (sum of weights) belongs to the E-step, which does not run for a split level
            for (uint32_t w = workOfs[c]; w < workOfs[c + 1]; ++w) acc += (double)partials[(size_t)w * stride + 4 * K + e];
        stats[(size_t)c * stride + 4 * K + e] = (float)acc;
    }
}

// Spatial refinement on the device (same rule and arithmetic as oracle_guiding.h guideSplit / the former host code):
// fold this update's cell statistics into the running headers, split every cell whose running sample count exceeds the
// threshold at the mean sample position along the axis of largest variance; cells are visited in index order, the
// left child keeps the parent's index, the right child gets index nCells0 + (number of splitting cells before it).
// One block of 1024 threads. counts = {nCells, nNodes, nWork, -}.
__global__ void __launch_bounds__(1024) k_split(uint4 *__restrict__ nodes, float4 *__restrict__ lobes, float4 *__restrict__ lobeStats,
                                                float2 *__restrict__ cells, uint32_t *__restrict__ cellLeaf, const float *__restrict__ stats,
                                                uint32_t *__restrict__ counts, int K, int stride, float maxCellSamples, uint32_t maxCells,
                                                int fold) {
    // fold = 0: a further split level of the same update (GuidingHost::end): the headers already hold this update's samples, and
    // `stats` carries only the cell slots (count, position moments) of the samples binned into the tree as it is now
    __shared__ uint32_t sm[33];
    const uint32_t nc0 = counts[0], nn0 = counts[1];
    const uint32_t per = (nc0 + 1023) / 1024;
    const uint32_t c0 = threadIdx.x * per, c1 = min(c0 + per, nc0);
    uint32_t mine = 0;
    for (uint32_t c = c0; c < c1; ++c) {
        const float *cs = stats + (size_t)stride * c + (size_t)K * 4;
        float2 h = cells[c];
        if (fold) {
            h.x = kGuideDecay * h.x + cs[0];
            h.y = kGuideDecay * h.y + cs[1];
            cells[c] = h;
        }
        bool split = false;
        if (h.x > maxCellSamples && cs[0] >= 2) {
            double best = 0.0;
            for (int a = 0; a < 3; ++a) {
                const float mean = cs[2 + a] / cs[0];
                const double var = (double)cs[5 + a] / (double)cs[0] - (double)mean * (double)mean;
                if (a == 0 || var > best) best = var;
            }
            split = best > 0;
        }
        mine += split ? 1u : 0u;
    }
    uint32_t total;
    uint32_t rank = blockExclusiveScan1024(mine, sm, total);
    const uint32_t room = maxCells > nc0 ? maxCells - nc0 : 0u;
    for (uint32_t c = c0; c < c1; ++c) {
        const float *cs = stats + (size_t)stride * c + (size_t)K * 4;
        const float2 h = cells[c];
        if (!(h.x > maxCellSamples && cs[0] >= 2)) continue;
        const float n = cs[0];
        float mean[3];
        double var[3];
        for (int a = 0; a < 3; ++a) {
            mean[a] = cs[2 + a] / n;
            var[a] = (double)cs[5 + a] / (double)n - (double)mean[a] * (double)mean[a];
        }
        int axis = 0;
        if (var[1] > var[axis]) axis = 1;
        if (var[2] > var[axis]) axis = 2;
        if (!(var[axis] > 0)) continue;
        const uint32_t r = rank++;
        if (r >= room) continue;  // field is at capacity
        const uint32_t leaf = cellLeaf[c], left = nn0 + 2 * r, newCell = nc0 + r;
        nodes[left] = make_uint4(3u, 0u, c, 0u);
        nodes[left + 1] = make_uint4(3u, 0u, newCell, 0u);
        nodes[leaf] = make_uint4((uint32_t)axis, __float_as_uint(mean[axis]), left, 0u);
        cellLeaf[c] = left;
        cellLeaf[newCell] = left + 1;
        const float2 hh = make_float2(h.x * 0.5f, h.y * 0.5f);
        cells[c] = hh;
        cells[newCell] = hh;
        for (int k = 0; k < K; ++k) {
            float4 st = lobeStats[(size_t)c * K + k];
            st.x *= 0.5f; st.y *= 0.5f; st.z *= 0.5f; st.w *= 0.5f;
            lobeStats[(size_t)c * K + k] = st;
            lobeStats[(size_t)newCell * K + k] = st;
            lobes[((size_t)newCell * K + k) * 2] = lobes[((size_t)c * K + k) * 2];
            lobes[((size_t)newCell * K + k) * 2 + 1] = lobes[((size_t)c * K + k) * 2 + 1];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = min(total, room);
        counts[0] = nc0 + done;
        counts[1] = nn0 + 2 * done;
    }
}

// ---- E-step ------------------------------------------------------------------------------------------
// One warp per work item (a chunk of <= kChunk consecutive samples of ONE cell, gathered into sorted order
// beforehand); lane = sample. The cell's K lobes sit in shared memory (broadcast reads), every lane evaluates all K
// lobes for its sample (same summation order as the oracle: total += p_k, k ascending) and keeps the 4 K
// sufficient statistics of its samples in registers; one butterfly reduction per work item at the end.
// ~9 issued instructions per sample and lobe-free shuffles in the inner loop, against ~35 for the earlier
// lane-per-lobe formulation (profiles/r01_v3_summary.txt).
// ---- TMA-style bulk copies (cp.async.bulk, 1-D) + mbarrier, the Blackwell way to stream a contiguous array through
// shared memory without staging registers
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smemAddr(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

static constexpr int kEStages = 4;    // ring depth per warp
static constexpr int kETile = 64;     // samples per stage (1 KB of positions + 1 KB of directions)
#ifndef PG_ESTEP_PAIR
#define PG_ESTEP_PAIR 1
#endif

// E-step. One warp per work item (<= kChunk consecutive samples of ONE cell, usable weights first); lane = sample.
// The warp streams its samples through a private ring of kEStages shared-memory tiles filled by 1-D bulk async copies
// (one elected lane arms the stage's mbarrier with the byte count and issues the two copies; all lanes wait on the
// barrier's phase), so 16 KB per warp are in flight ahead of the arithmetic instead of one sample per lane in registers.
template <int KMAX>
__global__ void __launch_bounds__(128) k_estep(GuideDevice G, const float4 *__restrict__ sPos, const float4 *__restrict__ sDir,
                                               const uint4 *__restrict__ work, uint32_t *__restrict__ counts,
                                               float *__restrict__ partials, int stride) {
    __shared__ float4 sLobe[4][KMAX * 2];
    __shared__ __align__(128) float4 sTile[4][kEStages][2][kETile];
    __shared__ __align__(8) uint64_t sBar[4][kEStages];
    const uint32_t nWork = counts[2];
    const uint32_t warp = threadIdx.x >> 5, ln = lane();
    const int K = G.K;
    float4 *myLobes = sLobe[warp];
    if (ln == 0) {
        for (int s = 0; s < kEStages; ++s) mbarInit(&sBar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;  // bit s = phase the next wait on stage s expects
    // dynamic work fetch (counts[3], zeroed before the launch): chunks differ in size once cells hold fewer samples than a chunk
    // (large fields), and a static round-robin leaves the kernel waiting for the warps that drew the full ones
    while (true) {
        uint32_t w = 0;
        if (ln == 0) w = atomicAdd(&counts[3], 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nWork) break;
        const uint4 item = work[w];
        __syncwarp();
        for (int k = (int)ln; k < 2 * K; k += 32) myLobes[k] = __ldg(G.lobes + (size_t)item.x * K * 2 + k);
        float S[KMAX], Rx[KMAX], Ry[KMAX], Rz[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) S[k] = Rx[k] = Ry[k] = Rz[k] = 0.0f;
        float cW = 0.0f;
        // samples with a usable weight were moved to the front of the chunk by k_gather_partition (item.w of them)
        const uint32_t nGood = item.w, nTiles = (nGood + kETile - 1) / kETile;
        auto issue = [&](uint32_t t) {  // lane 0 only
            const int st = (int)(t % kEStages);
            const uint32_t cnt = min((uint32_t)kETile, nGood - t * kETile), bytes = cnt * (uint32_t)sizeof(float4);
            mbarExpectTx(&sBar[warp][st], 2 * bytes);
            bulkLoad(&sTile[warp][st][0][0], sPos + item.y + (size_t)t * kETile, bytes, &sBar[warp][st]);
            bulkLoad(&sTile[warp][st][1][0], sDir + item.y + (size_t)t * kETile, bytes, &sBar[warp][st]);
        };
        if (ln == 0)
            for (uint32_t t = 0; t < min(nTiles, (uint32_t)kEStages); ++t) issue(t);
        __syncwarp();
        for (uint32_t t = 0; t < nTiles; ++t) {
            const int st = (int)(t % kEStages);
            mbarWait(&sBar[warp][st], (parity >> st) & 1u);
            parity ^= 1u << st;
#if PG_ESTEP_PAIR
            // the lane's two samples of the tile share every lobe fetch from shared memory (the kernel issued one LDS.128 + one
            // LDS.64 per lobe and SAMPLE; the shared-memory pipe, not the FP32 pipe, was the busier one). Same terms, same order of
            // the additions into every accumulator (sample 0 before sample 1) as the one-sample-at-a-time loop.
            static_assert(kETile == 64, "two samples per lane and tile");
            float3 dd[2];
            float sw[2], total[2] = {0.0f, 0.0f};
            bool live[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t local = r * 32 + ln;
                live[r] = t * kETile + local < nGood;
                const float4 p = sTile[warp][st][0][local], d = sTile[warp][st][1][local];
                sw[r] = p.w;
                dd[r] = f3(d.x, d.y, d.z);
            }
            float pk[2][KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                pk[0][k] = pk[1][k] = 0.0f;
                if (k < K) {
                    const float4 la = myLobes[2 * k], lb = myLobes[2 * k + 1];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        pk[r][k] = guideLobeTerm(la, lb, dd[r]);
                        total[r] += pk[r][k];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (!live[r] || !(total[r] > 0) || !isfinite(total[r])) continue;
                cW += sw[r];
                const float inv = 1.0f / total[r];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < K) {
                        const float g = sw[r] * (pk[r][k] * inv);
                        S[k] += g;
                        Rx[k] += g * dd[r].x;
                        Ry[k] += g * dd[r].y;
                        Rz[k] += g * dd[r].z;
                    }
                }
            }
#else
#pragma unroll
            for (int r = 0; r < kETile / 32; ++r) {
                const uint32_t local = r * 32 + ln;
                if (t * kETile + local >= nGood) break;
                const float4 p = sTile[warp][st][0][local], d = sTile[warp][st][1][local];
                const float sw = p.w;
                float pk[KMAX];
                float total = 0.0f;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    pk[k] = 0.0f;
                    if (k < K) {
                        const float4 la = myLobes[2 * k], lb = myLobes[2 * k + 1];
                        pk[k] = guideLobeTerm(la, lb, f3(d.x, d.y, d.z));
                        total += pk[k];
                    }
                }
                if (!(total > 0) || !isfinite(total)) continue;
                cW += sw;
                const float inv = 1.0f / total;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < K) {
                        const float g = sw * (pk[k] * inv);
                        S[k] += g;
                        Rx[k] += g * d.x;
                        Ry[k] += g * d.y;
                        Rz[k] += g * d.z;
                    }
                }
            }
#endif
            __syncwarp();  // every lane has read the stage: it can be refilled
            if (ln == 0 && t + kEStages < nTiles) issue(t + kEStages);
        }
        // butterfly reduction over the 32 lanes (fixed order -> deterministic); lane 0 holds the sums
        float *out = partials + (size_t)w * stride;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                float a = S[k], b = Rx[k], c = Ry[k], e = Rz[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    e += __shfl_xor_sync(0xffffffffu, e, o);
                }
                if (ln == 0) *reinterpret_cast<float4 *>(out + 4 * k) = make_float4(a, b, c, e);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cW += __shfl_xor_sync(0xffffffffu, cW, o);
        if (ln == 0) out[4 * K + 1] = cW;
    }
}


// Gather into sorted order, one warp per work chunk, with a stable partition inside the chunk: samples whose weight
// is positive and finite go to the front (count -> work[w].w), the others to the back in reverse order. Zero-weight
// samples (paths that found no light) only matter for the cell's sample count and position moments, so the E-step
// iterates over the dense front part with all lanes busy.
__global__ void __launch_bounds__(128) k_gather_partition(const float4 *__restrict__ sRec,
                                                          const uint32_t *__restrict__ perm, uint4 *__restrict__ work,
                                                          const uint32_t *__restrict__ counts, float4 *__restrict__ oPos,
                                                          float4 *__restrict__ oDir, float *__restrict__ partials, int stride, int K) {
    const uint32_t warp = threadIdx.x >> 5, ln = lane();
    const uint32_t nWork = counts[2];
    for (uint32_t w = blockIdx.x * 4 + warp; w < nWork; w += gridDim.x * 4) {
        const uint4 item = work[w];
        uint32_t nGood = 0, nBad = 0;
        float m[6] = {0, 0, 0, 0, 0, 0};  // position moments of the chunk (split rule), all samples
        for (uint32_t j0 = item.y; j0 < item.z; j0 += 32) {
            const uint32_t j = j0 + ln;
            const bool valid = j < item.z;
            float4 p = make_float4(0, 0, 0, 0), d = p;
            if (valid) {
                const uint32_t i = perm[j];
                const F8 r = ldStream256(sRec + 2 * (size_t)i);  // one 32-byte record per gathered sample
                p = r.a;
                d = r.b;
                m[0] += p.x; m[1] += p.y; m[2] += p.z;
                m[3] += p.x * p.x; m[4] += p.y * p.y; m[5] += p.z * p.z;
            }
            const bool good = valid && p.w > 0 && isfinite(p.w);
            const unsigned gm = __ballot_sync(0xffffffffu, good), bm = __ballot_sync(0xffffffffu, valid && !good);
            const unsigned lt = (1u << ln) - 1u;
            if (valid) {
                const uint32_t dst = good ? item.y + nGood + __popc(gm & lt) : item.z - 1 - (nBad + __popc(bm & lt));
                oPos[dst] = p;
                oDir[dst] = d;
            }
            nGood += __popc(gm);
            nBad += __popc(bm);
        }
        if (ln == 0) work[w].w = nGood;
        // moments: per-lane float partial sums (<= 64 samples each), combined across the lanes in double; written into the
        // cell-statistics slots of the partials buffer that k_estep leaves alone (slot 1, the weight sum, is k_estep's)
        double acc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            acc[i] = (double)m[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (ln == 0) {
            float *c = partials + (size_t)w * stride + 4 * K;
            c[0] = (float)(item.z - item.y);
#pragma unroll
            for (int i = 0; i < 6; ++i) c[2 + i] = (float)acc[i];
        }
    }
}

// work items of one cell are consecutive: workOfs[cell] .. workOfs[cell + 1]
__global__ void __launch_bounds__(256) k_reduce_partials(const float *__restrict__ partials, const uint32_t *__restrict__ workOfs,
                                                         uint32_t nCells, int stride, float *__restrict__ stats) {
    const size_t total = (size_t)nCells * stride;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t cell = (uint32_t)(i / stride), e = (uint32_t)(i % stride);
        double acc = 0.0;
        for (uint32_t w = workOfs[cell]; w < workOfs[cell + 1]; ++w) acc += (double)partials[(size_t)w * stride + e];
        stats[i] = (float)acc;
    }
}

// ---- M-step: lane k = lobe k -----------------------------------------------------------------------------
// PG_MSTEP_BODY: the refit of one cell by one warp from `st` (the cell's summed statistics) -- shared by both kernels.
#define PG_MSTEP_BODY(c, st)                                                                      \
    {                                                                                             \
        float4 a = make_float4(0, 0, 0, 0), b = a, s = a;                                         \
        float S = 0, R0 = 0, R1 = 0, R2 = 0;                                                      \
        if (k < K) {                                                                              \
            const float4 *L = lobes + ((size_t)(c) * K + k) * 2;                                  \
            a = L[0];                                                                             \
            b = L[1];                                                                             \
            s = lobeStats[(size_t)(c) * K + k];                                                   \
            S = kGuideDecay * s.x + (st)[4 * k];                                                  \
            R0 = kGuideDecay * s.y + (st)[4 * k + 1];                                             \
            R1 = kGuideDecay * s.z + (st)[4 * k + 2];                                             \
            R2 = kGuideDecay * s.w + (st)[4 * k + 3];                                             \
        }                                                                                         \
        float sumS = S;                                                                           \
        _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) sumS += __shfl_xor_sync(0xffffffffu, sumS, o); \
        if (k < K) {                                                                              \
            if (sumS > 0 && isfinite(sumS)) {                                                     \
                const float prior = kGuidePriorWeight * sumS / (float)K;                          \
                const float denom = 1.0f / (sumS + (float)K * prior);                             \
                a.x = (S + prior) * denom;                                                        \
                const float rl = sqrtf(R0 * R0 + R1 * R1 + R2 * R2);                              \
                float rbar = (rl + prior * kGuidePriorMeanCos) / (S + prior);                     \
                rbar = fminf(rbar, 0.9999f);                                                      \
                const float kappa = rbar * (3.0f - rbar * rbar) / (1.0f - rbar * rbar);           \
                b.x = fminf(kGuideKappaMax, fmaxf(kGuideKappaMin, kappa));                        \
                if (rl > 0) {                                                                     \
                    const float ir = 1.0f / rl;                                                   \
                    a.y = R0 * ir;                                                                \
                    a.z = R1 * ir;                                                                \
                    a.w = R2 * ir;                                                                \
                }                                                                                 \
                b.z = expf(-2.0f * b.x);                                                          \
                b.y = b.x / (2 * kPi * (1.0f - b.z));                                             \
            }                                                                                     \
            float4 *L = lobes + ((size_t)(c) * K + k) * 2;                                        \
            L[0] = a;                                                                             \
            L[1] = b;                                                                             \
            if (commit) lobeStats[(size_t)(c) * K + k] = make_float4(S, R0, R1, R2);              \
        }                                                                                         \
    }

// One warp per cell; the statistics are already summed per cell (multi-GPU / external-collective path).
__global__ void __launch_bounds__(256) k_mstep(float4 *__restrict__ lobes, float4 *__restrict__ lobeStats, float *__restrict__ stats,
                                               uint32_t nCells, int K, int stride, int commit) {
    const uint32_t warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    const int k = (int)lane();
    for (uint32_t c = warpGlobal; c < nCells; c += nWarps) {
        const float *st = stats + (size_t)c * stride;
        PG_MSTEP_BODY(c, st)
    }
}

// Single-GPU path: the per-cell sum of the chunk partials (k_reduce_partials' job) folded in -- one launch and one round
// trip of the statistics less per EM iteration. ONE BLOCK PER CELL: a cell owns between one and several hundred chunks
// (a few cells in front of the light hold most of the samples), and a single warp walking them one dependent load at a
// time made the whole launch wait for the heaviest cell (ncu launch list r01_v5: 120 us at 788 cells). The block's
// 8 warps take the chunks round-robin, 4 loads in flight each, in double; warp 0 adds the 8 partial sums in warp order
// (a fixed order: results do not depend on scheduling) and refits the cell.
static constexpr int kMstepWarps = 8;
__global__ void __launch_bounds__(kMstepWarps * 32) k_mstep_partials(float4 *__restrict__ lobes, float4 *__restrict__ lobeStats,
                                                                     float *__restrict__ stats, uint32_t nCells, int K, int stride,
                                                                     int commit, const float *__restrict__ partials,
                                                                     const uint32_t *__restrict__ workOfs) {
    __shared__ double sAcc[kMstepWarps][4 * kGuideMaxK + 8];
    const int k = (int)lane(), warp = (int)(threadIdx.x >> 5);
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        const uint32_t w0 = workOfs[c], w1 = workOfs[c + 1];
        for (int e = k; e < stride; e += 32) {
            double acc = 0.0;
            uint32_t w = w0 + (uint32_t)warp;
            for (; w + 3 * kMstepWarps < w1; w += 4 * kMstepWarps) {
                const float v0 = __ldcs(partials + (size_t)w * stride + e);
                const float v1 = __ldcs(partials + (size_t)(w + kMstepWarps) * stride + e);
                const float v2 = __ldcs(partials + (size_t)(w + 2 * kMstepWarps) * stride + e);
                const float v3 = __ldcs(partials + (size_t)(w + 3 * kMstepWarps) * stride + e);
                acc += (double)v0;
                acc += (double)v1;
                acc += (double)v2;
                acc += (double)v3;
            }
            for (; w < w1; w += kMstepWarps) acc += (double)__ldcs(partials + (size_t)w * stride + e);
            sAcc[warp][e] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            float *own = stats + (size_t)c * stride;
            for (int e = k; e < stride; e += 32) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < kMstepWarps; ++j) acc += sAcc[j][e];
                own[e] = (float)acc;
            }
            __syncwarp();
            const float *st = own;
            PG_MSTEP_BODY(c, st)
        }
        __syncthreads();
    }
}

// ---- fused cross-GPU sum + M-step over NVLink peer memory ---------------------------------------------------
// One kernel per EM iteration does: (1) arrival signal into every peer's flag row, (2) wait for every peer's signal,
// (3) per cell: statistics = sum over ranks r = 0..world-1 (fixed order -> bit-identical on every rank) of the
// peers' E-step results, read straight from their HBM over NVLink, (4) M-step. The statistics buffers are
// double-buffered by iteration parity, so no second barrier is needed: a rank can overwrite buffer b only after
// passing the barrier of the next iteration, which every peer reaches only after it finished reading buffer b.
//
// Large fields (cv.twoPhase): every rank reading every peer's whole buffer moves (world-1) x B bytes per rank. Instead,
// rank r sums only ITS slice of the cells (reads (world-1)/world x B) and pushes the sums into every rank's `sum`
// region (writes (world-1)/world x B), a second barrier follows, and the M-step reads the local `sum` region: the
// traffic of a reduce-scatter + all-gather, still one kernel and still bit-identical everywhere (every cell is summed
// by exactly one rank, in rank order).
// Exchange block of a rank: {buf0, buf1, sum: bufFloats floats each; 64 flags: [0,16) barrier 1, [16,32) barrier 2,
// [48] block counter of the intra-GPU grid barrier, [63] spare slot of the local-only microbenchmark}.
struct CommView {
    float *const *peers;   // world pointers to the ranks' exchange blocks
    int rank, world;
    int flagSlot;          // arrival slot this rank signals in (= rank; the local-only microbenchmark uses a spare slot)
    uint32_t epoch;        // arrival value of this iteration (monotonic)
    size_t bufFloats;      // floats per statistics buffer
    int buf;               // which of the two buffers this iteration uses
    int twoPhase;
    uint32_t *error;       // set when the wait timed out (a peer never arrived); sticky until the next connect
    unsigned long long timeoutNs;  // wall-clock bound of one barrier wait
};
static constexpr int kExchangeThreads = 1024;  // one block per SM (all blocks must be resident while they wait)
__device__ __forceinline__ uint32_t *commFlags(float *block, size_t bufFloats) {
    return reinterpret_cast<uint32_t *>(block + 3 * bufFloats);
}
__device__ __forceinline__ float4 ldPeer4(const float *p) {  // system-scope load that bypasses the (non-coherent) L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stPeer4(float *p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Sum of one float4 of statistics over n source blocks in source order; loads are issued four at a time.
__device__ __forceinline__ float4 sumSources(float *const *src, int n, size_t ofs) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int r = 0; r < n; r += 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (r + j < n) v[j] = ldPeer4(src[r + j] + ofs);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (r + j < n) {
                acc.x += v[j].x;
                acc.y += v[j].y;
                acc.z += v[j].z;
                acc.w += v[j].w;
            }
    }
    return acc;
}
// Cross-GPU barrier: block 0 signals `epoch` in slot (base + flagSlot) of every peer's flag row, every block waits until
// all `world` slots of the local row have reached it. The wait is bounded in WALL-CLOCK time (%globaltimer, nanoseconds,
// cv.timeoutNs from the host -- not in SM cycles, whose rate varies); a rank whose peer never arrives raises cv.error.
// Returns false (for the whole block) when the exchange has failed: the caller must then leave the field alone.
__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool commBarrier(const CommView &cv, int slotBase, bool signal, int *sFail) {
    if (signal && (int)threadIdx.x < cv.world) {
        __threadfence_system();
        uint32_t *peerRow = commFlags(cv.peers[threadIdx.x], cv.bufFloats) + slotBase + cv.flagSlot;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peerRow), "r"(cv.epoch) : "memory");
    }
    if ((int)threadIdx.x < cv.world) {
        const uint32_t *mine = commFlags(cv.peers[cv.rank], cv.bufFloats) + slotBase + threadIdx.x;
        const unsigned long long t0 = globalTimerNs();
        uint32_t v, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - cv.epoch) >= 0) break;
            if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > cv.timeoutNs) {  // a peer never arrived: fail loudly, do not hang
                atomicExch(cv.error, 1u);
                break;
            }
        } while (true);
    }
    __syncthreads();
    if (threadIdx.x == 0) *sFail = *reinterpret_cast<volatile uint32_t *>(cv.error) != 0u;
    __syncthreads();
    return *sFail == 0;
}

__global__ void __launch_bounds__(kExchangeThreads) k_mstep_allreduce(CommView cv, float4 *__restrict__ lobes,
                                                                       float4 *__restrict__ lobeStats, float *__restrict__ stats,
                                                                       uint32_t nCells, int K, int stride, int commit) {
    __shared__ float *sSrc[16];
    __shared__ int sLast, sFail;
    // A failed exchange is sticky: once a wait has timed out (this iteration or an earlier one) no block sums, refits or
    // pushes anything any more, so the field of the surviving ranks stays what it was; the host raises after the loop.
    if (*reinterpret_cast<volatile uint32_t *>(cv.error) != 0u) return;
    if (!commBarrier(cv, 0, blockIdx.x == 0, &sFail)) return;

    const uint32_t warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    const int k = (int)lane();
    const size_t bufOfs = (size_t)cv.buf * cv.bufFloats;
    int nSrc = cv.world;
    if ((int)threadIdx.x < cv.world) sSrc[threadIdx.x] = cv.peers[threadIdx.x] + bufOfs;
    __syncthreads();
    if (cv.twoPhase) {
        // ---- reduce-scatter + all-gather: this rank's slice of the cells, summed once and pushed to every rank
        const uint32_t per = (nCells + (uint32_t)cv.world - 1u) / (uint32_t)cv.world;
        const uint32_t c0 = min(nCells, per * (uint32_t)cv.rank), c1 = min(nCells, c0 + per);
        const size_t sumOfs = 2 * cv.bufFloats;
        for (uint32_t c = c0 + warpGlobal; c < c1; c += nWarps)
            for (int e = k; e < (stride >> 2); e += 32) {
                const size_t o = (size_t)c * stride + 4 * e;
                const float4 acc = sumSources(sSrc, cv.world, o);
                for (int r = 0; r < cv.world; ++r) stPeer4(cv.peers[r] + sumOfs + o, acc);
            }
        // intra-GPU grid barrier (every block's pushes are out), then the second cross-GPU barrier
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            uint32_t *gridDone = commFlags(cv.peers[cv.rank], cv.bufFloats) + 48;
            const uint32_t prev = atomicAdd(gridDone, 1u);
            sLast = prev == gridDim.x - 1;
            if (sLast) atomicExch(gridDone, 0u);
        }
        __syncthreads();
        if (!commBarrier(cv, 16, sLast != 0, &sFail)) return;
        if (threadIdx.x == 0) sSrc[0] = cv.peers[cv.rank] + sumOfs;
        __syncthreads();
        nSrc = 1;
    }
    for (uint32_t c = warpGlobal; c < nCells; c += nWarps) {
        // sum over the sources, float4 e = lane, lane + 32, ... of the cell's `stride` statistics
        // (16-byte accesses: stride = 4K + 8 is a multiple of 4 floats and the buffers are 256-byte aligned)
        float *own = stats + (size_t)c * stride;
        for (int e = k; e < (stride >> 2); e += 32)
            reinterpret_cast<float4 *>(own)[e] = sumSources(sSrc, nSrc, (size_t)c * stride + 4 * e);
        __syncwarp();
        const float *st = own;
        float4 a = make_float4(0, 0, 0, 0), b = a, s = a;
        float S = 0, R0 = 0, R1 = 0, R2 = 0;
        if (k < K) {
            const float4 *L = lobes + ((size_t)c * K + k) * 2;
            a = L[0];
            b = L[1];
            s = lobeStats[(size_t)c * K + k];
            S = kGuideDecay * s.x + st[4 * k];
            R0 = kGuideDecay * s.y + st[4 * k + 1];
            R1 = kGuideDecay * s.z + st[4 * k + 2];
            R2 = kGuideDecay * s.w + st[4 * k + 3];
        }
        float sumS = S;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sumS += __shfl_xor_sync(0xffffffffu, sumS, o);
        if (k < K) {
            if (sumS > 0 && isfinite(sumS)) {
                const float prior = kGuidePriorWeight * sumS / (float)K;
                const float denom = 1.0f / (sumS + (float)K * prior);
                a.x = (S + prior) * denom;
                const float rl = sqrtf(R0 * R0 + R1 * R1 + R2 * R2);
                float rbar = (rl + prior * kGuidePriorMeanCos) / (S + prior);
                rbar = fminf(rbar, 0.9999f);
                const float kappa = rbar * (3.0f - rbar * rbar) / (1.0f - rbar * rbar);
                b.x = fminf(kGuideKappaMax, fmaxf(kGuideKappaMin, kappa));
                if (rl > 0) {
                    const float ir = 1.0f / rl;
                    a.y = R0 * ir;
                    a.z = R1 * ir;
                    a.w = R2 * ir;
                }
                b.z = expf(-2.0f * b.x);
                b.y = b.x / (2 * kPi * (1.0f - b.z));
            }
            float4 *L = lobes + ((size_t)c * K + k) * 2;
            L[0] = a;
            L[1] = b;
            if (commit) lobeStats[(size_t)c * K + k] = make_float4(S, R0, R1, R2);
        }
    }
}

__global__ void __launch_bounds__(256) k_guide_query(GuideDevice G, const float *__restrict__ pos, const float *__restrict__ dir,
                                                     const float *__restrict__ u, uint32_t n, float *__restrict__ outPdf,
                                                     float *__restrict__ outDir, float *__restrict__ outSpdf,
                                                     uint32_t *__restrict__ outCell) {
    // the warp-cooperative routines the shade stage uses (all lanes of a warp stay in the loop together)
    __shared__ float4 sCoop[8 * kCoopFloat4PerWarp];
    float4 *sq = sCoop + kCoopFloat4PerWarp * (threadIdx.x >> 5);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < n;
        float3 p = f3(0.0f), w = f3(0.0f, 0.0f, 1.0f);
        float u0 = 0, u1 = 0, u2 = 0;
        uint32_t c = 0;
        if (valid) {
            p = ld3(pos + 3 * (size_t)i);
            w = ld3(dir + 3 * (size_t)i);
            u0 = u[3 * (size_t)i]; u1 = u[3 * (size_t)i + 1]; u2 = u[3 * (size_t)i + 2];
            c = guideLookup(G, p);
        }
        const int k = guideSelectCoop(G, sq, valid, c, u0);
        float3 d = f3(0.0f, 0.0f, 1.0f);
        if (valid) d = guideSampleLobe(G, c, k, u1, u2);
        float pw = 0, pd = 0;
        guidePdf2Coop(G, sq, valid, c, w, d, pw, pd);
        if (valid) {
            outCell[i] = c;
            outPdf[i] = pw;
            outDir[3 * (size_t)i] = d.x;
            outDir[3 * (size_t)i + 1] = d.y;
            outDir[3 * (size_t)i + 2] = d.z;
            outSpdf[i] = pd;
        }
    }
}

__global__ void __launch_bounds__(256) k_pack_samples(const float *pos, const float *dir, const float *weight, const float *pdf,
                                                      const float *dist, uint32_t n, float4 *sRec, float *sDist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        sRec[2 * (size_t)i] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], weight ? weight[i] : 0.0f);
        sRec[2 * (size_t)i + 1] = make_float4(dir ? dir[3 * (size_t)i] : 0.0f, dir ? dir[3 * (size_t)i + 1] : 0.0f, dir ? dir[3 * (size_t)i + 2] : 1.0f,
                              pdf ? pdf[i] : 1.0f);
        sDist[i] = dist ? dist[i] : 0.0f;
    }
}

// =============================================================================================
// host side
// =============================================================================================
static int gridFor(size_t n, int block) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    size_t want = (n + block - 1) / block;
    return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sms * 8));
}

static void finalizeLobe(GuideLobeHost &l) {
    l.eMin2K = std::exp(-2.0f * l.kappa);
    l.norm = l.kappa / (2 * 3.14159265358979323846f * (1.0f - l.eMin2K));
}

void GuidingHost::init(const B200pgIntegratorParams &P, const HostScene &H, cudaStream_t st) {
    stream = st;
    active = P.guiding != 0;
    K = std::min(std::max(P.guide_max_components > 0 ? P.guide_max_components : 16, 1), kGuideMaxK);
    alpha = P.guiding_probability > 0 ? P.guiding_probability : 0.5f;
    maxCellSamples = P.guide_max_cell_samples > 0 ? (float)P.guide_max_cell_samples : 32768.0f;
    maxVerts = P.max_depth > 0 ? std::max(1, P.max_depth - 1) : 16;
    maxVerts = std::min(maxVerts, 32);
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {  // slightly enlarged scene box
        float ext = H.sceneMax[a] - H.sceneMin[a];
        mn[a] = H.sceneMin[a] - 0.01f * ext - 1e-3f;
        mx[a] = H.sceneMax[a] + 0.01f * ext + 1e-3f;
    }
    if (active) {  // the field lives on the device at a fixed capacity; the host keeps a lazily refreshed mirror
        dNodes.allocExact(2 * kCommMaxCells);
        dLobes.allocExact(kCommMaxCells * (size_t)K * 2);
        dLobeStats.allocExact(kCommMaxCells * (size_t)K);
        dCells.allocExact(kCommMaxCells);
        dCellLeaf.allocExact(kCommMaxCells);
        dCounts.allocExact(4);
    }
    resetField(mn, mx);
    dSCount.alloc(1);
    CUDA_OK(cudaMemsetAsync(dSCount.p, 0, sizeof(uint32_t), stream));
}

void GuidingHost::resetField(const float *, const float *) {
    nodes.assign(1, GuideNodeHost{3u, 0.0f, 0u, 0u});
    cells.assign(1, GuideCellHost());
    std::memset(&cells[0], 0, sizeof(GuideCellHost));
    lobes.resize(K);
    for (int k = 0; k < K; ++k) {  // spherical Fibonacci directions, equal weights
        GuideLobeHost &l = lobes[k];
        std::memset(&l, 0, sizeof(l));
        float z = 1.0f - (2.0f * k + 1.0f) / (float)K;
        float r = std::sqrt(std::max(0.0f, 1.0f - z * z));
        float phi = 2.0f * 3.14159265358979323846f * (float)k * 0.6180339887f;
        l.weight = 1.0f / K;
        l.mux = r * std::cos(phi);
        l.muy = r * std::sin(phi);
        l.muz = z;
        l.kappa = kGuideInitKappa;
        finalizeLobe(l);
    }
    trained = false;
    uploadField();
}

// host mirror -> device (initial field, loaded snapshots)
void GuidingHost::uploadField() {
    keysValid = false;  // a different tree: cells stored with recorded samples no longer apply
    nCells = (uint32_t)cells.size();
    nNodes = (uint32_t)nodes.size();
    mirrorValid = true;
    if (!active && dNodes.n == 0) {  // query-only use (tests): size the buffers to the field
        dNodes.allocExact(std::max<size_t>(nodes.size(), 1));
        dLobes.allocExact(std::max<size_t>(lobes.size() * 2, 1));
        dLobeStats.allocExact(std::max<size_t>(lobes.size(), 1));
        dCells.allocExact(std::max<size_t>(cells.size(), 1));
        dCellLeaf.allocExact(std::max<size_t>(cells.size(), 1));
        dCounts.allocExact(4);
    }
    if (nodes.size() > dNodes.n || lobes.size() * 2 > dLobes.n || cells.size() > dCells.n) {
        // a snapshot larger than the current capacity: grow everything (keeps query-only integrators small)
        dNodes.release(); dLobes.release(); dLobeStats.release(); dCells.release(); dCellLeaf.release();
        dNodes.allocExact(std::max<size_t>(nodes.size(), 2 * kCommMaxCells));
        dLobes.allocExact(std::max<size_t>(lobes.size() * 2, kCommMaxCells * (size_t)K * 2));
        dLobeStats.allocExact(std::max<size_t>(lobes.size(), kCommMaxCells * (size_t)K));
        dCells.allocExact(std::max<size_t>(cells.size(), kCommMaxCells));
        dCellLeaf.allocExact(std::max<size_t>(cells.size(), kCommMaxCells));
    }
    stageQuery.resize(lobes.size() * 8);
    stageStats.resize(lobes.size() * 4);
    for (size_t i = 0; i < lobes.size(); ++i) {
        std::memcpy(&stageQuery[8 * i], &lobes[i].weight, 32);
        std::memcpy(&stageStats[4 * i], &lobes[i].statS, 16);
    }
    std::vector<float> hdr(cells.size() * 2);
    std::vector<uint32_t> leaf(cells.size(), 0);
    for (size_t c = 0; c < cells.size(); ++c) {
        hdr[2 * c] = cells[c].sampleCount;
        hdr[2 * c + 1] = cells[c].weightSum;
    }
    for (uint32_t n = 0; n < nodes.size(); ++n)
        if (nodes[n].axis == 3u) leaf[nodes[n].left] = n;
    const uint32_t counts[4] = {nCells, nNodes, 0u, 0u};
    CUDA_OK(cudaMemcpyAsync(dNodes.p, nodes.data(), nodes.size() * 16, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dLobes.p, stageQuery.data(), lobes.size() * 32, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dLobeStats.p, stageStats.data(), lobes.size() * 16, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dCells.p, hdr.data(), hdr.size() * 4, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dCellLeaf.p, leaf.data(), leaf.size() * 4, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dCounts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
}

// device -> host mirror (snapshots); the device copy is the source of truth after a training update
void GuidingHost::syncMirror() {
    if (mirrorValid) return;
    nodes.resize(nNodes);
    cells.resize(nCells);
    lobes.resize((size_t)nCells * K);
    stageQuery.resize(lobes.size() * 8);
    stageStats.resize(lobes.size() * 4);
    std::vector<float> hdr((size_t)nCells * 2);
    CUDA_OK(cudaMemcpyAsync(nodes.data(), dNodes.p, nodes.size() * 16, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(stageQuery.data(), dLobes.p, lobes.size() * 32, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(stageStats.data(), dLobeStats.p, lobes.size() * 16, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(hdr.data(), dCells.p, hdr.size() * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < lobes.size(); ++i) {
        std::memcpy(&lobes[i].weight, &stageQuery[8 * i], 32);
        std::memcpy(&lobes[i].statS, &stageStats[4 * i], 16);
    }
    for (size_t c = 0; c < cells.size(); ++c) {
        std::memset(&cells[c], 0, sizeof(GuideCellHost));
        cells[c].sampleCount = hdr[2 * c];
        cells[c].weightSum = hdr[2 * c + 1];
    }
    mirrorValid = true;
}

void GuidingHost::ensureBatch(size_t nPaths) {
    if (!active) return;
    const size_t nv = nPaths * (size_t)maxVerts;
    if (nv > vertCapacity) {
        dVRec.alloc(4 * nv);
        vertCapacity = nv;
    }
}

void GuidingHost::configure(ShadeArgs &A) {
    GuideDevice &G = A.G;
    std::memset(&G, 0, sizeof(G));
    G.nodes = dNodes.p;
    G.lobes = dLobes.p;
    G.lobeStats = dLobeStats.p;
    G.K = K;
    G.alpha = alpha;
    G.enabled = active && sampling && trained;
    G.record = active && recording;
    G.vRec = dVRec.p;
    G.maxVerts = maxVerts;
    G.sRec = dSRec.p; G.sDist = dSDist.p; G.sKey = dSKey.p;
    if (G.record && !G.enabled) keysValid = false;  // vertices recorded without a cell look-up
    G.sCount = dSCount.p;
    G.sCapacity = (uint32_t)sampleCapacity;
}

void GuidingHost::sortByCell(uint32_t n, bool storedKeys) {
    const uint32_t nBlocks = (n + kSortTile - 1) / kSortTile;
    dKeysA.alloc(n); dKeysB.alloc(n); dValsA.alloc(n); dValsB.alloc(n);
    dBlockHist.alloc((size_t)256 * std::max(nBlocks, 1u));
    GuideDevice G;
    std::memset(&G, 0, sizeof(G));
    G.nodes = dNodes.p;
    G.lobes = dLobes.p;
    G.lobeStats = dLobeStats.p;
    G.K = K;
    if (n) {
        if (storedKeys)
            k_keys_from_samples<<<gridFor(n, 256), 256, 0, stream>>>(dSKey.p, n, dKeysA.p, dValsA.p);
        else
            k_guide_cells<<<gridFor(n, 256), 256, 0, stream>>>(G, dSRec.p, n, dKeysA.p, dValsA.p);
        launches++;
    }
    uint32_t *kin = dKeysA.p, *vin = dValsA.p, *kout = dKeysB.p, *vout = dValsB.p;
    const int passes = numCells() > 65536 ? 3 : (numCells() > 256 ? 2 : 1);
    for (int pass = 0; pass < passes && n > 0; ++pass) {
        const int shift = 8 * pass;
        k_radix_hist<<<nBlocks, kSortThreads, 0, stream>>>(kin, n, shift, dBlockHist.p, nBlocks);
        const uint32_t m = 256 * nBlocks, nSeg = (m + kScanSeg - 1) / kScanSeg;
        dScanTotals.alloc(nSeg);
        k_scan_segments<<<nSeg, 256, 0, stream>>>(dBlockHist.p, m, dScanTotals.p);
        k_scan_totals<<<1, 1024, 0, stream>>>(dScanTotals.p, nSeg);
        k_scan_add<<<nSeg, 256, 0, stream>>>(dBlockHist.p, m, dScanTotals.p);
        k_radix_scatter<<<nBlocks, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, dBlockHist.p, nBlocks);
        launches += 5;
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    sortedCells = kin;
    sortedPerm = vin;
    // first sorted index of every cell that occurs (0xFFFFFFFF = empty cell)
    dCellStart.alloc(numCells() + 1);
    CUDA_OK(cudaMemsetAsync(dCellStart.p, 0xFF, numCells() * sizeof(uint32_t), stream));
    if (n) {
        k_cell_starts<<<gridFor(n, 256), 256, 0, stream>>>(sortedCells, n, dCellStart.p);
        launches++;
    }
}

// Work items, gather + partition, position moments: all enqueued without a host round trip; kernels read the number
// of work items from dCounts[2] and run on persistent grids.
void GuidingHost::buildWork() {
    if (numCells() > kCommMaxCells) throw std::runtime_error("guiding field exceeds its capacity of 65536 cells");
    const uint32_t nc = numCells();
    workBound = nSamples / kChunk + nc + 1;
    dOffsets.alloc(nc + 1);
    dWorkOfs.alloc(nc + 1);
    dWork.alloc(workBound);
    k_build_work<<<1, 1024, 0, stream>>>(dCellStart.p, nc, nSamples, (uint32_t)kChunk, dOffsets.p, dWorkOfs.p, dCounts.p);
    k_fill_work<<<gridFor(workBound, 256), 256, 0, stream>>>(dOffsets.p, dWorkOfs.p, nc, (uint32_t)kChunk, dCounts.p, dWork.p);
    launches += 2;
    dSortPos.alloc(nSamples);
    dSortDir.alloc(nSamples);
    dPartials.alloc((size_t)workBound * statsStride());
    dStats.alloc((size_t)nc * statsStride());
    int smCount = 148;
    cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, 0);
    const uint32_t grid = std::max(1u, std::min<uint32_t>((workBound + 3) / 4, (uint32_t)smCount * 16));
    k_gather_partition<<<grid, 128, 0, stream>>>(dSRec.p, sortedPerm, dWork.p, dCounts.p, dSortPos.p, dSortDir.p, dPartials.p,
                                                 (int)statsStride(), K);
    launches++;
}

void GuidingHost::begin() {
    if (pendingCount == 0xFFFFFFFFu) {  // not told by the integrator: ask the device
        uint32_t n = 0;
        CUDA_OK(cudaMemcpyAsync(&n, dSCount.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        pendingCount = n;
    }
    nSamples = (uint32_t)std::min<size_t>(pendingCount, sampleCapacity);
    pendingCount = 0xFFFFFFFFu;
    samplesTrained += nSamples;
    sortByCell(nSamples, keysValid);
    buildWork();
}

void GuidingHost::beginExternal(const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist,
                                size_t n) {
    if (n > sampleCapacity) {
        dSRec.alloc(2 * n); dSDist.alloc(n); dSKey.alloc(n);
        sampleCapacity = n;
    }
    DevBuf<float> a, b, c, d, e;
    a.upload(pos, 3 * n, stream);
    if (dir) b.upload(dir, 3 * n, stream);
    if (weight) c.upload(weight, n, stream);
    if (pdf) d.upload(pdf, n, stream);
    if (dist) e.upload(dist, n, stream);
    if (n) {
        k_pack_samples<<<gridFor(n, 256), 256, 0, stream>>>(a.p, dir ? b.p : nullptr, weight ? c.p : nullptr, pdf ? d.p : nullptr,
                                                            dist ? e.p : nullptr, (uint32_t)n, dSRec.p, dSDist.p);
        launches++;
    }
    CUDA_OK(cudaStreamSynchronize(stream));
    nSamples = (uint32_t)n;
    keysValid = false;  // these samples came from the host
    sortByCell(nSamples, false);
    buildWork();
}

void GuidingHost::accumulate() { accumulateInto(dStats.p); }

void GuidingHost::accumulateInto(float *statsOut) {
    estepOnly();
    const int stride = (int)statsStride();
    k_reduce_partials<<<gridFor((size_t)numCells() * stride, 256), 256, 0, stream>>>(dPartials.p, dWorkOfs.p, numCells(), stride, statsOut);
    launches++;
}

void GuidingHost::estepOnly() {
    GuideDevice G;
    std::memset(&G, 0, sizeof(G));
    G.lobes = dLobes.p;
    G.K = K;
    const int stride = (int)statsStride();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint32_t grid = std::max(1u, std::min<uint32_t>((workBound + 3) / 4, (uint32_t)sms * 4));  // one resident wave (128 registers)
    CUDA_OK(cudaMemsetAsync(dCounts.p + 3, 0, sizeof(uint32_t), stream));
    if (K <= 8)
        k_estep<8><<<grid, 128, 0, stream>>>(G, dSortPos.p, dSortDir.p, dWork.p, dCounts.p, dPartials.p, stride);
    else if (K <= 16)
        k_estep<16><<<grid, 128, 0, stream>>>(G, dSortPos.p, dSortDir.p, dWork.p, dCounts.p, dPartials.p, stride);
    else
        k_estep<32><<<grid, 128, 0, stream>>>(G, dSortPos.p, dSortDir.p, dWork.p, dCounts.p, dPartials.p, stride);
    launches++;
}

void GuidingHost::update(bool commit) {
    const int stride = (int)statsStride();
    k_mstep<<<gridFor((size_t)numCells() * 32, 256), 256, 0, stream>>>(dLobes.p, dLobeStats.p, dStats.p, numCells(), K, stride, commit ? 1 : 0);
    launches++;
}

// Fold the update's cell statistics into the running headers and split over-full cells -- on the device; the host
// only learns the new cell / node counts (8 bytes).
void GuidingHost::end() {
    const uint32_t cap = (uint32_t)std::min<size_t>(dCells.n, kCommMaxCells);
    uint32_t before = nCells;
    k_split<<<1, 1024, 0, stream>>>(dNodes.p, dLobes.p, dLobeStats.p, dCells.p, dCellLeaf.p, dStats.p, dCounts.p,
                                   K, (int)statsStride(), maxCellSamples, cap, 1);
    launches++;
    uint32_t counts[2] = {0, 0};
    CUDA_OK(cudaMemcpyAsync(counts, dCounts.p, sizeof(counts), cudaMemcpyDeviceToHost, stream));
    // Further split levels of the same update (oracle_guiding.h: guideTrain, splitLevels): bin the samples into the tree as it is
    // now, and split every cell whose halved running count still exceeds the threshold at the mean of ITS samples. One device only:
    // with connected peers every level would need its own cross-GPU sum of the cell moments.
    for (int level = 1; level < splitLevels && commWorld <= 1 && nSamples > 0; ++level) {
        CUDA_OK(cudaStreamSynchronize(stream));
        if (counts[0] == before) break;  // nothing split at the previous level
        before = nCells = counts[0];
        nNodes = counts[1];
        sortByCell(nSamples, false);
        buildWork();  // chunk list + gather: leaves the chunks' position moments in the partials buffer
        CUDA_OK(cudaMemsetAsync(dStats.p, 0, (size_t)nCells * statsStride() * sizeof(float), stream));
        k_cell_moments<<<gridFor((size_t)nCells * 32, 256), 256, 0, stream>>>(dPartials.p, dWorkOfs.p, nCells, K, (int)statsStride(), dStats.p);
        k_split<<<1, 1024, 0, stream>>>(dNodes.p, dLobes.p, dLobeStats.p, dCells.p, dCellLeaf.p, dStats.p, dCounts.p,
                                       K, (int)statsStride(), maxCellSamples, cap, 0);
        launches += 2;
        CUDA_OK(cudaMemcpyAsync(counts, dCounts.p, sizeof(counts), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_OK(cudaMemsetAsync(dSCount.p, 0, sizeof(uint32_t), stream));
    keysValid = true;  // the sample buffer is empty: what is recorded from now on is looked up in the tree as it is now
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaGetLastError());
    nCells = counts[0];
    nNodes = counts[1];
    mirrorValid = false;
    trained = true;
    sortedPerm = sortedCells = nullptr;
}

void GuidingHost::trainLocal() { train(emIterations); }

// One complete training update without host round trips between the EM iterations. With connected peers the
// E-step result of iteration `it` goes into exchange buffer it % 2 and k_mstep_allreduce sums it over the ranks.
void GuidingHost::train(int nIter) {
    begin();
    const int stride = (int)statsStride();
    if (commWorld > 1 && (size_t)numCells() * stride > commFloats)
        throw std::runtime_error("guiding field has more cells than the multi-GPU exchange buffer holds");
    for (int it = 0; it < nIter; ++it) {
        const bool commit = it == nIter - 1;
        if (commWorld > 1) {
            const int buf = (int)(commEpoch & 1u);
            accumulateInto(commBlock + (size_t)buf * commFloats);
            CommView cv;
            cv.peers = dCommPeers.p;
            cv.rank = commRank;
            cv.world = commWorld;
            cv.flagSlot = commRank;
            cv.epoch = ++commEpoch;
            cv.bufFloats = commFloats;
            cv.buf = buf;
            cv.twoPhase = exchangeTwoPhase(numCells()) ? 1 : 0;
            cv.error = dCommError.p;
            cv.timeoutNs = commTimeoutNs;
            k_mstep_allreduce<<<exchangeGrid(numCells()), kExchangeThreads, 0, stream>>>(cv, dLobes.p, dLobeStats.p, dStats.p, numCells(), K,
                                                                                       stride, commit ? 1 : 0);
            launches++;
        } else {  // single GPU: the per-cell sum of the partials is folded into the M-step kernel
            estepOnly();
            k_mstep_partials<<<gridFor((size_t)numCells() * kMstepWarps * 32, kMstepWarps * 32), kMstepWarps * 32, 0, stream>>>(
                dLobes.p, dLobeStats.p, dStats.p, numCells(), K, stride, commit ? 1 : 0, dPartials.p, dWorkOfs.p);
            launches++;
        }
    }
    if (commWorld > 1) {
        uint32_t err = 0;
        CUDA_OK(cudaMemcpyAsync(&err, dCommError.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        if (err) throw std::runtime_error("multi-GPU statistics exchange timed out: a peer rank never arrived");
    }
    end();
}

// every block of the exchange grid must be resident while it waits for the peers: at most one block per SM
int GuidingHost::exchangeGrid(uint32_t cells) const {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<size_t>(1, std::min<size_t>(((size_t)cells * 32 + kExchangeThreads - 1) / kExchangeThreads, (size_t)sms));
}
// reduce-scatter + all-gather form once the all-read form would pull more than 8 MB from the peers (measured cross-over on
// 4 and 8 B200s, profiles/r01_v5_em_exchange.jsonl; with two ranks both forms move the same bytes and the all-read form
// saves a barrier); depends only on numbers every rank shares
bool GuidingHost::exchangeTwoPhase(uint32_t cells) const {
    if (commForceMode >= 0) return commForceMode != 0 && commWorld > 1;
    return commWorld > 2 && (size_t)cells * statsStride() * sizeof(float) * (size_t)(commWorld - 1) > ((size_t)8 << 20);
}

void GuidingHost::commLocalHandle(void *out64) {
    if (!commBlock) {
        commFloats = kCommMaxCells * statsStride();
        const size_t bytes = 3 * commFloats * sizeof(float) + 64 * sizeof(uint32_t);  // {buf0, buf1, sum, flags}
        CUDA_OK(cudaMalloc(&commBlock, bytes));
        CUDA_OK(cudaMemset(commBlock, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, commBlock));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(out64, &h, 64);
}

void GuidingHost::commConnect(int rank, int world, const void *handles) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world) throw std::runtime_error("invalid rank / world size");
    if (!commBlock) throw std::runtime_error("b200pg_comm_local_handle must be called first");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            commPeers[r] = commBlock;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        commPeers[r] = (float *)ptr;
    }
    commRank = rank;
    commWorld = world;
    commIpc = true;
    finishConnect();
}

// Shared tail of the two connect paths: a (re)connect starts from a clean slate -- error flag cleared, the intra-GPU grid
// counter of the two-phase form zeroed (a failed exchange can leave it non-zero). The arrival flags are NOT cleared here:
// a peer that connected first may already have signalled; epochs restart at 0 on every rank and the flags only ever
// compare as "reached", so stale larger values from a previous session must be wiped by the owner BEFORE anyone connects
// (commLocalHandle / commLocalBlock zero the block when it is created; reconnecting ranks call commResetFlags first).
void GuidingHost::finishConnect() {
    commEpoch = 0;
    if (const char *e = std::getenv("B200PG_EXCHANGE_FORM")) commForceMode = std::atoi(e);  // tests: 0 all-read, 1 two-phase
    if (const char *e = std::getenv("B200PG_COMM_TIMEOUT_S")) commTimeoutNs = (unsigned long long)(std::atof(e) * 1e9);
    dCommPeers.upload(commPeers, (size_t)commWorld, stream);
    dCommError.alloc(1);
    CUDA_OK(cudaMemsetAsync(dCommError.p, 0, sizeof(uint32_t), stream));
    CUDA_OK(cudaMemsetAsync(reinterpret_cast<uint32_t *>(commBlock + 3 * commFloats) + 48, 0, sizeof(uint32_t), stream));
    CUDA_OK(cudaStreamSynchronize(stream));
}

// In-process variant (b200pg_render with a device list: one worker thread per GPU inside one process, where CUDA IPC handles
// cannot be opened): the ranks exchange plain device pointers; peer access between the devices must be enabled.
float *GuidingHost::commLocalBlock() {
    if (!commBlock) {
        commFloats = kCommMaxCells * statsStride();
        const size_t bytes = 3 * commFloats * sizeof(float) + 64 * sizeof(uint32_t);  // {buf0, buf1, sum, flags}
        CUDA_OK(cudaMalloc(&commBlock, bytes));
        CUDA_OK(cudaMemset(commBlock, 0, bytes));
    }
    return commBlock;
}
void GuidingHost::commConnectPointers(int rank, int world, float *const *blocks) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world) throw std::runtime_error("invalid rank / world size");
    if (!commBlock || blocks[rank] != commBlock) throw std::runtime_error("commLocalBlock must be called first");
    for (int r = 0; r < world; ++r) commPeers[r] = blocks[r];
    commRank = rank;
    commWorld = world;
    commIpc = false;
    finishConnect();
}

// ---- exchange microbenchmark ("EM-allreduce scaling, cells in {1 k, 8 k, 64 k} x K = 32") --------
// Times k_mstep_allreduce alone: synthetic statistics in both exchange buffers, scratch lobes (the field is not touched).
// Every connected rank must call it with the same arguments (the kernel contains the cross-GPU barrier).
// localOnly: the same kernel restricted to this rank's own buffer = the M-step share of the time.
__global__ void k_fill_stats(float *p, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed * 40503u;
        h ^= h >> 15;
        p[i] = 0.25f + (float)(h & 1023u) * (1.0f / 1024.0f);
    }
}
float GuidingHost::exchangeBench(uint32_t cells, int nIter, bool localOnly) {
    if (!commBlock || commWorld < 1 || !dCommPeers.p) throw std::runtime_error("b200pg_comm_connect must be called first");
    if (cells == 0 || cells > kCommMaxCells) throw std::runtime_error("cell count outside the exchange buffer");
    const int stride = (int)statsStride();
    DevBuf<float4> lobesT, statsT;
    DevBuf<float> sumT;
    lobesT.allocExact((size_t)cells * K * 2);
    statsT.allocExact((size_t)cells * K);
    sumT.allocExact((size_t)cells * stride);
    CUDA_OK(cudaMemsetAsync(lobesT.p, 0, (size_t)cells * K * 2 * sizeof(float4), stream));
    CUDA_OK(cudaMemsetAsync(statsT.p, 0, (size_t)cells * K * sizeof(float4), stream));
    k_fill_stats<<<592, 256, 0, stream>>>(commBlock, 2 * commFloats, (uint32_t)commRank + 1u);
    CUDA_OK(cudaStreamSynchronize(stream));
    const int grid = exchangeGrid(cells);
    float *selfPeer = commBlock;
    DevBuf<float *> dSelf;
    dSelf.upload(&selfPeer, 1, stream);
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const int warm = 3;
    for (int it = 0; it < warm + nIter; ++it) {
        if (it == warm) CUDA_OK(cudaEventRecord(e0, stream));
        CommView cv;
        cv.peers = localOnly ? dSelf.p : dCommPeers.p;
        cv.rank = localOnly ? 0 : commRank;
        cv.world = localOnly ? 1 : commWorld;
        cv.flagSlot = localOnly ? 63 : commRank;
        cv.buf = (int)(commEpoch & 1u);
        cv.epoch = localOnly ? 0u : ++commEpoch;  // epoch 0 is always "arrived"
        cv.bufFloats = commFloats;
        cv.twoPhase = (!localOnly && exchangeTwoPhase(cells)) ? 1 : 0;
        cv.error = dCommError.p;
        cv.timeoutNs = commTimeoutNs;
        k_mstep_allreduce<<<grid, kExchangeThreads, 0, stream>>>(cv, lobesT.p, statsT.p, sumT.p, cells, K, stride, 0);
        launches++;
    }
    CUDA_OK(cudaEventRecord(e1, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaGetLastError());
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    uint32_t err = 0;
    CUDA_OK(cudaMemcpy(&err, dCommError.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (err) throw std::runtime_error("multi-GPU statistics exchange timed out: a peer rank never arrived");
    return ms / (float)nIter;
}

// back to a single rank; the exchange block stays allocated for the next connect
void GuidingHost::commDisconnect() {
    for (int r = 0; r < commWorld; ++r)
        if (commIpc && r != commRank && commPeers[r]) cudaIpcCloseMemHandle(commPeers[r]);
    for (auto &q : commPeers) q = nullptr;
    commWorld = 1;
    commRank = 0;
    commEpoch = 0;
    if (commBlock) cudaMemset(reinterpret_cast<uint32_t *>(commBlock + 3 * commFloats), 0, 64 * sizeof(uint32_t));  // arrival flags
}

void GuidingHost::commClose() {
    for (int r = 0; r < commWorld; ++r)
        if (commIpc && r != commRank && commPeers[r]) cudaIpcCloseMemHandle(commPeers[r]);
    if (commBlock) cudaFree(commBlock);
    commBlock = nullptr;
    commWorld = 1;
    for (auto &q : commPeers) q = nullptr;
}

void GuidingHost::query(const float *pos, const float *dir, const float *u, size_t n, float *outPdf, float *outDir, float *outSpdf,
                        uint32_t *outCell) {
    DevBuf<float> dP, dD, dU, oP, oD, oS;
    DevBuf<uint32_t> oC;
    dP.upload(pos, 3 * n, stream);
    dD.upload(dir, 3 * n, stream);
    dU.upload(u, 3 * n, stream);
    oP.alloc(n); oD.alloc(3 * n); oS.alloc(n); oC.alloc(n);
    GuideDevice G;
    std::memset(&G, 0, sizeof(G));
    G.nodes = dNodes.p;
    G.lobes = dLobes.p;
    G.lobeStats = dLobeStats.p;
    G.K = K;
    if (n) k_guide_query<<<gridFor(n, 256), 256, 0, stream>>>(G, dP.p, dD.p, dU.p, (uint32_t)n, oP.p, oD.p, oS.p, oC.p);
    launches++;
    CUDA_OK(cudaMemcpyAsync(outPdf, oP.p, n * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(outDir, oD.p, 3 * n * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(outSpdf, oS.p, n * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(outCell, oC.p, n * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaGetLastError());
}

void GuidingHost::bin(const float *pos, size_t n, uint32_t *outCell, uint32_t *outPerm, uint32_t *outOffsets, uint32_t *nCells) {
    beginExternal(pos, nullptr, nullptr, nullptr, nullptr, n);
    if (n) {
        CUDA_OK(cudaMemcpyAsync(outPerm, sortedPerm, n * 4, cudaMemcpyDeviceToHost, stream));
        // cells in ORIGINAL order: keys buffer A still holds them only if an even number of passes ran; recompute instead
        GuideDevice G;
        std::memset(&G, 0, sizeof(G));
        G.nodes = dNodes.p;
        DevBuf<uint32_t> k, v;
        k.alloc(n); v.alloc(n);
        k_guide_cells<<<gridFor(n, 256), 256, 0, stream>>>(G, dSRec.p, (uint32_t)n, k.p, v.p);
        launches++;
        CUDA_OK(cudaMemcpyAsync(outCell, k.p, n * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
    }
    CUDA_OK(cudaMemcpyAsync(outOffsets, dOffsets.p, ((size_t)numCells() + 1) * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (nCells) *nCells = numCells();
    CUDA_OK(cudaGetLastError());
}

std::vector<uint32_t> GuidingHost::snapshot() {
    syncMirror();
    std::vector<uint32_t> w(8 + 4 * nodes.size() + 8 * cells.size() + 12 * lobes.size());
    w[0] = 0x47554944u;
    w[1] = (uint32_t)nodes.size();
    w[2] = (uint32_t)cells.size();
    w[3] = (uint32_t)K;
    w[4] = w[5] = w[6] = w[7] = 0;
    size_t o = 8;
    std::memcpy(&w[o], nodes.data(), nodes.size() * 16);
    o += 4 * nodes.size();
    std::memcpy(&w[o], cells.data(), cells.size() * 32);
    o += 8 * cells.size();
    std::memcpy(&w[o], lobes.data(), lobes.size() * 48);
    return w;
}

bool GuidingHost::load(const uint32_t *w, size_t n) {
    if (n < 8 || w[0] != 0x47554944u) return false;
    const size_t nn = w[1], nc = w[2];
    const int k = (int)w[3];
    if (k <= 0 || k > kGuideMaxK || n != 8 + 4 * nn + 8 * nc + 12 * nc * (size_t)k) return false;
    K = k;
    nodes.resize(nn);
    cells.resize(nc);
    lobes.resize(nc * (size_t)K);
    size_t o = 8;
    std::memcpy(nodes.data(), &w[o], nn * 16);
    o += 4 * nn;
    std::memcpy(cells.data(), &w[o], nc * 32);
    o += 8 * nc;
    std::memcpy(lobes.data(), &w[o], lobes.size() * 48);
    trained = true;
    uploadField();
    return true;
}

}  // namespace pg

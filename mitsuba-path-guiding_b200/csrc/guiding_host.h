// guiding_host.h -- host side of the guiding field: device buffers, training schedule, snapshots.
//
// A training update (north-star subsystem 3) runs between two progressions, where a guided version of
// ProgressiveMonteCarloIntegrator would use its postprogression() hook (progressiveintegrator.h:40-52):
//   begin()      count the recorded samples, look up their cells, stable radix sort by cell (binning)
//   accumulate() E-step: per-cell sufficient statistics of the local samples  -> stats buffer
//   [ external sum of the stats buffer over ranks: NCCL allreduce, see bench.py / INTEGRATION.md ]
//   update()     M-step from decayed running statistics + MAP priors (identical on every rank)
//   end()        spatial split of over-full cells (host, deterministic), re-upload of the field
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/b200pg.h"
#include "devbuf.h"
#include "host_scene.h"
#include "wavefront.cuh"

namespace pg {

struct GuideNodeHost {
    uint32_t axis;
    float split;
    uint32_t left;
    uint32_t pad;
};
struct GuideLobeHost {
    float weight, mux, muy, muz;
    float kappa, norm, eMin2K, pad0;
    float statS, statRx, statRy, statRz;
};
struct GuideCellHost {
    float sampleCount, weightSum;
    float pad[6];
};

struct GuidingHost {
    // configuration
    int K = 16;
    int maxVerts = 0;
    float alpha = 0.5f;
    float maxCellSamples = 32768;
    int emIterations = 4;
    bool active = false;    // guiding requested by the integrator parameters
    bool trained = false;   // the field has seen at least one training update
    bool recording = false; // the next progression records training vertices
    bool sampling = false;  // the next progression samples from the field
    cudaStream_t stream = nullptr;

    // field (host mirror + device copy)
    std::vector<GuideNodeHost> nodes;
    std::vector<GuideLobeHost> lobes;
    std::vector<GuideCellHost> cells;
    DevBuf<uint4> dNodes;
    DevBuf<float2> dCells;       // running per-cell headers {sample count, weight sum}
    DevBuf<uint32_t> dCellLeaf;  // leaf node of every cell
    DevBuf<uint32_t> dCounts;    // {nCells, nNodes, nWork, -}
    uint32_t nCells = 1, nNodes = 1; // host copy of the counts (refreshed at the end of every update)
    bool mirrorValid = true;     // nodes / lobes / cells below mirror the device copy
    uint32_t pendingCount = 0xFFFFFFFFu;  // recorded-sample count handed over by the integrator (saves a sync)
    uint64_t samplesTrained = 0; // cumulative number of samples that went through a training update
    uint32_t workBound = 1;      // upper bound of the number of work items (grid / buffer sizing)
    DevBuf<float4> dLobes, dLobeStats;
    std::vector<float> stageQuery, stageStats;  // host staging of the two device lobe arrays

    // training-vertex records and samples
    DevBuf<float4> dVRec;  // 4 x float4 per training vertex
    DevBuf<float4> dSRec, dSortPos, dSortDir;  // samples as recorded (32-byte records) / sorted by cell (two arrays for the E-step's bulk copies)
    DevBuf<float> dSDist;
    DevBuf<uint32_t> dSKey;      // per recorded sample: the guiding cell the shade stage found for its vertex
    // the stored keys are the binning keys as long as every sample in the buffer was recorded while sampling from the CURRENT tree
    // (false after a recording progression without sampling, a field load / reset; true again when the buffer is emptied)
    bool keysValid = false;
    int splitLevels = 1;         // spatial split levels per training update (set_option "split_levels"; > 1 on one device only)
    DevBuf<uint32_t> dSCount;
    size_t vertCapacity = 0, sampleCapacity = 0;

    // binning (radix sort) and EM scratch
    DevBuf<uint32_t> dKeysA, dKeysB, dValsA, dValsB, dBlockHist, dCellStart, dOffsets, dWorkOfs, dScanTotals;
    DevBuf<float> dStats, dPartials;
    DevBuf<uint4> dWork;  // (cell, begin, end, 0) chunks of the sorted sample range
    uint32_t nSamples = 0, nWork = 0;
    uint32_t *sortedPerm = nullptr;  // device pointer into dVals*, valid between begin() and end()
    uint32_t *sortedCells = nullptr;
    uint64_t launches = 0;

    void init(const B200pgIntegratorParams &P, const HostScene &H, cudaStream_t st);
    void resetField(const float *bmin, const float *bmax);
    void uploadField();
    void ensureBatch(size_t nPaths);
    void configure(ShadeArgs &A);
    void preprogression(int pass) { (void)pass; }
    uint32_t numCells() const { return nCells; }
    void syncMirror();
    size_t statsStride() const { return (size_t)K * 4 + 8; }

    // training update
    void begin();
    void beginExternal(const float *pos, const float *dir, const float *weight, const float *pdf, const float *dist, size_t n);
    void accumulate();
    void accumulateInto(float *statsOut);
    void estepOnly();
    void update(bool commit);
    void end();
    void trainLocal();  // begin + emIterations x (accumulate, update) + end
    void train(int nIter);  // same with an explicit iteration count; sums the statistics over the connected ranks

    // ---- multi-GPU: peer-memory exchange of the per-cell EM statistics (no NCCL call on the data path) -----------
    // Every rank owns one cudaMalloc'ed exchange block {2 x kCommCapacity floats of statistics (double-buffered by EM
    // iteration parity), kCommCapacity floats of pushed sums (large fields), 64 flags} that its peers map through CUDA IPC. k_mstep_allreduce signals arrival
    // in every peer's flag row, waits for all peers, then sums the cell's statistics over the ranks in rank order
    // (identical result on every rank) and runs the M-step in the same kernel.
    static constexpr size_t kCommMaxCells = 65536;
    int commRank = 0, commWorld = 1;
    uint32_t commEpoch = 0;
    float *commBlock = nullptr;            // own exchange block (device)
    size_t commFloats = 0;                 // floats per statistics buffer
    float *commPeers[16] = {nullptr};      // mapped blocks of all ranks (own included)
    DevBuf<float *> dCommPeers;
    DevBuf<uint32_t> dCommError;
    void commLocalHandle(void *out64);
    void commConnect(int rank, int world, const void *handles);
    float *commLocalBlock();
    void commConnectPointers(int rank, int world, float *const *blocks);
    void finishConnect();
    bool commIpc = true;                   // peers were mapped through CUDA IPC (closed with cudaIpcCloseMemHandle)
    unsigned long long commTimeoutNs = 10000000000ULL;  // bound of one cross-GPU barrier wait (B200PG_COMM_TIMEOUT_S)
    void commClose();
    void commDisconnect();
    // microbenchmark of the exchange step alone: avg ms per k_mstep_allreduce over `cells` synthetic cells
    float exchangeBench(uint32_t cells, int nIter, bool localOnly);
    int commForceMode = -1;  // -1: choose by size; 0: all-read form; 1: reduce-scatter + all-gather form (microbenchmark)
    int exchangeGrid(uint32_t cells) const;
    bool exchangeTwoPhase(uint32_t cells) const;
    ~GuidingHost() { commClose(); }

    // per-kernel entry points
    void query(const float *pos, const float *dir, const float *u, size_t n, float *outPdf, float *outDir, float *outSpdf,
               uint32_t *outCell);
    void bin(const float *pos, size_t n, uint32_t *outCell, uint32_t *outPerm, uint32_t *outOffsets, uint32_t *nCells);

    // snapshots (32-bit words; layout in oracle/oracle_guiding.h and DESIGN.md)
    std::vector<uint32_t> snapshot();
    bool load(const uint32_t *w, size_t n);

private:
    void sortByCell(uint32_t n, bool storedKeys);
    void buildWork();
};

}  // namespace pg

// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_guiding.h: CPU statement of this repo's guiding field (north-star subsystems 2 and 3).
//
// PARITY UNPINNED: the guided integrator and the guiding library (Intel Open PGL, unvendored and
// unpinned: build_dependencies.sh:10, superbuild/CMakeLists.txt:161) are NOT in the reference
// snapshot, so there is no reference arithmetic, call site, test or golden vector
// for anything in this file. The algorithm below is this repo's own design (documented in
// DESIGN.md "Guiding field"), written once here (double-precision accumulators, scalar loops) and
// once in CUDA; "parity" for guiding means CUDA vs this file.
//
// Specification shared with the CUDA implementation:
//  * spatial kd-tree over the scene box; node = {axis (3 = leaf), split, left | cell}; lookup walks
//    `p[axis] < split ? left : left + 1` from node 0.
//  * per cell a mixture of K von Mises-Fisher lobes: pdf(w) = sum_k pi_k * kappa_k / (2 pi (1 - exp(-2 kappa_k)))
//    * exp(kappa_k (mu_k . w - 1)).
//  * sampling: lobe by cdf walk over pi with u0; cos(theta) = 1 + log(u1 + (1 - u1) exp(-2 kappa)) / kappa,
//    phi = 2 pi u2, frame = coordinateSystem(mu) (util.cpp:594-603).
//  * training sample = {position, direction, weight = avg_rgb(Li estimate) / pdf, pdf, distance}.
//  * binning = stable sort of the samples by cell index.
//  * one EM iteration per cell over its samples in sorted order: gamma_ik = pi_k f_k(w_i) / sum_j pi_j f_j(w_i);
//    S_k = sum w_i gamma_ik, R_k = sum w_i gamma_ik w_i(dir); M-step with decayed running statistics and MAP priors
//    (constants below); spatial split at the sample mean along the axis of largest variance when a cell's
//    running sample count exceeds maxCellSamples.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle_math.h"

namespace orc {

static const int kGuideMaxK = 32;
static const Float kGuidePriorWeight = 0.01f;  // prior mass relative to the cell's total statistic
static const Float kGuidePriorMeanCos = 0.8f;  // mean cosine of the prior lobe (kappa ~ 5)
static const Float kGuideKappaMin = 0.01f, kGuideKappaMax = 5000.0f;
static const Float kGuideDecay = 0.25f;        // weight of the previous statistics in a new training update
static const Float kGuideInitKappa = 5.0f;

struct GuideNode {
    uint32_t axis;  // 0,1,2 inner; 3 leaf
    Float split;
    uint32_t left;  // inner: index of the left child (right = left + 1); leaf: cell index
    uint32_t pad;
};

struct GuideLobe {  // 12 floats, the same record the GPU stores (3 x float4)
    Float weight, mux, muy, muz;
    Float kappa, norm, eMin2K, pad0;
    Float statS, statRx, statRy, statRz;
};

struct GuideCellHeader {  // 8 floats
    Float sampleCount;  // running (decayed) number of samples
    Float weightSum;    // running sum of sample weights
    Float pad[6];
};

inline void lobeFinalize(GuideLobe &l) {
    l.norm = l.kappa / (2 * PI_F * (1.0f - std::exp(-2.0f * l.kappa)));
    l.eMin2K = std::exp(-2.0f * l.kappa);
}

struct GuideField {
    int K = 0;
    std::vector<GuideNode> nodes;
    std::vector<GuideLobe> lobes;  // cells * K
    std::vector<GuideCellHeader> cells;
    Float bmin[3], bmax[3];
    uint32_t numCells() const { return (uint32_t)cells.size(); }

    void init(int K_, const Float *mn, const Float *mx) {
        K = K_;
        for (int i = 0; i < 3; ++i) {
            bmin[i] = mn[i];
            bmax[i] = mx[i];
        }
        nodes.assign(1, GuideNode{3u, 0.0f, 0u, 0u});
        cells.assign(1, GuideCellHeader());
        std::memset(&cells[0], 0, sizeof(GuideCellHeader));
        lobes.resize(K);
        for (int k = 0; k < K; ++k) {  // spherical Fibonacci directions
            GuideLobe &l = lobes[k];
            std::memset(&l, 0, sizeof(l));
            Float z = 1.0f - (2.0f * k + 1.0f) / (Float)K;
            Float r = safe_sqrt(1.0f - z * z);
            Float phi = 2.0f * PI_F * (Float)k * 0.6180339887f;
            l.weight = 1.0f / K;
            l.mux = r * std::cos(phi);
            l.muy = r * std::sin(phi);
            l.muz = z;
            l.kappa = kGuideInitKappa;
            lobeFinalize(l);
        }
    }

    uint32_t lookup(const Vec3 &p) const {
        uint32_t n = 0;
        while (nodes[n].axis != 3u) n = p[(int)nodes[n].axis] < nodes[n].split ? nodes[n].left : nodes[n].left + 1;
        return nodes[n].left;
    }

    Float pdf(uint32_t cell, const Vec3 &w) const {
        const GuideLobe *L = &lobes[(size_t)cell * K];
        Float s = 0;
        for (int k = 0; k < K; ++k) {
            Float c = L[k].mux * w.x + L[k].muy * w.y + L[k].muz * w.z;
            s += L[k].weight * L[k].norm * std::exp(L[k].kappa * (c - 1.0f));
        }
        return s;
    }

    Vec3 sample(uint32_t cell, Float u0, Float u1, Float u2) const {
        const GuideLobe *L = &lobes[(size_t)cell * K];
        int k = 0;
        while (k < K - 1 && u0 >= L[k].weight) {
            u0 -= L[k].weight;
            ++k;
        }
        const GuideLobe &l = L[k];
        Float cosT = 1.0f + std::log(u1 + (1.0f - u1) * l.eMin2K) / l.kappa;
        cosT = std::min(1.0f, std::max(-1.0f, cosT));
        Float sinT = safe_sqrt(1.0f - cosT * cosT);
        Float phi = 2.0f * PI_F * u2;
        Vec3 mu(l.mux, l.muy, l.muz), s, t;
        coordinateSystem(mu, s, t);
        return s * (sinT * std::cos(phi)) + t * (sinT * std::sin(phi)) + mu * cosT;
    }

    // ---- snapshot (32-bit words): header[8], nodes[4 * nNodes], cell headers[8 * nCells], lobes[12 * nCells * K]
    std::vector<uint32_t> snapshot() const {
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
    bool load(const uint32_t *w, size_t n) {
        if (n < 8 || w[0] != 0x47554944u) return false;
        size_t nn = w[1], nc = w[2];
        K = (int)w[3];
        if (K <= 0 || K > kGuideMaxK || n != 8 + 4 * nn + 8 * nc + 12 * nc * (size_t)K) return false;
        nodes.resize(nn);
        cells.resize(nc);
        lobes.resize(nc * (size_t)K);
        size_t o = 8;
        std::memcpy(nodes.data(), &w[o], nn * 16);
        o += 4 * nn;
        std::memcpy(cells.data(), &w[o], nc * 32);
        o += 8 * nc;
        std::memcpy(lobes.data(), &w[o], lobes.size() * 48);
        return true;
    }
};

struct GuideSamples {
    std::vector<Vec3> pos, dir;
    std::vector<Float> weight, pdf, dist;
    size_t size() const { return pos.size(); }
    void push(const Vec3 &p, const Vec3 &d, Float w, Float pd, Float ds) {
        pos.push_back(p);
        dir.push_back(d);
        weight.push_back(w);
        pdf.push_back(pd);
        dist.push_back(ds);
    }
};

// Stable counting sort by cell (what the GPU radix sort must reproduce bit for bit).
inline void guideBin(const GuideField &F, const Vec3 *pos, size_t n, std::vector<uint32_t> &cell, std::vector<uint32_t> &perm,
                     std::vector<uint32_t> &offsets) {
    const uint32_t nc = F.numCells();
    cell.resize(n);
    offsets.assign(nc + 1, 0);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) cell[i] = F.lookup(pos[i]);
    for (size_t i = 0; i < n; ++i) offsets[cell[i] + 1]++;
    for (uint32_t c = 0; c < nc; ++c) offsets[c + 1] += offsets[c];
    perm.resize(n);
    std::vector<uint32_t> cur(offsets.begin(), offsets.end() - 1);
    for (size_t i = 0; i < n; ++i) perm[cur[cell[i]]++] = (uint32_t)i;
}

struct GuideCellStats {  // per cell: K * 4 lobe statistics + 8 cell statistics (same layout as the GPU stats buffer)
    // lobe k: S, Rx, Ry, Rz ; cell: n, W, P1x, P1y, P1z, P2x, P2y, P2z
};

// One E-step: stats buffer layout per cell = [K*4 lobe stats][8 cell stats], doubles here.
inline void guideEStep(const GuideField &F, const GuideSamples &smp, const std::vector<uint32_t> &perm,
                       const std::vector<uint32_t> &offsets, std::vector<double> &stats) {
    const int K = F.K;
    const size_t stride = (size_t)K * 4 + 8;
    stats.assign(stride * F.numCells(), 0.0);
    // cells are independent: parallel over cells (the CPU baseline uses every host core); the sums inside a cell stay
    // sequential in sorted order, so the result does not depend on the thread count
#pragma omp parallel for schedule(dynamic, 1)
    for (long long c = 0; c < (long long)F.numCells(); ++c) {
        double *st = &stats[stride * c];
        const GuideLobe *L = &F.lobes[(size_t)c * K];
        for (uint32_t j = offsets[c]; j < offsets[c + 1]; ++j) {
            const uint32_t i = perm[j];
            const Vec3 &w = smp.dir[i];
            const Float sw = smp.weight[i];
            double *cs = st + (size_t)K * 4;
            cs[0] += 1.0;
            cs[2] += smp.pos[i].x; cs[3] += smp.pos[i].y; cs[4] += smp.pos[i].z;
            cs[5] += (double)smp.pos[i].x * smp.pos[i].x; cs[6] += (double)smp.pos[i].y * smp.pos[i].y;
            cs[7] += (double)smp.pos[i].z * smp.pos[i].z;
            if (!(sw > 0) || !std::isfinite(sw)) continue;
            Float p[kGuideMaxK];
            Float total = 0;
            for (int k = 0; k < K; ++k) {
                Float cosv = L[k].mux * w.x + L[k].muy * w.y + L[k].muz * w.z;
                p[k] = L[k].weight * L[k].norm * std::exp(L[k].kappa * (cosv - 1.0f));
                total += p[k];
            }
            if (!(total > 0) || !std::isfinite(total)) continue;
            cs[1] += sw;
            const Float inv = 1.0f / total;
            for (int k = 0; k < K; ++k) {
                const Float g = sw * (p[k] * inv);
                st[4 * k + 0] += g;
                st[4 * k + 1] += (double)g * w.x;
                st[4 * k + 2] += (double)g * w.y;
                st[4 * k + 3] += (double)g * w.z;
            }
        }
    }
}

// M-step from (decay * running statistics + new statistics). `commit` folds the new statistics into the
// running ones (done once per training update, after the last EM iteration).
inline void guideMStep(GuideField &F, const std::vector<float> &stats, bool commit) {
    const int K = F.K;
    const size_t stride = (size_t)K * 4 + 8;
    for (uint32_t c = 0; c < F.numCells(); ++c) {
        const float *st = &stats[stride * c];
        GuideLobe *L = &F.lobes[(size_t)c * K];
        Float S[kGuideMaxK], R[kGuideMaxK][3], sumS = 0;
        for (int k = 0; k < K; ++k) {
            S[k] = kGuideDecay * L[k].statS + st[4 * k];
            R[k][0] = kGuideDecay * L[k].statRx + st[4 * k + 1];
            R[k][1] = kGuideDecay * L[k].statRy + st[4 * k + 2];
            R[k][2] = kGuideDecay * L[k].statRz + st[4 * k + 3];
            sumS += S[k];
        }
        if (sumS > 0 && std::isfinite(sumS)) {
            const Float a = kGuidePriorWeight * sumS / (Float)K;
            const Float denom = 1.0f / (sumS + (Float)K * a);
            for (int k = 0; k < K; ++k) {
                L[k].weight = (S[k] + a) * denom;
                const Float rl = std::sqrt(R[k][0] * R[k][0] + R[k][1] * R[k][1] + R[k][2] * R[k][2]);
                Float rbar = (rl + a * kGuidePriorMeanCos) / (S[k] + a);
                rbar = std::min(rbar, 0.9999f);
                Float kappa = rbar * (3.0f - rbar * rbar) / (1.0f - rbar * rbar);
                L[k].kappa = std::min(kGuideKappaMax, std::max(kGuideKappaMin, kappa));
                if (rl > 0) {
                    const Float ir = 1.0f / rl;
                    L[k].mux = R[k][0] * ir;
                    L[k].muy = R[k][1] * ir;
                    L[k].muz = R[k][2] * ir;
                }
                lobeFinalize(L[k]);
            }
        }
        if (commit) {
            for (int k = 0; k < K; ++k) {
                L[k].statS = S[k];
                L[k].statRx = R[k][0];
                L[k].statRy = R[k][1];
                L[k].statRz = R[k][2];
            }
            const float *cs = st + (size_t)K * 4;
            F.cells[c].sampleCount = kGuideDecay * F.cells[c].sampleCount + cs[0];
            F.cells[c].weightSum = kGuideDecay * F.cells[c].weightSum + cs[1];
        }
    }
}

// Spatial refinement: cells whose running sample count exceeds the threshold split at the mean position of this
// update's samples along the axis of largest variance; both children start from the parent's mixture with half
// of its statistics. One level per training update. Cells are visited in index order; the left child keeps the
// parent's index, the right child is appended.
inline void guideSplit(GuideField &F, const std::vector<float> &stats, Float maxCellSamples) {
    const int K = F.K;
    const size_t stride = (size_t)K * 4 + 8;
    const uint32_t nc0 = F.numCells();
    // leaf node of every cell
    std::vector<uint32_t> leafOf(nc0, 0);
    for (uint32_t n = 0; n < F.nodes.size(); ++n)
        if (F.nodes[n].axis == 3u) leafOf[F.nodes[n].left] = n;
    for (uint32_t c = 0; c < nc0; ++c) {
        if (!(F.cells[c].sampleCount > maxCellSamples)) continue;
        const float *cs = &stats[stride * c + (size_t)K * 4];
        const Float n = cs[0];
        if (!(n >= 2)) continue;
        Float mean[3];
        double var[3];
        for (int a = 0; a < 3; ++a) {
            mean[a] = cs[2 + a] / n;
            var[a] = (double)cs[5 + a] / (double)n - (double)mean[a] * (double)mean[a];
        }
        int axis = 0;
        if (var[1] > var[axis]) axis = 1;
        if (var[2] > var[axis]) axis = 2;
        if (!(var[axis] > 0)) continue;
        const uint32_t leaf = leafOf[c];
        const uint32_t left = (uint32_t)F.nodes.size();
        const uint32_t newCell = F.numCells();
        F.nodes.push_back(GuideNode{3u, 0.0f, c, 0u});
        F.nodes.push_back(GuideNode{3u, 0.0f, newCell, 0u});
        F.nodes[leaf] = GuideNode{(uint32_t)axis, mean[axis], left, 0u};
        GuideCellHeader h = F.cells[c];
        h.sampleCount *= 0.5f;
        h.weightSum *= 0.5f;
        F.cells[c] = h;
        F.cells.push_back(h);
        for (int k = 0; k < K; ++k) {
            GuideLobe &l = F.lobes[(size_t)c * K + k];
            l.statS *= 0.5f; l.statRx *= 0.5f; l.statRy *= 0.5f; l.statRz *= 0.5f;
        }
        for (int k = 0; k < K; ++k) F.lobes.push_back(F.lobes[(size_t)c * K + k]);
    }
}

// Position statistics of this update's samples per cell of the CURRENT tree (the cell-statistics slots of a stats buffer: count and
// the first / second position moments; the lobe slots stay zero): what a further split level needs after the tree has changed.
inline void guideCellMoments(const GuideField &F, const GuideSamples &smp, std::vector<float> &stats) {
    const size_t stride = (size_t)F.K * 4 + 8;
    std::vector<uint32_t> cell, perm, offsets;
    guideBin(F, smp.pos.data(), smp.size(), cell, perm, offsets);
    std::vector<double> st(stride * F.numCells(), 0.0);
    for (uint32_t c = 0; c < F.numCells(); ++c) {
        double *cs = &st[stride * c + (size_t)F.K * 4];
        for (uint32_t j = offsets[c]; j < offsets[c + 1]; ++j) {  // sorted order, as the E-step sums them
            const Vec3 &p = smp.pos[perm[j]];
            cs[0] += 1.0;
            cs[2] += p.x; cs[3] += p.y; cs[4] += p.z;
            cs[5] += (double)p.x * p.x; cs[6] += (double)p.y * p.y; cs[7] += (double)p.z * p.z;
        }
    }
    stats.assign(st.begin(), st.end());
}

// One complete training update: bin, nIter x (E-step, M-step), commit, split.
// splitLevels > 1: after the regular split the samples are binned into the new tree and every cell whose (halved) running sample
// count still exceeds the threshold splits again, at the mean of ITS samples -- up to splitLevels levels per update, so a field
// that starts from one cell reaches its size in a few updates instead of one level per update. Children keep inheriting the parent's
// mixture with half of its statistics; the mixtures themselves are refined by the next update.
inline void guideTrain(GuideField &F, const GuideSamples &smp, int nIter, Float maxCellSamples, int splitLevels = 1) {
    std::vector<uint32_t> cell, perm, offsets;
    guideBin(F, smp.pos.data(), smp.size(), cell, perm, offsets);
    std::vector<double> statsD;
    std::vector<float> stats;
    for (int it = 0; it < nIter; ++it) {
        guideEStep(F, smp, perm, offsets, statsD);
        stats.assign(statsD.begin(), statsD.end());
        guideMStep(F, stats, it == nIter - 1);
    }
    uint32_t before = F.numCells();
    guideSplit(F, stats, maxCellSamples);
    for (int level = 1; level < splitLevels && F.numCells() != before; ++level) {
        before = F.numCells();
        guideCellMoments(F, smp, stats);
        guideSplit(F, stats, maxCellSamples);
    }
}

}  // namespace orc

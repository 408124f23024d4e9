// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_scene.h: scene containers, SAH kd-tree (gkdtree.h rules), Havran traversal
// (sahkdtree3.h:178-308), TriAccel (triaccel.h:37-158), rectangle (rectangle.cpp:125-168)
// and intersection records (skdtree.h:343-428).
#pragma once
#include <functional>
#include <cstdio>
#include <memory>
#include <vector>

#include "../include/b200pg.h"
#include "oracle_math.h"

// GCC: keep a function free of fused multiply-add contraction (the translation unit is built with -march=x86-64-v3)
#define ORC_NO_CONTRACT __attribute__((optimize("fp-contract=off")))

namespace orc {

struct Ray {
    Vec3 o, d, dRcp;
    Float mint, maxt;
    Ray() : mint(Epsilon), maxt(std::numeric_limits<Float>::infinity()) {}
    Ray(const Vec3 &o_, const Vec3 &d_, Float mint_ = Epsilon, Float maxt_ = std::numeric_limits<Float>::infinity())
        : o(o_), d(d_), mint(mint_), maxt(maxt_) {
        dRcp = Vec3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // ray.h:  setDirection
    }
    Vec3 operator()(Float t) const { return o + d * t; }
};

struct AABB {
    Vec3 min, max;
    AABB() { reset(); }
    void reset() {
        min = Vec3(std::numeric_limits<Float>::infinity());
        max = Vec3(-std::numeric_limits<Float>::infinity());
    }
    void expandBy(const Vec3 &p) {
        for (int i = 0; i < 3; ++i) {
            min[i] = std::min(min[i], p[i]);
            max[i] = std::max(max[i], p[i]);
        }
    }
    void expandBy(const AABB &b) {
        expandBy(b.min);
        expandBy(b.max);
    }
    Float surfaceArea() const {
        Vec3 d = max - min;
        return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    // aabb.h:308-338
    bool rayIntersect(const Ray &ray, Float &nearT, Float &farT) const {
        nearT = -std::numeric_limits<Float>::infinity();
        farT = std::numeric_limits<Float>::infinity();
        for (int i = 0; i < 3; i++) {
            const Float origin = ray.o[i];
            const Float minVal = min[i], maxVal = max[i];
            if (ray.d[i] == 0) {
                if (origin < minVal || origin > maxVal) return false;
            } else {
                Float t1 = (minVal - origin) * ray.dRcp[i];
                Float t2 = (maxVal - origin) * ray.dRcp[i];
                if (t1 > t2) std::swap(t1, t2);
                nearT = std::max(t1, nearT);
                farT = std::min(t2, farT);
                if (!(nearT <= farT)) return false;
            }
        }
        return true;
    }
};

static const uint32_t KNoTriangleFlag = 0xFFFFFFFFu;

// triaccel.h:37-94
struct TriAccel {
    uint32_t k;
    Float n_u, n_v, n_d;
    Float a_u, a_v, b_nu, b_nv;
    Float c_nu, c_nv;
    uint32_t shapeIndex, primIndex;

    int load(const Vec3 &A, const Vec3 &B, const Vec3 &C) {
        static const int waldModulo[4] = {1, 2, 0, 1};
        Vec3 b = C - A, c = B - A, N = cross(c, b);
        k = 0;
        for (int j = 0; j < 3; j++)
            if (std::abs(N[j]) > std::abs(N[k])) k = j;
        uint32_t u = waldModulo[k], v = waldModulo[k + 1];
        const Float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
        if (denom == 0) {
            k = 3;
            return 1;
        }
        n_u = N[u] / n_k;
        n_v = N[v] / n_k;
        n_d = dot(A, N) / n_k;
        b_nu = b[u] / denom;
        b_nv = -b[v] / denom;
        a_u = A[u];
        a_v = A[v];
        c_nu = c[v] / denom;
        c_nv = -c[u] / denom;
        return 0;
    }

    // triaccel.h:96-158
    bool rayIntersect(const Ray &ray, Float mint, Float maxt, Float &u, Float &v, Float &t) const {
        Float o_u, o_v, o_k, d_u, d_v, d_k;
        switch (k) {
            case 0: o_u = ray.o[1]; o_v = ray.o[2]; o_k = ray.o[0]; d_u = ray.d[1]; d_v = ray.d[2]; d_k = ray.d[0]; break;
            case 1: o_u = ray.o[2]; o_v = ray.o[0]; o_k = ray.o[1]; d_u = ray.d[2]; d_v = ray.d[0]; d_k = ray.d[1]; break;
            case 2: o_u = ray.o[0]; o_v = ray.o[1]; o_k = ray.o[2]; d_u = ray.d[0]; d_v = ray.d[1]; d_k = ray.d[2]; break;
            default: return false;
        }
        t = (n_d - o_u * n_u - o_v * n_v - o_k) / (d_u * n_u + d_v * n_v + d_k);
        if (t < mint || t > maxt) return false;
        const Float hu = o_u + t * d_u - a_u;
        const Float hv = o_v + t * d_v - a_v;
        u = hv * b_nu + hu * b_nv;
        v = hu * c_nu + hv * c_nv;
        return u >= 0 && v >= 0 && u + v <= 1.0f;
    }
};

struct Shape {
    int type;
    int bsdf, emitter, interiorMedium, exteriorMedium;
    // rectangle
    Mat4 objectToWorld, worldToObject;
    Frame frame;
    Vec3 dpdu, dpdv;
    Float invSurfaceArea;
    // trimesh
    std::vector<Vec3> positions, normals;
    std::vector<Vec2> texcoords;
    std::vector<uint32_t> indices;
    std::vector<Vec3> uvTangents;  // per-triangle dpdu of TriMesh::computeUVTangents (trimesh.cpp:683-735); empty without texcoords
    std::vector<Float> areaCdf;  // DiscreteDistribution m_cdf (pmf.h)
    uint32_t primOffset;         // first global primitive id
    bool isMediumTransition() const { return interiorMedium >= 0 || exteriorMedium >= 0; }
};

struct Intersection {
    Float t;
    Vec3 p;
    Vec3 geoN;      // geoFrame.n
    Frame shFrame;
    Vec2 uv;
    Vec3 wi;
    Vec3 dpdu;
    int shape;
    uint32_t primIndex;
    bool isValid() const { return t != std::numeric_limits<Float>::infinity(); }
    Vec3 toWorld(const Vec3 &v) const { return shFrame.toWorld(v); }
    Vec3 toLocal(const Vec3 &v) const { return shFrame.toLocal(v); }
};

// DiscreteDistribution::sample / sampleReuse (pmf.h:124-188) over a normalized cdf
inline size_t cdfSample(const std::vector<Float> &cdf, Float sampleValue) {
    auto entry = std::lower_bound(cdf.begin(), cdf.end(), sampleValue);
    size_t index = std::min(cdf.size() - 2, (size_t)std::max((ptrdiff_t)0, (ptrdiff_t)(entry - cdf.begin()) - 1));
    while ((cdf[index + 1] - cdf[index]) == 0 && index < cdf.size() - 1) ++index;
    return index;
}
inline size_t cdfSampleReuse(const std::vector<Float> &cdf, Float &sampleValue, Float &pdf) {
    size_t index = cdfSample(cdf, sampleValue);
    pdf = cdf[index + 1] - cdf[index];
    sampleValue = (sampleValue - cdf[index]) / (cdf[index + 1] - cdf[index]);
    return index;
}
inline Float cdfNormalize(std::vector<Float> &cdf) {  // pmf.h:98-112
    Float sum = cdf[cdf.size() - 1];
    if (sum > 0) {
        Float normalization = 1.0f / sum;
        for (size_t i = 1; i < cdf.size(); ++i) cdf[i] *= normalization;
        cdf[cdf.size() - 1] = 1.0f;
    }
    return sum;
}

// Triangle::getClippedAABB (src/libcore/triangle.cpp:69-138): Sutherland-Hodgman clipping of the triangle against the six
// planes of `box`, in double precision ("the kd-tree code will frequently call this function with almost-collapsed
// AABBs"), bounds rounded outward (castflt_down / castflt_up), then clipped to the box.
inline int sutherlandHodgman(const double (*in)[3], int inCount, double (*out)[3], int axis, double splitPos, bool isMinimum) {
    if (inCount < 3) return 0;
    double cur[3] = {in[0][0], in[0][1], in[0][2]};
    const double sign = isMinimum ? 1.0 : -1.0;
    bool curIsInside = sign * (cur[axis] - splitPos) >= 0;
    int outCount = 0;
    for (int i = 0; i < inCount; ++i) {
        const int nextIdx = i + 1 == inCount ? 0 : i + 1;
        const double next[3] = {in[nextIdx][0], in[nextIdx][1], in[nextIdx][2]};
        const bool nextIsInside = sign * (next[axis] - splitPos) >= 0;
        if (curIsInside && nextIsInside) {
            for (int c = 0; c < 3; ++c) out[outCount][c] = next[c];
            ++outCount;
        } else if (curIsInside != nextIsInside) {
            const double t = (splitPos - cur[axis]) / (next[axis] - cur[axis]);
            for (int c = 0; c < 3; ++c) out[outCount][c] = cur[c] + (next[c] - cur[c]) * t;
            out[outCount][axis] = splitPos;  // avoid roundoff errors
            ++outCount;
            if (nextIsInside) {
                for (int c = 0; c < 3; ++c) out[outCount][c] = next[c];
                ++outCount;
            }
        }
        for (int c = 0; c < 3; ++c) cur[c] = next[c];
        curIsInside = nextIsInside;
    }
    return outCount;
}
inline AABB clippedTriangleAABB(const Vec3 &p0, const Vec3 &p1, const Vec3 &p2, const AABB &box) {
    double v1[10][3] = {{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, {p2.x, p2.y, p2.z}}, v2[10][3];
    int n = 3;
    for (int axis = 0; axis < 3; ++axis) {
        n = sutherlandHodgman(v1, n, v2, axis, box.min[axis], true);
        n = sutherlandHodgman(v2, n, v1, axis, box.max[axis], false);
    }
    AABB r;
    r.reset();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) {
            const double pos = v1[i][j];
            float lo = (float)pos, hi = (float)pos;
            if ((double)lo > pos) lo = std::nextafter(lo, -std::numeric_limits<float>::infinity());  // castflt_down
            if ((double)hi < pos) hi = std::nextafter(hi, std::numeric_limits<float>::infinity());   // castflt_up
            r.min[j] = std::min(r.min[j], lo);
            r.max[j] = std::max(r.max[j], hi);
        }
    for (int j = 0; j < 3; ++j) {  // AABB::clip
        r.min[j] = std::max(r.min[j], box.min[j]);
        r.max[j] = std::min(r.max[j], box.max[j]);
    }
    return r;
}

// ---------------------------------------------------------------------------
// kd-tree: 8-byte nodes (gkdtree.h:452-582): leaf = {start|0x80000000, end};
// inner = {(leftOffset << 2) | axis, split}; children adjacent (right = left + 1).
// ---------------------------------------------------------------------------
struct KDNode {
    uint32_t a;
    union {
        float split;
        uint32_t end;
    };
    bool isLeaf() const { return a & 0x80000000u; }
    uint32_t primStart() const { return a & 0x7FFFFFFFu; }
    uint32_t primEnd() const { return end; }
    int axis() const { return a & 3u; }
    uint32_t leftOffset() const { return (a & 0x7FFFFFFFu) >> 2; }
};

struct TraversalCounters {
    uint64_t nodes = 0, indices = 0, prims = 0;
    uint64_t leaves = 0;  // leaf visits among `nodes` (the reference's numTraversals counts the inner ones only, sahkdtree3.h:357)
};

class KDTree {
public:
    // build defaults, gkdtree.h:734-744
    Float traversalCost = 15, queryCost = 20, emptySpaceBonus = 0.9f;
    uint32_t stopPrims = 6, maxBadRefines = 3, exactPrimThreshold = 65536, minMaxBins = 128;
    int maxDepth = 0;

    std::vector<KDNode> nodes;
    std::vector<uint32_t> indices;
    AABB aabb, tightAABB;

    // clip(prim, box) = bounds of the part of primitive `prim` inside `box` (Triangle::getClippedAABB for triangles,
    // AABB intersection for other shapes, Shape::getClippedAABB); null = no clipping (m_clip = false)
    typedef std::function<AABB(uint32_t, const AABB &)> ClipFn;

    void build(const std::vector<AABB> &primBoxes, const ClipFn &clip = nullptr) {
        clipFn = clip;
        uint32_t n = (uint32_t)primBoxes.size();
        aabb.reset();
        for (uint32_t i = 0; i < n; ++i) aabb.expandBy(primBoxes[i]);
        tightAABB = aabb;
        if (maxDepth == 0) {
            int lg = 0;
            for (uint32_t v = n; v > 1; v >>= 1) ++lg;          // math::log2i
            maxDepth = (int)(8 + 1.3f * lg);                    // gkdtree.h:986
        }
        maxDepth = std::min(maxDepth, 48);                      // MTS_KD_MAXDEPTH
        nodes.clear();
        indices.clear();
        nodes.push_back(KDNode());
        primBoxesPtr = &primBoxes;
        // more primitives than the exact method takes: the reference builds in parallel on a multi-core host (gkdtree.h:980-981,
        // 1036-1037), and a subtree handed to a worker is never torn down again by its min-max ancestors (:1751-1752)
        parallelBuild = n > exactPrimThreshold;
        std::vector<uint32_t> all(n);
        for (uint32_t i = 0; i < n; ++i) all[i] = i;
        buildMinMax(0, all, aabb, aabb, 1, 0);  // depth 1 = root (gkdtree.h:1050)
        primBoxesPtr = nullptr;
        // enlarge after the build (gkdtree.h:1214-1219)
        const Float eps = 1e-3f;
        aabb.min = aabb.min - ((aabb.max - aabb.min) * eps + Vec3(eps));
        aabb.max = aabb.max + ((aabb.max - aabb.min) * eps + Vec3(eps));
        clipFn = nullptr;
    }

    uint64_t prunedPrims = 0;  // references removed by clipping ("perfect splits")
    uint64_t retractedSplits = 0;  // splits torn up again because the subtree did not beat the leaf cost

private:
    ClipFn clipFn;
    const std::vector<AABB> *primBoxesPtr = nullptr;
    bool parallelBuild = false;
    struct Item {
        uint32_t prim;
        AABB box;  // bounds of the primitive inside the node it currently belongs to
    };

    struct Event {
        Float pos;
        int type;  // 0 = end, 1 = planar, 2 = start (ordering of gkdtree.h:1331-1377)
        bool operator<(const Event &o) const { return pos < o.pos || (pos == o.pos && type < o.type); }
    };

    void makeLeaf(uint32_t nodeIdx, const std::vector<Item> &prims) {
        KDNode &nd = nodes[nodeIdx];
        nd.a = 0x80000000u | (uint32_t)indices.size();
        for (const Item &it : prims) indices.push_back(it.prim);
        nd.end = (uint32_t)indices.size();
    }
    static bool usable(const AABB &b) {  // isValid() && getSurfaceArea() > 0 (gkdtree.h:1563-1564, 2228)
        return b.min.x <= b.max.x && b.min.y <= b.max.y && b.min.z <= b.max.z && b.surfaceArea() > 0;
    }

    // SurfaceAreaHeuristic3 (sahkdtree3.h:39-84): probabilities = SA(child) / SA(node), in the reference's order of operations
    // (m_temp0 = (e1 e2) temp, m_temp1 = (e1 + e2) temp, p = m_temp0 + m_temp1 * width) and without contraction into fused
    // multiply-adds (the reference build has none): on a regular mesh whole families of candidate planes have the same cost in
    // exact arithmetic, and which one wins the strict `<` below is decided by the last bit. With this the port's tree IS the
    // reference's tree on every scene compared (tests/test_ref_pin.py::test_kd_tree_build_and_traversal_counters).
    ORC_NO_CONTRACT inline Float sahCost(const AABB &box, int axis, Float split, uint32_t nL, uint32_t nR) const {
        Vec3 d = box.max - box.min;
        int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        const Float temp = 1.0f / (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]);
        const Float temp0 = (d[a1] * d[a2]) * temp, temp1 = (d[a1] + d[a2]) * temp;
        const Float pL = temp0 + temp1 * (split - box.min[axis]);
        const Float pR = temp0 + temp1 * (box.max[axis] - split);
        Float cost = traversalCost + queryCost * (pL * (Float)nL + pR * (Float)nR);
        if (nL == 0 || nR == 0) cost *= emptySpaceBonus;  // gkdtree.h:2039-2041
        return cost;
    }

    void makeLeafIds(uint32_t nodeIdx, const std::vector<uint32_t> &ids) {
        KDNode &nd = nodes[nodeIdx];
        nd.a = 0x80000000u | (uint32_t)indices.size();
        indices.insert(indices.end(), ids.begin(), ids.end());
        nd.end = (uint32_t)indices.size();
    }

    // transitionToNLogN (gkdtree.h:1729-1761): clip the primitives to the node, continue with the exact builder
    Float transition(uint32_t nodeIdx, std::vector<uint32_t> &ids, const AABB &nodeBox, int depth, uint32_t badRefines) {
        std::vector<Item> items(ids.size());
        for (size_t i = 0; i < ids.size(); ++i) items[i] = Item{ids[i], (*primBoxesPtr)[ids[i]]};
        std::vector<uint32_t>().swap(ids);
        const Float cost = buildNode(nodeIdx, items, nodeBox, depth, badRefines, false);
        return parallelBuild ? -std::numeric_limits<Float>::infinity() : cost;
    }

    // buildTreeMinMax (gkdtree.h:1792-1925) with MinMaxBins (:2412-2600): min-max binning over the TIGHT bounds of the node on the
    // primitives' full boxes; the partition goes by bin index, the children's tight bounds are the unions of their primitives'
    // boxes cut at the split plane. Arithmetic in the reference's order and without contraction (see sahCost).
    ORC_NO_CONTRACT Float buildMinMax(uint32_t nodeIdx, std::vector<uint32_t> &ids, const AABB &nodeBox, const AABB &tightBox, int depth,
                                      uint32_t badRefines) {
        const uint32_t n = (uint32_t)ids.size();
        const Float leafCost = n * queryCost;
        if (n <= stopPrims || depth >= maxDepth) {
            makeLeafIds(nodeIdx, ids);
            return leafCost;
        }
        if (n <= exactPrimThreshold) return transition(nodeIdx, ids, nodeBox, depth, badRefines);
        const std::vector<AABB> &pb = *primBoxesPtr;
        const int B = (int)minMaxBins;
        float mn[3], binSize[3], invBinSize[3];
        for (int a = 0; a < 3; ++a) {  // MinMaxBins::setAABB
            mn[a] = tightBox.min[a];
            binSize[a] = (tightBox.max[a] - tightBox.min[a]) / B;
            invBinSize[a] = 1 / binSize[a];
        }
        auto computeIndex = [&](float pos, int a) {
            return (uint32_t)std::min((float)(B - 1), std::max(0.0f, (pos - mn[a]) * invBinSize[a]));
        };
        std::vector<uint32_t> minBins(3 * (size_t)B, 0u), maxBins(3 * (size_t)B, 0u);
        for (uint32_t id : ids)
            for (int a = 0; a < 3; ++a) {
                minBins[a * B + computeIndex(pb[id].min[a], a)]++;
                maxBins[a * B + computeIndex(pb[id].max[a], a)]++;
            }
        // minimizeCost: no empty-space bonus
        Float bestCost = std::numeric_limits<Float>::infinity();
        int bestAxis = -1, leftBin = -1;
        uint32_t bestLeft = 0, bestRight = 0;
        {
            Vec3 d = tightBox.max - tightBox.min;
            const Float temp = 1.0f / (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]);
            int binIdx = 0;
            for (int a = 0; a < 3; ++a) {
                const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                const Float temp0 = (d[a1] * d[a2]) * temp, temp1 = (d[a1] + d[a2]) * temp;
                uint32_t numLeft = 0, numRight = n;
                Float leftWidth = 0, rightWidth = tightBox.max[a] - tightBox.min[a];
                for (int i = 0; i < B - 1; ++i) {
                    numLeft += minBins[binIdx];
                    numRight -= maxBins[binIdx];
                    leftWidth += binSize[a];
                    rightWidth -= binSize[a];
                    const Float pL = temp0 + temp1 * leftWidth, pR = temp0 + temp1 * rightWidth;
                    const Float cost = traversalCost + queryCost * (pL * numLeft + pR * numRight);
                    if (cost < bestCost) {
                        bestCost = cost;
                        bestAxis = a;
                        bestLeft = numLeft;
                        bestRight = numRight;
                        leftBin = i;
                    }
                    binIdx++;
                }
                binIdx++;
            }
        }
        if (bestCost == std::numeric_limits<Float>::infinity())  // collapsed bounds: the exact method takes over (:1817-1832)
            return transition(nodeIdx, ids, nodeBox, depth, badRefines);
        if (bestCost >= leafCost) {
            if ((bestCost > 4 * leafCost && n < 16) || badRefines >= maxBadRefines) {
                makeLeafIds(nodeIdx, ids);
                return leafCost;
            }
            ++badRefines;
        }
        // MinMaxBins::partition
        std::vector<uint32_t> left, right;
        left.reserve(bestLeft);
        right.reserve(bestRight);
        AABB lb, rb;
        lb.reset();
        rb.reset();
        for (uint32_t id : ids) {
            const AABB &b = pb[id];
            const int startIdx = (int)computeIndex(b.min[bestAxis], bestAxis), endIdx = (int)computeIndex(b.max[bestAxis], bestAxis);
            if (endIdx <= leftBin) {
                lb.expandBy(b);
                left.push_back(id);
            } else if (startIdx > leftBin) {
                rb.expandBy(b);
                right.push_back(id);
            } else {
                lb.expandBy(b);
                rb.expandBy(b);
                left.push_back(id);
                right.push_back(id);
            }
        }
        for (int a = 0; a < 3; ++a) {  // AABB::clip(m_aabb)
            lb.min[a] = std::max(lb.min[a], tightBox.min[a]);
            lb.max[a] = std::min(lb.max[a], tightBox.max[a]);
            rb.min[a] = std::max(rb.min[a], tightBox.min[a]);
            rb.max[a] = std::min(rb.max[a], tightBox.max[a]);
        }
        const Float pos = mn[bestAxis] + binSize[bestAxis] * (leftBin + 1);
        lb.max[bestAxis] = std::min(lb.max[bestAxis], pos);
        rb.min[bestAxis] = std::max(rb.min[bestAxis], pos);
        std::vector<uint32_t>().swap(ids);
        const size_t nodePosBeforeSplit = nodes.size() + 2, indexPosBeforeSplit = indices.size();
        const uint32_t leftIdx = (uint32_t)nodes.size();
        nodes.push_back(KDNode());
        nodes.push_back(KDNode());
        nodes[nodeIdx].a = ((leftIdx - nodeIdx) << 2) | (uint32_t)bestAxis;
        nodes[nodeIdx].split = pos;
        AABB child = nodeBox;
        child.max[bestAxis] = pos;
        const Float leftCost = buildMinMax(leftIdx, left, child, lb, depth + 1, badRefines);
        child.min[bestAxis] = pos;
        child.max[bestAxis] = nodeBox.max[bestAxis];
        const Float rightCost = buildMinMax(leftIdx + 1, right, child, rb, depth + 1, badRefines);
        const Float finalCost = innerCost(nodeBox, bestAxis, pos, leftCost, rightCost);
        if (finalCost < leafCost) return finalCost;
        retract(nodeIdx, nodePosBeforeSplit - 2, indexPosBeforeSplit);
        return leafCost;
    }

    // Retraction (m_retract, gkdtree.h:1910-1921 / 2385-2394): in the end splitting did not reduce the cost. Tear up everything below
    // the node and make it a leaf over the primitives its subtree's leaves reference (createLeafAfterRetraction, :1665-1699: the
    // index entries written since the split, sorted, without duplicates).
    void retract(uint32_t nodeIdx, size_t nodePos, size_t indexPos) {
        nodes.resize(nodePos);
        std::sort(indices.begin() + indexPos, indices.end());
        indices.erase(std::unique(indices.begin() + indexPos, indices.end()), indices.end());
        nodes[nodeIdx].a = 0x80000000u | (uint32_t)indexPos;
        nodes[nodeIdx].end = (uint32_t)indices.size();
        ++retractedSplits;
    }

    // the final cost of an inner node from its children's (gkdtree.h:2363-2370), same arithmetic as sahCost
    ORC_NO_CONTRACT Float innerCost(const AABB &box, int axis, Float split, Float leftCost, Float rightCost) const {
        Vec3 d = box.max - box.min;
        int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        const Float temp = 1.0f / (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]);
        const Float temp0 = (d[a1] * d[a2]) * temp, temp1 = (d[a1] + d[a2]) * temp;
        const Float pL = temp0 + temp1 * (split - box.min[axis]);
        const Float pR = temp0 + temp1 * (box.max[axis] - split);
        return traversalCost + (pL * leftCost + pR * rightCost);
    }

    // returns the node's cost, as buildTreeMinMax / buildTreeSAH do (the parent's retraction decision needs it)
    Float buildNode(uint32_t nodeIdx, std::vector<Item> &prims, const AABB &box, int depth, uint32_t badRefines, bool exactStage) {
        uint32_t n = (uint32_t)prims.size();
        Float leafCost = n * queryCost;
        if (n <= stopPrims || depth >= maxDepth) {  // gkdtree.h:1797-1800
            makeLeaf(nodeIdx, prims);
            return leafCost;
        }
        Float bestCost = std::numeric_limits<Float>::infinity(), bestSplit = 0;
        int bestAxis = -1;
        bool bestPlanarLeft = true;
        const bool exact = true;  // buildTreeSAH and its recursion; the min-max stage is buildMinMax below

        if (!exactStage && clipFn) {
            // passing from min-max binning to the O(n log n) builder: clip every primitive to the node (createEventList,
            // gkdtree.h:1551-1592); primitives whose clipped bounds are invalid or have no area drop out
            std::vector<Item> kept;
            kept.reserve(n);
            for (const Item &it : prims) {
                AABB c = clipFn(it.prim, box);
                if (usable(c)) kept.push_back(Item{it.prim, c});
                else ++prunedPrims;
            }
            prims.swap(kept);
            n = (uint32_t)prims.size();
            leafCost = n * queryCost;  // buildTreeSAH starts over with the clipped primitive count (gkdtree.h:1958)
            if (n <= stopPrims) {
                makeLeaf(nodeIdx, prims);
                return leafCost;
            }
        }

        {
            // exact sweep over sorted edge events (gkdtree.h:1954-2405)
            std::vector<Event> ev;
            ev.reserve(2 * n);
            for (int axis = 0; axis < 3; ++axis) {
                ev.clear();
                for (const Item &it : prims) {
                    const AABB &b = it.box;
                    Float lo = std::max(b.min[axis], box.min[axis]), hi = std::min(b.max[axis], box.max[axis]);
                    if (lo == hi) {
                        ev.push_back({lo, 1});
                    } else {
                        ev.push_back({lo, 2});
                        ev.push_back({hi, 0});
                    }
                }
                std::sort(ev.begin(), ev.end());
                uint32_t nL = 0, nR = n;
                size_t i = 0;
                while (i < ev.size()) {
                    Float pos = ev[i].pos;
                    uint32_t pEnd = 0, pPlanar = 0, pStart = 0;
                    while (i < ev.size() && ev[i].pos == pos && ev[i].type == 0) { ++pEnd; ++i; }
                    while (i < ev.size() && ev[i].pos == pos && ev[i].type == 1) { ++pPlanar; ++i; }
                    while (i < ev.size() && ev[i].pos == pos && ev[i].type == 2) { ++pStart; ++i; }
                    nR -= pPlanar + pEnd;
                    if (pos > box.min[axis] && pos < box.max[axis]) {
                        Float cL = sahCost(box, axis, pos, nL + pPlanar, nR);
                        Float cR = sahCost(box, axis, pos, nL, nR + pPlanar);
                        if (cL < bestCost || cR < bestCost) {  // planar prims to the cheaper side (:2050-2078)
                            bestPlanarLeft = cL < cR;  // a tie goes to the right (gkdtree.h:2064-2074)
                            bestCost = std::min(cL, cR);
                            bestAxis = axis;
                            bestSplit = pos;
                        }
                    }
                    nL += pStart + pPlanar;
                }
            }
        }

        if (bestAxis < 0) {
            makeLeaf(nodeIdx, prims);
            return leafCost;
        }
        if (bestCost >= leafCost) {  // "bad refines", gkdtree.h:1835-1843 / 2105-2112
            if ((bestCost > 4 * leafCost && n < 16) || badRefines >= maxBadRefines) {
                makeLeaf(nodeIdx, prims);
                return leafCost;
            }
            ++badRefines;
        }

        AABB lb = box, rb = box;
        lb.max[bestAxis] = bestSplit;
        rb.min[bestAxis] = bestSplit;
        std::vector<Item> left, right;
        for (const Item &it : prims) {
            const AABB &b = it.box;
            Float lo = std::max(b.min[bestAxis], box.min[bestAxis]), hi = std::min(b.max[bestAxis], box.max[bestAxis]);
            if (lo == hi && lo == bestSplit) {
                (bestPlanarLeft ? left : right).push_back(it);
            } else if (lo < bestSplit && hi > bestSplit) {
                // the primitive overlaps the split plane: re-clip for each side ("perfect splits", gkdtree.h:2218-2265)
                if (exact && clipFn) {
                    AABB cl = clipFn(it.prim, lb), cr = clipFn(it.prim, rb);
                    if (usable(cl)) left.push_back(Item{it.prim, cl}); else ++prunedPrims;
                    if (usable(cr)) right.push_back(Item{it.prim, cr}); else ++prunedPrims;
                } else {
                    left.push_back(it);
                    right.push_back(it);
                }
            } else {
                if (lo < bestSplit) left.push_back(it);
                if (hi > bestSplit) right.push_back(it);
            }
        }
        if (left.size() == n && right.size() == n) {  // no progress at all
            makeLeaf(nodeIdx, prims);
            return leafCost;
        }
        std::vector<Item>().swap(prims);
        const size_t nodePosBeforeSplit = nodes.size(), indexPosBeforeSplit = indices.size();
        uint32_t leftIdx = (uint32_t)nodes.size();
        nodes.push_back(KDNode());
        nodes.push_back(KDNode());
        nodes[nodeIdx].a = ((leftIdx - nodeIdx) << 2) | (uint32_t)bestAxis;
        nodes[nodeIdx].split = bestSplit;
        const Float leftCost = buildNode(leftIdx, left, lb, depth + 1, badRefines, exact);
        const Float rightCost = buildNode(leftIdx + 1, right, rb, depth + 1, badRefines, exact);
        const Float finalCost = innerCost(box, bestAxis, bestSplit, leftCost, rightCost);
        if (finalCost < leafCost) return finalCost;
        retract(nodeIdx, nodePosBeforeSplit, indexPosBeforeSplit);
        return leafCost;
    }
};

struct IntersectionCache {  // skdtree.h:237-241
    uint32_t shapeIndex, primIndex;
    Float u, v;
};

struct Scene;  // fwd

}  // namespace orc

// host_scene.cpp -- see host_scene.h. Host-side "scene compiler" of the B200 path tracer.
#include "host_scene.h"

#include <dlfcn.h>
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace pg {

namespace {

struct V3 {
    float x, y, z;
    float operator[](int i) const { return (&x)[i]; }
    float &operator[](int i) { return (&x)[i]; }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float len(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) {
    float r = 1.0f / len(a);
    return v3(a.x * r, a.y * r, a.z * r);
}

// 4x4 inverse in double precision (adjugate); returns false when singular.
bool invert4(const float *m, float *out) {
    double a[16], inv[16];
    for (int i = 0; i < 16; ++i) a[i] = m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0) return false;
    det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out[i] = (float)(inv[i] * det);
    return true;
}
void mul4(const float *a, const float *b, float *o) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    std::memcpy(o, r, sizeof(r));
}
inline V3 xfPoint(const float *m, V3 p) {  // Transform::operator()(Point)
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return v3(x, y, z);
    float r = 1.0f / w;
    return v3(x * r, y * r, z * r);
}
inline V3 xfVector(const float *m, V3 v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
inline V3 xfNormalFromInverse(const float *inv, V3 n) {  // inverse transpose
    return v3(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
              inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
}
inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// ------------------------------------------------------------------ BVH builder (binned SAH)
struct BPrim {
    float bmin[3], bmax[3], c[3];
    uint32_t id;
};
struct Box {
    float mn[3], mx[3];
    void reset() {
        for (int i = 0; i < 3; ++i) {
            mn[i] = std::numeric_limits<float>::infinity();
            mx[i] = -std::numeric_limits<float>::infinity();
        }
    }
    void grow(const float *a, const float *b) {
        for (int i = 0; i < 3; ++i) {
            mn[i] = std::min(mn[i], a[i]);
            mx[i] = std::max(mx[i], b[i]);
        }
    }
    float area() const {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (!(dx >= 0)) return 0;
        return 2 * (dx * dy + dy * dz + dz * dx);
    }
};

struct TmpNode {
    Box box[2];
    int32_t child[2];
};

struct Builder {
    std::vector<BPrim> &P;
    std::vector<TmpNode> nodes;
    std::atomic<int> next{0};
    static const int kLeaf = 4, kBins = 16, kMedianDepth = 36;
    explicit Builder(std::vector<BPrim> &p) : P(p) { nodes.resize(std::max<size_t>(1, p.size())); }

    static int32_t leafCode(size_t begin, size_t count) { return ~(int32_t)((begin << 4) | count); }

    // returns child reference, fills box
    int32_t build(size_t begin, size_t end, Box &box, int depth) {
        box.reset();
        Box cb;
        cb.reset();
        for (size_t i = begin; i < end; ++i) {
            box.grow(P[i].bmin, P[i].bmax);
            cb.grow(P[i].c, P[i].c);
        }
        size_t n = end - begin;
        if (n <= (size_t)kLeaf) return leafCode(begin, n);

        int bestAxis = -1, bestBin = -1;
        float bestCost = std::numeric_limits<float>::infinity();
        for (int axis = 0; axis < 3; ++axis) {
            float lo = cb.mn[axis], hi = cb.mx[axis];
            if (!(hi > lo)) continue;
            Box bb[kBins];
            uint32_t cnt[kBins];
            for (int b = 0; b < kBins; ++b) {
                bb[b].reset();
                cnt[b] = 0;
            }
            float scale = kBins / (hi - lo);
            for (size_t i = begin; i < end; ++i) {
                int b = std::min(kBins - 1, std::max(0, (int)((P[i].c[axis] - lo) * scale)));
                bb[b].grow(P[i].bmin, P[i].bmax);
                cnt[b]++;
            }
            float rightArea[kBins];
            uint32_t rightCnt[kBins];
            Box acc;
            acc.reset();
            uint32_t c = 0;
            for (int b = kBins - 1; b > 0; --b) {
                if (cnt[b]) acc.grow(bb[b].mn, bb[b].mx);
                c += cnt[b];
                rightArea[b] = acc.area();
                rightCnt[b] = c;
            }
            acc.reset();
            c = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                if (cnt[b]) acc.grow(bb[b].mn, bb[b].mx);
                c += cnt[b];
                if (c == 0 || rightCnt[b + 1] == 0) continue;
                float cost = acc.area() * c + rightArea[b + 1] * rightCnt[b + 1];
                if (cost < bestCost) {
                    bestCost = cost;
                    bestAxis = axis;
                    bestBin = b;
                }
            }
        }
        size_t mid;
        if (depth >= kMedianDepth) {
            // depth bound for the traversal stack (device_scene.cuh: kTraceStack = 64): from here on the primitives are halved
            // along the widest centroid axis, so the subtree adds at most log2(n) <= 24 levels
            int axis = 0;
            for (int a = 1; a < 3; ++a)
                if (cb.mx[a] - cb.mn[a] > cb.mx[axis] - cb.mn[axis]) axis = a;
            mid = begin + n / 2;
            std::nth_element(&P[begin], &P[mid], &P[begin] + n, [=](const BPrim &a, const BPrim &b) { return a.c[axis] < b.c[axis]; });
        } else if (bestAxis < 0) {
            mid = begin + n / 2;  // coincident centroids: split by index
        } else {
            float lo = cb.mn[bestAxis], hi = cb.mx[bestAxis];
            float scale = kBins / (hi - lo);
            int axis = bestAxis, bin = bestBin;
            BPrim *m = std::partition(&P[begin], &P[begin] + n, [=](const BPrim &p) {
                int b = std::min(kBins - 1, std::max(0, (int)((p.c[axis] - lo) * scale)));
                return b <= bin;
            });
            mid = m - &P[0];
            if (mid == begin || mid == end) mid = begin + n / 2;
        }
        int idx = next.fetch_add(1);
        TmpNode &nd = nodes[idx];
        if (n > 200000 && depth < 8) {
            int32_t c0, c1;
            Box b0, b1;
#pragma omp task shared(c0, b0) firstprivate(begin, mid, depth)
            c0 = build(begin, mid, b0, depth + 1);
#pragma omp task shared(c1, b1) firstprivate(mid, end, depth)
            c1 = build(mid, end, b1, depth + 1);
#pragma omp taskwait
            nd.child[0] = c0;
            nd.child[1] = c1;
            nd.box[0] = b0;
            nd.box[1] = b1;
        } else {
            nd.child[0] = build(begin, mid, nd.box[0], depth + 1);
            nd.child[1] = build(mid, end, nd.box[1], depth + 1);
        }
        return idx;
    }
};

std::string dataDir() {
    const char *env = std::getenv("B200PG_DATA_DIR");
    if (env && *env) return env;
    Dl_info info;
    if (dladdr((void *)&dataDir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        if (s != std::string::npos) {
            const std::string dir = p.substr(0, s);
            // A/B variants of the library live one level down (_variants/) and share the package's tables
            if (FILE *f = std::fopen((dir + "/data/rtrans_beckmann.bin").c_str(), "rb")) {
                std::fclose(f);
                return dir + "/data";
            }
            return dir + "/../data";
        }
    }
    return "data";
}

}  // namespace

// ------------------------------------------------------------------ rough transmittance
namespace {
// Catmull-Rom tensor-product interpolation with one-sided end stencils, knots on [0,1]
// (the scheme src/libcore/spline.cpp uses for evalCubicInterp{1,2,3}D).
void cubicWeights(float x, int size, float w[4], int &knot) {
    float t = x * (size - 1);
    knot = std::min((int)t, size - 2);
    t -= (float)knot;
    float t2 = t * t, t3 = t2 * t;
    w[0] = 0;
    w[1] = 2 * t3 - 3 * t2 + 1;
    w[2] = -2 * t3 + 3 * t2;
    w[3] = 0;
    float d0 = t3 - 2 * t2 + t, d1 = t3 - t2;
    if (knot > 0) { w[2] += 0.5f * d0; w[0] -= 0.5f * d0; } else { w[2] += d0; w[1] -= d0; }
    if (knot + 2 < size) { w[3] += 0.5f * d1; w[1] -= 0.5f * d1; } else { w[2] += d1; w[1] -= d1; }
}
float interp3(const float *v, int nx, int ny, int nz, float x, float y, float z) {
    if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1 && z >= 0 && z <= 1)) return 0;
    float wx[4], wy[4], wz[4];
    int kx, ky, kz;
    cubicWeights(x, nx, wx, kx);
    cubicWeights(y, ny, wy, ky);
    cubicWeights(z, nz, wz, kz);
    float r = 0;
    for (int c = -1; c <= 2; ++c)
        for (int b = -1; b <= 2; ++b) {
            float wyz = wy[b + 1] * wz[c + 1];
            for (int a = -1; a <= 2; ++a) {
                float w = wx[a + 1] * wyz;
                if (w == 0) continue;
                r += v[((size_t)(kz + c) * ny + (ky + b)) * nx + kx + a] * w;
            }
        }
    return r;
}
float interp2(const float *v, int nx, int ny, float x, float y) {
    if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1)) return 0;
    float wx[4], wy[4];
    int kx, ky;
    cubicWeights(x, nx, wx, kx);
    cubicWeights(y, ny, wy, ky);
    float r = 0;
    for (int b = -1; b <= 2; ++b)
        for (int a = -1; a <= 2; ++a) {
            float w = wx[a + 1] * wy[b + 1];
            if (w == 0) continue;
            r += v[(size_t)(ky + b) * nx + kx + a] * w;
        }
    return r;
}
float interp1(const float *v, int n, float x) {
    if (!(x >= 0 && x <= 1)) return 0;
    float t = x * (n - 1);
    int k = std::max(0, std::min((int)t, n - 2));
    float f0 = v[k], f1 = v[k + 1];
    float d0 = k > 0 ? 0.5f * (v[k + 1] - v[k - 1]) : v[k + 1] - v[k];
    float d1 = k + 2 < n ? 0.5f * (v[k + 2] - v[k]) : v[k + 1] - v[k];
    t -= (float)k;
    float t2 = t * t, t3 = t2 * t;
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
}

struct RTable {
    int etaN = 0, alphaN = 0, thetaN = 0;
    float etaMin, etaMax, alphaMin, alphaMax;
    std::vector<float> trans, diff;
    bool loaded = false;
};
RTable g_rt[2];

bool loadRTable(int distr, std::string &err) {
    RTable &t = g_rt[distr];
    if (t.loaded) return true;
    std::string path = dataDir() + (distr == B200PG_DISTR_GGX ? "/rtrans_ggx.bin" : "/rtrans_beckmann.bin");
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open rough transmittance table " + path + " (set B200PG_DATA_DIR)";
        return false;
    }
    char magic[8];
    int32_t dims[3];
    float rng[4];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "B200RTR1", 8) == 0 &&
              std::fread(dims, 4, 3, f) == 3 && std::fread(rng, 4, 4, f) == 4;
    if (ok) {
        t.etaN = dims[0]; t.alphaN = dims[1]; t.thetaN = dims[2];
        t.etaMin = rng[0]; t.etaMax = rng[1]; t.alphaMin = rng[2]; t.alphaMax = rng[3];
        t.trans.resize((size_t)2 * t.etaN * t.alphaN * t.thetaN);
        t.diff.resize((size_t)2 * t.etaN * t.alphaN);
        ok = std::fread(t.trans.data(), 4, t.trans.size(), f) == t.trans.size() &&
             std::fread(t.diff.data(), 4, t.diff.size(), f) == t.diff.size();
    }
    std::fclose(f);
    if (!ok) {
        err = "malformed rough transmittance table " + path;
        return false;
    }
    t.loaded = true;
    return true;
}
}  // namespace

bool rtransReduce(int distribution, float eta, float alpha, float *extTrans100, float *extDiff, float *intDiff,
                  std::string &err) {
    if (distribution != B200PG_DISTR_BECKMANN && distribution != B200PG_DISTR_GGX) {
        err = "unsupported microfacet distribution";
        return false;
    }
    if (!loadRTable(distribution, err)) return false;
    const RTable &t = g_rt[distribution];
    if (t.thetaN != 100) {
        err = "rough transmittance table must have 100 theta samples";
        return false;
    }
    float chk = eta < 1 ? 1 / eta : eta;  // RoughTransmittance::checkEta / checkAlpha
    if (chk < t.etaMin || chk > t.etaMax) {
        err = "roughplastic: relative IOR outside the precomputed range";
        return false;
    }
    if (alpha < t.alphaMin || alpha > t.alphaMax) {
        err = "roughplastic: roughness outside the precomputed range";
        return false;
    }
    // Fix eta (3-D -> 2-D) for the external (eta) and internal (1/eta) interfaces.
    auto fixEta = [&](float e, std::vector<float> &tr2, std::vector<float> &df1) {
        const float *tr = t.trans.data(), *df = t.diff.data();
        if (e < 1) {
            tr += (size_t)t.etaN * t.alphaN * t.thetaN;
            df += (size_t)t.etaN * t.alphaN;
            e = 1.0f / e;
        }
        if (e < t.etaMin) e = t.etaMin;
        float wEta = std::pow((e - t.etaMin) / (t.etaMax - t.etaMin), 0.25f);
        tr2.resize((size_t)t.alphaN * t.thetaN);
        df1.resize(t.alphaN);
        float dA = 1.0f / (t.alphaN - 1), dT = 1.0f / (t.thetaN - 1);
        for (int i = 0; i < t.alphaN; ++i) {
            for (int j = 0; j < t.thetaN; ++j)
                tr2[(size_t)i * t.thetaN + j] = interp3(tr, t.thetaN, t.alphaN, t.etaN, j * dT, i * dA, wEta);
            df1[i] = interp2(df, t.alphaN, t.etaN, i * dA, wEta);
        }
    };
    std::vector<float> eT, eD, iT, iD;
    fixEta(eta, eT, eD);
    fixEta(1.0f / eta, iT, iD);
    float wAlpha = std::pow((alpha - t.alphaMin) / (t.alphaMax - t.alphaMin), 0.25f);
    float dT = 1.0f / (t.thetaN - 1);
    for (int i = 0; i < t.thetaN; ++i) extTrans100[i] = interp2(eT.data(), t.thetaN, t.alphaN, i * dT, wAlpha);
    *extDiff = interp1(eD.data(), t.alphaN, wAlpha);
    *intDiff = std::min(1.0f, std::max(0.0f, interp1(iD.data(), t.alphaN, wAlpha)));
    return true;
}

// ------------------------------------------------------------------ deep copy
bool HostScene::copyFrom(const B200pgSceneDesc *d, std::string &err) {
    if (!d || d->n_shapes <= 0 || !d->shapes) {
        err = "scene has no shapes";
        return false;
    }
    shapes.assign(d->shapes, d->shapes + d->n_shapes);
    bsdfs.assign(d->bsdfs, d->bsdfs + std::max(0, d->n_bsdfs));
    emitters.assign(d->emitters, d->emitters + std::max(0, d->n_emitters));
    media.assign(d->media, d->media + std::max(0, d->n_media));
    sensor = d->sensor;
    film = d->film;
    sampleCount = d->sample_count;
    seed = d->seed;
    for (auto &s : shapes) {
        if (s.type != B200PG_SHAPE_TRIMESH) {
            s.positions = s.normals = s.texcoords = nullptr;
            s.indices = nullptr;
            continue;
        }
        if (!s.positions || !s.indices || s.n_vertices == 0 || s.n_triangles == 0) {
            err = "trimesh without vertices/indices";
            return false;
        }
        ownedF.emplace_back(s.positions, s.positions + 3 * (size_t)s.n_vertices);
        s.positions = ownedF.back().data();
        if (s.normals) {
            ownedF.emplace_back(s.normals, s.normals + 3 * (size_t)s.n_vertices);
            s.normals = ownedF.back().data();
        }
        if (s.texcoords) {
            ownedF.emplace_back(s.texcoords, s.texcoords + 2 * (size_t)s.n_vertices);
            s.texcoords = ownedF.back().data();
        }
        ownedU.emplace_back(s.indices, s.indices + 3 * (size_t)s.n_triangles);
        s.indices = ownedU.back().data();
        for (size_t i = 0; i < 3 * (size_t)s.n_triangles; ++i)
            if (s.indices[i] >= s.n_vertices) {
                err = "trimesh index out of range";
                return false;
            }
    }
    for (auto &m : media) {
        size_t n = (size_t)m.res[0] * m.res[1] * m.res[2];
        if (!m.density || n == 0) {
            err = "medium without density grid";
            return false;
        }
        ownedF.emplace_back(m.density, m.density + n);
        m.density = ownedF.back().data();
    }
    b200pg_integrator_params_default(&xmlParams);
    refreshView();
    return true;
}

void HostScene::refreshView() {
    view.n_shapes = (int)shapes.size();
    view.n_bsdfs = (int)bsdfs.size();
    view.n_emitters = (int)emitters.size();
    view.n_media = (int)media.size();
    view.shapes = shapes.data();
    view.bsdfs = bsdfs.data();
    view.emitters = emitters.data();
    view.media = media.data();
    view.sensor = sensor;
    view.film = film;
    view.sample_count = sampleCount;
    view.seed = seed;
}

// ------------------------------------------------------------------ compile
bool HostScene::compile(std::string &err) {
    // ---- default BSDFs (Shape::configure, shape.cpp:48-70)
    int defHalf = -1, defBlack = -1, defNull = -1;
    auto addDefault = [&](int type, float refl) {
        B200pgBsdf b;
        std::memset(&b, 0, sizeof(b));
        b.type = type;
        for (int c = 0; c < 3; ++c) {
            b.reflectance[c] = refl;
            b.specular_reflectance[c] = b.specular_transmittance[c] = 1.0f;
        }
        b.int_ior = 1.5046f;
        b.ext_ior = 1.000277f;
        b.alpha_u = b.alpha_v = 0.1f;
        b.sample_visible = 1;
        bsdfs.push_back(b);
        return (int)bsdfs.size() - 1;
    };
    for (auto &s : shapes) {
        if (s.bsdf >= (int)bsdfs.size() || s.emitter >= (int)emitters.size() ||
            s.interior_medium >= (int)media.size() || s.exterior_medium >= (int)media.size()) {
            err = "shape references a missing bsdf/emitter/medium";
            return false;
        }
        if (s.bsdf < 0) {
            bool transition = s.interior_medium >= 0 || s.exterior_medium >= 0;
            if (s.emitter >= 0) {
                if (defBlack < 0) defBlack = addDefault(B200PG_BSDF_DIFFUSE, 0.0f);
                s.bsdf = defBlack;
            } else if (!transition) {
                if (defHalf < 0) defHalf = addDefault(B200PG_BSDF_DIFFUSE, 0.5f);
                s.bsdf = defHalf;
            } else {
                if (defNull < 0) defNull = addDefault(B200PG_BSDF_NULL, 0.0f);
                s.bsdf = defNull;
            }
        }
        // shape.cpp:72-76: index-matched BSDF on an emitter is an error
        if (bsdfs[s.bsdf].type == B200PG_BSDF_NULL && s.emitter >= 0) {
            err = "shape has an index-matched BSDF and an emitter at the same time";
            return false;
        }
    }

    // ---- BSDF records
    bsdfRecs.clear();
    for (auto &b : bsdfs) {
        if (b.type < 0 || b.type > B200PG_BSDF_NULL) {
            err = "unknown BSDF type";
            return false;
        }
        bool micro = b.type == B200PG_BSDF_ROUGHCONDUCTOR || b.type == B200PG_BSDF_ROUGHPLASTIC;
        if (micro && !b.sample_visible) {
            err = "microfacet BSDFs support sampleVisible=true only";
            return false;
        }
        if (b.type == B200PG_BSDF_ROUGHPLASTIC && b.alpha_u != b.alpha_v) {
            err = "The 'roughplastic' plugin currently does not support anisotropic microfacet distributions!";
            return false;
        }
        if ((b.type == B200PG_BSDF_DIELECTRIC || b.type == B200PG_BSDF_ROUGHPLASTIC) &&
            (b.int_ior <= 0 || b.ext_ior <= 0)) {
            err = "The interior and exterior indices of refraction must be positive!";
            return false;
        }
        if (b.type == B200PG_BSDF_ROUGHPLASTIC && b.int_ior == b.ext_ior) {
            err = "The interior and exterior indices of refraction must be positive and differ!";
            return false;
        }
        BsdfRecord r;
        std::memset(&r, 0, sizeof(r));
        r.type = b.type;
        r.twosided = b.twosided;
        for (int c = 0; c < 3; ++c) {
            r.reflectance[c] = b.reflectance[c];
            r.specRefl[c] = b.specular_reflectance[c];
            r.specTrans[c] = b.specular_transmittance[c];
            r.condEta[c] = b.eta[c];
            r.condK[c] = b.k[c];
        }
        r.eta = (b.type == B200PG_BSDF_DIELECTRIC || b.type == B200PG_BSDF_ROUGHPLASTIC) ? b.int_ior / b.ext_ior : 1.0f;
        r.invEta = 1 / r.eta;
        r.invEta2 = 1.0f / (r.eta * r.eta);
        r.distribution = b.distribution;
        r.alphaU = b.alpha_u;
        r.alphaV = b.alpha_v;
        r.nonlinear = b.nonlinear;
        float dAvg = b.reflectance[0] * 0.212671f + b.reflectance[1] * 0.715160f + b.reflectance[2] * 0.072169f;
        float sAvg = b.specular_reflectance[0] * 0.212671f + b.specular_reflectance[1] * 0.715160f +
                     b.specular_reflectance[2] * 0.072169f;
        r.specSamplingWeight = sAvg / (dAvg + sAvg);
        if (b.type == B200PG_BSDF_ROUGHPLASTIC) {
            if (b.rt_ext_diff == 0 && b.rt_ext_trans[99] == 0) {  // not supplied: reduce the tables here
                if (!rtransReduce(b.distribution, r.eta, b.alpha_u, b.rt_ext_trans, &b.rt_ext_diff, &b.rt_int_diff, err))
                    return false;
            }
            std::memcpy(r.rtExt, b.rt_ext_trans, sizeof(r.rtExt));
            r.rtIntDiff = b.rt_int_diff;
        }
        // bsdf.h:220-262 lobe flags
        const uint32_t ENull = 0x1, EDiffR = 0x2, EGlossyR = 0x8, EDeltaR = 0x20, EDeltaT = 0x40, EFront = 0x8000,
                       EBack = 0x10000;
        switch (b.type) {
            case B200PG_BSDF_DIFFUSE: r.typeFlags = EDiffR | EFront; break;
            case B200PG_BSDF_DIELECTRIC: r.typeFlags = EDeltaR | EDeltaT | EFront | EBack; break;
            case B200PG_BSDF_ROUGHCONDUCTOR: r.typeFlags = EGlossyR | EFront; break;
            case B200PG_BSDF_ROUGHPLASTIC: r.typeFlags = EGlossyR | EDiffR | EFront; break;
            case B200PG_BSDF_NULL: r.typeFlags = ENull | EFront | EBack; break;
        }
        if (b.twosided) {
            if (b.type == B200PG_BSDF_DIELECTRIC || b.type == B200PG_BSDF_NULL) {
                err = "twosided: only BRDFs (BSDFs without a transmission component) can be nested";
                return false;
            }
            r.typeFlags |= EBack;
        }
        bsdfRecs.push_back(r);
    }

    // ---- shapes, primitives (global ids numbered as in skdtree.cpp:53-104)
    std::vector<BPrim> bprims;
    std::vector<PrimRecord> flat;  // global-id order
    std::vector<PrimInfo> flatInfo;
    shapeRecs.clear();
    rects.clear();
    meshes.clear();
    positions.clear();
    normals.clear();
    texcoords.clear();
    indices.clear();
    areaCdf.clear();
    primCount = 0;
    for (size_t si = 0; si < shapes.size(); ++si) {
        const B200pgShape &s = shapes[si];
        ShapeRecord sr;
        std::memset(&sr, 0, sizeof(sr));
        sr.type = s.type;
        sr.bsdf = s.bsdf;
        sr.emitter = s.emitter;
        sr.interiorMedium = s.interior_medium;
        sr.exteriorMedium = s.exterior_medium;
        sr.primOffset = primCount;
        if (s.type == B200PG_SHAPE_RECTANGLE) {
            float inv[16];
            if (!invert4(s.to_world, inv)) {
                err = "rectangle: singular toWorld";
                return false;
            }
            V3 dpdu = xfVector(s.to_world, v3(2, 0, 0)), dpdv = xfVector(s.to_world, v3(0, 2, 0));
            V3 n = normalize(xfNormalFromInverse(inv, v3(0, 0, 1)));
            if (std::fabs(dot(normalize(dpdu), normalize(dpdv))) > 1e-4f) {
                err = "Error: 'toWorld' transformation contains shear!";  // rectangle.cpp:105-106
                return false;
            }
            RectRecord rr;
            std::memset(&rr, 0, sizeof(rr));
            for (int i = 0; i < 12; ++i) rr.q[i] = inv[i];
            rr.q[12] = n.x; rr.q[13] = n.y; rr.q[14] = n.z;
            rr.q[15] = 1.0f / (len(dpdu) * len(dpdv));
            rr.q[16] = dpdu.x; rr.q[17] = dpdu.y; rr.q[18] = dpdu.z;
            for (int i = 0; i < 12; ++i) rr.q[20 + i] = s.to_world[i];
            sr.meshOffset = (uint32_t)rects.size();
            rects.push_back(rr);
            PrimRecord pr;
            for (int i = 0; i < 12; ++i) pr.q[i] = inv[i];  // rows of worldToObject
            flat.push_back(pr);
            flatInfo.push_back(PrimInfo{(uint32_t)si, kNoTriangle});
            BPrim bp;
            Box bx;
            bx.reset();
            const float cs[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
            for (auto &c : cs) {
                V3 p = xfPoint(s.to_world, v3(c[0], c[1], 0));
                bx.grow(&p.x, &p.x);
            }
            for (int a = 0; a < 3; ++a) {
                bp.bmin[a] = bx.mn[a];
                bp.bmax[a] = bx.mx[a];
                bp.c[a] = 0.5f * (bx.mn[a] + bx.mx[a]);
            }
            bp.id = primCount;
            bprims.push_back(bp);
            primCount += 1;
        } else if (s.type == B200PG_SHAPE_TRIMESH) {
            MeshRecord mr;
            std::memset(&mr, 0, sizeof(mr));
            mr.vertexOffset = (uint32_t)(positions.size() / 3);
            mr.indexOffset = (uint32_t)(indices.size() / 3);
            mr.cdfOffset = (uint32_t)areaCdf.size();
            mr.nTriangles = s.n_triangles;
            mr.hasNormals = s.normals ? 1 : 0;
            mr.hasTexcoords = s.texcoords ? 1 : 0;
            positions.insert(positions.end(), s.positions, s.positions + 3 * (size_t)s.n_vertices);
            if (s.normals)
                normals.insert(normals.end(), s.normals, s.normals + 3 * (size_t)s.n_vertices);
            else
                normals.resize(normals.size() + 3 * (size_t)s.n_vertices, 0.0f);
            if (s.texcoords)
                texcoords.insert(texcoords.end(), s.texcoords, s.texcoords + 2 * (size_t)s.n_vertices);
            else
                texcoords.resize(texcoords.size() + 2 * (size_t)s.n_vertices, 0.0f);
            indices.insert(indices.end(), s.indices, s.indices + 3 * (size_t)s.n_triangles);
            sr.meshOffset = (uint32_t)meshes.size();
            sr.flags = mr.hasNormals;
            size_t cdf0 = areaCdf.size();
            areaCdf.push_back(0.0f);
            for (uint32_t t = 0; t < s.n_triangles; ++t) {
                const float *pa = s.positions + 3 * (size_t)s.indices[3 * t], *pb = s.positions + 3 * (size_t)s.indices[3 * t + 1],
                            *pc = s.positions + 3 * (size_t)s.indices[3 * t + 2];
                V3 A = v3(pa[0], pa[1], pa[2]), B = v3(pb[0], pb[1], pb[2]), C = v3(pc[0], pc[1], pc[2]);
                // Plane form of the triangle test: with e1 = B-A, e2 = C-A, N = e1 x e2,
                //   t = (N.A - N.o) / (N.d),  u = Nu.(P-A),  v = Nv.(P-A),  Nu = (e2 x N)/(e1.(e2 x N)),
                //   Nv = (N x e1)/(e2.(N x e1)) -- the same (t, u, v) TriAccel::rayIntersect returns
                // (triaccel.h:96-158; u weights vertex 1, v vertex 2). Constants are derived in double.
                PrimRecord pr;
                std::memset(&pr, 0, sizeof(pr));
                {
                    double a[3] = {A.x, A.y, A.z}, e1[3] = {(double)B.x - A.x, (double)B.y - A.y, (double)B.z - A.z},
                           e2[3] = {(double)C.x - A.x, (double)C.y - A.y, (double)C.z - A.z};
                    auto crs = [](const double *p, const double *q, double *r) {
                        r[0] = p[1] * q[2] - p[2] * q[1];
                        r[1] = p[2] * q[0] - p[0] * q[2];
                        r[2] = p[0] * q[1] - p[1] * q[0];
                    };
                    auto dt = [](const double *p, const double *q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
                    double N[3], nu[3], nv[3];
                    crs(e1, e2, N);
                    crs(e2, N, nu);
                    crs(N, e1, nv);
                    double du = dt(e1, nu), dv = dt(e2, nv), nn = std::sqrt(dt(N, N));
                    if (nn == 0 || du == 0 || dv == 0) {
                        pr.q[11] = 1.0f;  // degenerate (TriAccel::load failure, triaccel.h:80-83): t = -1/0, never hit
                    } else {
                        for (int j = 0; j < 3; ++j) {
                            pr.q[j] = (float)(nu[j] / du);
                            pr.q[4 + j] = (float)(nv[j] / dv);
                            pr.q[8 + j] = (float)(N[j] / nn);
                        }
                        pr.q[3] = (float)(-dt(nu, a) / du);
                        pr.q[7] = (float)(-dt(nv, a) / dv);
                        pr.q[11] = (float)(-dt(N, a) / nn);
                    }
                }
                flat.push_back(pr);
                flatInfo.push_back(PrimInfo{(uint32_t)si, t});
                BPrim bp;
                for (int a = 0; a < 3; ++a) {
                    bp.bmin[a] = std::min(A[a], std::min(B[a], C[a]));
                    bp.bmax[a] = std::max(A[a], std::max(B[a], C[a]));
                    bp.c[a] = 0.5f * (bp.bmin[a] + bp.bmax[a]);
                }
                bp.id = primCount + t;
                bprims.push_back(bp);
                float area = 0.5f * len(cross(sub(B, A), sub(C, A)));
                areaCdf.push_back(areaCdf.back() + area);
            }
            // DiscreteDistribution::normalize (pmf.h:98-112)
            float sum = areaCdf.back();
            if (sum > 0) {
                float norm = 1.0f / sum;
                for (size_t i = cdf0 + 1; i < areaCdf.size(); ++i) areaCdf[i] *= norm;
                areaCdf.back() = 1.0f;
            }
            mr.invSurfaceArea = 1.0f / sum;
            meshes.push_back(mr);
            primCount += s.n_triangles;
        } else {
            err = "unknown shape type";
            return false;
        }
        shapeRecs.push_back(sr);
    }
    if (positions.empty()) {  // keep device pools non-empty
        positions.resize(3, 0);
        normals.resize(3, 0);
        texcoords.resize(2, 0);
        indices.resize(3, 0);
        areaCdf.resize(2, 0);
        meshes.push_back(MeshRecord());
    }
    if (rects.empty()) rects.push_back(RectRecord());

    // ---- emitters (uniform unless samplingWeight set, scene.cpp:419-423)
    emitterRecs.clear();
    emitterCdf.assign(1, 0.0f);
    for (auto &e : emitters) {
        if (e.shape < 0 || e.shape >= (int)shapes.size()) {
            err = "emitter without a shape";
            return false;
        }
        EmitterRecord er;
        for (int c = 0; c < 3; ++c) er.radiance[c] = e.radiance[c];
        er.shape = e.shape;
        emitterRecs.push_back(er);
        emitterCdf.push_back(emitterCdf.back() + e.sampling_weight);
    }
    if (!emitters.empty() && emitterCdf.back() > 0) {
        float norm = 1.0f / emitterCdf.back();
        for (size_t i = 1; i < emitterCdf.size(); ++i) emitterCdf[i] *= norm;
        emitterCdf.back() = 1.0f;
    }
    if (emitterRecs.empty()) {
        emitterRecs.push_back(EmitterRecord{{0, 0, 0}, 0});
        emitterCdf.push_back(1.0f);
    }

    // ---- BVH
    {
        Builder bld(bprims);
        Box rootBox;
        int32_t root;
#pragma omp parallel
#pragma omp single
        root = bld.build(0, bprims.size(), rootBox, 0);
        int nInner = bld.next.load();
        for (int a = 0; a < 3; ++a) {
            sceneMin[a] = rootBox.mn[a];
            sceneMax[a] = rootBox.mx[a];
        }
        if (root < 0) {  // whole scene fits one leaf: synthesise a root with one real child
            TmpNode &nd = bld.nodes[0];
            nd.child[0] = root;
            nd.box[0] = rootBox;
            nd.child[1] = Builder::leafCode(0, 0);
            nd.box[1].reset();
            nInner = 1;
            root = 0;
        }
        // Primitive slots: every leaf starts at an EVEN slot, so that the plane rows of two neighbouring primitives of a leaf are
        // one aligned 256-bit load (device_scene.cuh: bvhLeafStep); odd-sized leaves leave a hole behind them.
        std::vector<uint32_t> slotOf(bprims.size() + 1, 0);
        uint32_t nSlots = 0;
        {
            std::vector<std::pair<uint32_t, uint32_t>> leaves;  // (first, count) in the builder's primitive order
            for (int i = 0; i < nInner; ++i)
                for (int c = 0; c < 2; ++c) {
                    const int32_t ch = bld.nodes[i].child[c];
                    if (ch >= 0) continue;
                    const uint32_t code = (uint32_t)(~ch);
                    if ((code & 15u) != 0) leaves.emplace_back(code >> 4, code & 15u);
                }
            std::sort(leaves.begin(), leaves.end());
            uint32_t cursor = 0;
            for (auto &lf : leaves) {
                cursor = (cursor + 1u) & ~1u;
                for (uint32_t k = 0; k < lf.second; ++k) slotOf[lf.first + k] = cursor + k;
                cursor += lf.second;
            }
            nSlots = ((cursor + 1u) & ~1u) + 2u;  // a leaf's second plane pair may be read one slot past its last primitive
        }
        // relayout in DFS order so that a node's first child follows it
        std::vector<int32_t> remap(nInner, -1), order;
        order.reserve(nInner);
        std::vector<int32_t> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            int32_t n = stack.back();
            stack.pop_back();
            remap[n] = (int32_t)order.size();
            order.push_back(n);
            if (bld.nodes[n].child[1] >= 0) stack.push_back(bld.nodes[n].child[1]);
            if (bld.nodes[n].child[0] >= 0) stack.push_back(bld.nodes[n].child[0]);
        }
        nodes.resize(order.size());
        for (size_t i = 0; i < order.size(); ++i) {
            const TmpNode &t = bld.nodes[order[i]];
            BvhNode &o = nodes[i];
            const Box &b0 = t.box[0], &b1 = t.box[1];
            o.q[0] = b0.mn[0]; o.q[1] = b0.mx[0]; o.q[2] = b0.mn[1]; o.q[3] = b0.mx[1];
            o.q[4] = b1.mn[0]; o.q[5] = b1.mx[0]; o.q[6] = b1.mn[1]; o.q[7] = b1.mx[1];
            o.q[8] = b0.mn[2]; o.q[9] = b0.mx[2]; o.q[10] = b1.mn[2]; o.q[11] = b1.mx[2];
            // leaves: re-encode as (first << 7) | (rectMask << 3) | count
            auto leaf = [&](int32_t c) {
                uint32_t code = (uint32_t)(~c);
                uint32_t first = code >> 4, count = code & 15u, mask = 0;
                for (uint32_t k = 0; k < count; ++k)
                    if (flatInfo[bprims[first + k].id].prim == kNoTriangle) mask |= 1u << k;
                return ~(int32_t)((slotOf[first] << kLeafShift) | (mask << 3) | count);
            };
            int32_t c0 = t.child[0] >= 0 ? remap[t.child[0]] : leaf(t.child[0]);
            int32_t c1 = t.child[1] >= 0 ? remap[t.child[1]] : leaf(t.child[1]);
            o.q[12] = u2f((uint32_t)c0);
            o.q[13] = u2f((uint32_t)c1);
            o.q[14] = o.q[15] = 0;
        }
        // ---- wide BVH (pg_types.h: WideNode): the binary tree collapsed to up to 8 children per node with the children's
        // boxes quantised to 8 bits on a per-node power-of-two grid. Large meshes only: three binary levels become one node
        // visit (one dependent round trip instead of three) and the node array shrinks ~4.5x (10 M triangles: 256 MB ->
        // 56 MB, inside the 126 MB L2). Small scenes stay on the binary tree, which lives in L1 there.
        wideNodes.clear();
        wideDepth = 0;
        {
            const char *env = std::getenv("B200PG_WIDE_MIN_PRIMS");
            // built only on request (B200PG_WIDE_MIN_PRIMS=n): measured slower than the binary tree on C4, see integrator.cu
            const size_t minPrims = env ? (size_t)std::atoll(env) : (size_t)-1;
            if (bprims.size() >= minPrims && media.empty()) {
                auto leafRef = [&](int32_t c) {  // same leaf code as the binary nodes
                    uint32_t code = (uint32_t)(~c);
                    uint32_t first = code >> 4, count = code & 15u, mask = 0;
                    for (uint32_t k = 0; k < count; ++k)
                        if (flatInfo[bprims[first + k].id].prim == kNoTriangle) mask |= 1u << k;
                    return ~(int32_t)((slotOf[first] << kLeafShift) | (mask << 3) | count);
                };
                const float diag = std::sqrt((rootBox.mx[0] - rootBox.mn[0]) * (rootBox.mx[0] - rootBox.mn[0]) +
                                             (rootBox.mx[1] - rootBox.mn[1]) * (rootBox.mx[1] - rootBox.mn[1]) +
                                             (rootBox.mx[2] - rootBox.mn[2]) * (rootBox.mx[2] - rootBox.mn[2]));
                // leaves below every binary node (children have larger indices than their parent is NOT guaranteed: recurse)
                std::vector<uint32_t> leafCount(bld.nodes.size(), 0);
                {
                    std::vector<int32_t> order2, st2;
                    st2.push_back(root);
                    while (!st2.empty()) {
                        const int32_t b = st2.back();
                        st2.pop_back();
                        order2.push_back(b);
                        for (int c = 0; c < 2; ++c)
                            if (bld.nodes[b].child[c] >= 0) st2.push_back(bld.nodes[b].child[c]);
                    }
                    for (size_t i = order2.size(); i-- > 0;) {  // children before parents
                        const TmpNode &t = bld.nodes[order2[i]];
                        uint32_t cnt = 0;
                        for (int c = 0; c < 2; ++c) cnt += t.child[c] >= 0 ? leafCount[t.child[c]] : (((uint32_t)(~t.child[c]) & 15u) ? 1u : 0u);
                        leafCount[order2[i]] = cnt;
                    }
                }
                struct Item { int32_t bin; int depth; };  // binary node to collapse into wide node #index-in-queue
                std::vector<Item> queue;
                queue.push_back(Item{root, 1});
                std::vector<WideNode> out;
                for (size_t qi = 0; qi < queue.size(); ++qi) {
                    const Item it = queue[qi];
                    wideDepth = std::max(wideDepth, it.depth);
                    // collapse: open children until 8 remain. Opening greedily by surface area alone leaves the bottom of the
                    // tree full of 2-3-child nodes (a balanced 16-leaf subtree becomes 8 two-child nodes: 3.4 children per
                    // node on the heightfield); so a subtree of <= 8 leaves is either inlined COMPLETELY, when all of its
                    // leaves fit, or kept whole as a child node that will then be (nearly) full.
                    int32_t ref[8];
                    Box box[8];
                    int n = 2;
                    ref[0] = bld.nodes[it.bin].child[0]; box[0] = bld.nodes[it.bin].box[0];
                    ref[1] = bld.nodes[it.bin].child[1]; box[1] = bld.nodes[it.bin].box[1];
                    while (n < 8) {
                        int best = -1;
                        float bestArea = -1;
                        for (int c = 0; c < n; ++c)  // 1. the largest subtree that is too big for one node
                            if (ref[c] >= 0 && leafCount[ref[c]] > 8 && box[c].area() > bestArea) { bestArea = box[c].area(); best = c; }
                        if (best < 0)
                            for (int c = 0; c < n; ++c)  // 2. a small subtree all of whose leaves fit into the free slots
                                if (ref[c] >= 0 && n - 1 + (int)leafCount[ref[c]] <= 8 && box[c].area() > bestArea) { bestArea = box[c].area(); best = c; }
                        if (best < 0) break;
                        const TmpNode &t = bld.nodes[ref[best]];
                        ref[best] = t.child[0]; box[best] = t.box[0];
                        ref[n] = t.child[1]; box[n] = t.box[1];
                        ++n;
                    }
                    // drop the empty child a synthesised root may carry
                    int m = 0;
                    for (int c = 0; c < n; ++c)
                        if (!(ref[c] < 0 && ((uint32_t)(~ref[c]) & 15u) == 0)) { ref[m] = ref[c]; box[m] = box[c]; ++m; }
                    n = m;
                    Box all;
                    all.reset();
                    for (int c = 0; c < n; ++c) all.grow(box[c].mn, box[c].mx);
                    WideNode W;
                    std::memset(&W, 0, sizeof(W));
                    uint32_t expo[3];
                    float origin[3], scale[3];
                    for (int a = 0; a < 3; ++a) {
                        // grid: origin = lo - delta, planes at origin + q * 2^e, q in [0, 255]. The device evaluates a plane's
                        // ray distance as fma(2^23 + q, 2^e * idir, (origin - o) * idir - 2^23 * 2^e * idir), which is off by at
                        // most half a grid step plus ~1e-7 |origin - o| in space: every child box is therefore widened by one
                        // grid step beyond floor / ceil, and delta adds 1e-6 of the scene diagonal.
                        const float ext = std::max(all.mx[a] - all.mn[a], 0.0f);
                        int e = (int)std::ceil(std::log2(std::max((ext * 1.02f + 4e-6f * diag) / 250.0f, 1e-30f)));
                        float sc, delta;
                        for (;; ++e) {
                            sc = std::ldexp(1.0f, e);
                            delta = sc + 1e-6f * diag;
                            if ((ext + 2 * delta) / sc <= 252.0f) break;
                        }
                        scale[a] = sc;
                        origin[a] = std::nextafter(all.mn[a] - delta, -std::numeric_limits<float>::infinity());
                        expo[a] = (uint32_t)(e + 127);
                    }
                    W.q[0] = origin[0]; W.q[1] = origin[1]; W.q[2] = origin[2];
                    W.q[3] = u2f(expo[0] | (expo[1] << 8) | (expo[2] << 16) | ((uint32_t)n << 24));
                    uint8_t *quant = reinterpret_cast<uint8_t *>(&W.q[12]);  // lox[8] loy[8] loz[8] hix[8] hiy[8] hiz[8]
                    for (int c = 0; c < 8; ++c) {
                        int32_t r = (int32_t)0x7FFFFFFF;  // empty slot: never a hit (inverted box), never pushed
                        if (c < n) {
                            if (ref[c] >= 0) {
                                r = (int32_t)queue.size();
                                queue.push_back(Item{ref[c], it.depth + 1});
                            } else {
                                r = leafRef(ref[c]);
                            }
                            for (int a = 0; a < 3; ++a) {
                                const double lo = ((double)box[c].mn[a] - origin[a]) / scale[a], hi = ((double)box[c].mx[a] - origin[a]) / scale[a];
                                const int qlo = std::max(0, (int)std::floor(lo) - 1), qhi = std::min(255, (int)std::ceil(hi) + 1);
                                quant[8 * a + c] = (uint8_t)qlo;
                                quant[24 + 8 * a + c] = (uint8_t)qhi;
                            }
                        } else {
                            for (int a = 0; a < 3; ++a) {
                                quant[8 * a + c] = 255;
                                quant[24 + 8 * a + c] = 0;
                            }
                        }
                        W.q[4 + c] = u2f((uint32_t)r);
                    }
                    out.push_back(W);
                }
                if (wideDepth <= kWideMaxDepth) {
                    wideNodes.swap(out);
                } else {
                    wideDepth = 0;  // degenerate tree: the traversal stack could not hold it -- stay on the binary tree
                }
            }
        }
        if (nSlots >= (1u << 24)) {
            err = "more than 16M primitive slots are not supported by the leaf encoding";
            return false;
        }
        primPlanes.assign((size_t)nSlots * 4, 0.0f);
        primRows.assign((size_t)nSlots * 8, 0.0f);
        primGlobalId.assign(nSlots, 0xFFFFFFFFu);
        primInfo.assign(nSlots, PrimInfo{0u, kNoTriangle});
        shadeTris.assign((size_t)nSlots * 24, 0.0f);
        nPrimitives = bprims.size();
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < bprims.size(); ++b) {
            const size_t i = slotOf[b];
            const PrimRecord &pr = flat[bprims[b].id];
            for (int k = 0; k < 8; ++k) primRows[i * 8 + k] = pr.q[k];        // rows 0 / 1: (u, v)
            for (int k = 0; k < 4; ++k) primPlanes[i * 4 + k] = pr.q[8 + k];  // row 2: the plane
            primInfo[i] = flatInfo[bprims[b].id];
            primGlobalId[i] = bprims[b].id;
            // shading record (pg_types.h: ShadeTri): the hit's vertex data in one place, in BVH order
            const PrimInfo pi = primInfo[i];
            const ShapeRecord &sr = shapeRecs[pi.shape];
            float *q = &shadeTris[i * 24];
            uint32_t flags = 0;
            if (pi.prim != kNoTriangle) {
                const MeshRecord &mr = meshes[sr.meshOffset];
                const uint32_t *idx = &indices[3 * ((size_t)mr.indexOffset + pi.prim)];
                flags = 1u | (mr.hasNormals ? 2u : 0u);
                for (int v = 0; v < 3; ++v) {
                    const size_t vi = (size_t)idx[v] + mr.vertexOffset;
                    for (int a = 0; a < 3; ++a) {
                        q[4 * v + a] = positions[3 * vi + a];
                        q[12 + 4 * v + a] = mr.hasNormals ? normals[3 * vi + a] : 0.0f;
                    }
                }
                q[19] = 1.0f;  // dpdu = first edge (skdtree.h:378-380) ...
                q[23] = 0.0f;
                if (mr.hasTexcoords) {  // ... or the UV tangent, TriMesh::computeUVTangents (trimesh.cpp:701-735)
                    const float *uv0 = &texcoords[2 * ((size_t)idx[0] + mr.vertexOffset)], *uv1 = &texcoords[2 * ((size_t)idx[1] + mr.vertexOffset)],
                                *uv2 = &texcoords[2 * ((size_t)idx[2] + mr.vertexOffset)];
                    const V3 dP1 = v3(q[4] - q[0], q[5] - q[1], q[6] - q[2]), dP2 = v3(q[8] - q[0], q[9] - q[1], q[10] - q[2]);
                    const float dU1x = uv1[0] - uv0[0], dU1y = uv1[1] - uv0[1], dU2x = uv2[0] - uv0[0], dU2y = uv2[1] - uv0[1];
                    const V3 n = cross(dP1, dP2);
                    const float nl = len(n);
                    if (nl != 0) {  // degenerate triangles are never hit
                        const float determinant = dU1x * dU2y - dU1y * dU2x;
                        if (determinant == 0) {
                            // degenerate parameterisation: coordinateSystem(n / |n|) picks the tangent (util.cpp); it lies in the
                            // triangle's plane, so it has coefficients in the (dP1, dP2) basis (normal equations, in double)
                            const V3 nn = v3(n.x / nl, n.y / nl, n.z / nl);
                            V3 tg;
                            if (std::fabs(nn.x) > std::fabs(nn.y)) {
                                const float invLen = 1.0f / std::sqrt(nn.x * nn.x + nn.z * nn.z);
                                tg = v3(nn.z * invLen, 0.0f, -nn.x * invLen);  // c, then b = cross(c, a)
                            } else {
                                const float invLen = 1.0f / std::sqrt(nn.y * nn.y + nn.z * nn.z);
                                tg = v3(0.0f, nn.z * invLen, -nn.y * invLen);
                            }
                            tg = cross(tg, nn);
                            const double g11 = dot(dP1, dP1), g12 = dot(dP1, dP2), g22 = dot(dP2, dP2), r1 = dot(tg, dP1), r2 = dot(tg, dP2);
                            const double det = g11 * g22 - g12 * g12;
                            if (det != 0) {
                                q[19] = (float)((r1 * g22 - r2 * g12) / det);
                                q[23] = (float)((r2 * g11 - r1 * g12) / det);
                            }
                        } else {
                            const float invDet = 1.0f / determinant;
                            q[19] = dU2y * invDet;
                            q[23] = -dU1y * invDet;
                        }
                        flags |= 4u;
                    }
                }
            }
            if (sr.emitter >= (1 << 24) - 1) {
#pragma omp critical
                err = "more than 16M emitters are not supported by the shading record";
            }
            flags |= (uint32_t)(sr.emitter + 1) << 8;
            q[3] = u2f(pi.shape);
            q[7] = u2f(flags);
            q[11] = u2f(pi.prim);
            q[15] = u2f((uint32_t)sr.bsdf);
        }
        if (!err.empty()) return false;
    }

    // ---- camera (perspective.cpp:126-155, no crop window)
    {
        if (film.width <= 0 || film.height <= 0) {
            err = "film size must be positive";
            return false;
        }
        float aspect = (float)film.width / (float)film.height;
        float xfov = sensor.fov;
        int axis = sensor.fov_axis;
        const float pi = 3.14159265358979323846f;
        if (axis == 3) axis = aspect > 1 ? 1 : 0;
        if (axis == 4) axis = aspect > 1 ? 0 : 1;
        if (axis == 1) {
            xfov = 2.0f * std::atan(std::tan(0.5f * sensor.fov * pi / 180.0f) * aspect) * 180.0f / pi;
        } else if (axis == 2) {
            float diagonal = 2 * std::tan(0.5f * sensor.fov * pi / 180.0f);
            float width = diagonal / std::sqrt(1.0f + 1.0f / (aspect * aspect));
            xfov = 2.0f * std::atan(width * 0.5f) * 180.0f / pi;
        }
        if (!(xfov > 0 && xfov < 180)) {
            err = "The horizontal field of view must be in the interval (0, 180)!";
            return false;
        }
        float recip = 1.0f / (sensor.far_clip - sensor.near_clip);
        float cot = 1.0f / std::tan((xfov / 2.0f) * pi / 180.0f);
        float persp[16] = {cot, 0, 0, 0, 0, cot, 0, 0, 0, 0, sensor.far_clip * recip, -sensor.near_clip * sensor.far_clip * recip, 0, 0, 1, 0};
        float tr[16] = {1, 0, 0, -1.0f, 0, 1, 0, -1.0f / aspect, 0, 0, 1, 0, 0, 0, 0, 1};
        float sc[16] = {-0.5f, 0, 0, 0, 0, -0.5f * aspect, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        float tmp[16], c2s[16];
        mul4(tr, persp, tmp);
        mul4(sc, tmp, c2s);
        if (!invert4(c2s, camera.sampleToCamera)) {
            err = "camera: singular projection";
            return false;
        }
        for (int i = 0; i < 12; ++i) camera.toWorld[i] = sensor.to_world[i];
        camera.nearClip = sensor.near_clip;
        camera.farClip = sensor.far_clip;
        camera.invResX = 1.0f / film.width;
        camera.invResY = 1.0f / film.height;
        camera.medium = sensor.medium;
    }
    // ---- film filter (gaussian.cpp:25-51, rfilter.cpp:38-56)
    {
        float stddev = film.filter_stddev > 0 ? film.filter_stddev : 0.5f;
        float radius = 4 * stddev;
        float alpha = -1.0f / (2.0f * stddev * stddev);
        float sum = 0;
        for (int i = 0; i < 31; ++i) {
            float x = (radius * i) / 31;
            float v = std::max(0.0f, std::exp(alpha * x * x) - std::exp(alpha * radius * radius));
            filmRec.values[i] = v;
            sum += v;
        }
        filmRec.values[31] = 0;
        sum *= 2 * radius / 31;
        float norm = 1.0f / sum;
        for (int i = 0; i < 31; ++i) filmRec.values[i] *= norm;
        filmRec.width = film.width;
        filmRec.height = film.height;
        filmRec.radius = radius;
        filmRec.scaleFactor = 31 / radius;
        if (radius > 2.5f) {
            err = "reconstruction filter radius above 2.5 pixels is not supported";
            return false;
        }
    }
    // ---- media (heterogeneous.cpp:228-262, gridvolume.cpp:188-198)
    mediumRecs.clear();
    densityPool.clear();
    for (auto &m : media) {
        for (int i = 0; i < 16; ++i)
            if (m.to_world[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) {  // b200pg.h: B200pgMedium::to_world, identity supported
                err = "medium: a non-identity volume toWorld is not supported";
                return false;
            }
        MediumRecord r;
        std::memset(&r, 0, sizeof(r));
        r.method = m.method;
        r.phaseType = m.phase_type;
        r.scale = m.scale;
        r.maxDensity = m.scale * 1.0f;  // getMaximumFloatValue() == 1, gridvolume.cpp:583-585
        r.invMaxDensity = 1.0f / r.maxDensity;
        r.g = m.phase_g;
        for (int c = 0; c < 3; ++c) {
            r.albedo[c] = m.albedo[c];
            r.res[c] = m.res[c];
            r.aabbMin[c] = m.aabb_min[c];
            r.aabbMax[c] = m.aabb_max[c];
        }
        float step = std::numeric_limits<float>::infinity();
        for (int c = 0; c < 3; ++c) {
            float ext = m.aabb_max[c] - m.aabb_min[c];
            step = std::min(step, 0.5f * ext / (float)(m.res[c] - 1));
            // worldToGrid = scale((res-1)/extent) * translate(-min)   (volume toWorld = identity)
            for (int k = 0; k < 4; ++k) r.worldToGrid[c * 4 + k] = 0;
            float sc = (m.res[c] - 1) / ext;
            r.worldToGrid[c * 4 + c] = sc;
            r.worldToGrid[c * 4 + 3] = sc * -m.aabb_min[c];
        }
        r.stepSize = step;
        r.densityOffset = densityPool.size();
        size_t n = (size_t)m.res[0] * m.res[1] * m.res[2];
        densityPool.insert(densityPool.end(), m.density, m.density + n);
        mediumRecs.push_back(r);
    }
    if (mediumRecs.empty()) {
        mediumRecs.push_back(MediumRecord());
        densityPool.push_back(0);
    }
    refreshView();
    return true;
}

}  // namespace pg

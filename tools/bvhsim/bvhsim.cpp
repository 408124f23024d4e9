// bvhsim.cpp -- ANALYSIS TOOL (not part of the product): replays the GPU's BVH2 while-while traversal on the host, ray by ray,
// and models how a 32-lane warp would execute a ray queue under different schedules (lock-step batches of 32, per-lane
// refill below a threshold, popped-node culling by entry distance). Used to decide kernel designs without GPU time.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../mitsuba-path-guiding_b200/csrc/host_scene.h"

namespace {
struct SceneHandle { pg::HostScene host; };
// the library keeps the primitive record split (plane row / (u, v) rows, one per slot: host_scene.h); the replay below walks the
// 48-byte form
static const pg::PrimRecord *primRecords(const pg::HostScene &H) {
    static std::vector<pg::PrimRecord> recs;
    static const pg::HostScene *built = nullptr;
    static std::mutex lock;  // called from inside OpenMP loops
    std::lock_guard<std::mutex> guard(lock);
    if (built != &H) {
        recs.resize(H.primGlobalId.size());
        for (size_t i = 0; i < recs.size(); ++i) {
            for (int k = 0; k < 8; ++k) recs[i].q[k] = H.primRows[i * 8 + k];
            for (int k = 0; k < 4; ++k) recs[i].q[8 + k] = H.primPlanes[i * 4 + k];
        }
        built = &H;
    }
    return recs.data();
}
const int kDone = (int)0x80000000;

struct RayState {
    float o[3], d[3], id[3], mint, tmax;
    int node, sp;
    int stack[64];
    float stackT[64];
    uint32_t prim; float t;
    bool any, cull;
    uint32_t nodeSteps = 0, primTests = 0, leaves = 0;
    void init(const float *r, bool anyHit, bool cullStack) {
        for (int a = 0; a < 3; ++a) { o[a] = r[a]; d[a] = r[4 + a]; id[a] = 1.0f / d[a]; }
        mint = r[3]; tmax = r[7]; node = 0; sp = 0; prim = 0xFFFFFFFFu; t = tmax; any = anyHit; cull = cullStack;
        nodeSteps = primTests = leaves = 0;
    }
    int pop() {
        while (sp) {
            --sp;
            if (!cull || stackT[sp] <= tmax) return stack[sp];
        }
        return kDone;
    }
    // one inner-node step
    void step(const pg::BvhNode *nodes) {
        const float *q = nodes[node].q;
        nodeSteps++;
        float tn[2], tf[2];
        for (int c = 0; c < 2; ++c) {
            float lox = (q[4 * c + 0] - o[0]) * id[0], hix = (q[4 * c + 1] - o[0]) * id[0];
            float loy = (q[4 * c + 2] - o[1]) * id[1], hiy = (q[4 * c + 3] - o[1]) * id[1];
            float loz = (q[8 + 2 * c] - o[2]) * id[2], hiz = (q[9 + 2 * c] - o[2]) * id[2];
            tn[c] = std::fmax(std::fmax(std::fmin(lox, hix), std::fmin(loy, hiy)), std::fmax(std::fmin(loz, hiz), mint));
            tf[c] = std::fmin(std::fmin(std::fmax(lox, hix), std::fmax(loy, hiy)), std::fmin(std::fmax(loz, hiz), tmax));
        }
        bool h0 = tn[0] <= tf[0] * 1.0000004f, h1 = tn[1] <= tf[1] * 1.0000004f;
        int c0, c1;
        std::memcpy(&c0, &q[12], 4); std::memcpy(&c1, &q[13], 4);
        float t0 = tn[0], t1 = tn[1];
        if (h0 && h1) {
            if (t1 < t0) { std::swap(c0, c1); std::swap(t0, t1); }
            stack[sp] = c1; stackT[sp] = t1; sp++;
            node = c0;
        } else if (h0 || h1) node = h0 ? c0 : c1;
        else node = pop();
    }
    // leaf: returns number of prims tested
    int leaf(const pg::PrimRecord *prims) {
        uint32_t code = (uint32_t)(~node);
        uint32_t first = code >> pg::kLeafShift, count = code & 7u, rectMask = (code >> 3) & 15u;
        leaves++;
        for (uint32_t i = 0; i < count; ++i) {
            const float *r = prims[first + i].q;
            primTests++;
            float loz = r[8] * o[0] + r[9] * o[1] + r[10] * o[2] + r[11];
            float ldz = r[8] * d[0] + r[9] * d[1] + r[10] * d[2];
            float tt = -loz / ldz;
            if (tt >= mint && tt <= tmax) {
                float lox = r[0] * o[0] + r[1] * o[1] + r[2] * o[2] + r[3], loy = r[4] * o[0] + r[5] * o[1] + r[6] * o[2] + r[7];
                float ldx = r[0] * d[0] + r[1] * d[1] + r[2] * d[2], ldy = r[4] * d[0] + r[5] * d[1] + r[6] * d[2];
                float u = lox + ldx * tt, v = loy + ldy * tt;
                bool isRect = (rectMask >> i) & 1u;
                bool ok = isRect ? (std::fabs(u) <= 1 && std::fabs(v) <= 1) : (u >= 0 && v >= 0 && u + v <= 1.0f);
                if (ok) { prim = first + i; tmax = tt; t = tt; if (any) { node = kDone; return (int)i + 1; } }
            }
        }
        node = pop();
        return (int)count;
    }
    bool done() const { return node == kDone; }
};
}  // namespace


// ---- host replay of the device's wide-node visit (device_scene.cuh: wideNodeStep), operation for operation
struct WideRay {
    float o[3], d[3], id[3], mint, tmax;
    int node, sp, maxSp = 0;
    struct E { int ref; float tn; } stack[256];
    uint32_t prim; float t;
    uint32_t nodeSteps = 0, primTests = 0;
    static float plane(uint32_t w, int byte, float s, float b) {
        uint32_t q = (w >> (8 * byte)) & 0xFFu, bits = 0x4B000000u | q;
        float f;
        std::memcpy(&f, &bits, 4);
        return std::fmaf(f, s, b);
    }
    int pop() {
        while (sp) { --sp; if (stack[sp].tn <= tmax) return stack[sp].ref; }
        return kDone;
    }
    void step(const pg::WideNode *nodes) {
        const float *q = nodes[node].q;
        nodeSteps++;
        uint32_t meta; std::memcpy(&meta, &q[3], 4);
        auto p2 = [](uint32_t e) { uint32_t b = e << 23; float f; std::memcpy(&f, &b, 4); return f; };
        float s[3] = {p2(meta & 0xFF) * id[0], p2((meta >> 8) & 0xFF) * id[1], p2((meta >> 16) & 0xFF) * id[2]};
        float b[3];
        for (int a = 0; a < 3; ++a) b[a] = std::fmaf(-8388608.0f, s[a], (q[a] - o[a]) * id[a]);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&q[12]);  // 12 words: lox0 lox1 loy0 loy1 loz0 loz1 hix0 hix1 hiy0 hiy1 hiz0 hiz1
        int best = kDone; float bestT = INFINITY;
        for (int c = 0; c < 8; ++c) {
            int ref; std::memcpy(&ref, &q[4 + c], 4);
            int wi = c >> 2, by = c & 3;
            float ax = plane(w[0 + wi], by, s[0], b[0]), cx = plane(w[6 + wi], by, s[0], b[0]);
            float ay = plane(w[2 + wi], by, s[1], b[1]), cy = plane(w[8 + wi], by, s[1], b[1]);
            float az = plane(w[4 + wi], by, s[2], b[2]), cz = plane(w[10 + wi], by, s[2], b[2]);
            float tn = std::fmax(std::fmax(std::fmin(ax, cx), std::fmin(ay, cy)), std::fmax(std::fmin(az, cz), mint));
            float tf = std::fmin(std::fmin(std::fmax(ax, cx), std::fmax(ay, cy)), std::fmin(std::fmax(az, cz), tmax));
            bool hit = ref != 0x7FFFFFFF && tn <= tf * 1.0000004f;
            bool better = hit && tn < bestT;
            int oref = better ? best : ref; float ot = better ? bestT : tn;
            if (hit && oref != kDone) { stack[sp].ref = oref; stack[sp].tn = ot; ++sp; maxSp = std::max(maxSp, sp); }
            best = better ? ref : best; bestT = better ? tn : bestT;
        }
        node = best != kDone ? best : pop();
    }
    void leaf(const pg::PrimRecord *prims) {
        uint32_t code = (uint32_t)(~node);
        uint32_t first = code >> pg::kLeafShift, count = code & 7u, rectMask = (code >> 3) & 15u;
        for (uint32_t i = 0; i < count; ++i) {
            const float *r = prims[first + i].q;
            primTests++;
            float loz = r[8] * o[0] + r[9] * o[1] + r[10] * o[2] + r[11];
            float ldz = r[8] * d[0] + r[9] * d[1] + r[10] * d[2];
            float tt = -loz / ldz;
            if (tt >= mint && tt <= tmax) {
                float lox = r[0] * o[0] + r[1] * o[1] + r[2] * o[2] + r[3], loy = r[4] * o[0] + r[5] * o[1] + r[6] * o[2] + r[7];
                float ldx = r[0] * d[0] + r[1] * d[1] + r[2] * d[2], ldy = r[4] * d[0] + r[5] * d[1] + r[6] * d[2];
                float u = lox + ldx * tt, v = loy + ldy * tt;
                bool isRect = (rectMask >> i) & 1u;
                bool ok = isRect ? (std::fabs(u) <= 1 && std::fabs(v) <= 1) : (u >= 0 && v >= 0 && u + v <= 1.0f);
                if (ok) { prim = first + i; tmax = tt; t = tt; }
            }
        }
        node = pop();
    }
};
extern "C" {
// Per-ray counters + hit (for generating secondary rays): out arrays may be null.
void bvhsim_trace(void *h, const float *rays, size_t n, int anyHit, int cull, uint32_t *nodeSteps, uint32_t *primTests,
                  uint32_t *leaves, float *tOut, uint32_t *primOut) {
    const pg::HostScene &H = ((SceneHandle *)h)->host;
#pragma omp parallel for schedule(dynamic, 1024)
    for (size_t i = 0; i < n; ++i) {
        RayState R;
        R.init(rays + 8 * i, anyHit, cull);
        while (!R.done()) {
            while (R.node >= 0) R.step(H.nodes.data());
            if (R.done()) break;
            R.leaf(primRecords(H));
        }
        if (nodeSteps) nodeSteps[i] = R.nodeSteps;
        if (primTests) primTests[i] = R.primTests;
        if (leaves) leaves[i] = R.leaves;
        if (tOut) tOut[i] = R.prim == 0xFFFFFFFFu ? INFINITY : R.t;
        if (primOut) primOut[i] = R.prim;
    }
}

// Model of the warp-cooperative kernel (kernels.cu: k_trace_tail): ONE ray served by 32 lanes from the root. Per round every lane
// pops one pending entry (up to 32; one above `wide` entries, the depth-first fallback), tests that node's two children or that
// leaf's primitives against the hit distance of the PREVIOUS round, and pushes what the ray enters. Reports per ray the number of
// rounds (= dependent memory round trips, against nodeSteps + leaves of the sequential walk), the node tests (work, against
// nodeSteps) and the largest stack. The hit must equal the sequential one (asserted by the caller on tOut / primOut).
void bvhsim_coop(void *h, const float *rays, size_t n, int anyHit, int wide, uint32_t *rounds, uint32_t *nodeTests, uint32_t *maxStack,
                 float *tOut, uint32_t *primOut) {
    const pg::HostScene &H = ((SceneHandle *)h)->host;
    const pg::PrimRecord *prims = primRecords(H);
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t i = 0; i < n; ++i) {
        RayState R;  // used for its node / leaf arithmetic only
        R.init(rays + 8 * i, anyHit != 0, false);
        std::vector<int> st;
        st.push_back(0);
        uint32_t nr = 0, nt = 0, ms = 1;
        float tmax = R.tmax, bestT = R.tmax;
        uint32_t bestPrim = 0xFFFFFFFFu;
        bool found = false;
        while (!st.empty() && !found) {
            const int take = (int)st.size() > wide ? 1 : std::min<int>((int)st.size(), 32);
            int entry[32];
            for (int l = 0; l < take; ++l) { entry[l] = st.back(); st.pop_back(); }
            std::vector<int> farC, nearC;
            float roundT = tmax;
            uint32_t roundPrim = bestPrim;
            for (int l = 0; l < take; ++l) {
                RayState L = R;
                L.tmax = tmax;  // every lane of the round prunes with the previous round's distance
                L.sp = 0;
                if (entry[l] >= 0) {
                    nt++;
                    L.node = entry[l];
                    L.step(H.nodes.data());  // leaves: node = nearer child, stack[0] = farther one (if both are entered)
                    if (L.sp == 1) farC.push_back(L.stack[0]);
                    if (L.node != kDone) nearC.push_back(L.node);
                } else {
                    L.node = entry[l];
                    L.prim = 0xFFFFFFFFu;
                    L.leaf(prims);
                    if (L.prim != 0xFFFFFFFFu) {
                        if (anyHit) { found = true; roundPrim = L.prim; }
                        else if (L.t < roundT || (L.t == roundT && L.prim < roundPrim)) { roundT = L.t; roundPrim = L.prim; }
                    }
                }
            }
            for (int c : farC) st.push_back(c);
            for (size_t k = nearC.size(); k-- > 0;) st.push_back(nearC[k]);
            if (roundPrim != bestPrim) { bestPrim = roundPrim; bestT = roundT; tmax = roundT; }
            ms = std::max<uint32_t>(ms, (uint32_t)st.size());
            nr++;
        }
        if (rounds) rounds[i] = nr;
        if (nodeTests) nodeTests[i] = nt;
        if (maxStack) maxStack[i] = ms;
        if (tOut) tOut[i] = bestPrim == 0xFFFFFFFFu ? INFINITY : bestT;
        if (primOut) primOut[i] = bestPrim;
    }
}

// Warp model. mode 0: lock-step batches of 32 consecutive rays (while-while). mode 1: persistent warps with per-lane refill when
// fewer than `thresh` lanes are active (checked once per outer iteration = descent + leaf). mode 2: as mode 1 but with
// speculative descent (a lane that reached a leaf postpones it and keeps descending until every lane holds a leaf).
// out[0] = lane node steps, out[1] = warp node steps (x32 = issued lane slots), out[2] = lane prim tests, out[3] = warp prim steps,
// out[4] = max node steps of one ray, out[5] = warp "rounds"
void bvhsim_warps(void *h, const float *rays, size_t n, int anyHit, int cull, int mode, int thresh, int nWarps, double *out) {
    const pg::HostScene &H = ((SceneHandle *)h)->host;
    double laneNode = 0, warpNode = 0, lanePrim = 0, warpPrim = 0, maxRay = 0, rounds = 0;
    if (mode == 0) {
        size_t nw = (n + 31) / 32;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : laneNode, warpNode, lanePrim, warpPrim, rounds) reduction(max : maxRay)
        for (size_t w = 0; w < nw; ++w) {
            RayState R[32];
            int m = (int)std::min<size_t>(32, n - 32 * w);
            for (int l = 0; l < m; ++l) R[l].init(rays + 8 * (32 * w + l), anyHit, cull);
            while (true) {
                int maxSteps = 0, maxPrims = 0, active = 0;
                for (int l = 0; l < m; ++l) {
                    if (R[l].done()) continue;
                    active++;
                    int s = 0;
                    while (R[l].node >= 0) { R[l].step(H.nodes.data()); s++; }
                    laneNode += s;
                    maxSteps = std::max(maxSteps, s);
                    if (!R[l].done()) { int p = R[l].leaf(primRecords(H)); lanePrim += p; maxPrims = std::max(maxPrims, p); }
                }
                if (!active) break;
                warpNode += maxSteps; warpPrim += maxPrims; rounds++;
            }
            for (int l = 0; l < m; ++l) maxRay = std::max<double>(maxRay, R[l].nodeSteps);
        }
    } else if (mode < 3) {
        // persistent warps, shared cursor, round-robin scheduling
        size_t cursor = 0;
        std::vector<RayState> R((size_t)nWarps * 32);
        std::vector<char> has((size_t)nWarps * 32, 0);
        std::vector<int> leafOf((size_t)nWarps * 32, 0);
        std::vector<char> alive(nWarps, 1);
        int nAlive = nWarps;
        while (nAlive) {
            for (int w = 0; w < nWarps; ++w) {
                if (!alive[w]) continue;
                RayState *L = &R[(size_t)w * 32];
                char *hs = &has[(size_t)w * 32];
                int active = 0;
                for (int l = 0; l < 32; ++l) active += hs[l];
                if (active < thresh && cursor < n) {
                    for (int l = 0; l < 32 && cursor < n; ++l)
                        if (!hs[l]) { L[l].init(rays + 8 * cursor, anyHit, cull); cursor++; hs[l] = 1; active++; }
                }
                if (!active) { alive[w] = 0; nAlive--; continue; }
                int maxSteps = 0, maxPrims = 0;
                if (mode == 1) {
                    for (int l = 0; l < 32; ++l) {
                        if (!hs[l]) continue;
                        int s = 0;
                        while (L[l].node >= 0) { L[l].step(H.nodes.data()); s++; }
                        laneNode += s; maxSteps = std::max(maxSteps, s);
                        if (!L[l].done()) { int p = L[l].leaf(primRecords(H)); lanePrim += p; maxPrims = std::max(maxPrims, p); }
                        if (L[l].done()) { hs[l] = 0; maxRay = std::max<double>(maxRay, L[l].nodeSteps); }
                    }
                } else {
                    // speculative: lock-step node steps; a lane holding a postponed leaf keeps descending until it meets a 2nd leaf
                    int *pl = &leafOf[(size_t)w * 32];
                    for (int l = 0; l < 32; ++l) pl[l] = 0;  // 0 = none (leaf codes are negative)
                    while (true) {
                        bool anyStepping = false, anyWithoutLeaf = false;
                        for (int l = 0; l < 32; ++l) {
                            if (!hs[l]) continue;
                            if (L[l].node >= 0) anyStepping = true;
                            if (pl[l] == 0 && !L[l].done()) anyWithoutLeaf = true;
                        }
                        if (!anyStepping || !anyWithoutLeaf) break;
                        maxSteps++;
                        for (int l = 0; l < 32; ++l) {
                            if (!hs[l] || L[l].node < 0) continue;
                            L[l].step(H.nodes.data()); laneNode++;
                            if (L[l].node < 0 && L[l].node != kDone && pl[l] == 0) { pl[l] = L[l].node; L[l].node = L[l].pop(); }
                        }
                    }
                    // leaves: postponed first, then a second leaf the lane may be standing on
                    for (int rep = 0; rep < 2; ++rep) {
                        int mp = 0;
                        for (int l = 0; l < 32; ++l) {
                            if (!hs[l]) continue;
                            int lf = 0;
                            if (pl[l] != 0) { lf = pl[l]; pl[l] = 0; }
                            else if (L[l].node < 0 && L[l].node != kDone) { lf = L[l].node; L[l].node = L[l].pop(); }
                            if (!lf) continue;
                            int save = L[l].node; L[l].node = lf;
                            // leaf() pops on its own: emulate by pushing back the saved continuation
                            if (save != kDone) { L[l].stack[L[l].sp] = save; L[l].stackT[L[l].sp] = -1e30f; L[l].sp++; }
                            int p = L[l].leaf(primRecords(H));
                            lanePrim += p; mp = std::max(mp, p);
                        }
                        maxPrims += mp;
                    }
                    for (int l = 0; l < 32; ++l)
                        if (hs[l] && L[l].done()) { hs[l] = 0; maxRay = std::max<double>(maxRay, L[l].nodeSteps); }
                }
                warpNode += maxSteps; warpPrim += maxPrims; rounds++;
            }
        }
    }

    if (mode == 3) {
        // fixed bursts: every round runs `thresh` lock-step node steps; a lane that reaches a leaf appends it to a small per-lane
        // leaf queue (capacity 4; a full queue stalls the lane) and keeps descending; after the burst all queued leaves are
        // tested (one per lane and iteration), then idle lanes are refilled.
        const int burst = thresh, cap = 4;
        laneNode = warpNode = lanePrim = warpPrim = maxRay = rounds = 0;
        size_t cursor = 0;
        std::vector<RayState> R((size_t)nWarps * 32);
        std::vector<char> has((size_t)nWarps * 32, 0);
        std::vector<int> lq((size_t)nWarps * 32 * cap, 0), lqn((size_t)nWarps * 32, 0);
        std::vector<char> alive(nWarps, 1);
        int nAlive = nWarps;
        while (nAlive) {
            for (int w = 0; w < nWarps; ++w) {
                if (!alive[w]) continue;
                RayState *L = &R[(size_t)w * 32];
                char *hs = &has[(size_t)w * 32];
                int *q = &lq[(size_t)w * 32 * cap], *qn = &lqn[(size_t)w * 32];
                int active = 0;
                for (int l = 0; l < 32; ++l) {
                    if (!hs[l] && cursor < n) { L[l].init(rays + 8 * cursor, anyHit, cull); cursor++; hs[l] = 1; qn[l] = 0; }
                    active += hs[l];
                }
                if (!active) { alive[w] = 0; nAlive--; continue; }
                int steps = 0;
                for (int b = 0; b < burst; ++b) {
                    bool any = false;
                    for (int l = 0; l < 32; ++l) {
                        if (!hs[l] || L[l].node < 0) continue;
                        any = true;
                        L[l].step(H.nodes.data()); laneNode++;
                        while (L[l].node < 0 && L[l].node != kDone && qn[l] < cap) { q[l * cap + qn[l]++] = L[l].node; L[l].node = L[l].pop(); }
                    }
                    if (!any) break;
                    steps++;
                }
                warpNode += steps;
                // leaves
                while (true) {
                    int mp = 0; bool any = false;
                    for (int l = 0; l < 32; ++l) {
                        if (!hs[l] || qn[l] == 0) continue;
                        any = true;
                        int lf = q[l * cap];
                        for (int k = 1; k < qn[l]; ++k) q[l * cap + k - 1] = q[l * cap + k];
                        qn[l]--;
                        int save = L[l].node; L[l].node = lf;
                        if (save != kDone) { L[l].stack[L[l].sp] = save; L[l].stackT[L[l].sp] = -1e30f; L[l].sp++; }
                        int p = L[l].leaf(primRecords(H));
                        if (anyHit && L[l].prim != 0xFFFFFFFFu) { L[l].node = kDone; qn[l] = 0; }
                        lanePrim += p; mp = std::max(mp, p);
                        // a leaf the lane was stalled on
                        while (L[l].node < 0 && L[l].node != kDone && qn[l] < cap) { q[l * cap + qn[l]++] = L[l].node; L[l].node = L[l].pop(); }
                    }
                    if (!any) break;
                    warpPrim += mp;
                }
                for (int l = 0; l < 32; ++l)
                    if (hs[l] && L[l].done() && qn[l] == 0) { hs[l] = 0; maxRay = std::max<double>(maxRay, L[l].nodeSteps); }
                rounds++;
            }
        }
    }

    if (mode >= 4) {
        // voting schedule: mode = 4 + 100 * TL + 10000 * TR; thresh = node burst length.
        //   every iteration: refill when >= TR lanes are idle; leaf phase (one queued leaf per lane) when >= TL lanes hold a leaf or
        //   no lane can take a node step; otherwise a burst of node steps.
        const int burst = thresh, cap = 4, TL = (mode / 100) % 100, TR = (mode / 10000) % 100;
        laneNode = warpNode = lanePrim = warpPrim = maxRay = rounds = 0;
        double refills = 0;
        size_t cursor = 0;
        std::vector<RayState> R((size_t)nWarps * 32);
        std::vector<char> has((size_t)nWarps * 32, 0);
        std::vector<int> lq((size_t)nWarps * 32 * cap, 0), lqn((size_t)nWarps * 32, 0);
        std::vector<char> alive(nWarps, 1);
        int nAlive = nWarps;
        while (nAlive) {
            for (int w = 0; w < nWarps; ++w) {
                if (!alive[w]) continue;
                RayState *L = &R[(size_t)w * 32];
                char *hs = &has[(size_t)w * 32];
                int *q = &lq[(size_t)w * 32 * cap], *qn = &lqn[(size_t)w * 32];
                int idle = 0;
                for (int l = 0; l < 32; ++l) idle += !hs[l];
                if (idle >= TR && cursor < n) {
                    refills++;
                    for (int l = 0; l < 32; ++l)
                        if (!hs[l] && cursor < n) { L[l].init(rays + 8 * cursor, anyHit, cull); cursor++; hs[l] = 1; qn[l] = 0; }
                }
                int nN = 0, nL = 0, act = 0;
                for (int l = 0; l < 32; ++l) {
                    if (!hs[l]) continue;
                    act++;
                    if (L[l].node >= 0) nN++;
                    if (qn[l]) nL++;
                }
                if (!act) { alive[w] = 0; nAlive--; continue; }
                if (nL >= TL || nN == 0) {
                    int mp = 0;
                    for (int l = 0; l < 32; ++l) {
                        if (!hs[l] || qn[l] == 0) continue;
                        int lf = q[l * cap];
                        for (int k = 1; k < qn[l]; ++k) q[l * cap + k - 1] = q[l * cap + k];
                        qn[l]--;
                        int save = L[l].node; L[l].node = lf;
                        if (save != kDone) { L[l].stack[L[l].sp] = save; L[l].stackT[L[l].sp] = -1e30f; L[l].sp++; }
                        int p = L[l].leaf(primRecords(H));
                        if (anyHit && L[l].prim != 0xFFFFFFFFu) { L[l].node = kDone; qn[l] = 0; }
                        lanePrim += p; mp = std::max(mp, p);
                        while (L[l].node < 0 && L[l].node != kDone && qn[l] < cap) { q[l * cap + qn[l]++] = L[l].node; L[l].node = L[l].pop(); }
                    }
                    warpPrim += mp;
                } else {
                    int steps = 0;
                    for (int b = 0; b < burst; ++b) {
                        bool any = false;
                        for (int l = 0; l < 32; ++l) {
                            if (!hs[l] || L[l].node < 0) continue;
                            any = true;
                            L[l].step(H.nodes.data()); laneNode++;
                            while (L[l].node < 0 && L[l].node != kDone && qn[l] < cap) { q[l * cap + qn[l]++] = L[l].node; L[l].node = L[l].pop(); }
                        }
                        if (!any) break;
                        steps++;
                    }
                    warpNode += steps;
                }
                for (int l = 0; l < 32; ++l)
                    if (hs[l] && L[l].done() && qn[l] == 0) { hs[l] = 0; maxRay = std::max<double>(maxRay, L[l].nodeSteps); }
                rounds++;
            }
        }
        out[6] = refills;
    }
    out[0] = laneNode; out[1] = warpNode; out[2] = lanePrim; out[3] = warpPrim; out[4] = maxRay; out[5] = rounds;
}

// closest hit through the wide tree; out: per-ray node visits, primitive tests, hit; returns the largest stack depth seen
int bvhsim_trace_wide(void *h, const float *rays, size_t n, uint32_t *nodeSteps, uint32_t *primTests, float *tOut, uint32_t *primOut) {
    const pg::HostScene &H = ((SceneHandle *)h)->host;
    if (H.wideNodes.empty()) return -1;
    int maxSp = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : maxSp)
    for (size_t i = 0; i < n; ++i) {
        WideRay R;
        const float *r = rays + 8 * i;
        for (int a = 0; a < 3; ++a) { R.o[a] = r[a]; R.d[a] = r[4 + a]; R.id[a] = 1.0f / R.d[a]; R.id[a] = std::copysign(std::fmin(std::fabs(R.id[a]), 1e25f), R.id[a]); }
        R.mint = r[3]; R.tmax = r[7]; R.node = 0; R.sp = 0; R.prim = 0xFFFFFFFFu; R.t = R.tmax;
        while (R.node != kDone) {
            while (R.node >= 0) R.step(H.wideNodes.data());
            if (R.node == kDone) break;
            R.leaf(primRecords(H));
        }
        if (nodeSteps) nodeSteps[i] = R.nodeSteps;
        if (primTests) primTests[i] = R.primTests;
        if (tOut) tOut[i] = R.prim == 0xFFFFFFFFu ? INFINITY : R.t;
        if (primOut) primOut[i] = R.prim;
        maxSp = std::max(maxSp, R.maxSp);
    }
    return maxSp;
}
void bvhsim_wide_info(void *h, uint64_t *out) {
    const pg::HostScene &H = ((SceneHandle *)h)->host;
    out[0] = H.wideNodes.size(); out[1] = (uint64_t)H.wideDepth; out[2] = H.nodes.size();
    uint64_t kids = 0;
    for (auto &w : H.wideNodes) { uint32_t m; std::memcpy(&m, &w.q[3], 4); kids += m >> 24; }
    out[3] = kids;
}
}

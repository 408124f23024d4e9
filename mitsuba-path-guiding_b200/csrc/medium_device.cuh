// medium_device.cuh -- heterogeneous medium over a dense grid volume (device routines).
//   gridLookup         GridDataSource::lookupFloat, src/volume/gridvolume.cpp:337-388 (trilinear, zero outside)
//   mediumSampleDistance / mediumTransmittance
//                      HeterogeneousMedium::sampleDistance / evalTransmittance, Woodcock branch
//                      (src/medium/heterogeneous.cpp:589-663, 546-587; sigma_max = scale * 1, gridvolume.cpp:583-585)
//   mediumIntegrateDensity / mediumInvertDensityIntegral
//                      method = simpson: composite Simpson quadrature and its Newton-bisection inversion
//                      (heterogeneous.cpp:301-376, 420-544; HETVOL_EARLY_EXIT on, :31)
//   phaseEval / phaseSample   src/phase/hg.cpp:74-110, src/phase/isotropic.cpp:62-78
#pragma once
#include "device_scene.cuh"
#include "guiding_device.cuh"

namespace pg {

PG_DEV float gridLookup(const MediumRecord &M, const float *__restrict__ density, float3 p) {
    const float px = M.worldToGrid[0] * p.x + M.worldToGrid[3], py = M.worldToGrid[5] * p.y + M.worldToGrid[7],
                pz = M.worldToGrid[10] * p.z + M.worldToGrid[11];
    const int x1 = (int)floorf(px), y1 = (int)floorf(py), z1 = (int)floorf(pz), x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
    const int rx = M.res[0], ry = M.res[1], rz = M.res[2];
    if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= rx || y2 >= ry || z2 >= rz) return 0.0f;
    const float fx = px - x1, fy = py - y1, fz = pz - z1, _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
    const float *f = density + M.densityOffset;
    const size_t r0 = ((size_t)z1 * ry + y1) * rx, r1 = ((size_t)z1 * ry + y2) * rx, r2 = ((size_t)z2 * ry + y1) * rx,
                 r3 = ((size_t)z2 * ry + y2) * rx;
    const float d000 = __ldg(f + r0 + x1), d001 = __ldg(f + r0 + x2), d010 = __ldg(f + r1 + x1), d011 = __ldg(f + r1 + x2);
    const float d100 = __ldg(f + r2 + x1), d101 = __ldg(f + r2 + x2), d110 = __ldg(f + r3 + x1), d111 = __ldg(f + r3 + x2);
    return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz +
           ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
}

// AABB::rayIntersect of the density box (include/mitsuba/core/aabb.h:308-338)
PG_DEV bool mediumClip(const MediumRecord &M, float3 o, float3 d, float &nearT, float &farT) {
    nearT = -kInf;
    farT = kInf;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float origin = comp(o, i), dir = comp(d, i), minVal = M.aabbMin[i], maxVal = M.aabbMax[i];
        if (dir == 0) {
            if (origin < minVal || origin > maxVal) return false;
        } else {
            const float rcp = 1.0f / dir;
            float t1 = (minVal - origin) * rcp, t2 = (maxVal - origin) * rcp;
            if (t1 > t2) {
                const float tmp = t1; t1 = t2; t2 = tmp;
            }
            nearT = fmaxf(t1, nearT);
            farT = fminf(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}

// heterogeneous.cpp:301-376
PG_DEV float mediumIntegrateDensity(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt) {
    float mint, maxt;
    if (!mediumClip(M, o, d, mint, maxt)) return 0.0f;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    const float length = maxt - mint;
    float3 p = o + d * mint;
    const float3 pLast = o + d * maxt;
    const float maxComp = fmaxf(fmaxf(fmaxf(fabsf(p.x), fabsf(pLast.x)), fmaxf(fabsf(p.y), fabsf(pLast.y))), fmaxf(fabsf(p.z), fabsf(pLast.z)));
    if (length < 1e-6f * maxComp) return 0.0f;
    uint32_t nSteps = (uint32_t)ceilf(length / M.stepSize);
    nSteps += nSteps % 2;
    const float step = length / nSteps;
    const float3 increment = d * step;
    float integrated = gridLookup(M, density, p) + gridLookup(M, density, pLast);
    const float stopValue = -logf(kEpsilon) * 3.0f / (step * M.scale);
    p = p + increment;
    float m = 4;
    for (uint32_t i = 1; i < nSteps; ++i) {
        integrated += m * gridLookup(M, density, p);
        m = 6 - m;
        if (integrated > stopValue) return kInf;
        const float3 next = p + increment;
        if (p.x == next.x && p.y == next.y && p.z == next.z) break;
        p = next;
    }
    return integrated * M.scale * step * (1.0f / 3.0f);
}

// heterogeneous.cpp:420-544
PG_DEV bool mediumInvertDensityIntegral(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt,
                                        float desiredDensity, float &integratedDensity, float &t, float &densityAtT) {
    integratedDensity = densityAtT = 0.0f;
    float mint, maxt;
    if (!mediumClip(M, o, d, mint, maxt)) return false;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    const float length = maxt - mint;
    float3 p = o + d * mint;
    const float3 pLast = o + d * maxt;
    const float maxComp = fmaxf(fmaxf(fmaxf(fabsf(p.x), fabsf(pLast.x)), fmaxf(fabsf(p.y), fabsf(pLast.y))), fmaxf(fabsf(p.z), fabsf(pLast.z)));
    if (length < 1e-6f * maxComp) return false;
    const uint32_t nSteps = (uint32_t)ceilf(length / (2 * M.stepSize));
    const float step = length / nSteps, multiplier = (1.0f / 6.0f) * step * M.scale;
    const float3 fullStep = d * step, halfStep = fullStep * 0.5f;
    float node1 = gridLookup(M, density, p);
    for (uint32_t i = 0; i < nSteps; ++i) {
        const float node2 = gridLookup(M, density, p + halfStep), node3 = gridLookup(M, density, p + fullStep);
        const float newDensity = integratedDensity + multiplier * (node1 + node2 * 4 + node3);
        if (newDensity >= desiredDensity) {
            float a = 0, b = step, x = a, fx = integratedDensity - desiredDensity;
            const float stepSqr = step * step, temp = M.scale / stepSqr;
            int it = 1;
            while (true) {
                const float dfx = temp * (node1 * stepSqr - (3 * node1 - 4 * node2 + node3) * step * x + 2 * (node1 - 2 * node2 + node3) * x * x);
                x -= fx / dfx;
                if (x <= a || x >= b || dfx == 0) x = 0.5f * (b + a);
                const float intval = integratedDensity + temp * (1.0f / 6.0f) *
                                     (x * (6 * node1 * stepSqr - 3 * (3 * node1 - 4 * node2 + node3) * step * x +
                                           4 * (node1 - 2 * node2 + node3) * x * x));
                fx = intval - desiredDensity;
                if (fabsf(fx) < 1e-6f) {
                    t = mint + step * i + x;
                    integratedDensity = intval;
                    densityAtT = temp * (node1 * stepSqr - (3 * node1 - 4 * node2 + node3) * step * x +
                                         2 * (node1 - 2 * node2 + node3) * x * x);
                    return true;
                } else if (++it > 30) {
                    return false;
                }
                if (fx > 0) b = x; else a = x;
            }
        }
        const float3 next = p + fullStep;
        if (p.x == next.x && p.y == next.y && p.z == next.z) break;
        integratedDensity = newDensity;
        node1 = node3;
        p = next;
    }
    return false;
}

struct MediumSample {
    float t;
    float3 p, sigmaS;
    float transmittance;
    float pdfSuccess, pdfFailure;  // 1 for Woodcock tracking (heterogeneous.cpp:616-618)
};

// method = simpson branch of sampleDistance (heterogeneous.cpp:594-611)
PG_DEV bool mediumSampleDistanceSimpson(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt,
                                        MediumSample &mRec, Rng &rng) {
    float integratedDensity, densityAtT;
    bool success = false;
    const float desiredDensity = -logf(1 - rng.next1D());
    if (mediumInvertDensityIntegral(M, density, o, d, rmint, rmaxt, desiredDensity, integratedDensity, mRec.t, densityAtT)) {
        mRec.p = o + d * mRec.t;
        success = true;
        mRec.sigmaS = ld3(M.albedo) * densityAtT;
    }
    const float expVal = expf(-integratedDensity);
    mRec.pdfFailure = expVal;
    mRec.pdfSuccess = expVal * densityAtT;
    mRec.transmittance = expVal;
    return success && mRec.pdfSuccess > 0;
}

PG_DEV bool mediumSampleDistance(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt,
                                 MediumSample &mRec, Rng &rng) {
    if (M.method == B200PG_MEDIUM_SIMPSON) return mediumSampleDistanceSimpson(M, density, o, d, rmint, rmaxt, mRec, rng);
    mRec.pdfSuccess = mRec.pdfFailure = 1.0f;
    mRec.transmittance = 1.0f;
    float mint, maxt;
    if (!mediumClip(M, o, d, mint, maxt)) return false;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    float t = mint;
    while (true) {
        t -= logf(1 - rng.next1D()) * M.invMaxDensity;
        if (t >= maxt) break;
        const float3 p = o + d * t;
        const float densityAtT = gridLookup(M, density, p) * M.scale;
        if (densityAtT * M.invMaxDensity > rng.next1D()) {
            mRec.t = t;
            mRec.p = p;
            mRec.sigmaS = ld3(M.albedo) * densityAtT;
            float tr = densityAtT != 0.0f ? 1.0f / densityAtT : 0.0f;
            if (!isfinite(tr)) tr = 0.0f;
            mRec.transmittance = tr;
            return true;
        }
    }
    return false;
}

// Guided free-flight sampling: weighted delta tracking with field-steered collision probabilities (this repo's design;
// CPU statement and derivation: oracle/oracle_medium.h sampleDistanceGuided, DESIGN.md "Guiding in media").
PG_DEV bool mediumSampleDistanceGuided(const MediumRecord &M, const float *density, const GuideDevice &G, float3 o, float3 d,
                                       float rmint, float rmaxt, MediumSample &mRec, float3 &weight, Rng &rng) {
    weight = f3(1.0f);
    float mint, maxt;
    if (!mediumClip(M, o, d, mint, maxt)) return false;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    const float3 albedo = ld3(M.albedo);
    const float albedoAvg = (albedo.x + albedo.y + albedo.z) * (1.0f / 3.0f);
    const float beta = 0.5f;
    float t = mint;
    while (true) {
        t -= logf(1 - rng.next1D()) * M.invMaxDensity;
        if (t >= maxt) break;
        const float3 p = o + d * t;
        const float a = fminf(gridLookup(M, density, p) * M.scale * M.invMaxDensity, 1.0f);
        const float u = rng.next1D();
        if (!(a > 0)) continue;
        const float g = 4 * kPi * guidePdf(G, guideLookup(G, p), d);
        const float num = a * albedoAvg, den = num + (1 - a) * g;
        const float pGuided = den > 0 ? num / den : a;
        const float pReal = (1 - beta) * a + beta * pGuided;
        if (u < pReal) {
            weight = weight * (albedo * (a / pReal));
            mRec.t = t;
            mRec.p = p;
            return true;
        }
        weight = weight * ((1 - a) / (1 - pReal));
    }
    return false;
}

PG_DEV float mediumTransmittance(const MediumRecord &M, const float *density, float3 o, float3 d, float rmint, float rmaxt, Rng &rng) {
    if (M.method == B200PG_MEDIUM_SIMPSON) return expf(-mediumIntegrateDensity(M, density, o, d, rmint, rmaxt));  // (:547-548)
    float mint, maxt;
    if (!mediumClip(M, o, d, mint, maxt)) return 1.0f;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    float result = 0;
    for (int i = 0; i < 2; ++i) {  // nSamples = 2 (heterogeneous.cpp:562)
        float t = mint;
        while (true) {
            t -= logf(1 - rng.next1D()) * M.invMaxDensity;
            if (t >= maxt) {
                result += 1;
                break;
            }
            const float dens = gridLookup(M, density, o + d * t) * M.scale;
            if (dens * M.invMaxDensity > rng.next1D()) break;
        }
    }
    return result / 2;
}

PG_DEV float phaseEval(const MediumRecord &M, float3 wi, float3 wo) {
    if (M.phaseType == B200PG_PHASE_HG) {
        const float g = M.g;
        const float temp = 1.0f + g * g + 2.0f * g * dot(wi, wo);
        return (0.07957747154594766788f) * (1 - g * g) / (temp * sqrtf(temp));
    }
    return 0.07957747154594766788f;
}

PG_DEV float3 phaseSample(const MediumRecord &M, float3 wi, float2 sample, float &pdf) {
    float3 wo;
    if (M.phaseType == B200PG_PHASE_HG) {
        const float g = M.g;
        float cosTheta;
        if (fabsf(g) < kEpsilon) {
            cosTheta = 1 - 2 * sample.x;
        } else {
            const float sqrTerm = (1 - g * g) / (1 - g + 2 * g * sample.x);
            cosTheta = (1 + g * g - sqrTerm * sqrTerm) / (2 * g);
        }
        const float sinTheta = safeSqrt(1.0f - cosTheta * cosTheta);
        float sp, cp;
        sincosf(2 * kPi * sample.y, &sp, &cp);
        const Frame f = frameFromNormal(-wi);
        wo = f.toWorld(f3(sinTheta * cp, sinTheta * sp, cosTheta));
    } else {
        wo = squareToUniformSphere(sample);
    }
    pdf = phaseEval(M, wi, wo);
    return wo;
}

// geometric normal from the triangle winding / rectangle frame, as ShapeKDTree::rayIntersect(ray, t, shape, n, uv)
// returns it (skdtree.cpp:163-200) -- not flipped towards the shading normal
PG_DEV float3 windingNormal(const DeviceScene &S, uint32_t primSlot, const ShapeRecord &sr, uint32_t primIdx) {
    if (sr.type == B200PG_SHAPE_TRIMESH) {
        const MeshRecord mr = S.meshes[sr.meshOffset];
        const uint32_t *idx = S.indices + 3 * ((size_t)mr.indexOffset + primIdx);
        const float3 p0 = ld3(S.positions + 3 * (size_t)(idx[0] + mr.vertexOffset)), p1 = ld3(S.positions + 3 * (size_t)(idx[1] + mr.vertexOffset)),
                     p2 = ld3(S.positions + 3 * (size_t)(idx[2] + mr.vertexOffset));
        return normalize(cross(p1 - p0, p2 - p0));
    }
    const float4 r3 = __ldg(S.rects + 8 * sr.meshOffset + 3);
    return f3(r3.x, r3.y, r3.z);
}

PG_DEV int targetMedium(const ShapeRecord &sr, float3 geoN, float3 d) {  // records.inl:81-86
    return dot(d, geoN) > 0 ? sr.exteriorMedium : sr.interiorMedium;
}

}  // namespace pg

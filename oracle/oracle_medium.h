// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_medium.h: heterogeneous medium over a dense grid volume, restated from
//   src/medium/heterogeneous.cpp:228-262 (configure), :546-587 (evalTransmittance), :589-663 (sampleDistance, Woodcock),
//   src/volume/gridvolume.cpp:188-198 (worldToGrid, step size), :337-388 (trilinear lookupFloat),
//   src/phase/hg.cpp:74-110, src/phase/isotropic.cpp:62-78.
// Both integration methods are restated: Woodcock / delta tracking (the default, heterogeneous.cpp:195-197) and composite
// Simpson quadrature with Newton-bisection inversion (:301-376 integrateDensity, :420-544 invertDensityIntegral; the
// HETVOL_EARLY_EXIT shortcut of :336-355 is on by default, :31, and is restated too).
#pragma once
#include <vector>

#include "../include/b200pg.h"
#include "oracle_math.h"

namespace orc {

struct MediumSample {  // MediumSamplingRecord fields used on this path
    Float t;
    Vec3 p;
    Vec3 sigmaS, transmittance;
    Float pdfSuccess, pdfFailure;
};

struct Medium {
    B200pgMedium d;
    std::vector<Float> density;
    Float maxDensity, invMaxDensity, stepSize;
    Float gridScale[3], gridOffset[3];  // worldToGrid = scale((res-1)/extent) * translate(-min)
    Vec3 bmin, bmax;

    static Medium fromDesc(const B200pgMedium &m) {
        Medium r;
        r.d = m;
        size_t n = (size_t)m.res[0] * m.res[1] * m.res[2];
        if (m.density) r.density.assign(m.density, m.density + n);
        r.d.density = nullptr;
        r.maxDensity = m.scale * 1.0f;  // getMaximumFloatValue() == 1 (gridvolume.cpp:583-585)
        r.invMaxDensity = 1.0f / r.maxDensity;
        for (int c = 0; c < 3; ++c) {
            Float ext = m.aabb_max[c] - m.aabb_min[c];
            r.gridScale[c] = (m.res[c] - 1) / ext;
            r.gridOffset[c] = r.gridScale[c] * -m.aabb_min[c];
        }
        // heterogeneous.cpp:244-256: stepSize property, else the density volume's (gridvolume.cpp:196-198; constvolume: inf)
        r.stepSize = m.step_size_multiplier;
        if (r.stepSize == 0) {
            r.stepSize = std::numeric_limits<Float>::infinity();
            for (int c = 0; c < 3; ++c) r.stepSize = std::min(r.stepSize, 0.5f * (m.aabb_max[c] - m.aabb_min[c]) / (Float)(m.res[c] - 1));
        }
        r.bmin = Vec3(m.aabb_min[0], m.aabb_min[1], m.aabb_min[2]);
        r.bmax = Vec3(m.aabb_max[0], m.aabb_max[1], m.aabb_max[2]);
        return r;
    }

    // GridDataSource::lookupFloat, gridvolume.cpp:337-388 (float32 data, zero outside)
    Float lookup(const Vec3 &_p) const {
        const Float px = gridScale[0] * _p.x + gridOffset[0], py = gridScale[1] * _p.y + gridOffset[1],
                    pz = gridScale[2] * _p.z + gridOffset[2];
        const int x1 = (int)std::floor(px), y1 = (int)std::floor(py), z1 = (int)std::floor(pz), x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
        const int rx = d.res[0], ry = d.res[1], rz = d.res[2];
        if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= rx || y2 >= ry || z2 >= rz) return 0;
        const Float fx = px - x1, fy = py - y1, fz = pz - z1, _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
        const Float *f = density.data();
        const Float d000 = f[((size_t)z1 * ry + y1) * rx + x1], d001 = f[((size_t)z1 * ry + y1) * rx + x2],
                    d010 = f[((size_t)z1 * ry + y2) * rx + x1], d011 = f[((size_t)z1 * ry + y2) * rx + x2],
                    d100 = f[((size_t)z2 * ry + y1) * rx + x1], d101 = f[((size_t)z2 * ry + y1) * rx + x2],
                    d110 = f[((size_t)z2 * ry + y2) * rx + x1], d111 = f[((size_t)z2 * ry + y2) * rx + x2];
        return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz +
               ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
    }

    // AABB::rayIntersect of the density box (aabb.h:308-338)
    bool clip(const Vec3 &o, const Vec3 &dir, Float &nearT, Float &farT) const {
        nearT = -std::numeric_limits<Float>::infinity();
        farT = std::numeric_limits<Float>::infinity();
        for (int i = 0; i < 3; i++) {
            const Float origin = o[i], minVal = bmin[i], maxVal = bmax[i];
            if (dir[i] == 0) {
                if (origin < minVal || origin > maxVal) return false;
            } else {
                const Float rcp = 1.0f / dir[i];
                Float t1 = (minVal - origin) * rcp, t2 = (maxVal - origin) * rcp;
                if (t1 > t2) std::swap(t1, t2);
                nearT = std::max(t1, nearT);
                farT = std::min(t2, farT);
                if (!(nearT <= farT)) return false;
            }
        }
        return true;
    }

    // heterogeneous.cpp:301-376: composite Simpson quadrature of the density along [rmint, rmaxt]
    Float integrateDensity(const Vec3 &o, const Vec3 &dir, Float rmint, Float rmaxt) const {
        Float mint, maxt;
        if (!clip(o, dir, mint, maxt)) return 0.0f;
        mint = std::max(mint, rmint);
        maxt = std::min(maxt, rmaxt);
        Float length = maxt - mint, maxComp = 0;
        Vec3 p = o + dir * mint, pLast = o + dir * maxt;
        for (int i = 0; i < 3; ++i) maxComp = std::max(std::max(maxComp, std::abs(p[i])), std::abs(pLast[i]));
        if (length < 1e-6f * maxComp) return 0.0f;
        uint32_t nSteps = (uint32_t)std::ceil(length / stepSize);
        nSteps += nSteps % 2;
        const Float step = length / nSteps;
        const Vec3 increment = dir * step;
        Float integrated = lookup(p) + lookup(pLast);
        const Float stopAfterDensity = -std::log(Epsilon);
        const Float stopValue = stopAfterDensity * 3.0f / (step * d.scale);
        p += increment;
        Float m = 4;
        for (uint32_t i = 1; i < nSteps; ++i) {
            integrated += m * lookup(p);
            m = 6 - m;
            if (integrated > stopValue) return std::numeric_limits<Float>::infinity();  // HETVOL_EARLY_EXIT
            Vec3 next = p + increment;
            if (p.x == next.x && p.y == next.y && p.z == next.z) break;
            p = next;
        }
        return integrated * d.scale * step * (1.0f / 3.0f);
    }

    // heterogeneous.cpp:420-544: solve int_{mint}^t density = desiredDensity (Simpson segments + Newton-bisection)
    bool invertDensityIntegral(const Vec3 &o, const Vec3 &dir, Float rmint, Float rmaxt, Float desiredDensity, Float &integratedDensity,
                               Float &t, Float &densityAtMinT, Float &densityAtT) const {
        integratedDensity = densityAtMinT = densityAtT = 0.0f;
        Float mint, maxt;
        if (!clip(o, dir, mint, maxt)) return false;
        mint = std::max(mint, rmint);
        maxt = std::min(maxt, rmaxt);
        Float length = maxt - mint, maxComp = 0;
        Vec3 p = o + dir * mint, pLast = o + dir * maxt;
        for (int i = 0; i < 3; ++i) maxComp = std::max(std::max(maxComp, std::abs(p[i])), std::abs(pLast[i]));
        if (length < 1e-6f * maxComp) return false;
        uint32_t nSteps = (uint32_t)std::ceil(length / (2 * stepSize));
        Float step = length / nSteps, multiplier = (1.0f / 6.0f) * step * d.scale;
        Vec3 fullStep = dir * step, halfStep = fullStep * 0.5f;
        Float node1 = lookup(p);
        densityAtMinT = rmint == mint ? node1 * d.scale : 0.0f;
        for (uint32_t i = 0; i < nSteps; ++i) {
            Float node2 = lookup(p + halfStep), node3 = lookup(p + fullStep);
            Float newDensity = integratedDensity + multiplier * (node1 + node2 * 4 + node3);
            if (newDensity >= desiredDensity) {
                Float a = 0, b = step, x = a, fx = integratedDensity - desiredDensity, stepSqr = step * step, temp = d.scale / stepSqr;
                int it = 1;
                while (true) {
                    Float dfx = temp * (node1 * stepSqr - (3 * node1 - 4 * node2 + node3) * step * x + 2 * (node1 - 2 * node2 + node3) * x * x);
                    x -= fx / dfx;
                    if (x <= a || x >= b || dfx == 0) x = 0.5f * (b + a);
                    Float intval = integratedDensity + temp * (1.0f / 6.0f) *
                                   (x * (6 * node1 * stepSqr - 3 * (3 * node1 - 4 * node2 + node3) * step * x +
                                         4 * (node1 - 2 * node2 + node3) * x * x));
                    fx = intval - desiredDensity;
                    if (std::abs(fx) < 1e-6f) {
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
            Vec3 next = p + fullStep;
            if (p.x == next.x && p.y == next.y && p.z == next.z) break;
            integratedDensity = newDensity;
            node1 = node3;
            p = next;
        }
        return false;
    }

    // heterogeneous.cpp:589-663. Ray interval [rmint, rmaxt].
    bool sampleDistance(const Vec3 &o, const Vec3 &dir, Float rmint, Float rmaxt, MediumSample &mRec, Rng &rng) const {
        if (d.method == B200PG_MEDIUM_SIMPSON) {  // (:594-611)
            Float integratedDensity, densityAtMinT, densityAtT;
            bool success = false;
            const Float desiredDensity = -std::log(1 - rng.next1D());
            if (invertDensityIntegral(o, dir, rmint, rmaxt, desiredDensity, integratedDensity, mRec.t, densityAtMinT, densityAtT)) {
                mRec.p = o + dir * mRec.t;
                success = true;
                mRec.sigmaS = Vec3(d.albedo[0], d.albedo[1], d.albedo[2]) * densityAtT;
            }
            const Float expVal = std::exp(-integratedDensity);
            mRec.pdfFailure = expVal;
            mRec.pdfSuccess = expVal * densityAtT;
            mRec.transmittance = Vec3(expVal);
            return success && mRec.pdfSuccess > 0;  // (:662)
        }
        mRec.pdfFailure = 1.0f;
        mRec.pdfSuccess = 1.0f;
        mRec.transmittance = Vec3(1.0f);
        Float mint, maxt;
        if (!clip(o, dir, mint, maxt)) return false;
        mint = std::max(mint, rmint);
        maxt = std::min(maxt, rmaxt);
        Float t = mint, densityAtT = 0;
        while (true) {
            t -= std::log(1 - rng.next1D()) * invMaxDensity;
            if (t >= maxt) break;
            Vec3 p = o + dir * t;
            densityAtT = lookup(p) * d.scale;
            if (densityAtT * invMaxDensity > rng.next1D()) {
                mRec.t = t;
                mRec.p = p;
                Vec3 albedo(d.albedo[0], d.albedo[1], d.albedo[2]);
                mRec.sigmaS = albedo * densityAtT;
                Float tr = densityAtT != 0.0f ? 1.0f / densityAtT : 0;
                if (!std::isfinite(tr)) tr = 0;
                mRec.transmittance = Vec3(tr);
                return true;  // pdfSuccess == 1 > 0
            }
        }
        return false;
    }

    // Guided free-flight sampling (this repo's design, DESIGN.md "Guiding in media"; no counterpart in the reference
    // snapshot): Woodcock tracking whose real-vs-null decision at every tentative collision is steered by the guiding
    // field and compensated by weights ("weighted delta tracking"). With a = sigma_t(x) / sigma_max the analog
    // probability of a real collision, and g = 4 pi * (mixture pdf of the cell at x in the ray direction) the ratio of
    // the radiance arriving along the ray to the mean incident radiance,
    //     P_guided = a * albedo / (a * albedo + (1 - a) * g),     P_real = (1 - beta) * a + beta * P_guided,
    // real collision: weight *= albedo * a / P_real; null collision: weight *= (1 - a) / (1 - P_real).
    // beta = 1/2 bounds both factors by 2 (and by 2 * albedo). `pdfDir(p)` returns the mixture pdf at p in direction dir.
    template <typename PdfDir>
    bool sampleDistanceGuided(const Vec3 &o, const Vec3 &dir, Float rmint, Float rmaxt, MediumSample &mRec, Vec3 &weight, Rng &rng,
                              const PdfDir &pdfDir) const {
        weight = Vec3(1.0f);
        Float mint, maxt;
        if (!clip(o, dir, mint, maxt)) return false;
        mint = std::max(mint, rmint);
        maxt = std::min(maxt, rmaxt);
        const Vec3 albedo(d.albedo[0], d.albedo[1], d.albedo[2]);
        const Float albedoAvg = (albedo.x + albedo.y + albedo.z) * (1.0f / 3.0f);
        const Float beta = 0.5f;
        Float t = mint;
        while (true) {
            t -= std::log(1 - rng.next1D()) * invMaxDensity;
            if (t >= maxt) break;
            const Vec3 p = o + dir * t;
            const Float a = std::min(lookup(p) * d.scale * invMaxDensity, 1.0f);
            const Float u = rng.next1D();
            if (!(a > 0)) continue;  // empty space: a null collision with probability one
            const Float g = 4 * PI_F * pdfDir(p);
            const Float num = a * albedoAvg, den = num + (1 - a) * g;
            const Float pGuided = den > 0 ? num / den : a;
            const Float pReal = (1 - beta) * a + beta * pGuided;
            if (u < pReal) {
                weight *= albedo * (a / pReal);
                mRec.t = t;
                mRec.p = p;
                return true;
            }
            weight *= (1 - a) / (1 - pReal);
        }
        return false;
    }

    // heterogeneous.cpp:546-587, Woodcock branch: 2 ratio-free tracking trials
    Float evalTransmittance(const Vec3 &o, const Vec3 &dir, Float rmint, Float rmaxt, Rng &rng) const {
        if (d.method == B200PG_MEDIUM_SIMPSON) return std::exp(-integrateDensity(o, dir, rmint, rmaxt));  // (:547-548)
        Float mint, maxt;
        if (!clip(o, dir, mint, maxt)) return 1.0f;
        mint = std::max(mint, rmint);
        maxt = std::min(maxt, rmaxt);
        const int nSamples = 2;
        Float result = 0;
        for (int i = 0; i < nSamples; ++i) {
            Float t = mint;
            while (true) {
                t -= std::log(1 - rng.next1D()) * invMaxDensity;
                if (t >= maxt) {
                    result += 1;
                    break;
                }
                Float dens = lookup(o + dir * t) * d.scale;
                if (dens * invMaxDensity > rng.next1D()) break;
            }
        }
        return result / nSamples;
    }

    // phase functions: wi = direction the light comes from (= -ray.d), wo = scattered direction
    Float phaseEval(const Vec3 &wi, const Vec3 &wo) const {
        if (d.phase_type == B200PG_PHASE_HG) {  // hg.cpp:103-106
            const Float g = d.phase_g;
            Float temp = 1.0f + g * g + 2.0f * g * dot(wi, wo);
            return INV_FOURPI * (1 - g * g) / (temp * std::sqrt(temp));
        }
        return INV_FOURPI;  // isotropic.cpp:76-78
    }
    Vec3 phaseSample(const Vec3 &wi, const Vec2 &sample, Float &pdf) const {
        Vec3 wo;
        if (d.phase_type == B200PG_PHASE_HG) {  // hg.cpp:74-95
            const Float g = d.phase_g;
            Float cosTheta;
            if (std::abs(g) < Epsilon) {
                cosTheta = 1 - 2 * sample.x;
            } else {
                Float sqrTerm = (1 - g * g) / (1 - g + 2 * g * sample.x);
                cosTheta = (1 + g * g - sqrTerm * sqrTerm) / (2 * g);
            }
            Float sinTheta = safe_sqrt(1.0f - cosTheta * cosTheta);
            Float phi = 2 * PI_F * sample.y;
            wo = Frame(-wi).toWorld(Vec3(sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta));
        } else {
            wo = squareToUniformSphere(sample);
        }
        pdf = phaseEval(wi, wo);
        return wo;
    }
};

}  // namespace orc

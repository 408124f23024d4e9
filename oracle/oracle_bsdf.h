// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_bsdf.h: the BSDF models on the hot path, restated from
//   src/bsdfs/diffuse.cpp:116-157, dielectric.cpp:235-394, roughconductor.cpp:268-426,
//   roughplastic.cpp:359-534, twosided.cpp:117-195, null.cpp, microfacet.h, rtrans.h.
// Directions are local (wi, wo point away from the surface, bsdf.h:138-150).
#pragma once
#include "oracle_math.h"
#include "../include/b200pg.h"

namespace orc {

// bsdf.h:220-262
enum {
    ENull = 0x00001,
    EDiffuseReflection = 0x00002,
    EGlossyReflection = 0x00008,
    EDeltaReflection = 0x00020,
    EDeltaTransmission = 0x00040,
    EDelta = ENull | EDeltaReflection | EDeltaTransmission,
    ESmooth = EDiffuseReflection | 0x4 | EGlossyReflection | 0x10,
    EFrontSide = 0x08000,
    EBackSide = 0x10000,
    ETransmission = 0x4 | 0x10 | EDeltaTransmission | 0x100 | ENull
};

// microfacet.h:45-721 (Beckmann and GGX; visible-normal sampling, the default)
struct MicrofacetDistribution {
    int type;
    Float alphaU, alphaV;
    MicrofacetDistribution(int type_, Float au, Float av) : type(type_), alphaU(au), alphaV(av) {
        alphaU = std::max(alphaU, 1e-4f);
        alphaV = std::max(alphaV, 1e-4f);
    }
    bool isIsotropic() const { return alphaU == alphaV; }

    Float eval(const Vec3 &m) const {  // microfacet.h:191-234
        if (Frame::cosTheta(m) <= 0) return 0.0f;
        Float cosTheta2 = Frame::cosTheta2(m);
        Float beckmannExponent = ((m.x * m.x) / (alphaU * alphaU) + (m.y * m.y) / (alphaV * alphaV)) / cosTheta2;
        Float result;
        if (type == B200PG_DISTR_BECKMANN) {
            result = std::exp(-beckmannExponent) / (PI_F * alphaU * alphaV * cosTheta2 * cosTheta2);
        } else {
            Float root = (1.0f + beckmannExponent) * cosTheta2;
            result = 1.0f / (PI_F * alphaU * alphaV * root * root);
        }
        if (result * Frame::cosTheta(m) < 1e-20f) result = 0;
        return result;
    }

    Float projectRoughness(const Vec3 &v) const {  // microfacet.h:533-544
        Float invSinTheta2 = 1 / Frame::sinTheta2(v);
        if (isIsotropic() || invSinTheta2 <= 0) return alphaU;
        Float cosPhi2 = v.x * v.x * invSinTheta2;
        Float sinPhi2 = v.y * v.y * invSinTheta2;
        return std::sqrt(cosPhi2 * alphaU * alphaU + sinPhi2 * alphaV * alphaV);
    }

    Float smithG1(const Vec3 &v, const Vec3 &m) const {  // microfacet.h:477-514
        if (dot(v, m) * Frame::cosTheta(v) <= 0) return 0.0f;
        Float tanTheta = std::abs(Frame::tanTheta(v));
        if (tanTheta == 0.0f) return 1.0f;
        Float alpha = projectRoughness(v);
        if (type == B200PG_DISTR_BECKMANN) {
            Float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            Float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        } else {
            Float root = alpha * tanTheta;
            return 2.0f / (1.0f + hypot2(1.0f, root));
        }
    }
    Float G(const Vec3 &wi, const Vec3 &wo, const Vec3 &m) const { return smithG1(wi, m) * smithG1(wo, m); }

    Vec2 sampleVisible11(Float thetaI, Vec2 sample) const {  // microfacet.h:573-697
        const Float SQRT_PI_INV = 1 / std::sqrt(PI_F);
        Vec2 slope;
        if (type == B200PG_DISTR_BECKMANN) {
            if (thetaI < 1e-4f) {
                Float r = std::sqrt(-std::log(1.0f - sample.x));
                Float phi = 2 * PI_F * sample.y;
                return Vec2(r * std::cos(phi), r * std::sin(phi));
            }
            Float tanThetaI = std::tan(thetaI);
            Float cotThetaI = 1 / tanThetaI;
            Float a = -1, c = mts_erf(cotThetaI);
            Float sample_x = std::max(sample.x, 1e-6f);
            Float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            Float b = c - (1 + c) * std::pow(1 - sample_x, fit);
            Float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * std::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                Float invErf = mts_erfinv(b);
                Float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * std::exp(-invErf * invErf)) - sample_x;
                Float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::abs(value) < 1e-5f) break;
                if (value > 0)
                    c = b;
                else
                    a = b;
                b -= value / derivative;
            }
            slope.x = mts_erfinv(b);
            slope.y = mts_erfinv(2.0f * std::max(sample.y, 1e-6f) - 1.0f);
        } else {
            if (thetaI < 1e-4f) {
                Float r = safe_sqrt(sample.x / (1 - sample.x));
                Float phi = 2 * PI_F * sample.y;
                return Vec2(r * std::cos(phi), r * std::sin(phi));
            }
            Float tanThetaI = std::tan(thetaI);
            Float a = 1 / tanThetaI;
            Float G1 = 2.0f / (1.0f + safe_sqrt(1.0f + 1.0f / (a * a)));
            Float A = 2.0f * sample.x / G1 - 1.0f;
            if (std::abs(A) == 1) A -= signum(A) * Epsilon;
            Float tmp = 1.0f / (A * A - 1.0f);
            Float B = tanThetaI;
            Float D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            Float slope_x_1 = B * tmp - D;
            Float slope_x_2 = B * tmp + D;
            slope.x = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
            Float S;
            if (sample.y > 0.5f) {
                S = 1.0f;
                sample.y = 2.0f * (sample.y - 0.5f);
            } else {
                S = -1.0f;
                sample.y = 2.0f * (0.5f - sample.y);
            }
            Float z = (sample.y * (sample.y * (sample.y * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) +
                       0.000152998850436920f) /
                      (sample.y * (sample.y * (sample.y * (sample.y * 0.169507819808272f - 0.397203533833404f) -
                                               0.232500544458471f) + 1.0f) - 0.539825872510702f);
            slope.y = S * z * std::sqrt(1.0f + slope.x * slope.x);
        }
        return slope;
    }

    Vec3 sampleVisible(const Vec3 &_wi, const Vec2 &sample) const {  // microfacet.h:421-459
        Vec3 wi = normalize(Vec3(alphaU * _wi.x, alphaV * _wi.y, _wi.z));
        Float theta = 0, phi = 0;
        if (wi.z < 0.99999f) {
            theta = std::acos(wi.z);
            phi = std::atan2(wi.y, wi.x);
        }
        Float sinPhi = std::sin(phi), cosPhi = std::cos(phi);
        Vec2 slope = sampleVisible11(theta, sample);
        slope = Vec2(cosPhi * slope.x - sinPhi * slope.y, sinPhi * slope.x + cosPhi * slope.y);
        slope.x *= alphaU;
        slope.y *= alphaV;
        Float normalization = 1.0f / std::sqrt(slope.x * slope.x + slope.y * slope.y + 1.0f);
        return Vec3(-slope.x * normalization, -slope.y * normalization, normalization);
    }
    Float pdfVisible(const Vec3 &wi, const Vec3 &m) const {  // microfacet.h:462-466
        if (Frame::cosTheta(wi) == 0) return 0.0f;
        return smithG1(wi, m) * absDot(wi, m) * eval(m) / std::abs(Frame::cosTheta(wi));
    }
    Vec3 sample(const Vec3 &wi, const Vec2 &s, Float &pdf) const {
        Vec3 m = sampleVisible(wi, s);
        pdf = pdfVisible(wi, m);
        return m;
    }
};

inline Vec3 reflectLocal(const Vec3 &wi) { return Vec3(-wi.x, -wi.y, wi.z); }
inline Vec3 reflectAbout(const Vec3 &wi, const Vec3 &m) { return m * (2 * dot(wi, m)) - wi; }

struct Bsdf {
    B200pgBsdf d;
    Float eta, invEta;               // dielectric / plastic
    Float specularSamplingWeight;    // roughplastic.cpp:283-286
    Float invEta2;

    void configure() {
        eta = d.int_ior / d.ext_ior;
        invEta = 1 / eta;
        invEta2 = 1.0f / (eta * eta);
        Vec3 dr(d.reflectance[0], d.reflectance[1], d.reflectance[2]);
        Vec3 sr(d.specular_reflectance[0], d.specular_reflectance[1], d.specular_reflectance[2]);
        Float dAvg = dr.luminance(), sAvg = sr.luminance();
        specularSamplingWeight = sAvg / (dAvg + sAvg);
    }
    Vec3 R() const { return Vec3(d.reflectance[0], d.reflectance[1], d.reflectance[2]); }
    Vec3 SR() const { return Vec3(d.specular_reflectance[0], d.specular_reflectance[1], d.specular_reflectance[2]); }
    Vec3 ST() const {
        return Vec3(d.specular_transmittance[0], d.specular_transmittance[1], d.specular_transmittance[2]);
    }

    unsigned typeFlags() const {
        unsigned t = 0;
        switch (d.type) {
            case B200PG_BSDF_DIFFUSE: t = EDiffuseReflection | EFrontSide; break;
            case B200PG_BSDF_DIELECTRIC: t = EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide; break;
            case B200PG_BSDF_ROUGHCONDUCTOR: t = EGlossyReflection | EFrontSide; break;
            case B200PG_BSDF_ROUGHPLASTIC: t = EGlossyReflection | EDiffuseReflection | EFrontSide; break;
            case B200PG_BSDF_NULL: t = ENull | EFrontSide | EBackSide; break;
        }
        if (d.twosided) t |= EBackSide;  // twosided.cpp:72-84
        return t;
    }

    // rtrans.h:183-198 with alpha and eta fixed (1-D table)
    Float rtransExt(Float cosTheta) const {
        Float warpedCosTheta = std::pow(std::abs(cosTheta), 0.25f);
        if (!(cosTheta >= 0)) return 0.f;
        Float result = evalCubicInterp1D(warpedCosTheta, d.rt_ext_trans, 100, 0.0f, 1.0f);
        return std::min(1.0f, std::max(0.0f, result));
    }

    // ---- single-sided models ----
    Vec3 evalInner(const Vec3 &wi, const Vec3 &wo) const {
        switch (d.type) {
            case B200PG_BSDF_DIFFUSE: {  // diffuse.cpp:116-124
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return Vec3(0.0f);
                return R() * (INV_PI * Frame::cosTheta(wo));
            }
            case B200PG_BSDF_ROUGHCONDUCTOR: {  // roughconductor.cpp:268-306
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return Vec3(0.0f);
                Vec3 H = normalize(wo + wi);
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_v);
                const Float D = distr.eval(H);
                if (D == 0) return Vec3(0.0f);
                const Vec3 F = fresnelConductorExact(dot(wi, H), Vec3(d.eta[0], d.eta[1], d.eta[2]),
                                                     Vec3(d.k[0], d.k[1], d.k[2])) * SR();
                const Float G = distr.G(wi, wo, H);
                Float model = D * G / (4.0f * Frame::cosTheta(wi));
                return F * model;
            }
            case B200PG_BSDF_ROUGHPLASTIC: {  // roughplastic.cpp:359-413
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return Vec3(0.0f);
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_u);
                Vec3 result(0.0f);
                {
                    const Vec3 H = normalize(wo + wi);
                    const Float D = distr.eval(H);
                    const Float F = fresnelDielectricExt(dot(wi, H), eta);
                    const Float G = distr.G(wi, wo, H);
                    Float value = F * D * G / (4.0f * Frame::cosTheta(wi));
                    result += SR() * value;
                }
                {
                    Vec3 diff = R();
                    Float T12 = rtransExt(Frame::cosTheta(wi));
                    Float T21 = rtransExt(Frame::cosTheta(wo));
                    Float Fdr = 1 - d.rt_int_diff;
                    if (d.nonlinear)
                        diff = diff / (Vec3(1.0f) - diff * Fdr);
                    else
                        diff = diff / (1 - Fdr);
                    result += diff * (INV_PI * Frame::cosTheta(wo) * T12 * T21 * invEta2);
                }
                return result;
            }
            default: return Vec3(0.0f);  // delta / null: zero w.r.t. solid angle
        }
    }

    Float plasticProbSpecular(const Vec3 &wi) const {  // roughplastic.cpp:437-445
        Float probSpecular = 1 - rtransExt(Frame::cosTheta(wi));
        probSpecular = (probSpecular * specularSamplingWeight) /
                       (probSpecular * specularSamplingWeight + (1 - probSpecular) * (1 - specularSamplingWeight));
        return probSpecular;
    }

    Float pdfInner(const Vec3 &wi, const Vec3 &wo) const {
        switch (d.type) {
            case B200PG_BSDF_DIFFUSE:  // diffuse.cpp:126-133
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return 0.0f;
                return squareToCosineHemispherePdf(wo);
            case B200PG_BSDF_ROUGHCONDUCTOR: {  // roughconductor.cpp:308-334
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return 0.0f;
                Vec3 H = normalize(wo + wi);
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_v);
                return distr.eval(H) * distr.smithG1(wi, H) / (4.0f * Frame::cosTheta(wi));
            }
            case B200PG_BSDF_ROUGHPLASTIC: {  // roughplastic.cpp:415-466
                if (Frame::cosTheta(wi) <= 0 || Frame::cosTheta(wo) <= 0) return 0.0f;
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_u);
                const Vec3 H = normalize(wo + wi);
                Float probSpecular = plasticProbSpecular(wi);
                Float probDiffuse = 1 - probSpecular;
                const Float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
                const Float prob = distr.pdfVisible(wi, H);
                Float result = prob * dwh_dwo * probSpecular;
                result += probDiffuse * squareToCosineHemispherePdf(wo);
                return result;
            }
            default: return 0.0f;
        }
    }

    // returns weight = f*cos/pdf; fills wo, pdf, eta, sampledType
    Vec3 sampleInner(const Vec3 &wi, const Vec2 &sample_, Vec3 &wo, Float &pdf, Float &outEta, unsigned &sampledType) const {
        outEta = 1.0f;
        sampledType = 0;
        pdf = 0;
        switch (d.type) {
            case B200PG_BSDF_DIFFUSE: {  // diffuse.cpp:147-157
                if (Frame::cosTheta(wi) <= 0) return Vec3(0.0f);
                wo = squareToCosineHemisphere(sample_);
                sampledType = EDiffuseReflection;
                pdf = squareToCosineHemispherePdf(wo);
                return R();
            }
            case B200PG_BSDF_DIELECTRIC: {  // dielectric.cpp:289-340
                Float cosThetaT;
                Float F = fresnelDielectricExt(Frame::cosTheta(wi), cosThetaT, eta);
                if (sample_.x <= F) {
                    sampledType = EDeltaReflection;
                    wo = reflectLocal(wi);
                    outEta = 1.0f;
                    pdf = F;
                    return SR();
                } else {
                    sampledType = EDeltaTransmission;
                    Float scale = -(cosThetaT < 0 ? invEta : eta);  // dielectric.cpp:222-225
                    wo = Vec3(scale * wi.x, scale * wi.y, cosThetaT);
                    outEta = cosThetaT < 0 ? eta : invEta;
                    pdf = 1 - F;
                    Float factor = cosThetaT < 0 ? invEta : eta;  // ERadiance mode
                    return ST() * (factor * factor);
                }
            }
            case B200PG_BSDF_ROUGHCONDUCTOR: {  // roughconductor.cpp:383-426
                if (Frame::cosTheta(wi) < 0) return Vec3(0.0f);
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_v);
                Vec3 m = distr.sample(wi, sample_, pdf);
                if (pdf == 0) return Vec3(0.0f);
                wo = reflectAbout(wi, m);
                sampledType = EGlossyReflection;
                if (Frame::cosTheta(wo) <= 0) return Vec3(0.0f);
                Vec3 F = fresnelConductorExact(dot(wi, m), Vec3(d.eta[0], d.eta[1], d.eta[2]),
                                               Vec3(d.k[0], d.k[1], d.k[2])) * SR();
                Float weight = distr.smithG1(wo, m);
                pdf /= 4.0f * dot(wo, m);
                return F * weight;
            }
            case B200PG_BSDF_ROUGHPLASTIC: {  // roughplastic.cpp:468-534
                if (Frame::cosTheta(wi) <= 0) return Vec3(0.0f);
                bool choseSpecular = true;
                Vec2 sample(sample_);
                MicrofacetDistribution distr(d.distribution, d.alpha_u, d.alpha_u);
                Float probSpecular = plasticProbSpecular(wi);
                if (sample.y < probSpecular) {
                    sample.y /= probSpecular;
                } else {
                    sample.y = (sample.y - probSpecular) / (1 - probSpecular);
                    choseSpecular = false;
                }
                if (choseSpecular) {
                    Vec3 m = distr.sampleVisible(wi, sample);
                    wo = reflectAbout(wi, m);
                    sampledType = EGlossyReflection;
                    if (Frame::cosTheta(wo) <= 0) return Vec3(0.0f);
                } else {
                    sampledType = EDiffuseReflection;
                    wo = squareToCosineHemisphere(sample);
                }
                pdf = pdfInner(wi, wo);
                if (pdf == 0) return Vec3(0.0f);
                return evalInner(wi, wo) / pdf;
            }
            case B200PG_BSDF_NULL: {  // null.cpp
                wo = -wi;
                sampledType = ENull;
                pdf = 1;
                return Vec3(1.0f);
            }
        }
        return Vec3(0.0f);
    }

    // ---- twosided adapter (twosided.cpp:117-195) ----
    Vec3 eval(Vec3 wi, Vec3 wo) const {
        if (d.twosided && !(Frame::cosTheta(wi) > 0)) {
            wi.z *= -1;
            wo.z *= -1;
        }
        return evalInner(wi, wo);
    }
    Float pdf(Vec3 wi, Vec3 wo) const {
        if (d.twosided && !(wi.z > 0)) {
            wi.z *= -1;
            wo.z *= -1;
        }
        return pdfInner(wi, wo);
    }
    Vec3 sample(Vec3 wi, const Vec2 &s, Vec3 &wo, Float &pdf, Float &outEta, unsigned &sampledType) const {
        bool flipped = false;
        if (d.twosided && Frame::cosTheta(wi) < 0) {
            wi.z *= -1;
            flipped = true;
        }
        Vec3 result = sampleInner(wi, s, wo, pdf, outEta, sampledType);
        if (flipped && !result.isZero() && pdf != 0) wo.z *= -1;
        return result;
    }
};

}  // namespace orc

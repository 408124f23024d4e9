// device_bsdf.cuh -- BSDF evaluation / pdf / sampling for the shade stage.
// Models (reference plugin semantics, local frame, wi/wo pointing away from the surface):
//   diffuse        src/bsdfs/diffuse.cpp:116-157
//   dielectric     src/bsdfs/dielectric.cpp:235-394      (delta lobes, never guided)
//   roughconductor src/bsdfs/roughconductor.cpp:268-426  (visible-normal sampling)
//   roughplastic   src/bsdfs/roughplastic.cpp:359-534    (tabulated rough transmittance, rtrans.h)
//   twosided       src/bsdfs/twosided.cpp:117-195        (flag on the record)
//   null           src/bsdfs/null.cpp                    (index-matched medium boundary)
#pragma once
#include "device_math.cuh"
#include "pg_types.h"

namespace pg {

// BSDF::EBSDFType bits (include/mitsuba/render/bsdf.h:220-262)
enum : uint32_t {
    kNull = 0x00001,
    kDiffuseReflection = 0x00002,
    kGlossyReflection = 0x00008,
    kDeltaReflection = 0x00020,
    kDeltaTransmission = 0x00040,
    kDelta = kNull | kDeltaReflection | kDeltaTransmission,
    kSmooth = 0x00002 | 0x00004 | 0x00008 | 0x00010,
    kFrontSide = 0x08000,
    kBackSide = 0x10000,
    kTransmission = 0x00004 | 0x00010 | 0x00040 | 0x00100 | 0x00001
};

// ---- Fresnel (src/libcore/util.cpp:653-683, 741-763)
PG_DEV float fresnelDielectricExt(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1) {
        cosThetaT_ = -cosThetaI_;
        return 0.0f;
    }
    float scale = (cosThetaI_ > 0) ? 1 / eta : eta;
    float cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) {
        cosThetaT_ = 0.0f;
        return 1.0f;
    }
    float cosThetaI = fabsf(cosThetaI_);
    float cosThetaT = sqrtf(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
PG_DEV float fresnelConductorExact1(float cosThetaI, float eta, float k) {
    float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    float temp1 = eta * eta - k * k - sinThetaI2;
    float a2pb2 = safeSqrt(temp1 * temp1 + k * k * eta * eta * 4);
    float a = safeSqrt((a2pb2 + temp1) * 0.5f);
    float term1 = a2pb2 + cosThetaI2, term2 = a * (2 * cosThetaI);
    float Rs2 = (term1 - term2) / (term1 + term2);
    float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}

// ---- Mitsuba's erf / erfinv approximations (src/libcore/math.cpp:25-72)
PG_DEV float mtsErfinv(float x) {
    float w = -logf((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = 3.43273939e-07f + p * w;
        p = -3.5233877e-06f + p * w;
        p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w;
        p = -0.00125372503f + p * w;
        p = -0.00417768164f + p * w;
        p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = 0.000100950558f + p * w;
        p = 0.00134934322f + p * w;
        p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w;
        p = -0.0076224613f + p * w;
        p = 0.00943887047f + p * w;
        p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}
PG_DEV float mtsErf(float x) {
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f,
                p = 0.3275911f;
    float sign = signum(x);
    x = fabsf(x);
    float t = 1.0f / (1.0f + p * x);
    float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-x * x);
    return sign * y;
}
PG_DEV float hypot2(float a, float b) {  // math.cpp:74-88
    float r;
    if (fabsf(a) > fabsf(b)) {
        r = b / a;
        r = fabsf(a) * sqrtf(1.0f + r * r);
    } else if (b != 0.0f) {
        r = a / b;
        r = fabsf(b) * sqrtf(1.0f + r * r);
    } else {
        r = 0.0f;
    }
    return r;
}

// ---- microfacet distribution (src/bsdfs/microfacet.h), Beckmann + GGX
#ifndef PG_MF_NOINLINE
#define PG_MF_NOINLINE 0
#endif
#if PG_MF_NOINLINE
#define PG_MF __device__ __noinline__
#else
#define PG_MF PG_DEV
#endif
struct Microfacet {
    int type;
    float au, av;
    PG_DEV Microfacet(int t, float a, float b) : type(t), au(fmaxf(a, 1e-4f)), av(fmaxf(b, 1e-4f)) {}

    PG_MF float eval(float3 m) const {  // microfacet.h:191-234
        if (m.z <= 0) return 0.0f;
        float cosTheta2 = m.z * m.z;
        float beckmannExponent = ((m.x * m.x) / (au * au) + (m.y * m.y) / (av * av)) / cosTheta2;
        float result;
        if (type == 0) {
            result = expf(-beckmannExponent) / (kPi * au * av * cosTheta2 * cosTheta2);
        } else {
            float root = (1.0f + beckmannExponent) * cosTheta2;
            result = 1.0f / (kPi * au * av * root * root);
        }
        if (result * m.z < 1e-20f) result = 0;
        return result;
    }
    PG_DEV float projectRoughness(float3 v) const {  // microfacet.h:533-544
        float invSinTheta2 = 1 / (1.0f - v.z * v.z);
        if (au == av || invSinTheta2 <= 0) return au;
        float cosPhi2 = v.x * v.x * invSinTheta2;
        float sinPhi2 = v.y * v.y * invSinTheta2;
        return sqrtf(cosPhi2 * au * au + sinPhi2 * av * av);
    }
    PG_MF float smithG1(float3 v, float3 m) const {  // microfacet.h:477-514
        if (dot(v, m) * v.z <= 0) return 0.0f;
        float temp = 1 - v.z * v.z;  // Frame::tanTheta
        float tanTheta = temp <= 0.0f ? 0.0f : fabsf(sqrtf(temp) / v.z);
        if (tanTheta == 0.0f) return 1.0f;
        float alpha = projectRoughness(v);
        if (type == 0) {
            float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        } else {
            float root = alpha * tanTheta;
            return 2.0f / (1.0f + hypot2(1.0f, root));
        }
    }
    PG_DEV float G(float3 wi, float3 wo, float3 m) const { return smithG1(wi, m) * smithG1(wo, m); }

    PG_MF float2 sampleVisible11(float thetaI, float2 sample) const {  // microfacet.h:573-697
        const float SQRT_PI_INV = 1 / sqrtf(kPi);
        float2 slope;
        if (type == 0) {
            if (thetaI < 1e-4f) {
                float r = sqrtf(-logf(1.0f - sample.x));
                float s, c;
                sincosf(2 * kPi * sample.y, &s, &c);
                return make_float2(r * c, r * s);
            }
            float tanThetaI = tanf(thetaI);
            float cotThetaI = 1 / tanThetaI;
            float a = -1, c = mtsErf(cotThetaI);
            float sample_x = fmaxf(sample.x, 1e-6f);
            float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * powf(1 - sample_x, fit);
            float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * expf(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                float invErf = mtsErfinv(b);
                float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * expf(-invErf * invErf)) - sample_x;
                float derivative = normalization * (1 - invErf * tanThetaI);
                if (fabsf(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slope.x = mtsErfinv(b);
            slope.y = mtsErfinv(2.0f * fmaxf(sample.y, 1e-6f) - 1.0f);
        } else {
            if (thetaI < 1e-4f) {
                float r = safeSqrt(sample.x / (1 - sample.x));
                float s, c;
                sincosf(2 * kPi * sample.y, &s, &c);
                return make_float2(r * c, r * s);
            }
            float tanThetaI = tanf(thetaI);
            float a = 1 / tanThetaI;
            float G1 = 2.0f / (1.0f + safeSqrt(1.0f + 1.0f / (a * a)));
            float A = 2.0f * sample.x / G1 - 1.0f;
            if (fabsf(A) == 1) A -= signum(A) * kEpsilon;
            float tmp = 1.0f / (A * A - 1.0f);
            float B = tanThetaI;
            float D = safeSqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            float slope_x_1 = B * tmp - D;
            float slope_x_2 = B * tmp + D;
            slope.x = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
            float S;
            if (sample.y > 0.5f) {
                S = 1.0f;
                sample.y = 2.0f * (sample.y - 0.5f);
            } else {
                S = -1.0f;
                sample.y = 2.0f * (0.5f - sample.y);
            }
            float z = (sample.y * (sample.y * (sample.y * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) +
                       0.000152998850436920f) /
                      (sample.y * (sample.y * (sample.y * (sample.y * 0.169507819808272f - 0.397203533833404f) -
                                               0.232500544458471f) + 1.0f) - 0.539825872510702f);
            slope.y = S * z * sqrtf(1.0f + slope.x * slope.x);
        }
        return slope;
    }
    PG_DEV float3 sampleVisible(float3 _wi, float2 sample) const {  // microfacet.h:421-459
        float3 wi = normalize(f3(au * _wi.x, av * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) {
            theta = acosf(wi.z);
            phi = atan2f(wi.y, wi.x);
        }
        float sinPhi, cosPhi;
        sincosf(phi, &sinPhi, &cosPhi);
        float2 slope = sampleVisible11(theta, sample);
        slope = make_float2(cosPhi * slope.x - sinPhi * slope.y, sinPhi * slope.x + cosPhi * slope.y);
        slope.x *= au;
        slope.y *= av;
        float normalization = 1.0f / sqrtf(slope.x * slope.x + slope.y * slope.y + 1.0f);
        return f3(-slope.x * normalization, -slope.y * normalization, normalization);
    }
    PG_DEV float pdfVisible(float3 wi, float3 m) const {  // microfacet.h:462-466
        if (wi.z == 0) return 0.0f;
        return smithG1(wi, m) * fabsf(dot(wi, m)) * eval(m) / fabsf(wi.z);
    }
};

// Catmull-Rom lookup of the reduced 100-entry rough-transmittance table
// (RoughTransmittance::eval with alpha and eta fixed, rtrans.h:183-198; spline.cpp:23-60)
PG_DEV float rtransExt(const BsdfRecord &b, float cosTheta) {
    if (!(cosTheta >= 0)) return 0.f;
    float x = powf(fabsf(cosTheta), 0.25f);
    if (!(x >= 0.0f && x <= 1.0f)) return 0.0f;
    const int size = 100;
    float t = (x * (size - 1)) / 1.0f;
    int k = max(0, min((int)t, size - 2));
    const float *v = b.rtExt;
    float f0 = v[k], f1 = v[k + 1];
    float d0 = k > 0 ? 0.5f * (v[k + 1] - v[k - 1]) : v[k + 1] - v[k];
    float d1 = k + 2 < size ? 0.5f * (v[k + 2] - v[k]) : v[k + 1] - v[k];
    t = t - (float)k;
    float t2 = t * t, t3 = t2 * t;
    float result = (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
    return fminf(1.0f, fmaxf(0.0f, result));
}

PG_DEV float plasticProbSpecular(const BsdfRecord &b, float3 wi) {  // roughplastic.cpp:437-445
    float probSpecular = 1 - rtransExt(b, wi.z);
    float w = b.specSamplingWeight;
    return (probSpecular * w) / (probSpecular * w + (1 - probSpecular) * (1 - w));
}

PG_DEV float3 reflectAbout(float3 wi, float3 m) { return m * (2 * dot(wi, m)) - wi; }

// ---- one-sided models
PG_DEV float3 bsdfEvalInner(const BsdfRecord &b, float3 wi, float3 wo) {
    if (b.type == B200PG_BSDF_DIFFUSE) {
        if (wi.z <= 0 || wo.z <= 0) return f3(0.0f);
        return ld3(b.reflectance) * (kInvPi * wo.z);
    } else if (b.type == B200PG_BSDF_ROUGHCONDUCTOR) {
        if (wi.z <= 0 || wo.z <= 0) return f3(0.0f);
        float3 H = normalize(wo + wi);
        Microfacet distr(b.distribution, b.alphaU, b.alphaV);
        float D = distr.eval(H);
        if (D == 0) return f3(0.0f);
        float c = dot(wi, H);
        float3 F = f3(fresnelConductorExact1(c, b.condEta[0], b.condK[0]), fresnelConductorExact1(c, b.condEta[1], b.condK[1]),
                      fresnelConductorExact1(c, b.condEta[2], b.condK[2])) * ld3(b.specRefl);
        float G = distr.G(wi, wo, H);
        float model = D * G / (4.0f * wi.z);
        return F * model;
    } else if (b.type == B200PG_BSDF_ROUGHPLASTIC) {
        if (wi.z <= 0 || wo.z <= 0) return f3(0.0f);
        Microfacet distr(b.distribution, b.alphaU, b.alphaU);
        float3 H = normalize(wo + wi);
        float D = distr.eval(H);
        float ct;
        float F = fresnelDielectricExt(dot(wi, H), ct, b.eta);
        float G = distr.G(wi, wo, H);
        float value = F * D * G / (4.0f * wi.z);
        float3 result = ld3(b.specRefl) * value;
        float3 diff = ld3(b.reflectance);
        float T12 = rtransExt(b, wi.z);
        float T21 = rtransExt(b, wo.z);
        float Fdr = 1 - b.rtIntDiff;
        if (b.nonlinear)
            diff = diff / (f3(1.0f) - diff * Fdr);
        else
            diff = diff / (1 - Fdr);
        result += diff * (kInvPi * wo.z * T12 * T21 * b.invEta2);
        return result;
    }
    return f3(0.0f);
}

PG_DEV float bsdfPdfInner(const BsdfRecord &b, float3 wi, float3 wo) {
    if (b.type == B200PG_BSDF_DIFFUSE) {
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        return kInvPi * wo.z;
    } else if (b.type == B200PG_BSDF_ROUGHCONDUCTOR) {
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        float3 H = normalize(wo + wi);
        Microfacet distr(b.distribution, b.alphaU, b.alphaV);
        return distr.eval(H) * distr.smithG1(wi, H) / (4.0f * wi.z);
    } else if (b.type == B200PG_BSDF_ROUGHPLASTIC) {
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        Microfacet distr(b.distribution, b.alphaU, b.alphaU);
        float3 H = normalize(wo + wi);
        float probSpecular = plasticProbSpecular(b, wi);
        float probDiffuse = 1 - probSpecular;
        float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
        float prob = distr.pdfVisible(wi, H);
        float result = prob * dwh_dwo * probSpecular;
        result += probDiffuse * (kInvPi * wo.z);
        return result;
    }
    return 0.0f;
}

// returns weight = f*cos/pdf (zero = failed sample)
PG_DEV float3 bsdfSampleInner(const BsdfRecord &b, float3 wi, float2 s, float3 &wo, float &pdf, float &outEta,
                              uint32_t &sampledType) {
    outEta = 1.0f;
    sampledType = 0;
    pdf = 0;
    wo = f3(0.0f);
    if (b.type == B200PG_BSDF_DIFFUSE) {
        if (wi.z <= 0) return f3(0.0f);
        wo = squareToCosineHemisphere(s);
        sampledType = kDiffuseReflection;
        pdf = kInvPi * wo.z;
        return ld3(b.reflectance);
    } else if (b.type == B200PG_BSDF_DIELECTRIC) {
        float cosThetaT;
        float F = fresnelDielectricExt(wi.z, cosThetaT, b.eta);
        if (s.x <= F) {
            sampledType = kDeltaReflection;
            wo = f3(-wi.x, -wi.y, wi.z);
            pdf = F;
            return ld3(b.specRefl);
        } else {
            sampledType = kDeltaTransmission;
            float scale = -(cosThetaT < 0 ? b.invEta : b.eta);
            wo = f3(scale * wi.x, scale * wi.y, cosThetaT);
            outEta = cosThetaT < 0 ? b.eta : b.invEta;
            pdf = 1 - F;
            float factor = cosThetaT < 0 ? b.invEta : b.eta;  // ERadiance transport
            return ld3(b.specTrans) * (factor * factor);
        }
    } else if (b.type == B200PG_BSDF_ROUGHCONDUCTOR) {
        if (wi.z < 0) return f3(0.0f);
        Microfacet distr(b.distribution, b.alphaU, b.alphaV);
        float3 m = distr.sampleVisible(wi, s);
        pdf = distr.pdfVisible(wi, m);
        if (pdf == 0) return f3(0.0f);
        wo = reflectAbout(wi, m);
        sampledType = kGlossyReflection;
        if (wo.z <= 0) return f3(0.0f);
        float c = dot(wi, m);
        float3 F = f3(fresnelConductorExact1(c, b.condEta[0], b.condK[0]), fresnelConductorExact1(c, b.condEta[1], b.condK[1]),
                      fresnelConductorExact1(c, b.condEta[2], b.condK[2])) * ld3(b.specRefl);
        float weight = distr.smithG1(wo, m);
        pdf /= 4.0f * dot(wo, m);
        return F * weight;
    } else if (b.type == B200PG_BSDF_ROUGHPLASTIC) {
        if (wi.z <= 0) return f3(0.0f);
        bool choseSpecular = true;
        Microfacet distr(b.distribution, b.alphaU, b.alphaU);
        float probSpecular = plasticProbSpecular(b, wi);
        if (s.y < probSpecular) {
            s.y /= probSpecular;
        } else {
            s.y = (s.y - probSpecular) / (1 - probSpecular);
            choseSpecular = false;
        }
        if (choseSpecular) {
            float3 m = distr.sampleVisible(wi, s);
            wo = reflectAbout(wi, m);
            sampledType = kGlossyReflection;
            if (wo.z <= 0) return f3(0.0f);
        } else {
            sampledType = kDiffuseReflection;
            wo = squareToCosineHemisphere(s);
        }
        pdf = bsdfPdfInner(b, wi, wo);
        if (pdf == 0) return f3(0.0f);
        return bsdfEvalInner(b, wi, wo) / pdf;
    } else {  // null
        wo = -wi;
        sampledType = kNull;
        pdf = 1;
        return f3(1.0f);
    }
}

// ---- twosided adapter
PG_DEV float3 bsdfEval(const BsdfRecord &b, float3 wi, float3 wo) {
    if (b.twosided && !(wi.z > 0)) {
        wi.z = -wi.z;
        wo.z = -wo.z;
    }
    return bsdfEvalInner(b, wi, wo);
}
PG_DEV float bsdfPdf(const BsdfRecord &b, float3 wi, float3 wo) {
    if (b.twosided && !(wi.z > 0)) {
        wi.z = -wi.z;
        wo.z = -wo.z;
    }
    return bsdfPdfInner(b, wi, wo);
}
PG_DEV float3 bsdfSample(const BsdfRecord &b, float3 wi, float2 s, float3 &wo, float &pdf, float &outEta,
                         uint32_t &sampledType) {
    bool flipped = false;
    if (b.twosided && wi.z < 0) {
        wi.z = -wi.z;
        flipped = true;
    }
    float3 r = bsdfSampleInner(b, wi, s, wo, pdf, outEta, sampledType);
    if (flipped && !isZero(r) && pdf != 0) wo.z = -wo.z;
    return r;
}

}  // namespace pg

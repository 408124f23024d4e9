// TEST INFRASTRUCTURE ONLY -- CPU oracle (restatement of the reference's algorithm).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may build, link or execute anything under oracle/. The product path never does.
//
// oracle_math.h: small vector math, frames, warps, Fresnel and special functions,
// restated from the reference (file:line cited per function; paths relative to the
// reference root).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

typedef float Float;

static const Float Epsilon = 1e-4f;        // include/mitsuba/core/constants.h:28
static const Float ShadowEpsilon = 1e-3f;  // constants.h:29
static const Float DeltaEpsilon = 1e-3f;   // constants.h:31
static const Float PI_F = 3.14159265358979323846f;
static const Float INV_PI = 0.31830988618379067154f;
static const Float INV_TWOPI = 0.15915494309189533577f;
static const Float INV_FOURPI = 0.07957747154594766788f;

struct Vec3 {
    Float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    explicit Vec3(Float v) : x(v), y(v), z(v) {}
    Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    Float operator[](int i) const { return (&x)[i]; }
    Float &operator[](int i) { return (&x)[i]; }
    Vec3 operator+(const Vec3 &b) const { return Vec3(x + b.x, y + b.y, z + b.z); }
    Vec3 operator-(const Vec3 &b) const { return Vec3(x - b.x, y - b.y, z - b.z); }
    Vec3 operator*(const Vec3 &b) const { return Vec3(x * b.x, y * b.y, z * b.z); }
    Vec3 operator/(const Vec3 &b) const { return Vec3(x / b.x, y / b.y, z / b.z); }
    Vec3 operator*(Float s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 operator/(Float s) const {
        Float r = 1.0f / s;  // TVector3::operator/ multiplies by the reciprocal (vector.h)
        return Vec3(x * r, y * r, z * r);
    }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    Vec3 &operator+=(const Vec3 &b) { x += b.x; y += b.y; z += b.z; return *this; }
    Vec3 &operator*=(const Vec3 &b) { x *= b.x; y *= b.y; z *= b.z; return *this; }
    Vec3 &operator*=(Float s) { x *= s; y *= s; z *= s; return *this; }
    Vec3 &operator/=(Float s) { Float r = 1.0f / s; x *= r; y *= r; z *= r; return *this; }
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
    Float maxc() const { return std::max(x, std::max(y, z)); }
    Float average() const { return (x + y + z) * (1.0f / 3.0f); }
    // Spectrum::getLuminance, src/libcore/spectrum.cpp:231-236 (Rec.709)
    Float luminance() const { return x * 0.212671f + y * 0.715160f + z * 0.072169f; }
};
inline Vec3 operator*(Float s, const Vec3 &v) { return v * s; }
inline Float dot(const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float absDot(const Vec3 &a, const Vec3 &b) { return std::abs(dot(a, b)); }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Float length(const Vec3 &a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(const Vec3 &a) { return a / length(a); }
inline Float safe_sqrt(Float v) { return std::sqrt(std::max(0.0f, v)); }  // math.h:260-267
inline Float signum(Float v) { return v < 0 ? -1.0f : (v > 0 ? 1.0f : 0.0f); }

struct Vec2 {
    Float x, y;
    Vec2() : x(0), y(0) {}
    Vec2(Float x_, Float y_) : x(x_), y(y_) {}
};

// util.cpp:594-603
inline void coordinateSystem(const Vec3 &a, Vec3 &b, Vec3 &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
        Float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z);
        c = Vec3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        Float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z);
        c = Vec3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// include/mitsuba/core/frame.h
struct Frame {
    Vec3 s, t, n;
    Frame() {}
    explicit Frame(const Vec3 &n_) : n(n_) { coordinateSystem(n, s, t); }
    Frame(const Vec3 &s_, const Vec3 &t_, const Vec3 &n_) : s(s_), t(t_), n(n_) {}
    Vec3 toLocal(const Vec3 &v) const { return Vec3(dot(v, s), dot(v, t), dot(v, n)); }
    Vec3 toWorld(const Vec3 &v) const { return s * v.x + t * v.y + n * v.z; }
    static Float cosTheta(const Vec3 &v) { return v.z; }
    static Float cosTheta2(const Vec3 &v) { return v.z * v.z; }
    static Float sinTheta2(const Vec3 &v) { return 1.0f - v.z * v.z; }
    static Float tanTheta(const Vec3 &v) {
        Float temp = 1 - v.z * v.z;
        if (temp <= 0.0f) return 0.0f;
        return std::sqrt(temp) / v.z;
    }
};

// util.cpp:605-610
inline void computeShadingFrame(const Vec3 &n, const Vec3 &dpdu, Frame &frame) {
    frame.n = n;
    frame.s = normalize(dpdu - frame.n * dot(frame.n, dpdu));
    frame.t = cross(frame.n, frame.s);
}

// 4x4 row-major affine helpers (transform.h)
struct Mat4 {
    Float m[4][4];
    static Mat4 identity() {
        Mat4 r;
        std::memset(r.m, 0, sizeof(r.m));
        for (int i = 0; i < 4; ++i) r.m[i][i] = 1;
        return r;
    }
    static Mat4 fromArray(const float *a) {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = a[i * 4 + j];
        return r;
    }
    Mat4 operator*(const Mat4 &b) const {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                Float s = 0;
                for (int k = 0; k < 4; ++k) s += m[i][k] * b.m[k][j];
                r.m[i][j] = s;
            }
        return r;
    }
    // Transform::operator()(Point): full projective (transform.h:124-139)
    Vec3 point(const Vec3 &p) const {
        Float x = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
        Float y = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
        Float z = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
        Float w = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
        if (w == 1.0f) return Vec3(x, y, z);
        return Vec3(x, y, z) / w;
    }
    Vec3 pointAffine(const Vec3 &p) const {
        return Vec3(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3],
                    m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
                    m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3]);
    }
    Vec3 vector(const Vec3 &v) const {
        return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                    m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
    }
    // Normals transform with the inverse transpose: call on the INVERSE matrix.
    Vec3 normalFromInverse(const Vec3 &v) const {
        return Vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                    m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
    }
    bool invert(Mat4 &out) const {  // Gauss-Jordan with partial pivoting (matrix.inl)
        double a[4][8];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                a[i][j] = m[i][j];
                a[i][j + 4] = (i == j) ? 1.0 : 0.0;
            }
        for (int c = 0; c < 4; ++c) {
            int piv = c;
            for (int r = c + 1; r < 4; ++r)
                if (std::abs(a[r][c]) > std::abs(a[piv][c])) piv = r;
            if (a[piv][c] == 0) return false;
            if (piv != c)
                for (int j = 0; j < 8; ++j) std::swap(a[piv][j], a[c][j]);
            double inv = 1.0 / a[c][c];
            for (int j = 0; j < 8; ++j) a[c][j] *= inv;
            for (int r = 0; r < 4; ++r)
                if (r != c) {
                    double f = a[r][c];
                    if (f != 0)
                        for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
                }
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) out.m[i][j] = (Float)a[i][j + 4];
        return true;
    }
};

// ---------------------------------------------------------------------------
// Counter-based RNG shared (by specification, not by code) with the CUDA path:
// PCG32 (XSH-RR), stream = pixel index, seeded per (seed, sample index).
// The reference uses per-pixel SFMT-19937 samplers (independent.cpp:52-104,
// random.cpp:551-640); that state (2.5 KB/pixel) is not replicated -- see DESIGN.md.
// Consumption ORDER follows the reference: next2D = (next1D, next1D).
// ---------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state, inc;
    void init(uint64_t seed, uint32_t pixel, uint32_t sample) {
        inc = ((uint64_t)pixel << 1) | 1ULL;
        state = 0;
        nextU32();
        state += mix64(seed + (uint64_t)sample * 0x9E3779B97F4A7C15ULL);
        nextU32();
    }
    uint32_t nextU32() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    Float next1D() { return (Float)(nextU32() >> 8) * (1.0f / 16777216.0f); }
    // independent child stream (transmittance estimates draw from a fork so that the main stream's consumption
    // does not depend on how many medium interactions a connection crosses); consumes two words of the parent
    Rng fork() {
        Rng r;
        uint64_t a = nextU32();
        uint64_t b = nextU32();
        r.inc = inc;
        r.state = mix64((a << 32) | b);
        return r;
    }
    Vec2 next2D() {
        Float a = next1D();
        Float b = next1D();
        return Vec2(a, b);
    }
};

// ---------------------------------------------------------------------------
// warp.cpp
// ---------------------------------------------------------------------------
inline Vec2 squareToUniformDiskConcentric(const Vec2 &sample) {  // warp.cpp:79-100
    Float r1 = 2.0f * sample.x - 1.0f;
    Float r2 = 2.0f * sample.y - 1.0f;
    Float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (PI_F / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (PI_F / 2.0f) - (r1 / r2) * (PI_F / 4.0f);
    }
    Float cosPhi = std::cos(phi), sinPhi = std::sin(phi);
    return Vec2(r * cosPhi, r * sinPhi);
}
inline Vec3 squareToCosineHemisphere(const Vec2 &sample) {  // warp.cpp:41-49
    Vec2 p = squareToUniformDiskConcentric(sample);
    Float z = safe_sqrt(1.0f - p.x * p.x - p.y * p.y);
    if (z == 0) z = 1e-10f;
    return Vec3(p.x, p.y, z);
}
inline Float squareToCosineHemispherePdf(const Vec3 &d) { return INV_PI * Frame::cosTheta(d); }  // warp.h
inline Vec2 squareToUniformTriangle(const Vec2 &sample) {  // warp.cpp:74-77
    Float a = safe_sqrt(1.0f - sample.x);
    return Vec2(1 - a, a * sample.y);
}
inline Vec3 squareToUniformSphere(const Vec2 &sample) {  // warp.cpp:24-31
    Float z = 1.0f - 2.0f * sample.y;
    Float r = safe_sqrt(1.0f - z * z);
    Float phi = 2.0f * PI_F * sample.x;
    return Vec3(r * std::cos(phi), r * std::sin(phi), z);
}

// ---------------------------------------------------------------------------
// Fresnel (util.cpp:653-683, 741-763)
// ---------------------------------------------------------------------------
inline Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta) {
    if (eta == 1) {
        cosThetaT_ = -cosThetaI_;
        return 0.0f;
    }
    Float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) {
        cosThetaT_ = 0.0f;
        return 1.0f;
    }
    Float cosThetaI = std::abs(cosThetaI_);
    Float cosThetaT = std::sqrt(cosThetaTSqr);
    Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
inline Float fresnelDielectricExt(Float cosThetaI, Float eta) {
    Float ct;
    return fresnelDielectricExt(cosThetaI, ct, eta);
}
inline Float fresnelConductorExact(Float cosThetaI, Float eta, Float k) {
    Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Float temp1 = eta * eta - k * k - sinThetaI2, a2pb2 = safe_sqrt(temp1 * temp1 + 4 * k * k * eta * eta),
          a = safe_sqrt(0.5f * (a2pb2 + temp1));
    Float term1 = a2pb2 + cosThetaI2, term2 = 2 * a * cosThetaI;
    Float Rs2 = (term1 - term2) / (term1 + term2);
    Float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    Float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
inline Vec3 fresnelConductorExact(Float cosThetaI, const Vec3 &eta, const Vec3 &k) {
    // The Spectrum overload (util.cpp:741-763) evaluates `k*k*eta*eta*4` per channel
    Vec3 r;
    for (int c = 0; c < 3; ++c) {
        Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
        Float temp1 = eta[c] * eta[c] - k[c] * k[c] - sinThetaI2;
        Float a2pb2 = safe_sqrt(temp1 * temp1 + k[c] * k[c] * eta[c] * eta[c] * 4);
        Float a = safe_sqrt((a2pb2 + temp1) * 0.5f);
        Float term1 = a2pb2 + cosThetaI2, term2 = a * (2 * cosThetaI);
        Float Rs2 = (term1 - term2) / (term1 + term2);
        Float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
        Float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
        r[c] = 0.5f * (Rp2 + Rs2);
    }
    return r;
}

// ---------------------------------------------------------------------------
// Mitsuba's own special functions (src/libcore/math.cpp:25-72), NOT libm.
// ---------------------------------------------------------------------------
inline Float mts_erfinv(Float x) {
    Float w = -std::log((1.0f - x) * (1.0f + x));
    Float p;
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
        w = std::sqrt(w) - 3.0f;
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
inline Float mts_erf(Float x) {
    Float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f,
          p = 0.3275911f;
    Float sign = signum(x);
    x = std::abs(x);
    Float t = 1.0f / (1.0f + p * x);
    Float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * std::exp(-x * x);
    return sign * y;
}
inline Float hypot2(Float a, Float b) {  // math.cpp:74-88
    Float r;
    if (std::abs(a) > std::abs(b)) {
        r = b / a;
        r = std::abs(a) * std::sqrt(1.0f + r * r);
    } else if (b != 0.0f) {
        r = a / b;
        r = std::abs(b) * std::sqrt(1.0f + r * r);
    } else {
        r = 0.0f;
    }
    return r;
}

// Catmull-Rom 1-D lookup, src/libcore/spline.cpp:23-60
inline Float evalCubicInterp1D(Float x, const Float *values, size_t size, Float min, Float max) {
    if (!(x >= min && x <= max)) return 0.0f;
    Float t = ((x - min) * (size - 1)) / (max - min);
    size_t k = std::max((size_t)0, std::min((size_t)t, size - 2));
    Float f0 = values[k], f1 = values[k + 1], d0, d1;
    if (k > 0)
        d0 = 0.5f * (values[k + 1] - values[k - 1]);
    else
        d0 = values[k + 1] - values[k];
    if (k + 2 < size)
        d1 = 0.5f * (values[k + 2] - values[k]);
    else
        d1 = values[k + 1] - values[k];
    t = t - (Float)k;
    Float t2 = t * t, t3 = t2 * t;
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
}

// N-dimensional variant used to reduce the rough-transmittance tables
// (spline.cpp:236-304 for 2-D, :379-450 for 3-D; same weights, nested loops).
template <int DIM>
inline Float evalCubicInterpND(const Float *p, const Float *values, const size_t *size) {
    Float knotWeights[DIM][4];
    size_t knot[DIM];
    for (int dim = 0; dim < DIM; ++dim) {
        Float *weights = knotWeights[dim];
        if (!(p[dim] >= 0.0f && p[dim] <= 1.0f)) return 0.0f;
        Float t = ((p[dim] - 0.0f) * (size[dim] - 1)) / (1.0f - 0.0f);
        knot[dim] = std::min((size_t)t, size[dim] - 2);
        t = t - (Float)knot[dim];
        Float t2 = t * t, t3 = t2 * t;
        weights[0] = 0.0f;
        weights[1] = 2 * t3 - 3 * t2 + 1;
        weights[2] = -2 * t3 + 3 * t2;
        weights[3] = 0.0f;
        Float d0 = t3 - 2 * t2 + t, d1 = t3 - t2;
        if (knot[dim] > 0) {
            weights[2] += 0.5f * d0;
            weights[0] -= 0.5f * d0;
        } else {
            weights[2] += d0;
            weights[1] -= d0;
        }
        if (knot[dim] + 2 < size[dim]) {
            weights[3] += 0.5f * d1;
            weights[1] -= 0.5f * d1;
        } else {
            weights[2] += d1;
            weights[1] -= d1;
        }
    }
    Float result = 0.0f;
    if (DIM == 2) {
        for (int y = -1; y <= 2; ++y) {
            Float wy = knotWeights[1][y + 1];
            for (int x = -1; x <= 2; ++x) {
                Float wxy = knotWeights[0][x + 1] * wy;
                if (wxy == 0) continue;
                size_t pos = (knot[1] + y) * size[0] + knot[0] + x;
                result += values[pos] * wxy;
            }
        }
    } else {
        for (int z = -1; z <= 2; ++z) {
            Float wz = knotWeights[2][z + 1];
            for (int y = -1; y <= 2; ++y) {
                Float wyz = knotWeights[1][y + 1] * wz;
                for (int x = -1; x <= 2; ++x) {
                    Float wxyz = knotWeights[0][x + 1] * wyz;
                    if (wxyz == 0) continue;
                    size_t pos = ((knot[2] + z) * size[1] + (knot[1] + y)) * size[0] + knot[0] + x;
                    result += values[pos] * wxyz;
                }
            }
        }
    }
    return result;
}

}  // namespace orc

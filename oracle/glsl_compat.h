// oracle/glsl_compat.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// A tiny GLSL/glm-compatible math shim so the reference's GLSL shaders
// (/root/reference/data/shaders/**) can be restated line by line in C++.
// glm itself (reference dependency, glm 0.9.7.4, README.md:23) is not installed
// in this image, so the handful of types/functions the hot path uses are
// restated here with glm's evaluation order. Everything is IEEE fp32; build with
// -ffp-contract=off so no FMA contraction happens (SURVEY.md Appendix C, Q14:
// "oracle = IEEE fp32, no contraction; this is the normative definition").
//
// Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may include or link this file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef uint32_t uint;

struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} explicit vec2(float a) : x(a), y(a) {} };
struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit vec3(float a) : x(a), y(a), z(a) {} };
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec4(const vec2& v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
    explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec3 xyz() const { return vec3(x, y, z); }
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} };
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };

// ---- component-wise arithmetic -------------------------------------------------
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator-(vec4 a, float s) { return vec4(a.x - s, a.y - s, a.z - s, a.w - s); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

// ---- geometric (glm evaluation order: ((x*x)+(y*y))+(z*z)) ---------------------
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline float length(vec2 a) { return sqrtf(dot(a, a)); }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline float length(vec4 a) { return sqrtf(dot(a, a)); }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
// glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }

// ---- common -------------------------------------------------------------------
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline int min(int a, int b) { return (b < a) ? b : a; }
inline int max(int a, int b) { return (a < b) ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float exp2i(int e) { return ldexpf(1.0f, e); }  // pow(2, e) for integer e: exact

// ---- bit casts / packing ---------------------------------------------------------
inline uint floatBitsToUint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }

// float -> int / uint, truncating like GLSL int()/uint(); NaN -> 0 and saturating
// (matches CUDA __float2int_rz / __float2uint_rz so both sides agree on garbage input).
inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
inline uint f2u(float f) {
    if (f != f) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    if (f <= 0.0f) return 0u;
    return (uint)f;
}

// packUnorm4x8: round(clamp(c,0,1)*255), x in the low byte. round = nearest-even.
inline uint packUnorm4x8(vec4 c) {
    uint r = 0;
    for (int i = 0; i < 4; i++) {
        float v = clamp(c[i], 0.0f, 1.0f) * 255.0f;
        r |= ((uint)rintf(v) & 0xFFu) << (8 * i);
    }
    return r;
}
inline vec4 unpackUnorm4x8(uint p) {
    return vec4(float(p & 0xFFu) / 255.0f, float((p >> 8) & 0xFFu) / 255.0f,
                float((p >> 16) & 0xFFu) / 255.0f, float((p >> 24) & 0xFFu) / 255.0f);
}

// IEEE binary16 conversion, round-to-nearest-even (what GL / CUDA __float2half_rn do).
inline uint16_t f32_to_f16(float f) {
    uint x = floatBitsToUint(f);
    uint sign = (x >> 16) & 0x8000u;
    uint mant = x & 0x007FFFFFu;
    int exp = (int)((x >> 23) & 0xFFu);
    if (exp == 0xFF) {  // Inf / NaN
        if (mant) return (uint16_t)0x7FFFu;  // canonical NaN: CUDA __float2half_rn returns 0x7FFF, sign dropped
        return (uint16_t)(sign | 0x7C00u);
    }
    int e = exp - 127 + 15;
    if (e >= 0x1F) return (uint16_t)(sign | 0x7C00u);  // overflow -> Inf
    if (e <= 0) {                                       // subnormal half or zero
        if (e < -10) return (uint16_t)sign;
        mant |= 0x00800000u;
        int shift = 14 - e;                             // 14..24
        uint hm = mant >> shift;
        uint rem = mant & ((1u << shift) - 1u);
        uint half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1u))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint hm = mant >> 13;
    uint rem = mant & 0x1FFFu;
    uint h = (uint)(e << 10) | hm;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;  // may carry into exponent (correct)
    return (uint16_t)(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    uint sign = ((uint)h & 0x8000u) << 16;
    uint exp = (h >> 10) & 0x1Fu;
    uint mant = h & 0x3FFu;
    uint x;
    if (exp == 0) {
        if (mant == 0) x = sign;
        else {
            int e = -1;
            do { mant <<= 1; e++; } while (!(mant & 0x400u));
            x = sign | ((uint)(127 - 15 - e) << 23) | ((mant & 0x3FFu) << 13);
        }
    } else if (exp == 0x1F) {
        x = sign | 0x7F800000u | (mant << 13);
    } else {
        x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
    }
    return uintBitsToFloat(x);
}
inline uint packHalf2x16(vec2 v) { return (uint)f32_to_f16(v.x) | ((uint)f32_to_f16(v.y) << 16); }
inline vec2 unpackHalf2x16(uint p) { return vec2(f16_to_f32((uint16_t)(p & 0xFFFFu)), f16_to_f32((uint16_t)(p >> 16))); }

// ---- matrices (column-major, m[col][row] like GLSL/glm) --------------------------
struct mat3 {
    vec3 c[3];
    mat3() {}
    mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 transpose(const mat3& m) {
    return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z));
}
// glm mat3*vec3: m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z (left to right)
inline vec3 operator*(const mat3& m, vec3 v) {
    return vec3((m[0].x * v.x + m[1].x * v.y) + m[2].x * v.z,
                (m[0].y * v.x + m[1].y * v.y) + m[2].y * v.z,
                (m[0].z * v.x + m[1].z * v.y) + m[2].z * v.z);
}

struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
    explicit mat4(const float* p) { for (int i = 0; i < 4; i++) c[i] = vec4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    void store(float* p) const { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) p[4 * i + j] = c[i][j]; }
};
// glm mat4*vec4: (m0*x + m1*y) + (m2*z + m3*w)
inline vec4 operator*(const mat4& m, vec4 v) {
    vec4 mul0 = m[0] * v.x, mul1 = m[1] * v.y, mul2 = m[2] * v.z, mul3 = m[3] * v.w;
    return (mul0 + mul1) + (mul2 + mul3);
}
// glm mat4*mat4: column j = ((A0*b0 + A1*b1) + A2*b2) + A3*b3
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int j = 0; j < 4; j++)
        r[j] = ((a[0] * b[j].x + a[1] * b[j].y) + a[2] * b[j].z) + a[3] * b[j].w;
    return r;
}
// glm::inverse(mat4) (cofactor expansion, glm/detail/func_matrix.inl compute_inverse<4,4>)
inline mat4 inverse(const mat4& m) {
    float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

    vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
    vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
    vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
    vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
    vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
    vec4 Fac5(Coef20, Coef20, Coef22, Coef23);

    vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
    vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
    vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
    vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);

    vec4 Inv0 = (Vec1 * Fac0 - Vec2 * Fac1) + Vec3 * Fac2;
    vec4 Inv1 = (Vec0 * Fac0 - Vec2 * Fac3) + Vec3 * Fac4;
    vec4 Inv2 = (Vec0 * Fac1 - Vec1 * Fac3) + Vec3 * Fac5;
    vec4 Inv3 = (Vec0 * Fac2 - Vec1 * Fac4) + Vec2 * Fac5;

    vec4 SignA(+1, -1, +1, -1);
    vec4 SignB(-1, +1, -1, +1);
    mat4 Inverse;
    Inverse[0] = Inv0 * SignA;
    Inverse[1] = Inv1 * SignB;
    Inverse[2] = Inv2 * SignA;
    Inverse[3] = Inv3 * SignB;

    vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
    vec4 Dot0 = m[0] * Row0;
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    mat4 r;
    for (int i = 0; i < 4; i++) r[i] = Inverse[i] * OneOverDeterminant;
    return r;
}

}  // namespace glsl

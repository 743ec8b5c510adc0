// glsl_rt.h -- a small GLSL 4.30 runtime for C++20: enough of the language's types, swizzles, built-ins,
// images / samplers and compute built-in variables that the reference's shader sources (read from
// /root/reference/data/shaders at build time, lightly rewritten by build_ref.py, never copied into the repo)
// compile with g++ and run on the CPU. TEST INFRASTRUCTURE ONLY: it exists to pin the oracle restatement
// (oracle/oracle.cpp) against the reference's own source text.
//
// Floating-point evaluation order is the one glm uses (and therefore the one oracle/glsl_compat.h uses), so the
// two can be compared bit for bit: dot = ((x*x + y*y) + z*z), mat4*vec4 = (c0*x + c1*y) + (c2*z + c3*w), ...
// GLSL leaves this order to the implementation; what the comparison pins is the algorithm (control flow, constants,
// formulas, data layout), not a particular GPU's rounding.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../glsl_compat.h"   // scalar packing helpers (half, unorm) shared with the oracle

namespace glsl_rt {
typedef unsigned int uint;

template <class T, int N> struct tvec;
template <class T, int M, int... I> struct swz;

template <class A, class B>
using prom = std::conditional_t<std::is_floating_point_v<A> || std::is_floating_point_v<B>, float,
                                std::conditional_t<std::is_same_v<A, uint> || std::is_same_v<B, uint>, uint,
                                                   std::conditional_t<std::is_same_v<A, bool> && std::is_same_v<B, bool>, bool, int>>>;

template <class X> struct vtraits { static constexpr int N = 0; using E = void; };
template <class T, int N_> struct vtraits<tvec<T, N_>> { static constexpr int N = N_; using E = T; };
template <class T, int M, int... I> struct vtraits<swz<T, M, I...>> { static constexpr int N = sizeof...(I); using E = T; };
template <class X> concept VecLike = (vtraits<std::remove_cvref_t<X>>::N > 0);
template <class X> concept Scalar = std::is_arithmetic_v<std::remove_cvref_t<X>>;
template <class X> concept Operand = VecLike<X> || Scalar<X>;
template <class X> using elem_t = std::conditional_t<VecLike<X>, typename vtraits<std::remove_cvref_t<X>>::E, std::remove_cvref_t<X>>;
template <class X> constexpr int size_of = VecLike<X> ? vtraits<std::remove_cvref_t<X>>::N : 1;

// ---- swizzle proxy: lives inside the parent's union, reads / writes the selected components ----
template <class T, int M, int... I> struct swz {
    T d[M];
    using V = tvec<T, (int)sizeof...(I)>;
    static constexpr int idx[sizeof...(I)] = {I...};
    operator V() const { V v; for (int k = 0; k < (int)sizeof...(I); k++) v[k] = d[idx[k]]; return v; }
    T operator[](int k) const { return d[idx[k]]; }
    swz& operator=(const V& v) { for (int k = 0; k < (int)sizeof...(I); k++) d[idx[k]] = v[k]; return *this; }
    swz& operator=(const swz& o) { return *this = V(o); }
    template <Operand U> swz& operator+=(const U& u) { return *this = V(V(*this) + u); }
    template <Operand U> swz& operator-=(const U& u) { return *this = V(V(*this) - u); }
    template <Operand U> swz& operator*=(const U& u) { return *this = V(V(*this) * u); }
    template <Operand U> swz& operator/=(const U& u) { return *this = V(V(*this) / u); }
};

// component access helpers used by the flattening constructors
template <class T> inline int flat_put(T* out, int k, int cap, bool) = delete;
template <class T, Scalar S> inline int flat_put(T* out, int k, int cap, const S& s) { if (k < cap) out[k] = T(s); return k + 1; }
template <class T, VecLike X> inline int flat_put(T* out, int k, int cap, const X& x) {
    for (int i = 0; i < size_of<X>; i++) { if (k < cap) out[k] = T(x[i]); k++; }
    return k;
}

#define GLSL_RT_COMMON(N)                                                                                         \
    tvec() { for (int i = 0; i < N; i++) (&x)[i] = T(0); }                                                         \
    tvec(const tvec& o) { std::memcpy((void*)this, (const void*)&o, sizeof(T) * N); }                              \
    tvec& operator=(const tvec& o) { std::memmove((void*)this, (const void*)&o, sizeof(T) * N); return *this; }    \
    template <Scalar S> explicit tvec(S s) { for (int i = 0; i < N; i++) (&x)[i] = T(s); }                         \
    template <VecLike X> requires (size_of<X> >= N) tvec(const X& v) { for (int i = 0; i < N; i++) (&x)[i] = T(v[i]); } \
    template <Operand A, Operand B, Operand... R> tvec(const A& a, const B& b, const R&... r) {                    \
        int k = 0; k = flat_put<T>(&x, k, N, a); k = flat_put<T>(&x, k, N, b); ((k = flat_put<T>(&x, k, N, r)), ...); \
    }                                                                                                              \
    T& operator[](int i) { return (&x)[i]; }                                                                       \
    const T& operator[](int i) const { return (&x)[i]; }                                                           \
    template <Operand U> tvec& operator+=(const U& u) { return *this = tvec(*this + u); }                          \
    template <Operand U> tvec& operator-=(const U& u) { return *this = tvec(*this - u); }                          \
    template <Operand U> tvec& operator*=(const U& u) { return *this = tvec(*this * u); }                          \
    template <Operand U> tvec& operator/=(const U& u) { return *this = tvec(*this / u); }                          \
    template <Operand U> tvec& operator|=(const U& u) { return *this = tvec(*this | u); }                          \
    template <Operand U> tvec& operator&=(const U& u) { return *this = tvec(*this & u); }

#define SW2(a, b, A, B) swz<T, NN, A, B> a##b;
#define SW3(a, b, c, A, B, C) swz<T, NN, A, B, C> a##b##c;
#define SW4(a, b, c, d, A, B, C, D) swz<T, NN, A, B, C, D> a##b##c##d;

template <class T> struct tvec<T, 2> {
    static constexpr int NN = 2;
    union {
        struct { T x, y; };
        struct { T r, g; };
        struct { T s, t; };
        SW2(x, y, 0, 1) SW2(y, x, 1, 0) SW2(x, x, 0, 0) SW2(y, y, 1, 1)
        SW2(r, g, 0, 1) SW2(g, r, 1, 0)
        SW2(s, t, 0, 1)
    };
    GLSL_RT_COMMON(2)
};
template <class T> struct tvec<T, 3> {
    static constexpr int NN = 3;
    union {
        struct { T x, y, z; };
        struct { T r, g, b; };
        SW2(x, y, 0, 1) SW2(x, z, 0, 2) SW2(y, z, 1, 2) SW2(y, x, 1, 0) SW2(z, x, 2, 0) SW2(z, y, 2, 1)
        SW2(r, g, 0, 1) SW2(g, b, 1, 2) SW2(r, b, 0, 2)
        SW3(x, y, z, 0, 1, 2) SW3(z, y, x, 2, 1, 0) SW3(x, z, y, 0, 2, 1) SW3(y, z, x, 1, 2, 0)
        SW3(r, g, b, 0, 1, 2) SW3(b, g, r, 2, 1, 0)
    };
    GLSL_RT_COMMON(3)
};
template <class T> struct tvec<T, 4> {
    static constexpr int NN = 4;
    union {
        struct { T x, y, z, w; };
        struct { T r, g, b, a; };
        SW2(x, y, 0, 1) SW2(x, z, 0, 2) SW2(y, z, 1, 2) SW2(z, w, 2, 3) SW2(x, w, 0, 3) SW2(y, x, 1, 0)
        SW2(r, g, 0, 1) SW2(g, b, 1, 2) SW2(b, a, 2, 3) SW2(r, b, 0, 2)
        SW3(x, y, z, 0, 1, 2) SW3(y, z, w, 1, 2, 3) SW3(x, y, w, 0, 1, 3)
        SW3(r, g, b, 0, 1, 2) SW3(g, b, a, 1, 2, 3)
        SW4(x, y, z, w, 0, 1, 2, 3) SW4(r, g, b, a, 0, 1, 2, 3)
    };
    GLSL_RT_COMMON(4)
};
#undef SW2
#undef SW3
#undef SW4

typedef tvec<float, 2> vec2; typedef tvec<float, 3> vec3; typedef tvec<float, 4> vec4;
typedef tvec<int, 2> ivec2; typedef tvec<int, 3> ivec3; typedef tvec<int, 4> ivec4;
typedef tvec<uint, 2> uvec2; typedef tvec<uint, 3> uvec3; typedef tvec<uint, 4> uvec4;
typedef tvec<bool, 2> bvec2; typedef tvec<bool, 3> bvec3; typedef tvec<bool, 4> bvec4;

template <class X> inline auto comp(const X& x, int i) { if constexpr (VecLike<X>) return x[i]; else return x; }

// ---- component-wise binary operators with GLSL's implicit conversions (int -> uint -> float) ----
#define GLSL_RT_BINOP(OP)                                                                                  \
    template <Operand L, Operand R> requires (VecLike<L> || VecLike<R>)                                     \
    inline auto operator OP(const L& l, const R& r) {                                                       \
        constexpr int N = size_of<L> > size_of<R> ? size_of<L> : size_of<R>;                                \
        static_assert((size_of<L> == N || size_of<L> == 1) && (size_of<R> == N || size_of<R> == 1), "vector size mismatch"); \
        using P = prom<elem_t<L>, elem_t<R>>;                                                               \
        tvec<P, N> out;                                                                                     \
        for (int i = 0; i < N; i++) out[i] = P(P(comp(l, i)) OP P(comp(r, i)));                             \
        return out;                                                                                         \
    }
GLSL_RT_BINOP(+) GLSL_RT_BINOP(-) GLSL_RT_BINOP(*) GLSL_RT_BINOP(/) GLSL_RT_BINOP(%) GLSL_RT_BINOP(&) GLSL_RT_BINOP(|) GLSL_RT_BINOP(^)
#undef GLSL_RT_BINOP
#define GLSL_RT_SHIFT(OP)                                                                                  \
    template <VecLike L, Operand R> inline auto operator OP(const L& l, const R& r) {                       \
        constexpr int N = size_of<L>; tvec<elem_t<L>, N> out;                                               \
        for (int i = 0; i < N; i++) out[i] = elem_t<L>(comp(l, i) OP comp(r, i));                           \
        return out;                                                                                         \
    }
GLSL_RT_SHIFT(>>) GLSL_RT_SHIFT(<<)
#undef GLSL_RT_SHIFT
template <VecLike L> inline auto operator-(const L& l) { tvec<elem_t<L>, size_of<L>> out; for (int i = 0; i < size_of<L>; i++) out[i] = -l[i]; return out; }
template <VecLike L, VecLike R> requires (size_of<L> == size_of<R>) inline bool operator==(const L& l, const R& r) {
    for (int i = 0; i < size_of<L>; i++) if (!(l[i] == r[i])) return false;
    return true;
}
template <VecLike L, VecLike R> requires (size_of<L> == size_of<R>) inline bool operator!=(const L& l, const R& r) { return !(l == r); }

// ---- scalar built-ins (implicit int -> float conversion of the arguments, as GLSL does) ----
template <Scalar A, Scalar B> inline prom<A, B> min(A a, B b) { using P = prom<A, B>; return (P(b) < P(a)) ? P(b) : P(a); }
template <Scalar A, Scalar B> inline prom<A, B> max(A a, B b) { using P = prom<A, B>; return (P(a) < P(b)) ? P(b) : P(a); }
template <Scalar A, Scalar B, Scalar C> inline auto clamp(A x, B lo, C hi) { return min(max(x, lo), hi); }
template <Scalar A> inline float sqrt(A a) { return ::sqrtf(float(a)); }
template <Scalar A> inline float inversesqrt(A a) { return 1.0f / ::sqrtf(float(a)); }
template <Scalar A, Scalar B> inline float pow(A a, B b) { return ::powf(float(a), float(b)); }
template <Scalar A> inline float exp2(A a) { return ::exp2f(float(a)); }
template <Scalar A> inline float log2(A a) { return ::log2f(float(a)); }
template <Scalar A> inline float sin(A a) { return ::sinf(float(a)); }
template <Scalar A> inline float cos(A a) { return ::cosf(float(a)); }
template <Scalar A> inline float floor(A a) { return ::floorf(float(a)); }
template <Scalar A> inline float ceil(A a) { return ::ceilf(float(a)); }
template <Scalar A> inline float fract(A a) { return float(a) - ::floorf(float(a)); }
template <Scalar A, Scalar B> inline float mod(A a, B b) { return float(a) - float(b) * ::floorf(float(a) / float(b)); }
inline float abs(float a) { return ::fabsf(a); }
inline int abs(int a) { return a < 0 ? -a : a; }
inline float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
template <Scalar A, Scalar B> inline float step(A edge, B x) { return float(x) < float(edge) ? 0.0f : 1.0f; }
template <Scalar A, Scalar B, Scalar C> inline float smoothstep(A e0, B e1, C x) {
    float t = clamp((float(x) - float(e0)) / (float(e1) - float(e0)), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
template <Scalar A, Scalar B, Scalar C> inline float mix(A a, B b, C t) { return float(a) * (1.0f - float(t)) + float(b) * float(t); }
inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline int floatBitsToInt(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float dot(float a, float b) { return a * b; }

// ---- vector built-ins ----
#define GLSL_RT_MAP1(NAME)                                                                                 \
    template <VecLike X> inline auto NAME(const X& x) { tvec<decltype(NAME(x[0])), size_of<X>> o; for (int i = 0; i < size_of<X>; i++) o[i] = NAME(x[i]); return o; }
GLSL_RT_MAP1(sqrt) GLSL_RT_MAP1(floor) GLSL_RT_MAP1(ceil) GLSL_RT_MAP1(fract) GLSL_RT_MAP1(abs) GLSL_RT_MAP1(sign) GLSL_RT_MAP1(exp2) GLSL_RT_MAP1(log2)
GLSL_RT_MAP1(floatBitsToUint) GLSL_RT_MAP1(uintBitsToFloat)
#undef GLSL_RT_MAP1
#define GLSL_RT_MAP2(NAME)                                                                                 \
    template <Operand L, Operand R> requires (VecLike<L> || VecLike<R>) inline auto NAME(const L& l, const R& r) { \
        constexpr int N = size_of<L> > size_of<R> ? size_of<L> : size_of<R>;                                \
        tvec<decltype(NAME(comp(l, 0), comp(r, 0))), N> o; for (int i = 0; i < N; i++) o[i] = NAME(comp(l, i), comp(r, i)); return o; }
GLSL_RT_MAP2(min) GLSL_RT_MAP2(max) GLSL_RT_MAP2(pow) GLSL_RT_MAP2(mod) GLSL_RT_MAP2(step)
#undef GLSL_RT_MAP2
template <VecLike X, Operand A, Operand B> inline auto clamp(const X& x, const A& lo, const B& hi) { return min(max(x, lo), hi); }
template <VecLike X, VecLike Y, Operand T> inline auto mix(const X& a, const Y& b, const T& t) {
    tvec<float, size_of<X>> o; for (int i = 0; i < size_of<X>; i++) o[i] = mix(a[i], b[i], comp(t, i)); return o;
}
template <VecLike A, VecLike B> requires (size_of<A> == size_of<B>) inline float dot(const A& a, const B& b) {
    if constexpr (size_of<A> == 2) return float(a[0]) * float(b[0]) + float(a[1]) * float(b[1]);
    else if constexpr (size_of<A> == 3) return (float(a[0]) * float(b[0]) + float(a[1]) * float(b[1])) + float(a[2]) * float(b[2]);
    else return (float(a[0]) * float(b[0]) + float(a[1]) * float(b[1])) + (float(a[2]) * float(b[2]) + float(a[3]) * float(b[3]));
}
template <VecLike A> inline float length(const A& a) { return ::sqrtf(dot(a, a)); }
template <VecLike A, VecLike B> inline float distance(const A& a, const B& b) { return length(a - b); }
template <VecLike A> inline auto normalize(const A& a) { float inv = 1.0f / ::sqrtf(dot(a, a)); return tvec<float, size_of<A>>(a * inv); }
template <VecLike A, VecLike B> inline vec3 cross(const A& a, const B& b) {
    return vec3(a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]);
}
template <VecLike A> inline bool any(const A& a) { for (int i = 0; i < size_of<A>; i++) if (a[i]) return true; return false; }
template <VecLike A> inline bool all(const A& a) { for (int i = 0; i < size_of<A>; i++) if (!a[i]) return false; return true; }

inline uint packUnorm4x8(const vec4& c) { return glsl::packUnorm4x8(glsl::vec4(c.x, c.y, c.z, c.w)); }
inline vec4 unpackUnorm4x8(uint p) { glsl::vec4 v = glsl::unpackUnorm4x8(p); return vec4(v.x, v.y, v.z, v.w); }
inline uint packHalf2x16(const vec2& v) { return glsl::packHalf2x16(glsl::vec2(v.x, v.y)); }
inline vec2 unpackHalf2x16(uint p) { glsl::vec2 v = glsl::unpackHalf2x16(p); return vec2(v.x, v.y); }
inline uint packUnorm2x16(const vec2& v) {
    auto q = [](float f) { return (uint)::lrintf(clamp(f, 0.0f, 1.0f) * 65535.0f); };
    return q(v.x) | (q(v.y) << 16);
}
inline vec2 unpackUnorm2x16(uint p) { return vec2(float(p & 0xFFFFu) / 65535.0f, float(p >> 16) / 65535.0f); }

// ---- matrices (column-major, like GLSL) ----
struct mat3 {
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 transpose(const mat3& m) {
    return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z));
}
template <VecLike X> requires (size_of<X> == 3) inline vec3 operator*(const mat3& m, const X& v_) {
    vec3 v(v_);
    return vec3((m[0].x * v.x + m[1].x * v.y) + m[2].x * v.z, (m[0].y * v.x + m[1].y * v.y) + m[2].y * v.z,
                (m[0].z * v.x + m[1].z * v.y) + m[2].z * v.z);
}
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(const float* p) { for (int j = 0; j < 4; j++) c[j] = vec4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
template <VecLike X> requires (size_of<X> == 4) inline vec4 operator*(const mat4& m, const X& v_) {
    vec4 v(v_), r;
    for (int i = 0; i < 4; i++) r[i] = (m[0][i] * v.x + m[1][i] * v.y) + (m[2][i] * v.z + m[3][i] * v.w);
    return r;
}

// ---- resources ----
enum Format { F_R32F, F_RGBA32F, F_RGB10A2, F_RGBA8, F_R16, F_R32UI, F_R16UI };
struct Image {            // a 2D (depth == 1) or 3D texture level / image; out-of-bounds loads return 0, stores are dropped
    void* data = nullptr; int w = 0, h = 0, depth = 1; Format fmt = F_R32F;
    bool in(int x, int y, int z = 0) const { return x >= 0 && y >= 0 && z >= 0 && x < w && y < h && z < depth; }
    size_t at(int x, int y, int z = 0) const { return ((size_t)z * h + y) * w + x; }
    vec4 loadf(int x, int y) const {
        if (!in(x, y)) return vec4(0.0f);
        const size_t i = at(x, y);
        switch (fmt) {
            case F_R32F: return vec4(((const float*)data)[i], 0.0f, 0.0f, 1.0f);
            case F_RGBA32F: { const float* p = (const float*)data + 4 * i; return vec4(p[0], p[1], p[2], p[3]); }
            case F_RGB10A2: { uint32_t u = ((const uint32_t*)data)[i]; return vec4(float(u & 0x3FFu) / 1023.0f, float((u >> 10) & 0x3FFu) / 1023.0f, float((u >> 20) & 0x3FFu) / 1023.0f, float(u >> 30) / 3.0f); }
            case F_RGBA8: { uint32_t u = ((const uint32_t*)data)[i]; return vec4(float(u & 0xFFu) / 255.0f, float((u >> 8) & 0xFFu) / 255.0f, float((u >> 16) & 0xFFu) / 255.0f, float(u >> 24) / 255.0f); }
            case F_R16: return vec4(float(((const uint16_t*)data)[i]) / 65535.0f, 0.0f, 0.0f, 1.0f);
            default: return vec4(0.0f);
        }
    }
    uvec4 loadu(int x, int y, int z = 0) const {
        if (!in(x, y, z)) return uvec4(0u);
        const size_t i = at(x, y, z);
        if (fmt == F_R32UI) return uvec4(((const uint32_t*)data)[i], 0u, 0u, 1u);
        return uvec4((uint)((const uint16_t*)data)[i], 0u, 0u, 1u);
    }
    void storef(int x, int y, const vec4& v) {
        if (!in(x, y)) return;
        const size_t i = at(x, y);
        if (fmt == F_RGBA32F) { float* p = (float*)data + 4 * i; p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
        else if (fmt == F_R32F) ((float*)data)[i] = v.x;
        else if (fmt == F_R16) ((uint16_t*)data)[i] = (uint16_t)::lrintf(clamp(v.x, 0.0f, 1.0f) * 65535.0f);   // GL 4.3 spec 2.3.4.1: round to nearest
    }
    void storeu(int x, int y, int z, uint v) {
        if (!in(x, y, z)) return;
        if (fmt == F_R32UI) ((uint32_t*)data)[at(x, y, z)] = v; else ((uint16_t*)data)[at(x, y, z)] = (uint16_t)v;
    }
};
struct image2D : Image {}; struct uimage2D : Image {}; struct uimage3D : Image {}; struct uimage1D : Image {}; struct imageBuffer : Image {};
struct sampler2D : Image {}; struct usampler2D : Image {}; struct sampler3D : Image {}; struct samplerCube : Image {};

template <VecLike C> inline vec4 imageLoad(const image2D& im, const C& c) { return im.loadf(int(c[0]), int(c[1])); }
template <VecLike C> inline uvec4 imageLoad(const uimage2D& im, const C& c) { return im.loadu(int(c[0]), int(c[1])); }
template <VecLike C> inline uvec4 imageLoad(const uimage3D& im, const C& c) { return im.loadu(int(c[0]), int(c[1]), int(c[2])); }
template <VecLike C, VecLike V> inline void imageStore(image2D& im, const C& c, const V& v) { im.storef(int(c[0]), int(c[1]), vec4(v)); }
template <VecLike C, VecLike V> inline void imageStore(uimage2D& im, const C& c, const V& v) { im.storeu(int(c[0]), int(c[1]), 0, uint(v[0])); }
template <VecLike C, VecLike V> inline void imageStore(uimage3D& im, const C& c, const V& v) { im.storeu(int(c[0]), int(c[1]), int(c[2]), uint(v[0])); }
template <VecLike C> inline uint imageAtomicMin(uimage2D& im, const C& c, uint v) {     // invocations run one at a time
    if (!im.in(int(c[0]), int(c[1]))) return 0u;
    uint32_t& t = ((uint32_t*)im.data)[im.at(int(c[0]), int(c[1]))];
    const uint old = t; if (v < t) t = v; return old;
}
template <Scalar C> inline uvec4 imageLoad(const uimage1D& im, C c) { return im.loadu(int(c), 0); }
template <Scalar C, VecLike V> inline void imageStore(uimage1D& im, C c, const V& v) { im.storeu(int(c), 0, 0, uint(v[0])); }
template <Scalar C> inline vec4 imageLoad(const imageBuffer& im, C c) { return im.loadf(int(c), 0); }
template <Scalar C, VecLike V> inline void imageStore(imageBuffer& im, C c, const V& v) { im.storef(int(c), 0, vec4(v)); }
inline ivec2 imageSize(const uimage2D& im) { return ivec2(im.w, im.h); }
inline ivec2 imageSize(const image2D& im) { return ivec2(im.w, im.h); }
struct size1 { int x; };
inline size1 imageSize(const imageBuffer& im) { return size1{im.w}; }
template <VecLike A, VecLike B> inline auto lessThan(const A& a, const B& b) { tvec<bool, size_of<A>> o; for (int i = 0; i < size_of<A>; i++) o[i] = a[i] < b[i]; return o; }
inline uint atomicAdd(uint& m, uint v) { const uint old = m; m += v; return old; }      // invocations never run concurrently
inline int atomicAdd(int& m, int v) { const int old = m; m += v; return old; }
template <class S> inline ivec2 textureSize(const S& s, int) { return ivec2(s.w, s.h); }
template <VecLike C> inline vec4 texelFetch(const sampler2D& s, const C& c, int) { return s.loadf(int(c[0]), int(c[1])); }
template <VecLike C> inline uvec4 texelFetch(const usampler2D& s, const C& c, int) { return s.loadu(int(c[0]), int(c[1])); }
// GL_NEAREST + GL_CLAMP_TO_EDGE (the sampling state of every texture this path reads through texture()/textureLod())
template <VecLike C> inline vec4 textureLod(const sampler2D& s, const C& uv, float) {
    int x = (int)::floorf(float(uv[0]) * float(s.w)), y = (int)::floorf(float(uv[1]) * float(s.h));
    x = x < 0 ? 0 : (x >= s.w ? s.w - 1 : x); y = y < 0 ? 0 : (y >= s.h ? s.h - 1 : y);
    return s.loadf(x, y);
}
template <VecLike C> inline vec4 texture(const sampler2D& s, const C& uv, float = 0.0f) { return textureLod(s, uv, 0.0f); }
template <VecLike C> inline vec4 texture(const sampler3D&, const C&) { return vec4(0.0f); }
template <VecLike C> inline vec4 texture(const samplerCube&, const C&) { return vec4(0.0f); }   // not on this path (omni shadow maps)

// `layout(local_size_x = X, ...) in;` becomes `static const wgsize<X, Y, Z> gl_WorkGroupSize;` (its components must be
// usable as array bounds)
template <uint X, uint Y, uint Z> struct wgsize {
    static constexpr uint x = X, y = Y, z = Z;
    static inline const uvec2 xy = uvec2(X, Y);
    static inline const uvec3 xyz = uvec3(X, Y, Z);
    operator uvec3() const { return uvec3(X, Y, Z); }
};

// ---- compute built-ins and the dispatcher ----
struct Builtins { uvec3 NumWorkGroups, WorkGroupID, LocalInvocationID, GlobalInvocationID; uint LocalInvocationIndex; };
inline Builtins g_builtins;
#define gl_NumWorkGroups (::glsl_rt::g_builtins.NumWorkGroups)
#define gl_WorkGroupID (::glsl_rt::g_builtins.WorkGroupID)
#define gl_LocalInvocationID (::glsl_rt::g_builtins.LocalInvocationID)
#define gl_GlobalInvocationID (::glsl_rt::g_builtins.GlobalInvocationID)
#define gl_LocalInvocationIndex (::glsl_rt::g_builtins.LocalInvocationIndex)

// a work group's invocations run one after the other; shaders that call barrier() are run by dispatch_phased (below)
template <class F> inline void dispatch(F shader_main, const uvec3& local, uint gx, uint gy, uint gz) {
    Builtins& b = g_builtins;
    b.NumWorkGroups = uvec3(gx, gy, gz);
    for (uint wz = 0; wz < gz; wz++) for (uint wy = 0; wy < gy; wy++) for (uint wx = 0; wx < gx; wx++) {
        b.WorkGroupID = uvec3(wx, wy, wz);
        for (uint lz = 0; lz < local.z; lz++) for (uint ly = 0; ly < local.y; ly++) for (uint lx = 0; lx < local.x; lx++) {
            b.LocalInvocationID = uvec3(lx, ly, lz);
            b.GlobalInvocationID = uvec3(wx * local.x + lx, wy * local.y + ly, wz * local.z + lz);
            b.LocalInvocationIndex = (lz * local.y + ly) * local.x + lx;
            shader_main();
        }
    }
}
}  // namespace glsl_rt
#include <ucontext.h>
namespace glsl_rt {
// Shaders that call barrier(): every invocation of a work group is a fiber (ucontext); barrier() yields to the scheduler,
// which resumes the group's fibers in invocation order until all have returned. Invocations still never run
// concurrently, so shared-memory atomics are plain read-modify-writes and their order is the invocation order.
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = false; uvec3 lid; uint lidx = 0; };
struct Scheduler { ucontext_t main; Fiber* cur = nullptr; void (*entry)() = nullptr; };
inline Scheduler g_sched;
inline void barrier() { if (g_sched.cur) swapcontext(&g_sched.cur->ctx, &g_sched.main); }
inline void fiber_trampoline() { g_sched.entry(); g_sched.cur->done = true; swapcontext(&g_sched.cur->ctx, &g_sched.main); }
inline void dispatch_phased(void (*shader_main)(), const uvec3& local, uint gx, uint gy, uint gz) {
    Builtins& b = g_builtins;
    b.NumWorkGroups = uvec3(gx, gy, gz);
    const uint n = local.x * local.y * local.z;
    std::vector<Fiber> fibers(n);
    for (auto& f : fibers) f.stack.resize(256 * 1024);
    g_sched.entry = shader_main;
    for (uint wz = 0; wz < gz; wz++) for (uint wy = 0; wy < gy; wy++) for (uint wx = 0; wx < gx; wx++) {
        uint k = 0;
        for (uint lz = 0; lz < local.z; lz++) for (uint ly = 0; ly < local.y; ly++) for (uint lx = 0; lx < local.x; lx++, k++) {
            Fiber& f = fibers[k];
            f.done = false; f.lid = uvec3(lx, ly, lz); f.lidx = k;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = &g_sched.main;
            makecontext(&f.ctx, fiber_trampoline, 0);
        }
        bool any_left = true;
        while (any_left) {
            any_left = false;
            for (auto& f : fibers) {
                if (f.done) continue;
                b.WorkGroupID = uvec3(wx, wy, wz);
                b.LocalInvocationID = f.lid;
                b.GlobalInvocationID = uvec3(wx * local.x + f.lid.x, wy * local.y + f.lid.y, wz * local.z + f.lid.z);
                b.LocalInvocationIndex = f.lidx;
                g_sched.cur = &f;
                swapcontext(&g_sched.main, &f.ctx);
                g_sched.cur = nullptr;
                if (!f.done) any_left = true;
            }
        }
    }
}
// tessellation / geometry stage built-ins (ism.tesc, ism.tese, ism.geom): plain variables, filled in by the driver
struct PerVertex { vec4 gl_Position; float gl_PointSize = 1.0f; };
inline PerVertex gl_in[3], gl_out[3];
inline vec4 gl_Position;
inline float gl_PointSize = 1.0f;
inline int gl_InvocationID = 0, gl_PrimitiveIDIn = 0;
inline vec3 gl_TessCoord;
inline float gl_TessLevelOuter[4], gl_TessLevelInner[2];
inline int g_emitted = 0;
inline void EmitVertex() { g_emitted++; }
inline void EndPrimitive() {}
inline uint atomicAdd(uint& m, int v) { const uint old = m; m += (uint)v; return old; }
inline vec4 g_FragCoord;
#define gl_FragCoord (::glsl_rt::g_FragCoord)
inline void memoryBarrierShared() {}
inline void memoryBarrier() {}
inline void groupMemoryBarrier() {}

}  // namespace glsl_rt

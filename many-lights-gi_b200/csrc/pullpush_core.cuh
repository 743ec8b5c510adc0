// pullpush_core.cuh -- the arithmetic of pull.comp / push.comp for one 64x64 ISM held in shared memory, written as
// per-thread "phases" (everything between two CTA barriers is one phase; inside a phase threads only communicate through
// the list counter). The kernel in exact.cu runs phase(tid) on 256 threads with __syncthreads() in between; the test
// harness tests/hostemu compiles the same phases for the host and runs tid = 0..255 in a loop, so the logic of this file is
// checked bit for bit against the oracle on a machine without a GPU. Every operation is an IEEE fp32 operation with one
// rounding (no FMA contraction: -fmad=false / -ffp-contract=off), so host and device agree.
//
// Bit-exact contract: pull.comp:20-126, push.comp:44-165, ImperfectShadowmap.cpp:121-182 (R16 result).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#else
#define PP_HD inline
#endif

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace mlgi {
namespace pp {

constexpr int kThreads = 256;
constexpr int kThrPitch = 18;                      // padded threshold plane: coarse level of push 5..1 (<= 16 wide) plus a one-texel ring

// float4 offset of pyramid level l (1..6) inside PY: sum_{k<l} 4^(6-k)
PP_HD constexpr int lvl_off(int l) { return (4096 - (4096 >> (2 * (l - 1)))) / 3; }

struct Shared {
    uint32_t L0[64 * 64];                          // R32UI tile in, R16 results (one per word) out
    float4 PY[1024 + 256 + 64 + 16 + 4 + 1];       // RGBA32F levels 1..6 (pull results, overwritten in place by push results)
    union {                                        // never live at the same time:
        float thr[2][kThrPitch * kThrPitch];       //   push: thresholds of the coarse level, ping-pong by (consumer level & 1)
        uint16_t list[1024];                       //   pull level 1: outputs whose 2x2 block holds a point
    };
    float lut_r[256], lut_md[256];                 // level-0 decode of the 8-bit radius (pull.comp:39,57)
    float rcp[20];                                 // RN(1 / n), n = 1..16: exact division by a small integer (div_small)
    float4 empty1;                                 // pull level 1 of an all-empty 2x2 block
    int count;                                     // number of entries of `list`
};

// ---- bit casts / conversions with host equivalents -------------------------------------------------------------------
PP_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
PP_HD uint32_t f2u_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PP_HD float pinf() { return u2f(0x7f800000u); }
PP_HD float min_glsl(float a, float b) { return (b < a) ? b : a; }
PP_HD float max_glsl(float a, float b) { return (a < b) ? b : a; }
PP_HD uint32_t pack_half2x(float x, float y) {
    return (uint32_t)__half_as_ushort(__float2half_rn(x)) | ((uint32_t)__half_as_ushort(__float2half_rn(y)) << 16);
}
PP_HD float half_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); }
PP_HD float half_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }
PP_HD uint32_t r16_unorm(float d) {                // push.comp:29-41: imageStore to R16 (clamp, scale, round to nearest even)
    const float c = min_glsl(max_glsl(d, 0.0f), 1.0f) * 65535.0f;
#if defined(__CUDA_ARCH__)
    return __float2uint_rn(c);
#else
    return (uint32_t)nearbyintf(c);
#endif
}
PP_HD float sqrt_rn(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}

// Exact square-root elimination. For a correctly rounded sqrtf and rs >= 0:  sqrtf(x) > rs  <=>  sqrt(x) > m, where
// m = (rs + nextafter(rs)) / 2 is the rounding boundary above rs  <=>  x > m*m. m has 25 significant bits, so m*m is
// exact in fp64 and is never itself a float; hence  sqrtf(x) > rs  <=>  x > T  with T = round_down_to_float(m*m).
// One threshold per coarse texel replaces one square root per (pixel, tap). rs = inf / NaN give T = NaN (never greater,
// like the original comparison); a negative rs (cannot occur: radii are means of non-negative values) keeps
// "always greater" through T = -1.
PP_HD float sqrt_gt_threshold(float rs) {
    const float next = u2f(f2u_bits(rs) + 1u);
    const double m = ((double)rs + (double)next) * 0.5;
    const double mm = m * m;
#if defined(__CUDA_ARCH__)
    const float T = __double2float_rd(mm);
#else
    float T = (float)mm;                           // round to nearest, then step down if that rounded up
    if ((double)T > mm) T = nextafterf(T, -INFINITY);
#endif
    return rs < 0.0f ? -1.0f : T;
}

// x / n for a small positive integer n (2 <= n <= 16 as a float, y = RN(1/n) from Shared::rcp), correctly rounded without
// the generic division sequence: q0 = RN(x*y) is within 2 ulp of x/n, the residual r = x - q0*n is then exactly
// representable (a small multiple of ulp(q0)) and is delivered exactly by the FMA, and q1 = RN(q0 + r*y) is the correctly
// rounded quotient (Markstein's correction step: the quotient of two floats is never a rounding midpoint, and for an
// integer n <= 16 its distance to the nearest midpoint is >= ulp/32, far beyond the 2^-24 ulp perturbation of r*y).
// Holds for x = 0 and for every normal x whose quotient is normal -- all values on this path are 0 or in [1e-8, 1e3];
// anything else (inf, NaN, huge, tiny) takes the IEEE division. n == 1 is exact trivially (q0 = x, r = 0).
PP_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
PP_HD float div_small_unchecked(float x, float n, float y) {    // caller guarantees x == 0 or 1e-30 <= |x| <= 1e30
    const float q0 = x * y;
    const float r = fma_rn(-q0, n, x);
    return fma_rn(r, y, q0);
}
PP_HD float div_small(float x, float n, float y) {
    const float ax = fabsf(x);
    if (!(ax < 1e30f) || (ax != 0.0f && ax < 1e-30f)) return x / n;
    return div_small_unchecked(x, n, y);
}

// ---- pull -------------------------------------------------------------------------------------------------------------
struct PullSample { float depth, maxDepth, radius, dx, dy; };

// level-0 decode, pull.comp:37-41,55-65 (ismPixelSize = A / S, ImperfectShadowmap.cpp:187-191)
PP_HD void decode_l0(float lut_r, float lut_md, uint32_t u, float zFar, float ismPixelSize, PullSample& smp) {
    const float depthSample = float(u >> 8) / 16777216.0f;
    float radius = lut_r;                                               // float(u & 0xFF) / 10      :39
    const float maxDepth = depthSample + lut_md;                        // + (radius*2)/zFar*0.6     :57
    const float distToCamera = depthSample * zFar;                      // :60
    radius = radius / distToCamera / 3.14f * ismPixelSize;              // :61
    radius *= 1.3f;                                                     // :63
    radius = min_glsl(radius, 15.0f);                                   // :65
    smp.depth = depthSample; smp.maxDepth = maxDepth; smp.radius = radius; smp.dx = 0.0f; smp.dy = 0.0f;
}

// pull.comp:67-123 given the four decoded inputs (absolute input coords ix0.., output coord ox,oy). Coordinates are GLOBAL
// atlas coordinates: the fp32 rounding of (coord + 0.5 + displacement) depends on their magnitude. rcp = Shared::rcp.
PP_HD float4 pull_combine(const PullSample (&smp)[4], int ix0, int iy0, int ox, int oy, int level, const float* rcp) {
    const int offx[4] = {0, 0, 1, 1}, offy[4] = {0, 1, 0, 1};   // pull.comp:24
    float e[4], ndx[4], ndy[4];
    bool valid[4];
    const float scale = u2f((uint32_t)(127 - level) << 23);      // 2^-level, exact (:72)
    const float fox = float(ox) + 0.5f, foy = float(oy) + 0.5f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ndx[i] = (float(ix0 + offx[i]) + 0.5f + smp[i].dx) / 2.0f - fox;   // :69
        ndy[i] = (float(iy0 + offy[i]) + 0.5f + smp[i].dy) / 2.0f - foy;
        const bool radiusCheckPassed = sqrt_rn(ndx[i] * ndx[i] + ndy[i] * ndy[i]) <= smp[i].radius * scale;   // :70-74
        valid[i] = (smp[i].depth != 1.0f) && radiusCheckPassed;            // :80-81
        e[i] = valid[i] ? smp[i].depth : pinf();
    }
    // :84-92: maxDepth is taken from the LAST sample whose depth equals the running minimum, i.e. the last index that
    // attains the minimum over the valid samples (depths are finite and >= +0, so fminf == the GLSL ternary min)
    const float mn = fminf(fminf(e[0], e[1]), fminf(e[2], e[3]));
    float maxDepth = smp[0].maxDepth;
#pragma unroll
    for (int i = 1; i < 4; i++) if (e[i] == mn) maxDepth = smp[i].maxDepth;
    float depthAcc = 0.0f, radiusAcc = 0.0f, maxValidMaxDepth = 0.0f, dax = 0.0f, day = 0.0f;
    int numValid = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {                                 // :94-113
        if (!valid[i] || smp[i].depth > maxDepth) continue;
        depthAcc += smp[i].depth;
        dax += ndx[i]; day += ndy[i];
        radiusAcc += smp[i].radius;
        maxValidMaxDepth = fmaxf(maxValidMaxDepth, smp[i].maxDepth);
        numValid++;
    }
    if (numValid > 0) {                                           // :116-123
        // depth sums are 0 or >= 2^-24 and <= 4; displacement sums are multiples of 2^-25 below 2^13 (coordinates < 2^13 plus halves
        // with a 2^-24 grid): all inside div_small's exact range, so only the radius -- which may be inf or NaN for a zero-depth
        // texel (radius / distToCamera, pull.comp:61) -- takes the checked form
        const float nv = float(numValid), y = rcp[numValid];
        return make_float4(div_small_unchecked(depthAcc, nv, y), maxValidMaxDepth, div_small(radiusAcc, nv, y),
                           u2f(pack_half2x(div_small_unchecked(dax, nv, y), div_small_unchecked(day, nv, y))));
    }
    return make_float4(1.0f, 0.0f, 0.0f, 0.0f);
}

// ---- push -------------------------------------------------------------------------------------------------------------
// The four coarse texels of one push thread T (push.comp:46-62), in the shader's order {0,0},{0,1},{1,1},{1,0}.
struct PushTaps {
    float depth[4], maxDepth[4], radius[4];
    float thr[4];           // threshold of the radius check; -1 for a tap that may not contribute at all (other ISM :77-86,
                            // or depth == 1.0 :88-93): dot(d,d) >= 0 > -1 zeroes its weight like the shader does
    float d2x[4], d2y[4];   // coarse displacement * 2                             (:101)
};

// One output pixel of thread T. OP is the shader's outputPixel (:65-67). displacementVectors[i] = (coarserTexCoord -
// thisTexCoord) + disp*2 (:98-101): the bracket is +-0.5 or +-1.5 for every T (differences of small half-integers are
// exact), so it is a compile-time constant per (OP, tap). Weights are {9,3,1,3} rotated by OP (:70-74). The weighted means
// are evaluated only when the texel is replaced (:159-161), and at level 0 only the depth is consumed (R16 store), so the
// other means are not evaluated there; skipped work never feeds a stored value.
template <bool LEVEL0>
PP_HD float4 push_pixel(int OP, const PushTaps& c, float4 orig, const float* rcp) {
    const int offx[4] = {0, 0, 1, 1}, offy[4] = {0, 1, 1, 0};              // tap offsets (:48); only indexed by constants
    const int opx = OP >> 1, opy = (OP == 1 || OP == 2) ? 1 : 0;           // offsets[OP]
    float dvx[4], dvy[4], e[4], w[4];
    bool p[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {                                          // :96-110
        const int d = (i - OP) & 3;                                        // weightsX[(i - outputPixel + 4) % 4], weightsX = {9,3,1,3}
        w[i] = d == 0 ? 9.0f : (d == 2 ? 1.0f : 3.0f);
        const float kx = float(2 * offx[i] - opx) - 0.5f, ky = float(2 * offy[i] - opy) - 0.5f;
        dvx[i] = kx + c.d2x[i];
        dvy[i] = ky + c.d2y[i];
        p[i] = !(dvx[i] * dvx[i] + dvy[i] * dvy[i] > c.thr[i]);            // == !(length(d) > radius * 2^-level)
        e[i] = p[i] ? c.depth[i] : pinf();
    }
    // :113-126 depth range check: maxDepth of the last tap attaining the minimum depth over the surviving taps
    const float mn = fminf(fminf(e[0], e[1]), fminf(e[2], e[3]));
    float maxDepth = c.maxDepth[0];
#pragma unroll
    for (int i = 1; i < 4; i++) if (e[i] == mn) maxDepth = c.maxDepth[i];
    bool any = false, occluded = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        p[i] = p[i] && !(c.depth[i] > maxDepth);
        any = any || p[i];                                                 // weightAcc > 0                     (:158)
        occluded = occluded || (p[i] && orig.x > c.maxDepth[i]);           // :153-156
    }
    const bool invalid = orig.x == 1.0f;                                   // :151 (never true at level 0: depth24 / 2^24 < 1)
    if (!any || (!invalid && !occluded)) return orig;                      // :159-161
    // weighted means (:129-147); a zero weight adds +-0 to a sum that starts at +0, which leaves it unchanged, so the
    // terms of dropped taps are skipped
    float depthAcc = 0.0f;
    int wi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) if (p[i]) { const int d = (i - OP) & 3; wi += d == 0 ? 9 : (d == 2 ? 1 : 3); depthAcc += c.depth[i] * w[i]; }
    const float wa = float(wi), y = rcp[wi];
    if (LEVEL0) return make_float4(div_small_unchecked(depthAcc, wa, y), 0.0f, 0.0f, 0.0f);
    float radiusAcc = 0.0f, maxDepthAcc = 0.0f, dax = 0.0f, day = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) if (p[i]) {
        maxDepthAcc += c.maxDepth[i] * w[i];
        radiusAcc += c.radius[i] * w[i];
        dax += dvx[i] * w[i]; day += dvy[i] * w[i];
    }
    // (same ranges as in pull_combine: depth / maxDepth sums are 0 or in [2^-24, 40], displacement sums multiples of 2^-25)
    return make_float4(div_small_unchecked(depthAcc, wa, y), div_small_unchecked(maxDepthAcc, wa, y), div_small(radiusAcc, wa, y),
                       u2f(pack_half2x(div_small_unchecked(dax, wa, y), div_small_unchecked(day, wa, y))));
}
template <int OP, bool LEVEL0>
PP_HD float4 push_pixel_c(const PushTaps& c, float4 orig, const float* rcp) { return push_pixel<LEVEL0>(OP, c, orig, rcp); }

// Level 0 (push.comp with LEVEL_ZERO): only the depth of the result is consumed, as R16 (:29-41). The squared components of
// displacementVectors take two values per tap and axis over the four pixels of a thread (the constant part is +-0.5 / +-1.5),
// so they are squared once per thread: sq[i][0..1] for offsets[OP] = 0 / 1. Returns the R16 code.
//  * orig = depth24 / 2^24 < 1, so `invalid` (:151) is never true and any surviving tap with orig > maxDepth replaces it;
//  * the weighted mean of depths in [0, 1) evaluated in fp32 lies in [0, 1] (partial sums <= 16, quotient <= 1), so the
//    clamp of the unorm conversion is the identity and is not evaluated; depthAcc is 0 or >= 2^-28, inside div_small's range.
struct PushTaps0 { float depth[4], maxDepth[4], thr[4], sqx[4][2], sqy[4][2]; };
template <int OP>
PP_HD uint32_t push_pixel0(const PushTaps0& c, uint32_t atlas_texel, const float* rcp) {
    constexpr int opx = OP >> 1, opy = (OP == 1 || OP == 2) ? 1 : 0;       // offsets[OP]
    const float orig = float(atlas_texel >> 8) / 16777216.0f;              // :19-27
    float e[4];
    bool p[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {                                          // :96-110
        p[i] = !(c.sqx[i][opx] + c.sqy[i][opy] > c.thr[i]);
        e[i] = p[i] ? c.depth[i] : pinf();
    }
    const float mn = fminf(fminf(e[0], e[1]), fminf(e[2], e[3]));          // :113-126
    float maxDepth = c.maxDepth[0];
#pragma unroll
    for (int i = 1; i < 4; i++) if (e[i] == mn) maxDepth = c.maxDepth[i];
    bool occluded = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        p[i] = p[i] && !(c.depth[i] > maxDepth);
        occluded = occluded || (p[i] && orig > c.maxDepth[i]);             // :153-156
    }
    float res = orig;                                                      // :159-161
    if (occluded) {
        float depthAcc = 0.0f;
        int wi = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) if (p[i]) {
            constexpr float WX[4] = {9.0f, 3.0f, 1.0f, 3.0f};
            const float w = WX[(i - OP + 4) & 3];
            wi += (int)w; depthAcc += c.depth[i] * w;
        }
        res = div_small_unchecked(depthAcc, float(wi), rcp[wi]);           // :144
    }
#if defined(__CUDA_ARCH__)
    return __float2uint_rn(res * 65535.0f);
#else
    return (uint32_t)nearbyintf(res * 65535.0f);
#endif
}

// threshold a texel carries when it is used as a coarse tap by the push pass that writes `consumer_level`
PP_HD float tap_threshold(float depth, float radius, int consumer_level) {
    if (depth == 1.0f) return -1.0f;                                        // push.comp:88-93
    return sqrt_gt_threshold(radius * u2f((uint32_t)(127 - consumer_level) << 23));   // radius * 2^-level (:106)
}

// ======================================================================================================================
// phases (tid in [0, kThreads)); the caller puts a CTA barrier between consecutive phases
// ======================================================================================================================
// tables + the result of an all-empty 2x2 block. The latter does not depend on the block's position: every term of
// pull.comp:69 is exact, so the four displacement vectors are (+-0.25, +-0.25) everywhere; it is evaluated once per CTA
// by the general code (one thread; none of this reads the tile, so it runs while the tile load is in flight).
PP_HD void phase_tables(Shared& s, int tid, float zFar, float ismPixelSize, int gx0, int gy0) {
    const float r = float(tid) / 10.0f;                // kThreads == 256 entries
    s.lut_r[tid] = r;
    s.lut_md[tid] = (r * 2.0f) / zFar * 0.6f;
    if (tid >= 1 && tid <= 16) s.rcp[tid] = 1.0f / float(tid);
    if (tid == 0) s.count = 0;
    if (tid == 32) {                                   // first lane of the second warp
        const float r255 = 255.0f / 10.0f;
        PullSample smp[4];
        float rc[5];
        for (int i = 1; i <= 4; i++) rc[i] = 1.0f / float(i);
        for (int i = 0; i < 4; i++) decode_l0(r255, (r255 * 2.0f) / zFar * 0.6f, 0xFFFFFFFFu, zFar, ismPixelSize, smp[i]);
        const int ox = gx0 >> 1, oy = gy0 >> 1;
        s.empty1 = pull_combine(smp, ox * 2, oy * 2, ox, oy, 1, rc);
    }
}

PP_HD void phase_pull1_classify(Shared& s, int tid) {
    for (int o = tid; o < 1024; o += kThreads) {
        const int lx = o & 31, ly = o >> 5;
        const uint32_t* r0 = &s.L0[(ly * 2) * 64 + lx * 2];
        const uint32_t* r1 = r0 + 64;
        const bool empty = (r0[0] & r0[1] & r1[0] & r1[1]) == 0xFFFFFFFFu;
        if (empty) s.PY[lvl_off(1) + o] = s.empty1;
#if defined(__CUDA_ARCH__)
        // one shared-memory atomic per warp (the trip count is warp-uniform: 1024 outputs / 256 threads); the order of the list is
        // immaterial, every entry is processed on its own
        const unsigned m = __ballot_sync(0xFFFFFFFFu, !empty);
        if (m) {
            const int lane = tid & 31, leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s.count, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!empty) s.list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)o;
        }
#else
        if (!empty) s.list[s.count++] = (uint16_t)o;
#endif
    }
}

PP_HD void phase_pull1_dense(Shared& s, int tid, float zFar, float ismPixelSize, int gx0, int gy0) {
    const int n = s.count;
    for (int q = tid; q < n; q += kThreads) {
        const int o = s.list[q];
        const int lx = o & 31, ly = o >> 5;
        PullSample smp[4];
        const int offx[4] = {0, 0, 1, 1}, offy[4] = {0, 1, 0, 1};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t u = s.L0[(ly * 2 + offy[i]) * 64 + lx * 2 + offx[i]];
            decode_l0(s.lut_r[u & 0xFFu], s.lut_md[u & 0xFFu], u, zFar, ismPixelSize, smp[i]);
        }
        const int ox = (gx0 >> 1) + lx, oy = (gy0 >> 1) + ly;
        s.PY[lvl_off(1) + o] = pull_combine(smp, ox * 2, oy * 2, ox, oy, 1, s.rcp);
    }
}

// ring of a threshold plane (a neighbouring ISM, push.comp:77-86; those CTAs own their texels): -1
PP_HD void fill_thr_ring(float* plane, int csz, int tid) {
    const int pitch = csz + 2;
    for (int q = tid; q < 4 * (csz + 1); q += kThreads) {
        const int side = q / (csz + 1), k = q - side * (csz + 1);
        const int px = side == 0 ? k : side == 1 ? csz + 1 : side == 2 ? csz + 1 - k : 0;
        const int py = side == 0 ? 0 : side == 1 ? k : side == 2 ? csz + 1 : csz + 1 - k;
        plane[py * kThrPitch + px] = -1.0f;
        (void)pitch;
    }
}

PP_HD void phase_pull(Shared& s, int tid, int level, int gx0, int gy0) {      // level 2..6 from level - 1
    const int osz = 64 >> level, isz = osz * 2;
    const float4* in = s.PY + lvl_off(level - 1);
    for (int o = tid; o < osz * osz; o += kThreads) {
        const int lx = o & (osz - 1), ly = o >> (6 - level);
        PullSample smp[4];
        const int offx[4] = {0, 0, 1, 1}, offy[4] = {0, 1, 0, 1};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float4 rd = in[(ly * 2 + offy[i]) * isz + lx * 2 + offx[i]];   // pull.comp:43-47
            smp[i].depth = rd.x; smp[i].maxDepth = rd.y; smp[i].radius = rd.z;
            const uint32_t pk = f2u_bits(rd.w);
            smp[i].dx = half_lo(pk); smp[i].dy = half_hi(pk);
        }
        const int ox = (gx0 >> level) + lx, oy = (gy0 >> level) + ly;
        const float4 res = pull_combine(smp, ox * 2, oy * 2, ox, oy, level, s.rcp);
        s.PY[lvl_off(level) + o] = res;
        if (level == 6) s.thr[5 & 1][1 * kThrPitch + 1] = tap_threshold(res.x, res.z, 5);   // coarse tap of push level 5
    }
    if (level == 6) fill_thr_ring(s.thr[5 & 1], 1, tid);
}

// One push pass over the tile, levels 5..1. Besides the texel, the pass stores the threshold the texel carries as a coarse
// tap of the next pass: in the ping-pong threshold plane for levels >= 2, and in place of the radius (.z, which push level
// 0 never reads) for level 1.
PP_HD void load_taps(const Shared& s, int level, int tlx, int tly, PushTaps& c) {
    const int csz = 32 >> level;
    const float4* coarse = s.PY + lvl_off(level + 1);
    const float* thr = s.thr[level & 1];
    const int x0 = tlx > 0 ? tlx - 1 : 0, x1 = tlx < csz ? tlx : csz - 1;   // clamped: an out-of-tile tap carries thr = -1
    const int y0 = tly > 0 ? tly - 1 : 0, y1 = tly < csz ? tly : csz - 1;
    const int cxs[4] = {x0, x0, x1, x1}, cys[4] = {y0, y1, y1, y0};         // offsets {0,0},{0,1},{1,1},{1,0}  (:48)
    const int pxs[4] = {tlx, tlx, tlx + 1, tlx + 1}, pys[4] = {tly, tly + 1, tly + 1, tly};   // ring-padded coordinates
#pragma unroll
    for (int i = 0; i < 4; i++) {                     // :55-62
        const float4 cs = coarse[cys[i] * csz + cxs[i]];
        c.depth[i] = cs.x; c.maxDepth[i] = cs.y; c.radius[i] = cs.z;
        const uint32_t pk = f2u_bits(cs.w);
        c.d2x[i] = half_lo(pk) * 2.0f; c.d2y[i] = half_hi(pk) * 2.0f;
        c.thr[i] = thr[pys[i] * kThrPitch + pxs[i]];
    }
}
PP_HD void store_mid(Shared& s, int level, int plx, int ply, float4 res) {
    const int size = 64 >> level;
    const float T = tap_threshold(res.x, res.z, level - 1);
    if (level == 1) res.z = T;
    else s.thr[(level - 1) & 1][(ply + 1) * kThrPitch + plx + 1] = T;
    s.PY[lvl_off(level) + ply * size + plx] = res;
}
// levels 5..2 are small, so the pass is latency-bound: one thread per fine pixel (four times the threads of the shader's
// 4-pixels-per-thread layout, a quarter of the dependent work per thread). The same layout for level 1 (-DPP_PUSH1_PER_PIXEL: four
// balanced rounds of 256 threads instead of 289 threads, i.e. a second round with 33 of 256 threads working) was measured: 5.5 %
// more instructions, same time (171.8 vs 171.2 us) -- ncu attributes 38 % of the kernel's warp time to CTA-barrier waits, but those
// are the idle warps of the serial top of the pyramid (pull 3..6, push 5..3: at most 64 texels per phase), not this round.
PP_HD void phase_push_top(Shared& s, int tid, int level) {
    const int size = 64 >> level;
    for (int it = tid; it < size * size; it += kThreads) {
        const int plx = it & (size - 1), ply = it >> (6 - level);
        const int tlx = (plx + 1) >> 1, tly = (ply + 1) >> 1;     // the shader thread T that owns this pixel (:67)
        const int opx = (plx + 1) & 1, opy = (ply + 1) & 1;       // offsets[outputPixel]
        const int OP = opx == 0 ? (opy == 0 ? 0 : 1) : (opy == 1 ? 2 : 3);
        PushTaps c;
        load_taps(s, level, tlx, tly, c);
        store_mid(s, level, plx, ply, push_pixel<false>(OP, c, s.PY[lvl_off(level) + it], s.rcp));
    }
    if (level > 1) fill_thr_ring(s.thr[(level - 1) & 1], size, tid);     // (level 1 carries its thresholds in .z, no plane)
}
// level 1 (1024 texels): the shader's layout, taps loaded and unpacked once per four pixels
PP_HD void phase_push1(Shared& s, int tid) {
    const int nT = 17;
    for (int it = tid; it < nT * nT; it += kThreads) {
        const int tlx = it % nT, tly = it / nT;
        PushTaps c;
        load_taps(s, 1, tlx, tly, c);
        const int bx = tlx * 2 - 1, by = tly * 2 - 1;
#define PP_ONE(OP, OX, OY)                                                                                        \
        {                                                                                                         \
            const int plx = bx + OX, ply = by + OY;                                                               \
            if (plx >= 0 && ply >= 0 && plx < 32 && ply < 32)                                                     \
                store_mid(s, 1, plx, ply, push_pixel_c<OP, false>(c, s.PY[lvl_off(1) + ply * 32 + plx], s.rcp));   \
        }
        PP_ONE(0, 0, 0) PP_ONE(1, 0, 1) PP_ONE(2, 1, 1) PP_ONE(3, 1, 0)
#undef PP_ONE
    }
}

// push level 0: thread T owns the (up to) four fine pixels that share its four coarse texels (the shader's layout: the taps
// are loaded, unpacked and squared once per four pixels). Level-1 texels carry their threshold in .z.
PP_HD void phase_push0(Shared& s, int tid) {
    const int nT = 33;
    const float4* coarse = s.PY + lvl_off(1);
    for (int it = tid; it < nT * nT; it += kThreads) {
        const int tlx = it % nT, tly = it / nT;           // local T; coarserLowerLeftPixel = T - 1 (push.comp:46)
        const bool inx0 = tlx > 0, inx1 = tlx < 32, iny0 = tly > 0, iny1 = tly < 32;
        const int x0 = inx0 ? tlx - 1 : 0, x1 = inx1 ? tlx : 31;
        const int y0 = iny0 ? tly - 1 : 0, y1 = iny1 ? tly : 31;
        PushTaps0 c;
        const int cxs[4] = {x0, x0, x1, x1}, cys[4] = {y0, y1, y1, y0};         // offsets {0,0},{0,1},{1,1},{1,0}  (:48)
        const int offx[4] = {0, 0, 1, 1}, offy[4] = {0, 1, 1, 0};
        const bool ins[4] = {inx0 && iny0, inx0 && iny1, inx1 && iny1, inx1 && iny0};
#pragma unroll
        for (int i = 0; i < 4; i++) {                     // :55-62
            const float4 cs = coarse[cys[i] * 32 + cxs[i]];
            c.depth[i] = cs.x; c.maxDepth[i] = cs.y;
            c.thr[i] = ins[i] ? cs.z : -1.0f;             // a tap in a neighbouring ISM never contributes (:77-86)
            const uint32_t pk = f2u_bits(cs.w);
            const float d2x = half_lo(pk) * 2.0f, d2y = half_hi(pk) * 2.0f;     // coarse displacement * 2  (:101)
#pragma unroll
            for (int o = 0; o < 2; o++) {                 // (coarserTexCoord - thisTexCoord) + disp*2, squared
                const float dx = (float(2 * offx[i] - o) - 0.5f) + d2x, dy = (float(2 * offy[i] - o) - 0.5f) + d2y;
                c.sqx[i][o] = dx * dx; c.sqy[i][o] = dy * dy;
            }
        }
        const int bx = tlx * 2 - 1, by = tly * 2 - 1;     // tile-local pixel of outputPixel 0 (:67)
#define PP_ONE(OP, OX, OY)                                                                                        \
        {                                                                                                         \
            const int plx = bx + OX, ply = by + OY;                                                               \
            if (plx >= 0 && ply >= 0 && plx < 64 && ply < 64)         /* else: a neighbouring ISM's CTA owns it */ \
                s.L0[ply * 64 + plx] = push_pixel0<OP>(c, s.L0[ply * 64 + plx], s.rcp);                           \
        }
        PP_ONE(0, 0, 0) PP_ONE(1, 0, 1) PP_ONE(2, 1, 1) PP_ONE(3, 1, 0)
#undef PP_ONE
    }
}

}  // namespace pp
}  // namespace mlgi

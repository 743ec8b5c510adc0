// common.cuh -- shared declarations of libmlgi (B200 / sm_100a).
//
// Numerical contract (DESIGN.md "bit-exactness"): the translation unit exact.cu is
// compiled with -fmad=false and only uses the helpers below, which fix the evaluation
// order (glm's: ((x*x)+(y*y))+(z*z), (m0*x+m1*y)+(m2*z+m3*w)), IEEE division/sqrt
// (nvcc defaults -prec-div=true -prec-sqrt=true) and GLSL-style min/max ternaries, so the
// integer outputs (atlas, ISM R16, cluster ids, light lists) agree bit for bit with the CPU
// oracle. gather.cu (radiance, 1e-3 tolerance) is compiled with FMA contraction on.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

#include "../../include/mlgi.h"

namespace mlgi {

// ---------------------------------------------------------------------------------------
// column-major 4x4 passed by value to kernels
struct Mat4 { float m[16]; };

// Hoisted per-VPL record used by the gathering kernel (80 B, 5 x float4, in the order the kernel
// consumes it; lookAtRH rows s,u from ism_utils.glsl:5-13 and the atlas tile origin :35-38 are pure
// functions of the VPL):
//   r0 = (-pos.x, -pos.y, -pos.z, nrm.x)   r1 = (nrm.y, nrm.z, h s.x, h u.x)   r2 = (h s.y, h u.y, h s.z, h u.z)
//   r3 = (col.r, col.g, col.b, 0)          r4 = (Cx, Cy, ox, oy)
//   h = half the ISM edge in texels (32 by default; IsmMode::px / 2), (ox, oy) = ISM tile origin + h in texels,
//   Cx = 1.5*2^23 + ox, Cy likewise: a round-down FMA onto it yields floor(texel coordinate).
struct VplRecord { float4 r[5]; };

// n3: the scaled-ISM experiment knobs (ism.geom:37-46, ism.comp:31-40, final_gathering.comp:40-46), resolved on the host.
// Defaults: sampled = ism_count = N, S2 = ism_indices_1d(N), px = 64.
struct IsmMode {
    int N;              // totalVplCount
    int vpl_start;      // vplStartIndex
    int vpl_count;      // vplEndIndex - vplStartIndex
    int sampled;        // sampledVplCount = pointsOnlyIntoScaledISMs ? vplCount : totalVplCount   (number of point buckets)
    int ism_count;      // ismCount = scaleISMs ? vplCount : totalVplCount
    int S2;             // ismIndices1d(ismCount): ISMs per atlas row in this mode
    int scale, points_only;
    float px;           // ISM edge in texels = atlas edge / S2 (ImperfectShadowmap.cpp:187-191); a power of two
};

struct DevLists {              // what the gathering kernel needs to find a pixel's list
    const uint16_t* light_list_ids;  // clustered: [16][Y][X]
    const uint16_t* lists;           // [list][16][L+1]
    int X, Y;                        // clustered tile grid
    int TX;                          // tiled: tiles per row
};

// ---------------------------------------------------------------------------------------
// exact helpers (device)
#ifdef __CUDACC__
__device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }   // GLSL min
__device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }   // GLSL max
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 mul3s(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 div3s(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float3 neg3(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float3 normalize3(float3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return mul3s(a, inv); }
__device__ __forceinline__ float4 mat4_mul_vec4(const Mat4& M, float x, float y, float z, float w) {
    float4 r;
    r.x = (M.m[0] * x + M.m[4] * y) + (M.m[8] * z + M.m[12] * w);
    r.y = (M.m[1] * x + M.m[5] * y) + (M.m[9] * z + M.m[13] * w);
    r.z = (M.m[2] * x + M.m[6] * y) + (M.m[10] * z + M.m[14] * w);
    r.w = (M.m[3] * x + M.m[7] * y) + (M.m[11] * z + M.m[15] * w);
    return r;
}
__device__ __forceinline__ int f2i(float f) { return __float2int_rz(f); }          // NaN -> 0, saturating
__device__ __forceinline__ uint32_t f2u(float f) { return __float2uint_rz(f); }    // NaN/neg -> 0, saturating

// floatpacking.glsl:46-66
__device__ __forceinline__ float3 unpack3SN(float packed) {
    uint32_t u = __float_as_uint(packed);
    float a = float(u & 0x3FFu) / 1023.0f;
    float b = float((u >> 10) & 0x3FFu) / 1023.0f;
    float c = (float((u >> 20) & 0x3FFu) - 1.5f) / 1021.0f;
    return f3(a * 2.0f - 1.0f, b * 2.0f - 1.0f, c * 2.0f - 1.0f);
}
// floatpacking.glsl:26-44
__device__ __forceinline__ float pack3SN(float3 n) {
    float cx = n.x * 0.5f + 0.5f, cy = n.y * 0.5f + 0.5f, cz = n.z * 0.5f + 0.5f;
    uint32_t u = f2u(cx * 1023.0f + 0.5f);
    u |= f2u(cy * 1023.0f + 0.5f) << 10;
    u |= f2u(cz * 1021.0f + 1.5f) << 20;
    return __uint_as_float(u);
}
// RGB10_A2 unorm fetch * 2 - 1
__device__ __forceinline__ float3 decode_rgb10_sn(uint32_t n10) {
    return f3(float(n10 & 0x3FFu) / 1023.0f * 2.0f - 1.0f, float((n10 >> 10) & 0x3FFu) / 1023.0f * 2.0f - 1.0f,
              float((n10 >> 20) & 0x3FFu) / 1023.0f * 2.0f - 1.0f);
}
__device__ __forceinline__ uint32_t pack_unorm8(float c) { return __float2uint_rn(gclamp(c, 0.0f, 1.0f) * 255.0f) & 0xFFu; }
__device__ __forceinline__ uint32_t pack_half2(float x, float y) {
    return (uint32_t)__half_as_ushort(__float2half_rn(x)) | ((uint32_t)__half_as_ushort(__float2half_rn(y)) << 16);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t p) {
    return make_float2(__half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))), __half2float(__ushort_as_half((unsigned short)(p >> 16))));
}
// reprojection.glsl:6-14
__device__ __forceinline__ float linear_depth(float depthSample, float A, float B) {
    float z_n = 2.0f * depthSample - 1.0f;
    return -B / (A + z_n);
}
// reprojection.glsl:16-24
__device__ __forceinline__ float log_depth(float z, float A, float B) {
    float z_n = (1.0f / z * -B) - A;
    return (z_n + 1.0f) / 2.0f;
}
// fp32 -> unsigned small float (5-bit exponent, bias 15, MBITS mantissa bits, no sign), round to nearest even;
// negative and -0 -> 0, NaN -> NaN, values above the largest finite clamp to it (GL 4.3 spec 2.3.4.3 / 2.3.4.4)
template <int MBITS>
__host__ __device__ inline uint32_t pack_ufloat(uint32_t x) {
    const uint32_t max_finite = (30u << MBITS) | ((1u << MBITS) - 1u);
    if (x & 0x80000000u) return ((x & 0x7FFFFFFFu) > 0x7F800000u) ? ((31u << MBITS) | 1u) : 0u;   // negative: 0 (NaN stays NaN)
    const uint32_t e = x >> 23, m = x & 0x7FFFFFu;
    if (e == 0xFF) return m ? ((31u << MBITS) | 1u) : (31u << MBITS);
    const int eh = (int)e - 127 + 15;
    if (eh >= 31) return max_finite;
    uint32_t r;
    if (eh <= 0) {                                   // subnormal result
        if (eh < -MBITS) return 0u;
        const uint32_t mm = m | 0x800000u;
        const int shift = (23 - MBITS) + 1 - eh;     // >= 24 - MBITS
        r = mm >> shift;
        const uint32_t rem = mm & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1u))) r++;
    } else {
        const int shift = 23 - MBITS;
        r = ((uint32_t)eh << MBITS) | (m >> shift);
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1u))) r++;
    }
    return r > max_finite ? max_finite : r;
}
template <int MBITS>
__host__ __device__ inline uint32_t unpack_ufloat_bits(uint32_t v) {   // returns fp32 bits
    const uint32_t e = v >> MBITS, m = v & ((1u << MBITS) - 1u);
    if (e == 0) {
        if (m == 0) return 0u;
        int sh = 0; uint32_t mm = m;
        while (!(mm & (1u << MBITS))) { mm <<= 1; sh++; }
        return ((uint32_t)(127 - 15 + 1 - sh) << 23) | ((mm & ((1u << MBITS) - 1u)) << (23 - MBITS));
    }
    if (e == 31) return 0x7F800000u | (m << (23 - MBITS));
    return ((e - 15 + 127) << 23) | (m << (23 - MBITS));
}
// branch-free form of unpack_ufloat_bits for the hot consumers (blur, de-interleave): v holds only the 5 + MBITS bits.
// Normal codes re-bias the exponent with one add; subnormal codes (e == 0) are m * 2^-(14 + MBITS), an exact int -> float
// conversion times a power of two (no normalisation loop); e == 31 (inf / NaN) moves on to fp32's all-ones exponent.
// Checked against unpack_ufloat_bits for every code by tests/hostemu.
template <int MBITS>
__host__ __device__ inline float unpack_ufloat_fast(uint32_t v) {
    const uint32_t e = v >> MBITS, m = v & ((1u << MBITS) - 1u);
    uint32_t bits = (v << (23 - MBITS)) + (112u << 23);
    if (e == 31u) bits += 112u << 23;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(bits);
#else
    memcpy(&f, &bits, 4);
#endif
    const float sub = float(m) * (MBITS == 6 ? 9.5367431640625e-07f : 1.9073486328125e-06f);   // 2^-20, 2^-19
    return e == 0u ? sub : f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t pack_r11g11b10(float r, float g, float b) {
    return pack_ufloat<6>(__float_as_uint(r)) | (pack_ufloat<6>(__float_as_uint(g)) << 11) | (pack_ufloat<5>(__float_as_uint(b)) << 22);
}
#endif
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// kernel launchers (defined in exact.cu / gather.cu, called from api.cpp). All return the
// number of kernels they launched so the ctx can keep mlgi_launch_count honest.
struct SliceTable { float b[17]; };

int launch_vpl_process(const mlgi_rsm& rsm, int rsm_w, int rsm_h, int N, const Mat4& inv_biased_light_vp,
                       float lightIntensity, int shuffle, const int32_t* shuffled, const IsmMode& mode, int compact,
                       float* vpl, float* packed, VplRecord* records, cudaStream_t st);

int launch_point_generate(const mlgi_mesh& mesh, const IsmMode& mode, float tessLevelFactor, int max_points, float4* points, uint32_t* point_bucket,
                          uint32_t* generated, cudaStream_t st);
// inspection / injection helpers: bucket-grouped view <-> (point, bucket id) list
int launch_group_points(const float4* points, const uint32_t* point_bucket, const uint32_t* generated, int max_points, int N,
                        uint32_t* bucket_hist, uint32_t* bucket_cursor, uint32_t* bucket_start, float4* points_grouped, cudaStream_t st);
int launch_bucket_expand(const uint32_t* bucket_start, int N, uint32_t* point_bucket, cudaStream_t st);
void launch_atlas_clear(uint32_t* atlas, int S, int vpl_lo, int vpl_hi, cudaStream_t st);
int launch_splat(const float4* points, const uint32_t* point_bucket, const uint32_t* generated, int max_points, const VplRecord* records,
                 const IsmMode& mode, int S, float zFar, uint32_t* atlas, int vpl_lo, int vpl_hi, bool clear, cudaStream_t st);
int launch_sqrt_thresholds(const float* rs, int n, float* out, cudaStream_t st);
int launch_pullpush(const uint32_t* atlas, int S, float zFar, float ism_pixel_size, uint16_t* ism, int vpl_lo, int vpl_hi, cudaStream_t st);

int launch_clustering(const float* depth, int W, int H, float projA, float projB, const SliceTable& tab,
                      uint32_t* tile_masks, uint32_t* compact_ids, uint16_t* light_list_ids, uint32_t* num_used,
                      cudaStream_t st);
int launch_light_lists(const VplRecord* records, int N, const uint32_t* compact_ids, const uint32_t* num_used,
                       const Mat4& inv_vp, float projA, float projB, int W, int H,
                       const SliceTable& tab, int class_lo, int class_hi, uint16_t* lists, cudaStream_t st);
// a8 (tiled culling) = tile_bounds (14 numbers per 64x64 tile, streamed from the G-buffer) + tile_lists (per-class VPL lists from
// the bounds) [+ tile_order]. bounds: float[tiles][16]. Across GPUs every rank bounds its own tile rows and stores them into the
// peers' bounds buffers as well (BoundsOut::peers), double-buffered by frame parity like the class planes.
struct BoundsOut {
    float* local;                  // this rank's bounds buffer (parity 0)
    int num_peers;                 // 0 = single GPU
    float* peers[8];               // the OTHER ranks' bounds buffers (parity 0), peer-mapped
    size_t parity_stride;          // floats from parity buffer 0 to parity buffer 1
    const uint32_t* frame_ctr;     // null = no parity (always buffer 0)
};
int launch_tile_bounds(const float* depth, const uint32_t* face_normal, int W, int H, const Mat4& inv_vp, int ty0, int ystride, int nrows,
                       const BoundsOut& bo, cudaStream_t st);
int launch_tile_lists(const float* bounds, size_t parity_stride, const uint32_t* frame_ctr, int W, const VplRecord* records, int N, int class_lo, int class_hi, int ty0, int ystride, int nrows,
                      uint16_t* lists, uint32_t* tile_weight, cudaStream_t st);
int launch_tile_order(const uint32_t* tile_weight, int W, int N, int class_lo, int class_hi, int ty0, int nrows, uint32_t* tile_order, cudaStream_t st);
int launch_tile_order_ids(const uint32_t* tile_weight, const uint32_t* ids, int first, int ntiles, int N, int class_lo, int class_hi, uint32_t* tile_order, cudaStream_t st);

struct GatherArgs {
    const float* depth; const uint32_t* face_normal; int W, H;
    const VplRecord* records; int N; int S;
    const uint16_t* ism;
    cudaTextureObject_t ism_tex;   // R16 unorm, point sampling, clamp: the same atlas through the texture unit (0 = use LDG)
    DevLists lists; int cull_mode;
    Mat4 inv_vp;            // viewInverted * projectionInverted (GIStage.cpp:289)
    float projA, projB;     // P[2][2], P[3][2]
    SliceTable tab;
    float zFar, giScale /* giIntensityFactor / vplCount * 16 */, clampv;
    int enable_shadowing;
    int tex_float_coord;    // texture path, A/B: 1 = hand the unfloored texel coordinate to the unit (MLGI_ISM_FETCH=texf), 2 = integer texel coordinates (texi)
    int bulk_staging;       // stage the VPL records with cp.async.bulk + mbarrier, double-buffered (A/B: MLGI_FG_BULK)
    int max_ctas_per_sm;    // 0 = as compiled (10); smaller: cap the resident CTAs per SM (host-input pipeline, see launch_fg_mode)
    // host-input sharded frame: the G-buffer is not replicated -- tile row t lives in the memory of rank t mod gb_world
    // (peer-mapped) and a gathering CTA reads its tile's pixels from the owner over NVLink (8 B per pixel, once)
    int gb_world;                        // 0 = depth / face_normal above hold the whole image
    const float* gb_depth[8]; const uint32_t* gb_normal[8];
    const uint32_t* tile_order;   // optional permutation of the tiles (longest list first); null = row-major
    int tile_begin, tile_count;   // 64x64 tiles [tile_begin, tile_begin + tile_count) in row-major order (whole image by default)
    int class_lo, class_hi;
    float4* out_raster;     // may be null
    float4* out_planar;     // may be null; plane of class c starts at (c - planar_class0) * PH * PW
    int planar_class0;
    uint32_t* out_packed;          // may be null: R11F_G11F_B10F raster
    uint32_t* out_planar_packed;   // may be null: R11F_G11F_B10F class planes
    // multi-GPU exchange: the class planes are stored into exchange buffers -- this rank's own one (pull mode, num_peers = 1)
    // or every rank's, peer-mapped over NVLink (push mode). The buffers are double-buffered by frame parity; the frame
    // counter lives in device memory so that a captured CUDA graph of the frame is valid for every frame.
    int num_peers;                 // 0 = off
    int peer_packed;
    void* peer_planar[8];          // parity buffer 0; element = uint32 (gi_packed) or float4
    size_t peer_parity_stride;     // bytes from parity buffer 0 to parity buffer 1
    const uint32_t* frame_ctr;     // frames completed so far; the frame being produced writes buffer (ctr + 1) & 1
    unsigned long long* stats;  // null, or 2 counters (pairs listed / lit)
};
int launch_final_gather(const GatherArgs& a, cudaStream_t st);
int launch_composite_ambient(const float4* gi, const uint32_t* diffuse_rgba8, const float* occlusion, int n, float exposure,
                             uint32_t* out_srgb8, cudaStream_t st);
// gi (fp32 raster) or gi_packed (GL_R11F_G11F_B10F raster, takes precedence when not null)
int launch_gi_blur(const float4* gi, const uint32_t* gi_packed, const float* depth, const uint32_t* face_normal, int W, int H, float projA,
                   float projB, float4* out, cudaStream_t st);
float measure_fp32_peak_tflops(int device, cudaStream_t st);   // FFMA micro-benchmark: the FP32-pipe roofline denominator
int launch_planar_to_raster(const float4* planar, int W, int H, int class_lo, int class_hi, float4* raster, cudaStream_t st);
// one kernel = signal + wait + (pull over NVLink) + de-interleave; see gather.cu
struct ExchangeArgs {
    int W, H, world, rank, packed;
    const char* src[8];            // parity buffer 0 of the rank that owns a class: peer-mapped (pull) or this rank's own (push mode)
    size_t parity_stride;
    uint32_t* peer_flags[8];       // flag array of every rank (peer-mapped): slot [rank] <- frame id
    const uint32_t* flags;         // this rank's flag array
    uint32_t* frame_ctr;           // device frame counter (incremented by the last CTA)
    uint32_t* ticket;              // CTA completion counter
    uint32_t* error;               // set when a wait times out (a peer died)
    uint32_t* raster_packed;       // may be null (may be a mapped pinned-host pointer: mlgi_frame_sharded_host)
    float4* raster;                // may be null
    // image rows to assemble: row_blocks blocks of row_block rows, the first at row_begin, row_stride apart, clipped to H
    // (whole image: one block of H rows; a rank of the host-input sharded frame: its 64-row tile rows, world * 64 apart)
    int row_begin, row_block, row_stride, row_blocks;
    int signal_wait;               // 1: publish "my planes are written" (slot [rank]) and wait for every rank first (the one-kernel exchange)
    int advance;                   // 1: the last CTA advances the frame counter
};
int launch_exchange(const ExchangeArgs& x, cudaStream_t st);
// host-input sharded frame: G-buffer slabs are pushed to the peers by the copy engines; these two order them
int launch_flag_signal(uint32_t* const* peer_flags, int world, int slot, const uint32_t* frame_ctr, uint32_t mul, uint32_t add, cudaStream_t st);
int launch_flag_wait(const uint32_t* flags, int world, const uint32_t* frame_ctr, uint32_t* error, uint32_t mul, uint32_t add, cudaStream_t st);
int launch_planar_packed_to_raster(const uint32_t* planar, int W, int H, uint32_t* raster_packed, float4* raster, cudaStream_t st);

constexpr int kTile = 64;  // tiled-culling tile edge in pixels (one gathering CTA = one class of one tile)

}  // namespace mlgi

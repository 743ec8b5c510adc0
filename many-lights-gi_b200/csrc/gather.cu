// gather.cu -- a9 final gathering (final_gathering.comp:63-135), n1 blur (gi_blur.frag),
// and the class-planar <-> raster remap. Radiance is under a 1e-3 relative contract, so this
// translation unit is compiled with FMA contraction on and uses rsqrt / rcp.
//
// Gathering design (DESIGN.md "final gathering"):
//   * one CTA = the 256 pixels of ONE interleave class inside one 64x64-pixel tile, so every
//     lane walks the same VPL list (final_gathering.comp:65-68 uses 8x8 pixels per workgroup);
//   * 128 threads x 2 pixels per thread: the VPL record is read once from shared memory
//     (broadcast) and used for two independent pixel chains;
//   * the list's hoisted VPL records (pos, nrm, lookAtRH rows, colour, atlas tile origin) are
//     staged through shared memory 64 at a time;
//   * ISM shadow lookups are 2-byte nearest-texel gathers straight from the L2-resident atlas;
//   * clustered mode walks the (<= 16) depth slices present in the tile and masks pixels.
#include "common.cuh"

namespace mlgi {

static constexpr int FG_THREADS = 128;
static constexpr int FG_CHUNK = 64;
static constexpr unsigned FULL = 0xFFFFFFFFu;

struct Pixel {
    float3 P;       // world position
    float3 n;       // face normal (10-bit decoded, not renormalised)
    float3 acc;
    int slice;      // clustered mode
    bool valid;
};

__device__ __forceinline__ void shade_pair(Pixel& px, const float4& r0, const float4& r1, const float4& r2, const float4& r3,
                                           const float4& r4, const uint16_t* __restrict__ ism, int A, float invZFar,
                                           float clampv, bool shadow, bool active, unsigned long long& listed,
                                           unsigned long long& lit, bool stats) {
    const float dx = px.P.x - r0.x, dy = px.P.y - r0.y, dz = px.P.z - r0.z;      // final_gathering.comp:102
    const float a = r1.x * dx + r1.y * dy + r1.z * dz;                             // dist * dot(vpl.normal, w)
    const float b = -(px.n.x * dx + px.n.y * dy + px.n.z * dz);                    // dist * dot(fragNormal, -w)
    if (stats && active) listed++;
    if (active && a > 0.0f && b > 0.0f) {                                          // :112-114
        if (stats) lit++;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float inv = rsqrtf(d2);
        const float inv2 = inv * inv;
        float g = fminf(a * b * inv2 * inv2 * inv2, clampv);                       // :115-117 (angle/d^2 * 1/d^4)
        if (shadow) {                                                              // :119-127
            const float dist = d2 * inv;
            const float sx = r2.x * dx + r2.y * dy + r2.z * dz;
            const float sy = r3.x * dx + r3.y * dy + r3.z * dz;
            const float q = __fdividef(32.0f, dist + a);                           // v.xy/(1 - v.z), v.z = -a/dist
            int tx = __float2int_rd(fmaf(sx, q, r0.w));                            // tile_x*64 + 32 + 32*px
            int ty = __float2int_rd(fmaf(sy, q, r1.w));
            tx = min(max(tx, 0), A - 1);                                           // CLAMP_TO_EDGE, NEAREST
            ty = min(max(ty, 0), A - 1);
            const float occ = float(__ldg(ism + (size_t)ty * A + tx)) * (1.0f / 65535.0f);
            float t = (0.02f - dist * invZFar + occ) * 50.0f;                      // smoothstep(0.98, 1, 1 - (z - occ))
            t = __saturatef(t);
            g *= t * t * (3.0f - 2.0f * t);
        }
        px.acc.x = fmaf(r4.x, g, px.acc.x);                                        // :129
        px.acc.y = fmaf(r4.y, g, px.acc.y);
        px.acc.z = fmaf(r4.z, g, px.acc.z);
    }
}

template <int MODE, bool SHADOW, bool STATS>
__global__ void __launch_bounds__(FG_THREADS) final_gather_kernel(const GatherArgs a) {
    __shared__ float4 srec[FG_CHUNK * 5];
    __shared__ uint32_t s_slices;

    const int ncls = a.class_hi - a.class_lo;
    const int cls = a.class_lo + (int)(blockIdx.x % ncls);
    const int tile = blockIdx.x / ncls;
    const int TX = (a.W + kTile - 1) / kTile;
    const int tile_x = tile % TX, tile_y = tile / TX;
    const int cxo = cls & 3, cyo = cls >> 2;                    // class = (x&3) + 4*(y&3), final_gathering.comp:67,91
    const int L = a.N / 16;
    const int A = 64 * a.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pi = 8 * (warp & 1) + (lane & 7);
    const int pj0 = 8 * (warp >> 1) + (lane >> 3);

    if (MODE == MLGI_CULL_CLUSTERED) { if (threadIdx.x == 0) s_slices = 0; __syncthreads(); }

    Pixel px[2];
    int sx[2], sy[2];
    uint32_t my_slices = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int x = tile_x * kTile + 4 * pi + cxo, y = tile_y * kTile + 4 * (pj0 + 4 * k) + cyo;
        sx[k] = x; sy[k] = y;
        Pixel& p = px[k];
        p.valid = x < a.W && y < a.H;
        p.acc = make_float3(0.0f, 0.0f, 0.0f);
        p.slice = 0;
        p.P = make_float3(0.0f, 0.0f, 0.0f);
        p.n = make_float3(0.0f, 0.0f, 0.0f);
        if (p.valid) {
            const float depthSample = a.depth[(size_t)y * a.W + x];                // :73
            // World position and normal are evaluated with the oracle's exact operation order (IEEE
            // mul/add/div, no FMA): unprojecting a window depth near 1 cancels catastrophically, so a
            // differently rounded P would move every pair of this pixel by far more than 1e-3.
            const float nx = __fadd_rn(__fmul_rn(__fdiv_rn(float(x), float(a.W)), 2.0f), -1.0f);   // :70,74
            const float ny = __fadd_rn(__fmul_rn(__fdiv_rn(float(y), float(a.H)), 2.0f), -1.0f);
            const float nz = __fadd_rn(__fmul_rn(depthSample, 2.0f), -1.0f);
            const float* m = a.inv_vp.m;
            float w4[4];
#pragma unroll
            for (int r = 0; r < 4; r++)                                            // :75  (m0*x + m1*y) + (m2*z + m3*1)
                w4[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[r], nx), __fmul_rn(m[4 + r], ny)), __fadd_rn(__fmul_rn(m[8 + r], nz), m[12 + r]));
            p.P = make_float3(__fdiv_rn(w4[0], w4[3]), __fdiv_rn(w4[1], w4[3]), __fdiv_rn(w4[2], w4[3]));   // :76
            const uint32_t n10 = a.face_normal[(size_t)y * a.W + x];
            p.n = make_float3(__fadd_rn(__fmul_rn(__fdiv_rn(float(n10 & 0x3FFu), 1023.0f), 2.0f), -1.0f),
                              __fadd_rn(__fmul_rn(__fdiv_rn(float((n10 >> 10) & 0x3FFu), 1023.0f), 2.0f), -1.0f),
                              __fadd_rn(__fmul_rn(__fdiv_rn(float((n10 >> 20) & 0x3FFu), 1023.0f), 2.0f), -1.0f));   // :79
            if (MODE == MLGI_CULL_CLUSTERED) {
                // the slice decides WHICH list is read: keep it exact (same table + IEEE ops as clustering)
                const float zlin = __fdiv_rn(-a.projB, __fadd_rn(a.projA, __fadd_rn(__fmul_rn(2.0f, depthSample), -1.0f)));   // :84
                const float d = -zlin;
                int s = 0;
#pragma unroll
                for (int i = 1; i <= 15; i++) if (d >= a.tab.b[i]) s = i;          // :85
                p.slice = s;
                my_slices |= 1u << s;
            }
        }
    }
    uint32_t slices = 1u;
    if (MODE == MLGI_CULL_CLUSTERED) {
        my_slices = __reduce_or_sync(FULL, my_slices);
        if (lane == 0 && my_slices) atomicOr(&s_slices, my_slices);
        __syncthreads();
        slices = s_slices;
    }

    unsigned long long listed = 0, lit = 0;
    const float invZFar = 1.0f / a.zFar;

    while (slices) {
        const int s = __ffs(slices) - 1;
        slices &= slices - 1;
        // ---- locate this (volume, class) sub-list ----
        const uint16_t* sub = nullptr;
        int count = L;
        if (MODE == MLGI_CULL_CLUSTERED) {
            const int ccx = (tile_x * kTile) / 128, ccy = (tile_y * kTile) / 128;   // :88
            const uint32_t listId = a.lists.light_list_ids[((size_t)s * a.lists.Y + ccy) * a.lists.X + ccx];   // :89
            sub = a.lists.lists + ((size_t)listId * 16 + cls) * (L + 1);
            count = (int)sub[0] - 1;                                               // :94
        } else if (MODE == MLGI_CULL_TILED) {
            sub = a.lists.lists + ((size_t)tile * 16 + cls) * (L + 1);
            count = (int)sub[0] - 1;
        }
        const bool act0 = px[0].valid && (MODE != MLGI_CULL_CLUSTERED || px[0].slice == s);
        const bool act1 = px[1].valid && (MODE != MLGI_CULL_CLUSTERED || px[1].slice == s);

        for (int base = 0; base < count; base += FG_CHUNK) {
            const int n = min(FG_CHUNK, count - base);
            __syncthreads();                                    // previous chunk fully consumed
            // stage n records: 5 float4 each, 128 threads copy float4 granules
            for (int q = threadIdx.x; q < n * 5; q += FG_THREADS) {
                const int e = q / 5, part = q - e * 5;
                const uint32_t id = sub ? (uint32_t)sub[1 + base + e] : (uint32_t)(cls * L + base + e);   // :98
                srec[q] = __ldg(&a.records[id].r[part]);
            }
            __syncthreads();
#pragma unroll 2
            for (int e = 0; e < n; e++) {                       // :97
                const float4 r0 = srec[e * 5 + 0], r1 = srec[e * 5 + 1], r2 = srec[e * 5 + 2], r3 = srec[e * 5 + 3], r4 = srec[e * 5 + 4];
                shade_pair(px[0], r0, r1, r2, r3, r4, a.ism, A, invZFar, a.clampv, SHADOW, act0, listed, lit, STATS);
                shade_pair(px[1], r0, r1, r2, r3, r4, a.ism, A, invZFar, a.clampv, SHADOW, act1, listed, lit, STATS);
            }
        }
    }

    // ---- epilogue: result = acc * giIntensityFactor / vplCount * 16  (:132) ----
    const int PW = (a.W + 3) >> 2, PH = (a.H + 3) >> 2;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (!px[k].valid) continue;
        const float4 o = make_float4(px[k].acc.x * a.giScale, px[k].acc.y * a.giScale, px[k].acc.z * a.giScale, 0.0f);
        if (a.out_raster) a.out_raster[(size_t)sy[k] * a.W + sx[k]] = o;
        if (a.out_planar) a.out_planar[((size_t)(cls - a.planar_class0) * PH + (sy[k] >> 2)) * PW + (sx[k] >> 2)] = o;
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            listed += __shfl_xor_sync(FULL, listed, o);
            lit += __shfl_xor_sync(FULL, lit, o);
        }
        if (lane == 0) { atomicAdd(&a.stats[0], listed); atomicAdd(&a.stats[1], lit); }
    }
}

template <int MODE>
static void launch_fg_mode(const GatherArgs& a, dim3 grid, cudaStream_t st) {
    const bool sh = a.enable_shadowing != 0, stt = a.stats != nullptr;
    if (sh && !stt) final_gather_kernel<MODE, true, false><<<grid, FG_THREADS, 0, st>>>(a);
    else if (sh && stt) final_gather_kernel<MODE, true, true><<<grid, FG_THREADS, 0, st>>>(a);
    else if (!sh && !stt) final_gather_kernel<MODE, false, false><<<grid, FG_THREADS, 0, st>>>(a);
    else final_gather_kernel<MODE, false, true><<<grid, FG_THREADS, 0, st>>>(a);
}

int launch_final_gather(const GatherArgs& a, cudaStream_t st) {
    const int TX = (a.W + kTile - 1) / kTile, TY = (a.H + kTile - 1) / kTile;
    const int ncls = a.class_hi - a.class_lo;
    dim3 grid((unsigned)(TX * TY * ncls));
    if (a.cull_mode == MLGI_CULL_CLUSTERED) launch_fg_mode<MLGI_CULL_CLUSTERED>(a, grid, st);
    else if (a.cull_mode == MLGI_CULL_TILED) launch_fg_mode<MLGI_CULL_TILED>(a, grid, st);
    else launch_fg_mode<MLGI_CULL_NONE>(a, grid, st);
    return 1;
}

// =====================================================================================
// class-planar [cls][ceil(H/4)][ceil(W/4)] -> raster W x H (consumer-side de-interleave after
// the multi-GPU allgather)
// =====================================================================================
__global__ void planar_to_raster_kernel(const float4* __restrict__ planar, int W, int H, int class_lo, int class_hi, float4* __restrict__ raster) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const int cls = (x & 3) + 4 * (y & 3);
    if (cls < class_lo || cls >= class_hi) return;
    const int PW = (W + 3) >> 2, PH = (H + 3) >> 2;
    raster[(size_t)y * W + x] = planar[((size_t)(cls - class_lo) * PH + (y >> 2)) * PW + (x >> 2)];
}
int launch_planar_to_raster(const float4* planar, int W, int H, int class_lo, int class_hi, float4* raster, cudaStream_t st) {
    planar_to_raster_kernel<<<dim3((W + 255) / 256, H), 256, 0, st>>>(planar, W, H, class_lo, class_hi, raster);
    return 1;
}

// =====================================================================================
// FP32-pipe peak: 8 independent FFMA chains per thread, enough CTAs to fill every SM.
// =====================================================================================
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456f) out[0] = r;
}
float measure_fp32_peak_tflops(cudaStream_t st) {
    float* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0f;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = sms * 8;
    fp32_peak_kernel<<<blocks, 256, 0, st>>>(d, 1024, 0.999f, 0.001f);   // warm-up
    float best = 0.0f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, st);
        fp32_peak_kernel<<<blocks, 256, 0, st>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * double(iters) * 256.0 * double(blocks);
        best = fmaxf(best, float(flops / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

// =====================================================================================
// n1: gi_blur.frag:24-72, two separable 7-tap passes (GIStage.cpp:308-344). Out-of-image taps
// are skipped (texelFetch outside the image is undefined in GL; same choice as the oracle).
// =====================================================================================
__global__ void gi_blur_kernel(const float4* __restrict__ gi, const float* __restrict__ depth, const uint32_t* __restrict__ face_normal,
                               int W, int H, float projA, float projB, int dxs, int dys, float4* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    auto lin = [&](float d) { return -projB / (projA + (2.0f * d - 1.0f)); };          // reprojection.glsl:6-14
    auto nrm = [&](uint32_t n10) {
        return make_float3(float(n10 & 0x3FFu) * (2.0f / 1023.0f) - 1.0f, float((n10 >> 10) & 0x3FFu) * (2.0f / 1023.0f) - 1.0f,
                           float((n10 >> 20) & 0x3FFu) * (2.0f / 1023.0f) - 1.0f);
    };
    const size_t c = (size_t)y * W + x;
    const float d = lin(depth[c]);                              // :58
    const float3 Nn = nrm(face_normal[c]);                      // :59
    float4 g0 = gi[c];
    float3 acc = make_float3(g0.x, g0.y, g0.z);                 // :60
    float factorAcc = 1.0f;
#pragma unroll
    for (int sgn = -1; sgn <= 1; sgn += 2)
#pragma unroll
        for (int i = 1; i <= 3; i++) {                          // :63-69
            const int tx = x + sgn * dxs * i, ty = y + sgn * dys * i;
            if (tx < 0 || ty < 0 || tx >= W || ty >= H) continue;
            const size_t t = (size_t)ty * W + tx;
            const float4 g = gi[t];
            const float3 ns = nrm(face_normal[t]);
            const float normalFactor = 1.0f - fmaxf(0.0f, Nn.x * ns.x + Nn.y * ns.y + Nn.z * ns.z);   // :44
            const float depthDiff = lin(depth[t]) - d;
            const float nd = normalFactor * 5.0f;
            const float w = fmaxf(0.0f, (3.0f - float(i) + 1.0f) / 4.0f - depthDiff * depthDiff - nd * nd);   // :24-35
            factorAcc += w;
            acc.x += g.x * w; acc.y += g.y * w; acc.z += g.z * w;
        }
    out[c] = make_float4(acc.x / factorAcc, acc.y / factorAcc, acc.z / factorAcc, 0.0f);   // :71
}
int launch_gi_blur(const float4* gi, const float* depth, const uint32_t* face_normal, int W, int H, float projA, float projB,
                   float4* tmp, float4* out, cudaStream_t st) {
    dim3 grid((W + 255) / 256, H);
    gi_blur_kernel<<<grid, 256, 0, st>>>(gi, depth, face_normal, W, H, projA, projB, 1, 0, tmp);   // GIStage.cpp:427 DIRECTION (1,0)
    gi_blur_kernel<<<grid, 256, 0, st>>>(tmp, depth, face_normal, W, H, projA, projB, 0, 1, out);  // :431 DIRECTION (0,1)
    return 2;
}

}  // namespace mlgi

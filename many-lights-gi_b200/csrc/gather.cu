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
#include "bulk_async.cuh"

#include <algorithm>

namespace mlgi {

// tuning constants (A/B measured at 4K/4096 on B200): 10 resident CTAs per SM (48 registers, no spills) 1.571 vs 1.593 ms at 9;
// 128-VPL staging chunks (half as many CTA barriers) 1.563 vs 1.578 ms at 64
#ifndef FG_MINB
#define FG_MINB 10
#endif
// record staging buffers of 64 VPLs each (bulk path): a ring; a buffer is refilled behind a CTA barrier once every warp is done
// with it. With 4 buffers a list of up to 256 VPLs (N <= 4096) is staged without any reuse, i.e. without a CTA barrier in the
// VPL loop at all -- measured (ncu with 2 buffers: 9.5 % of the warp time in barrier stalls; A/B at 4K/4096): 1.460 / 1.461 /
// 1.458 ms with 2 / 3 / 4 buffers. The stalled warps were not needed: the kernel is bound by issue slots and the FMA pipe, so
// the smaller footprint (10 KB per CTA, more L1 left for the ISM texels) stays the default.
#ifndef FG_NBUF
#define FG_NBUF 2
#endif
static constexpr int FG_CHUNK = 64 * FG_NBUF;   // also the chunk of the LDG -> STS staging path
// unroll factor of the VPL loop: the loop control (uniform-datapath counter, compare, branch, address update) is amortised over the
// unrolled iterations. Measured at 4K/4096: 1.485 / 1.459 / 1.436 / 1.429 / 1.440 / 1.429 / 1.441 ms for 1 / 2 / 3 / 4 / 6 / 8 / 16.
#ifndef FG_UNROLL
#define FG_UNROLL 4
#endif
#define FG_STR2(x) #x
#define FG_STR(x) FG_STR2(x)
#define FG_UNROLL_PRAGMA _Pragma(FG_STR(unroll FG_UNROLL))
static constexpr int FG_REC_Q = 5;       // float4 per staged VPL (same layout as VplRecord)
static constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- packed fp32x2 helpers (Blackwell FFMA2 / FADD2 / FMUL2: one issue slot for two pixels) ----
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 dup(float v) { return make_float2(v, v); }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsq_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float r; asm("fma.rn.sat.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float2 lo2(const float4& q) { return make_float2(q.x, q.y); }
__device__ __forceinline__ float2 hi2(const float4& q) { return make_float2(q.z, q.w); }

// -t*t*(3 - 2t) / 2 = t*t*(t - 1.5): the addition takes an immediate, where 3 - 2t needs a register holding the 2 (a
// per-iteration IMAD.MOV at 48 registers). Negating and halving are exact (t - 1.5 is the correctly rounded 3 - 2t scaled by
// -1/2, and round-to-nearest is symmetric), so the shadowed loop accumulates exactly -1/2 of the reference's sums and the
// epilogue undoes it.
__device__ __forceinline__ float2 neg_half_smoothstep(float2 t) { return mul2(mul2(t, t), add2(t, dup(-1.5f))); }

// nearest-texel fetch with INTEGER texel coordinates (PTX tex.2d with .s32 coordinates: unnormalised, no filtering, the
// sampler's clamp-to-edge applies): the floored coordinate leaves the round-down FMA as the bit pattern 0x4B400000 + k, so
// k costs one integer subtraction on the ALU pipe instead of a packed FP32 subtraction on the FMA pipe (A/B: MLGI_ISM_FETCH=texi)
__device__ __forceinline__ float tex2d_s32(cudaTextureObject_t t, int x, int y) {
    float r, g, b, a;
    asm("tex.2d.v4.f32.s32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(r), "=f"(g), "=f"(b), "=f"(a) : "l"(t), "r"(x), "r"(y));
    return r;
}

struct Normal2 { float2 x, y, z; };   // negated face normals of the two pixels; zero for a pixel that is not active

// Two pixels of one thread, component-wise packed: .x = pixel 0, .y = pixel 1
struct PixelPair {
    float2 Px, Py, Pz;      // world position
    float2 Nx, Ny, Nz;      // NEGATED face normal (10-bit decoded, not renormalised): b = N . d
    float2 ar, ag, ab;      // accumulators
    int slice0, slice1;     // clustered mode
    bool valid0, valid1;
};

// One staged VPL = VplRecord (5 float4 in shared memory, read by broadcast):
//   r0 = (-px,-py,-pz,nx) r1 = (ny,nz,32sx,32ux) r2 = (32sy,32uy,32sz,32uz) r3 = (cr,cg,cb,0) r4 = (Cx,Cy,0,0)
//   Cx = 1.5*2^23 + tile_x*64 + 32: adding it with round-down leaves floor(texel x) in the mantissa.
// VPL scalars enter the packed instructions through their scalar-broadcast operand form (R.F32).
template <bool SHADOW, bool STATS, int TEX, int NP>
__device__ __forceinline__ void shade_vpl(PixelPair (&px)[NP], const float4* __restrict__ rec, const uint16_t* __restrict__ ism,
                                          cudaTextureObject_t ism_tex, int A, float kz, float clampv, const Normal2 (&N)[NP], int nact,
                                          unsigned long long& listed, unsigned long long& lit) {
    const float4 r0 = rec[0], r1 = rec[1];
    float2 dx[NP], dy[NP], dz[NP], a[NP], b[NP];
    float m0[NP], m1[NP];
    float mmax = -1.0f;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        dx[p] = add2(px[p].Px, dup(r0.x)); dy[p] = add2(px[p].Py, dup(r0.y)); dz[p] = add2(px[p].Pz, dup(r0.z));   // final_gathering.comp:102
        a[p] = fma2(dup(r1.y), dz[p], fma2(dup(r1.x), dy[p], mul2(dup(r0.w), dx[p])));   // dist * dot(vpl.normal, w)
        b[p] = fma2(N[p].z, dz[p], fma2(N[p].y, dy[p], mul2(N[p].x, dx[p])));            // dist * dot(fragNormal, -w)
        // angleFactor > 0  <=>  min(a, b) > 0   (:112-114); inactive pixels carry a zero normal, so b == 0
        m0[p] = fminf(a[p].x, b[p].x); m1[p] = fminf(a[p].y, b[p].y);
        mmax = fmaxf(mmax, fmaxf(m0[p], m1[p]));
        if (STATS) lit += (m0[p] > 0.0f ? 1 : 0) + (m1[p] > 0.0f ? 1 : 0);
    }
    if (STATS) listed += nact;
    // The body runs when ANY pixel of the warp is lit. Written as a vote, the branch is warp-uniform for the compiler too: no
    // BSSY / BSYNC reconvergence pair around the body, and the header loads the record's second quad whole (5 instead of 6 LDS per
    // VPL): 68 -> 66 instructions per VPL, 1.430 -> 1.415 ms. Lanes without a lit pixel now execute the body as well (they were
    // idle before, the warp pays the same); their contribution is zeroed by the same select that zeroes the unlit pixel of a
    // lit pair. -DFG_DIVERGENT_BRANCH restores the per-thread branch.
#ifndef FG_DIVERGENT_BRANCH
    if (__any_sync(FULL, mmax > 0.0f)) {
#else
    if (mmax > 0.0f) {
#endif
        const float4 r3 = rec[3];
        float2 su_x, su_y, su_z, cxy, cxy_f;
        if (SHADOW) {
            const float4 r2 = rec[2];
            su_x = make_float2(r1.z, r1.w); su_y = make_float2(r2.x, r2.y); su_z = make_float2(r2.z, r2.w);
            if (TEX == 2) cxy_f = *(reinterpret_cast<const float2*>(rec + 4) + 1);     // tile origin + 32
            else cxy = *reinterpret_cast<const float2*>(rec + 4);                      // 1.5*2^23 + tile origin + 32
        }
#pragma unroll
        for (int p = 0; p < NP; p++) {                  // NP independent chains: the compiler interleaves them (ILP)
            const float2 d2 = fma2(dz[p], dz[p], fma2(dy[p], dy[p], mul2(dx[p], dx[p])));
            const float2 inv = make_float2(rsq_approx(d2.x), rsq_approx(d2.y));
            const float2 inv2 = mul2(inv, inv);
            const float2 inv4 = mul2(inv2, inv2);
            float2 g = mul2(mul2(a[p], b[p]), mul2(inv4, inv2));                          // :115-116  angle/d^2 * 1/d^4
            g.x = fminf(g.x, clampv); g.y = fminf(g.y, clampv);                           // :117
            if (SHADOW) {                                                                 // :119-127
                const float2 dist = mul2(d2, inv);
                // paraboloid projection, packed by COORDINATE (lane x = 32 s.d, lane y = 32 u.d) with the pixel's offset as
                // the broadcast operand: (texel x, texel y) of one pixel land in one register pair, which is what TEX takes
                const float2 p0 = fma2(su_z, dup(dz[p].x), fma2(su_y, dup(dy[p].x), mul2(su_x, dup(dx[p].x))));
                const float2 p1 = fma2(su_z, dup(dz[p].y), fma2(su_y, dup(dy[p].y), mul2(su_x, dup(dx[p].y))));
                const float2 den = add2(dist, a[p]);                                      // dist * (1 - v.z), v.z = -a/dist
                const float q0 = rcp_approx(den.x), q1 = rcp_approx(den.y);
                const float2 tb = fma2(dist, dup(kz), dup(1.0f));                         // (0.02 - dist/zFar) * 50
                float2 t;
                if (TEX == 2) {
                    // the unfloored texel coordinate goes to the texture unit, which floors it itself (one packed FMA per
                    // pixel instead of a round-down FMA plus a subtraction)
                    const float2 x0 = fma2(p0, dup(q0), cxy_f), x1 = fma2(p1, dup(q1), cxy_f);
                    const float o0 = tex2DLod<float>(ism_tex, x0.x, x0.y, 0.0f);
                    const float o1 = tex2DLod<float>(ism_tex, x1.x, x1.y, 0.0f);
                    t.x = fma_sat(o0, 50.0f, tb.x); t.y = fma_sat(o1, 50.0f, tb.y);
                    g = mul2(g, neg_half_smoothstep(t));
                    g.x = m0[p] > 0.0f ? g.x : 0.0f; g.y = m1[p] > 0.0f ? g.y : 0.0f;
                    px[p].ar = fma2(dup(r3.x), g, px[p].ar);
                    px[p].ag = fma2(dup(r3.y), g, px[p].ag);
                    px[p].ab = fma2(dup(r3.z), g, px[p].ab);
                    continue;
                }
                const float2 f0 = __ffma2_rd(p0, dup(q0), cxy);                           // 1.5*2^23 + floor(texel x | y)
                const float2 f1 = __ffma2_rd(p1, dup(q1), cxy);
                if (TEX == 4) {
                    // measurement only (MLGI_ISM_FETCH=none, wrong image): no fetch at all, the coordinates still feed the result --
                    // what the kernel would cost if the shadow lookup had no latency and used no texture bandwidth
                    t.x = fma_sat(f0.x + f0.y, 1e-30f, tb.x); t.y = fma_sat(f1.x + f1.y, 1e-30f, tb.y);
                } else if (TEX == 3) {
                    const float o0 = tex2d_s32(ism_tex, __float_as_int(f0.x) - 0x4B400000, __float_as_int(f0.y) - 0x4B400000);
                    const float o1 = tex2d_s32(ism_tex, __float_as_int(f1.x) - 0x4B400000, __float_as_int(f1.y) - 0x4B400000);
                    t.x = fma_sat(o0, 50.0f, tb.x); t.y = fma_sat(o1, 50.0f, tb.y);
                } else if (TEX) {
                    // textureLod(ismDepthSampler, v.xy, 0) with NEAREST + CLAMP_TO_EDGE (final_gathering.comp:122) on the
                    // texture unit: address, clamp and R16-unorm -> float conversion cost no SM instructions. The coordinate
                    // handed over is the already floored texel index (an exact integer), so the unit's own rounding cannot
                    // pick a neighbouring texel.
                    const float2 k0 = add2(f0, dup(-12582912.0f)), k1 = add2(f1, dup(-12582912.0f));
                    // both fetches are unconditional: the pixel of the pair that is not lit has its g zeroed below and the
                    // texture unit clamps whatever coordinate it is given (no predicates, no per-branch handle reload)
                    // (tex2DLod with level 0 = TEX.LZ: no LOD register to materialise per fetch)
                    const float o0 = tex2DLod<float>(ism_tex, k0.x, k0.y, 0.0f);
                    const float o1 = tex2DLod<float>(ism_tex, k1.x, k1.y, 0.0f);
                    // smoothstep(0.98, 1, 1 - (z - occ)); scalar FFMA.SAT: the packed form has no saturate modifier
                    t.x = fma_sat(o0, 50.0f, tb.x); t.y = fma_sat(o1, 50.0f, tb.y);
                } else {
                    int tx0 = __float_as_int(f0.x) - 0x4B400000, ty0 = __float_as_int(f0.y) - 0x4B400000;
                    int tx1 = __float_as_int(f1.x) - 0x4B400000, ty1 = __float_as_int(f1.y) - 0x4B400000;
                    tx0 = min(max(tx0, 0), A - 1); ty0 = min(max(ty0, 0), A - 1);         // NEAREST, CLAMP_TO_EDGE
                    tx1 = min(max(tx1, 0), A - 1); ty1 = min(max(ty1, 0), A - 1);
                    const uint32_t u0 = __ldg(ism + (uint32_t)(ty0 * A + tx0));
                    const uint32_t u1 = __ldg(ism + (uint32_t)(ty1 * A + tx1));
                    // u16 -> float without the conversion pipe: 2^23 + u as bits, minus 2^23
                    const float2 occ = add2(make_float2(__uint_as_float(0x4B000000u | u0), __uint_as_float(0x4B000000u | u1)), dup(-8388608.0f));
                    t = fma2(occ, dup(50.0f / 65535.0f), tb);
                    t.x = __saturatef(t.x); t.y = __saturatef(t.y);
                }
                g = mul2(g, neg_half_smoothstep(t));                                             // :125-126
            }
            g.x = m0[p] > 0.0f ? g.x : 0.0f; g.y = m1[p] > 0.0f ? g.y : 0.0f;
            px[p].ar = fma2(dup(r3.x), g, px[p].ar);                                      // :129
            px[p].ag = fma2(dup(r3.y), g, px[p].ag);
            px[p].ab = fma2(dup(r3.z), g, px[p].ab);
        }
    }
}

// BULK: the list's records are staged by the copy engine instead of LDG -> STS -- one 80-byte cp.async.bulk per record,
// issued by one thread each and completing on an mbarrier, double-buffered in two 64-record halves of the staging buffer:
// chunk c + 1 lands while chunk c is shaded, the first two chunks of a tiled list are already in flight during the pixel
// prologue, and a chunk costs one CTA barrier instead of two.
// RGB (remote G-buffer): the tile's pixels are read from the rank that uploaded them (GatherArgs::gb_*); a separate
// instantiation, so that the single-GPU kernel's register budget (48, no spills) is untouched.
template <int MODE, bool SHADOW, bool STATS, int TEX, int NP, bool BULK, bool RGB = false>
__global__ void __launch_bounds__(256 / (2 * NP), FG_MINB) final_gather_kernel(const GatherArgs a) {
    constexpr int THREADS = 256 / (2 * NP);     // 256 pixels of one class per CTA, 2*NP pixels per thread
    constexpr int NPX = 2 * NP;
    constexpr int BCH = 64, NBUF = FG_NBUF;     // records per bulk-staged chunk, number of chunk buffers
    static_assert(!BULK || (THREADS >= BCH && NBUF >= 2 && NBUF <= 8), "bulk staging: one issuing thread per record, a ring of buffers");
    __shared__ __align__(16) float4 srec[FG_CHUNK * FG_REC_Q];
    __shared__ uint32_t s_slices;
    __shared__ __align__(8) uint64_t s_full[FG_NBUF];

    const int ncls = a.class_hi - a.class_lo;
    const int TX = (a.W + kTile - 1) / kTile;
    // CTA order: tile-major, class fastest (the class-CTAs of a tile share its G-buffer sectors in L2); the tiles themselves
    // in the order given by tile_order (longest list first) when the cull kernel provided one
    const int cls = a.class_lo + (int)(blockIdx.x % ncls);
    const int slot = a.tile_begin + (int)(blockIdx.x / ncls);
    const int tile = a.tile_order ? (int)a.tile_order[slot] : slot;
    const int tile_x = tile % TX, tile_y = tile / TX;
    const int cxo = cls & 3, cyo = cls >> 2;                    // class = (x&3) + 4*(y&3), final_gathering.comp:67,91
    const int L = a.N / 16;
    const int A = 64 * a.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // class-tile pixel (i, j) in 16x16. A warp owns 8 columns x 4 rows (an 8x4 patch = 32x16 screen pixels);
    // the 2*NP pixels of a thread are 4 rows apart. NP=1: 4 warps (2x2 quadrants). NP=2: 2 warps (8-column strips).
    const int pi = 8 * (warp & 1) + (lane & 7);
    const int pj0 = (NP == 1 ? 8 * (warp >> 1) : 0) + (lane >> 3);

    if (MODE == MLGI_CULL_CLUSTERED) { if (threadIdx.x == 0) s_slices = 0; }
    if (BULK && threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; b++) tma::mbar_init(&s_full[b], 1);
        tma::fence_barrier_init();
    }
    if (MODE == MLGI_CULL_CLUSTERED || BULK) __syncthreads();
    // bulk staging of chunk c of a list into buffer c & 1 (sub == nullptr: the class's whole VPL range, MLGI_CULL_NONE)
    auto issue_chunk = [&](const uint16_t* sub, int count, int c) {
        const int n = min(BCH, count - c * BCH);
        float4* dst = srec + (c % NBUF) * BCH * FG_REC_Q;
        if (threadIdx.x == 0) tma::mbar_arrive_expect_tx(&s_full[c % NBUF], (uint32_t)n * (uint32_t)sizeof(VplRecord));
        if ((int)threadIdx.x < n) {
            const uint32_t id = sub ? (uint32_t)sub[1 + c * BCH + threadIdx.x] : (uint32_t)(cls * L + c * BCH + threadIdx.x);   // :98
            tma::bulk_g2s(dst + threadIdx.x * FG_REC_Q, &a.records[id], (uint32_t)sizeof(VplRecord), &s_full[c % NBUF]);
        }
    };
    uint32_t full_phase = 0;                    // bit b: parity of the next completion of s_full[b]
    const uint16_t* sub_first = nullptr;
    int count_first = L;
    if (BULK && MODE != MLGI_CULL_CLUSTERED) {  // one list per CTA, known before any pixel work: start fetching it now
        if (MODE == MLGI_CULL_TILED) { sub_first = a.lists.lists + ((size_t)tile * 16 + cls) * (L + 1); count_first = (int)sub_first[0] - 1; }
#pragma unroll
        for (int c = 0; c < NBUF; c++) if (count_first > c * BCH) issue_chunk(sub_first, count_first, c);
    }

    // the tile's G-buffer: this rank's copy, or (host-input sharded frame) the memory of the rank that uploaded these rows
    const float* gdepth = a.depth;
    const uint32_t* gnormal = a.face_normal;
    if (RGB) { const int g = tile_y % a.gb_world; gdepth = a.gb_depth[g]; gnormal = a.gb_normal[g]; }   // tile row t lives on rank t mod world
    float P[NPX][3], Nn[NPX][3];
    int sx[NPX], sy[NPX], slice[NPX];
    bool valid[NPX];
    uint32_t my_slices = 0;
#pragma unroll
    for (int k = 0; k < NPX; k++) {
        const int x = tile_x * kTile + 4 * pi + cxo, y = tile_y * kTile + 4 * (pj0 + 4 * k) + cyo;
        sx[k] = x; sy[k] = y;
        valid[k] = x < a.W && y < a.H;
        slice[k] = 0;
        P[k][0] = P[k][1] = P[k][2] = 0.0f;
        Nn[k][0] = Nn[k][1] = Nn[k][2] = 0.0f;
        if (valid[k]) {
            const float depthSample = gdepth[(size_t)y * a.W + x];                 // :73
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
            P[k][0] = __fdiv_rn(w4[0], w4[3]); P[k][1] = __fdiv_rn(w4[1], w4[3]); P[k][2] = __fdiv_rn(w4[2], w4[3]);   // :76
            const uint32_t n10 = gnormal[(size_t)y * a.W + x];                     // :79 (negated: b = -n . d)
            Nn[k][0] = -__fadd_rn(__fmul_rn(__fdiv_rn(float(n10 & 0x3FFu), 1023.0f), 2.0f), -1.0f);
            Nn[k][1] = -__fadd_rn(__fmul_rn(__fdiv_rn(float((n10 >> 10) & 0x3FFu), 1023.0f), 2.0f), -1.0f);
            Nn[k][2] = -__fadd_rn(__fmul_rn(__fdiv_rn(float((n10 >> 20) & 0x3FFu), 1023.0f), 2.0f), -1.0f);
            if (MODE == MLGI_CULL_CLUSTERED) {
                // the slice decides WHICH list is read: keep it exact (same table + IEEE ops as clustering)
                const float zlin = __fdiv_rn(-a.projB, __fadd_rn(a.projA, __fadd_rn(__fmul_rn(2.0f, depthSample), -1.0f)));   // :84
                const float d = -zlin;
                int s = 0;
#pragma unroll
                for (int i = 1; i <= 15; i++) if (d >= a.tab.b[i]) s = i;          // :85
                slice[k] = s;
                my_slices |= 1u << s;
            }
        }
    }
    PixelPair px[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const int k0 = 2 * p, k1 = 2 * p + 1;
        px[p].Px = make_float2(P[k0][0], P[k1][0]); px[p].Py = make_float2(P[k0][1], P[k1][1]); px[p].Pz = make_float2(P[k0][2], P[k1][2]);
        px[p].Nx = make_float2(Nn[k0][0], Nn[k1][0]); px[p].Ny = make_float2(Nn[k0][1], Nn[k1][1]); px[p].Nz = make_float2(Nn[k0][2], Nn[k1][2]);
        px[p].ar = px[p].ag = px[p].ab = make_float2(0.0f, 0.0f);
        px[p].valid0 = valid[k0]; px[p].valid1 = valid[k1]; px[p].slice0 = slice[k0]; px[p].slice1 = slice[k1];
    }

    uint32_t slices = 1u;
    if (MODE == MLGI_CULL_CLUSTERED) {
        my_slices = __reduce_or_sync(FULL, my_slices);
        if (lane == 0 && my_slices) atomicOr(&s_slices, my_slices);
        __syncthreads();
        slices = s_slices;
    }

    unsigned long long listed = 0, lit = 0;
    const float kz = -50.0f / a.zFar;

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
        Normal2 N[NP];   // a zero normal makes b == 0, i.e. "not lit", for pixels outside the viewport / other slices
        int nact = 0;
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const bool act0 = px[p].valid0 && (MODE != MLGI_CULL_CLUSTERED || px[p].slice0 == s);
            const bool act1 = px[p].valid1 && (MODE != MLGI_CULL_CLUSTERED || px[p].slice1 == s);
            N[p].x = make_float2(act0 ? px[p].Nx.x : 0.0f, act1 ? px[p].Nx.y : 0.0f);
            N[p].y = make_float2(act0 ? px[p].Ny.x : 0.0f, act1 ? px[p].Ny.y : 0.0f);
            N[p].z = make_float2(act0 ? px[p].Nz.x : 0.0f, act1 ? px[p].Nz.y : 0.0f);
            nact += (act0 ? 1 : 0) + (act1 ? 1 : 0);
        }

        if (BULK) {
            const int nchunks = (count + BCH - 1) / BCH;
            if (MODE == MLGI_CULL_CLUSTERED) {                  // (the tiled / unculled list was started before the prologue)
                __syncthreads();                                // the previous slice's list is consumed by every warp
#pragma unroll
                for (int c = 0; c < NBUF; c++) if (nchunks > c) issue_chunk(sub, count, c);
            }
            for (int c = 0; c < nchunks; c++) {
                const int b = c % NBUF, n = min(BCH, count - c * BCH);
                tma::mbar_wait(&s_full[b], (full_phase >> b) & 1u);
                full_phase ^= 1u << b;
                const float4* recs = srec + b * BCH * FG_REC_Q;
FG_UNROLL_PRAGMA
                for (int e = 0; e < n; e++)                     // :97
                    shade_vpl<SHADOW, STATS, TEX, NP>(px, recs + e * FG_REC_Q, a.ism, a.ism_tex, A, kz, a.clampv, N, nact, listed, lit);
                if (c + NBUF < nchunks) {                       // lists longer than the ring (N > 64 * NBUF * 16): refill buffer b
                    __syncthreads();                            // once every warp is done with it
                    issue_chunk(sub, count, c + NBUF);
                }
            }
            continue;
        }
        for (int base = 0; base < count; base += FG_CHUNK) {
            const int n = min(FG_CHUNK, count - base);
            __syncthreads();                                    // previous chunk fully consumed
            // stage n records (5 float4 each) cooperatively
            for (int q = threadIdx.x; q < n * 5; q += THREADS) {
                const int e = q / 5, part = q - e * 5;
                const uint32_t id = sub ? (uint32_t)sub[1 + base + e] : (uint32_t)(cls * L + base + e);   // :98
                srec[q] = __ldg(&a.records[id].r[part]);
            }
            __syncthreads();
FG_UNROLL_PRAGMA
            for (int e = 0; e < n; e++)                         // :97
                shade_vpl<SHADOW, STATS, TEX, NP>(px, srec + e * FG_REC_Q, a.ism, a.ism_tex, A, kz, a.clampv, N, nact, listed, lit);
        }
    }

    // ---- epilogue: result = acc * giIntensityFactor / vplCount * 16  (:132) ----
    const float giScale = SHADOW ? 2.0f * a.giScale : a.giScale;     // the shadowed loop accumulates -1/2 of the sums (neg_half_smoothstep)
    const int PW = (a.W + 3) >> 2, PH = (a.H + 3) >> 2;
    const size_t pboff = a.frame_ctr ? (size_t)((*a.frame_ctr + 1u) & 1u) * a.peer_parity_stride : 0;
    // Multi-GPU exchange, packed planes, CTA fully inside the image: the CTA's 16x16 plane entries are staged in shared memory
    // (the record buffer is free by now) and written as 16 rows of 64 contiguous bytes with 16-byte stores -- to this rank's
    // exchange region (pull mode) or to every rank's region over NVLink (push mode: the transfer overlaps the gathering of
    // the remaining tiles). 4 store instructions per thread and peer instead of 2 scattered 4-byte stores per pixel and peer.
    const bool rows_path = a.num_peers && a.peer_packed && !a.out_raster && !a.out_packed && (PW & 3) == 0 &&
                           tile_x * kTile + kTile <= a.W && tile_y * kTile + kTile <= a.H;
    if (rows_path) {
        uint32_t* s_out = reinterpret_cast<uint32_t*>(srec);
        __syncthreads();                                        // every warp is done with the staged records
#pragma unroll
        for (int k = 0; k < NPX; k++) {
            const PixelPair& pp = px[k >> 1];
            float r = (k & 1) ? pp.ar.y : pp.ar.x, g = (k & 1) ? pp.ag.y : pp.ag.x, b = (k & 1) ? pp.ab.y : pp.ab.x;
            if (SHADOW) { r = 0.0f - r; g = 0.0f - g; b = 0.0f - b; }      // (0 - (+0) = +0: an unlit pixel stays +0)
            s_out[(pj0 + 4 * k) * 16 + pi] = pack_r11g11b10(r * giScale, g * giScale, b * giScale);
        }
        __syncthreads();
        const size_t row0 = ((size_t)cls * PH + (size_t)tile_y * 16) * PW + (size_t)tile_x * 16;     // plane entry of the CTA's corner
        for (int item = threadIdx.x; item < a.num_peers * 64; item += THREADS) {
            const int g = item >> 6, row = (item >> 2) & 15, quad = item & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(s_out + row * 16 + quad * 4);
            uint32_t* dst = reinterpret_cast<uint32_t*>(static_cast<char*>(a.peer_planar[g]) + pboff) + row0 + (size_t)row * PW + quad * 4;
            *reinterpret_cast<uint4*>(dst) = v;
        }
    } else {
#pragma unroll
    for (int k = 0; k < NPX; k++) {
        if (!valid[k]) continue;
        const PixelPair& pp = px[k >> 1];
        float r = (k & 1) ? pp.ar.y : pp.ar.x, g = (k & 1) ? pp.ag.y : pp.ag.x, b = (k & 1) ? pp.ab.y : pp.ab.x;
        if (SHADOW) { r = 0.0f - r; g = 0.0f - g; b = 0.0f - b; }
        const float4 o = make_float4(r * giScale, g * giScale, b * giScale, 0.0f);
        if (a.out_raster) a.out_raster[(size_t)sy[k] * a.W + sx[k]] = o;
        if (a.out_planar) a.out_planar[((size_t)(cls - a.planar_class0) * PH + (sy[k] >> 2)) * PW + (sx[k] >> 2)] = o;
        if (a.out_packed) a.out_packed[(size_t)sy[k] * a.W + sx[k]] = pack_r11g11b10(o.x, o.y, o.z);
        if (a.out_planar_packed) a.out_planar_packed[((size_t)cls * PH + (sy[k] >> 2)) * PW + (sx[k] >> 2)] = pack_r11g11b10(o.x, o.y, o.z);
        if (a.num_peers) {
            // general shapes: this pixel's plane entry goes to the exchange region(s) directly
            const size_t off = ((size_t)cls * PH + (sy[k] >> 2)) * PW + (sx[k] >> 2);
            if (a.peer_packed) {
                const uint32_t v = pack_r11g11b10(o.x, o.y, o.z);
#pragma unroll 1
                for (int g = 0; g < a.num_peers; g++) reinterpret_cast<uint32_t*>(static_cast<char*>(a.peer_planar[g]) + pboff)[off] = v;
            } else {
#pragma unroll 1
                for (int g = 0; g < a.num_peers; g++) reinterpret_cast<float4*>(static_cast<char*>(a.peer_planar[g]) + pboff)[off] = o;
            }
        }
    }
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
    // occupancy cap (host-input pipeline): unused dynamic shared memory sized so that at most max_ctas_per_sm CTAs fit an SM, which
    // leaves registers and thread slots for the culling kernels of the next slab that run beside the gathering
    const size_t dyn = (a.max_ctas_per_sm > 0 && a.max_ctas_per_sm < FG_MINB) ? (size_t)(232448 / (a.max_ctas_per_sm + 1) + 64 - (FG_CHUNK * FG_REC_Q * 16 + 64) - 1024) : 0;
    const bool sh = a.enable_shadowing != 0, stt = a.stats != nullptr;
    const bool tex = sh && a.ism_tex != 0;      // the ISM is fetched through the texture unit whenever a texture object exists
    if (a.gb_world) {                            // host-input sharded frame: G-buffer rows live on the rank that uploaded them
        if (!sh) final_gather_kernel<MODE, false, false, 0, 1, false, true><<<grid, 128, dyn, st>>>(a);
        else if (!tex) final_gather_kernel<MODE, true, false, 0, 1, false, true><<<grid, 128, dyn, st>>>(a);
        else final_gather_kernel<MODE, true, false, 1, 1, true, true><<<grid, 128, dyn, st>>>(a);
        return;
    }
    if (!stt && tex && a.tex_float_coord == 1) { final_gather_kernel<MODE, true, false, 2, 1, false><<<grid, 128, dyn, st>>>(a); return; }
    if (!stt && tex && a.tex_float_coord == 3) { final_gather_kernel<MODE, true, false, 4, 1, true><<<grid, 128, dyn, st>>>(a); return; }
    if (!stt && tex && a.tex_float_coord == 2) { final_gather_kernel<MODE, true, false, 3, 1, true><<<grid, 128, dyn, st>>>(a); return; }
    if (stt) {                                   // counting variant (work counters for the roofline accounting)
        if (sh) final_gather_kernel<MODE, true, true, 0, 1, false><<<grid, 128, dyn, st>>>(a);
        else final_gather_kernel<MODE, false, true, 0, 1, false><<<grid, 128, dyn, st>>>(a);
    } else if (!sh) final_gather_kernel<MODE, false, false, 0, 1, false><<<grid, 128, dyn, st>>>(a);
    else if (!tex) final_gather_kernel<MODE, true, false, 0, 1, false><<<grid, 128, dyn, st>>>(a);
    else if (a.bulk_staging) final_gather_kernel<MODE, true, false, 1, 1, true><<<grid, 128, dyn, st>>>(a);
    else final_gather_kernel<MODE, true, false, 1, 1, false><<<grid, 128, dyn, st>>>(a);
}

int launch_final_gather(const GatherArgs& a, cudaStream_t st) {
    const int ncls = a.class_hi - a.class_lo;
    dim3 grid((unsigned)(a.tile_count * ncls));
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
// ---- multi-GPU exchange: ONE kernel = signal + wait + pull + de-interleave ----------------------------------------------
// Runs after the gathering kernel on the same stream, so this rank's class planes are complete (in its own exchange buffer
// in pull mode; in every rank's buffer in push mode). CTA 0 publishes the frame id into slot [rank] of every rank's flag
// array (system-scope fence first); every CTA then waits until all `world` slots of the local array show the frame id
// (bounded spin: a dead peer raises an error instead of hanging the GPU) and de-interleaves its share of the image,
// reading the planes of class c from the rank that owns c -- over NVLink with 16-byte volatile loads (never cached
// locally: the parity buffer is rewritten two frames later) in pull mode. Each thread transposes 4 planes x 4 entries
// into 16 consecutive raster pixels, so both the (remote) reads and the writes are full 512 B / 2 KB warp transactions.
// The last CTA to finish advances the device-resident frame counter, which makes the kernel graph-replayable.
__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 unpack_gi(uint32_t v) {
    return make_float4(__uint_as_float(unpack_ufloat_bits<6>(v & 0x7FFu)), __uint_as_float(unpack_ufloat_bits<6>((v >> 11) & 0x7FFu)),
                       __uint_as_float(unpack_ufloat_bits<5>(v >> 22)), 0.0f);
}
__device__ __forceinline__ bool wait_flags(const uint32_t* flags, int world, uint32_t frame, uint32_t* error) {
    // the first `world` threads of the CTA poll one slot each
    if (threadIdx.x < (unsigned)world) {
        const volatile uint32_t* f = flags + threadIdx.x;
        long long spins = 0;
        while ((int)(*f - frame) < 0) {                     // wrap-safe "not yet"
            __nanosleep(100);
            if (++spins > 20000000LL) { atomicExch(error, 1u + threadIdx.x); break; }   // seconds: a peer died; do not hang the GPU
        }
        __threadfence_system();
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(256) exchange_kernel(const ExchangeArgs x) {
    const uint32_t frame = *reinterpret_cast<volatile uint32_t*>(x.frame_ctr) + 1u;
    if (x.signal_wait) {
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)x.world) {
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t*>(x.peer_flags[threadIdx.x] + x.rank) = frame;
            __threadfence_system();
        }
        wait_flags(x.flags, x.world, frame, x.error);
    }
    const int PW = (x.W + 3) >> 2, PH = (x.H + 3) >> 2;
    const int nrows = x.row_blocks * x.row_block;               // rows = row_blocks blocks of row_block rows, row_stride apart (clipped to H)
    const size_t par = (size_t)(frame & 1u) * x.parity_stride;
    const int per = 16 / x.world;                               // classes per rank
    if (x.packed && (PW & 3) == 0 && (x.W & 15) == 0) {
        // unit = (raster row y, 16-pixel group i4): planes 4*cy .. 4*cy+3, entries [py][4*i4 .. 4*i4+3]
        const int groups = PW >> 2;
        const long long units = (long long)nrows * groups;
        for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
            const int ri = (int)(u / groups), i4 = (int)(u - (long long)ri * groups);
            const int y = x.row_begin + (ri / x.row_block) * x.row_stride + ri % x.row_block;
            if (y >= x.H) continue;
            const int cy = y & 3, py = y >> 2;
            uint4 pl[4];
#pragma unroll
            for (int cx = 0; cx < 4; cx++) {
                const int cls = 4 * cy + cx;
                const char* base = x.src[cls / per] + par;
                pl[cx] = ld_volatile_u4(reinterpret_cast<const uint32_t*>(base) + ((size_t)cls * PH + py) * PW + 4 * i4);
            }
            const uint4 o0 = make_uint4(pl[0].x, pl[1].x, pl[2].x, pl[3].x), o1 = make_uint4(pl[0].y, pl[1].y, pl[2].y, pl[3].y);
            const uint4 o2 = make_uint4(pl[0].z, pl[1].z, pl[2].z, pl[3].z), o3 = make_uint4(pl[0].w, pl[1].w, pl[2].w, pl[3].w);
            const size_t r = (size_t)y * x.W + 16 * (size_t)i4;
            if (x.raster_packed) {
                uint4* dst = reinterpret_cast<uint4*>(x.raster_packed + r);
                dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3;
            }
            if (x.raster) {
                const uint32_t v[16] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y, o2.z, o2.w, o3.x, o3.y, o3.z, o3.w};
#pragma unroll
                for (int i = 0; i < 16; i++) x.raster[r + i] = unpack_gi(v[i]);
            }
        }
    } else {
        // general shapes / fp32 planes: one pixel per thread
        const long long n = (long long)x.W * nrows;
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
            const int ri = (int)(j / x.W), xx = (int)(j - (long long)ri * x.W);
            const int y = x.row_begin + (ri / x.row_block) * x.row_stride + ri % x.row_block;
            if (y >= x.H) continue;
            const long long i = (long long)y * x.W + xx;
            const int cls = (xx & 3) + 4 * (y & 3);
            const char* base = x.src[cls / per] + par;
            const size_t e = ((size_t)cls * PH + (y >> 2)) * PW + (xx >> 2);
            if (x.packed) {
                const uint32_t v = *(reinterpret_cast<const volatile uint32_t*>(base) + e);
                if (x.raster_packed) x.raster_packed[i] = v;
                if (x.raster) x.raster[i] = unpack_gi(v);
            } else {
                const uint4 q = ld_volatile_u4(reinterpret_cast<const float4*>(base) + e);
                if (x.raster) x.raster[i] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
            }
        }
    }
    // last CTA out: advance the frame counter, re-arm the ticket (a phased frame advances in its last exchange kernel only)
    if (!x.advance) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(x.ticket, 1u) == gridDim.x - 1) {
            *x.ticket = 0u;
            *reinterpret_cast<volatile uint32_t*>(x.frame_ctr) = frame;
            __threadfence();
        }
    }
}
int launch_exchange(const ExchangeArgs& x, cudaStream_t st) {
    // every CTA must be resident (all of them spin on the flags before any of them can finish): at most one wave
    const long long work = ((long long)x.W * x.row_blocks * x.row_block) / 16 + 1;
    const int grid = (int)std::max(1LL, std::min<long long>(148 * 4, (work + 255) / 256));
    exchange_kernel<<<grid, 256, 0, st>>>(x);
    return 1;
}

// one-warp signal / bounded-wait kernels that order the cross-rank exchanges outside the exchange kernel: the tile bounds of the
// sharded frames (slots 16 + rank), and the per-phase bounds and class planes of the host-input sharded frame (32 + rank, 48 + rank)
// flag value = (frames completed + 1) * mul + add: mul = 1, add = 0 for one signal per frame; a phased frame (host-input
// sharded frame) uses mul = 16, add = phase, which stays monotonic over phases and frames
__global__ void flag_signal_kernel(uint32_t* const* peer_flags, int world, int slot, const uint32_t* frame_ctr, uint32_t mul, uint32_t add) {
    const uint32_t frame = (*reinterpret_cast<const volatile uint32_t*>(frame_ctr) + 1u) * mul + add;
    if (threadIdx.x < (unsigned)world) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(peer_flags[threadIdx.x] + slot) = frame;
        __threadfence_system();
    }
}
__global__ void flag_wait_kernel(const uint32_t* flags, int world, const uint32_t* frame_ctr, uint32_t* error, uint32_t mul, uint32_t add) {
    const uint32_t frame = (*reinterpret_cast<const volatile uint32_t*>(frame_ctr) + 1u) * mul + add;
    wait_flags(flags, world, frame, error);
}
int launch_flag_signal(uint32_t* const* peer_flags, int world, int slot, const uint32_t* frame_ctr, uint32_t mul, uint32_t add, cudaStream_t st) {
    flag_signal_kernel<<<1, 32, 0, st>>>(peer_flags, world, slot, frame_ctr, mul, add);
    return 1;
}
int launch_flag_wait(const uint32_t* flags, int world, const uint32_t* frame_ctr, uint32_t* error, uint32_t mul, uint32_t add, cudaStream_t st) {
    flag_wait_kernel<<<1, 32, 0, st>>>(flags, world, frame_ctr, error, mul, add);
    return 1;
}

__global__ void planar_packed_to_raster_kernel(const uint32_t* __restrict__ planar, int W, int H, uint32_t* __restrict__ raster_packed,
                                               float4* __restrict__ raster) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const int cls = (x & 3) + 4 * (y & 3);
    const int PW = (W + 3) >> 2, PH = (H + 3) >> 2;
    const uint32_t v = planar[((size_t)cls * PH + (y >> 2)) * PW + (x >> 2)];
    if (raster_packed) raster_packed[(size_t)y * W + x] = v;
    if (raster) raster[(size_t)y * W + x] = make_float4(__uint_as_float(unpack_ufloat_bits<6>(v & 0x7FFu)), __uint_as_float(unpack_ufloat_bits<6>((v >> 11) & 0x7FFu)),
                                                         __uint_as_float(unpack_ufloat_bits<5>(v >> 22)), 0.0f);
}
int launch_planar_packed_to_raster(const uint32_t* planar, int W, int H, uint32_t* raster_packed, float4* raster, cudaStream_t st) {
    planar_packed_to_raster_kernel<<<dim3((W + 255) / 256, H), 256, 0, st>>>(planar, W, H, raster_packed, raster);
    return 1;
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
float measure_fp32_peak_tflops(int device, cudaStream_t st) {
    float* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0f;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
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
// n1: gi_blur.frag:24-72, the two separable 7-tap passes of GIStage::blur (GIStage.cpp:308-344) as ONE kernel. A CTA owns a
// 32x26-pixel tile and stages the tile plus a 3-pixel halo (38x32 texels: GI colour, linearised depth, decoded normal -- each
// texel is fetched and decoded once instead of once per tap and pass) in shared memory; the intermediate image of the
// two-pass form never leaves the SM. HBM traffic per pixel: 4 B (GL_R11F_G11F_B10F GI, the reference's giBuffer format) or
// 16 B (fp32 GI) + 4 B depth + 4 B normal in, 16 B out.
// The passes are bound by shared-memory traffic (7 values per tap), so each thread computes a RUN of 4 consecutive outputs
// along the pass direction from a 10-texel window held in registers (2.5 instead of 7 texel reads per output), and the
// weight of a texel pair -- symmetric in its two texels (:24-47) -- is evaluated once and used for both outputs (18 instead
// of 24 weights per run). Lanes run ACROSS the pass direction (horizontal pass: lane = row; vertical pass: lane = column) and
// the planes have odd pitches, so every shared-memory access of a warp is conflict-free.
// Out-of-image taps are skipped by the reference (texelFetch outside the image is undefined in GL; same choice as the
// oracle): such texels are staged with depth 1e18, which drives the weight to max(0, -huge) = 0 exactly.
// =====================================================================================
namespace blur {
constexpr int TW = 32, TH = 26, R = 3, SW = TW + 2 * R, SH = TH + 2 * R, PITCH = SW + 1, TPITCH = TW + 1, RUN = 4, WIN = RUN + 2 * R;
static_assert(SH == 32 && TW == 32, "lanes map to staged rows (horizontal pass) and tile columns (vertical pass)");
struct Texel { float r, g, b, d, nx, ny, nz; };
// weight between two texels i apart (gi_blur.frag:24-47); symmetric in (c, t)
__device__ __forceinline__ float weight(const Texel& c, const Texel& t, int i) {
    const float normalFactor = 1.0f - fmaxf(0.0f, c.nx * t.nx + c.ny * t.ny + c.nz * t.nz);   // :44
    const float depthDiff = t.d - c.d;
    const float nd = normalFactor * 5.0f;
    return fmaxf(0.0f, (3.0f - float(i) + 1.0f) / 4.0f - depthDiff * depthDiff - nd * nd);
}
// one run: outputs are window positions R .. R + RUN - 1; acc starts at the centre texel with weight 1 (:60)
__device__ __forceinline__ void run_pass(const Texel (&w)[WIN], float (&ar)[RUN], float (&ag)[RUN], float (&ab)[RUN]) {
    float fa[RUN];
#pragma unroll
    for (int j = 0; j < RUN; j++) { ar[j] = w[R + j].r; ag[j] = w[R + j].g; ab[j] = w[R + j].b; fa[j] = 1.0f; }
#pragma unroll
    for (int p = 0; p < WIN; p++)
#pragma unroll
        for (int i = 1; i <= R; i++) {                                                         // :63-69, both directions at once
            const int q = p + i;
            if (q >= WIN) continue;
            const bool op = p >= R && p < R + RUN, oq = q >= R && q < R + RUN;
            if (!op && !oq) continue;
            const float wt = weight(w[p], w[q], i);
            if (op) { fa[p - R] += wt; ar[p - R] += w[q].r * wt; ag[p - R] += w[q].g * wt; ab[p - R] += w[q].b * wt; }
            if (oq) { fa[q - R] += wt; ar[q - R] += w[p].r * wt; ag[q - R] += w[p].g * wt; ab[q - R] += w[p].b * wt; }
        }
#pragma unroll
    for (int j = 0; j < RUN; j++) {                                                            // :71
        const float inv = rcp_approx(fa[j]);                                                   // fa >= 1 (1 ulp; the blur's contract is 1e-4)
        ar[j] *= inv; ag[j] *= inv; ab[j] *= inv;
    }
}
}  // namespace blur
template <bool PACKED>
__global__ void __launch_bounds__(256, 4) gi_blur_fused_kernel(const float4* __restrict__ gi, const uint32_t* __restrict__ gi_packed,
                                                               const float* __restrict__ depth, const uint32_t* __restrict__ face_normal,
                                                               int W, int H, float projA, float projB, float4* __restrict__ out) {
    using namespace blur;
    __shared__ float s_c[3][SH][PITCH], s_d[SH][PITCH], s_n[3][SH][PITCH];
    __shared__ float s_t[3][SH][TPITCH];
    const int x0 = blockIdx.x * TW - R, y0 = blockIdx.y * TH - R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // staging in two steps, so that all of a thread's global loads (3 per texel, 5 texels) are in flight before the first use
    constexpr int NQ = (SW * SH + 255) / 256;
    uint32_t gv[NQ], nv[NQ];
    float4 gc[NQ];
    float dv[NQ];
    bool in[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        const int q = threadIdx.x + k * 256, lx = q % SW, ly = q / SW, x = x0 + lx, y = y0 + ly;
        in[k] = q < SW * SH && x >= 0 && y >= 0 && x < W && y < H;
        gv[k] = 0u; nv[k] = 0u; dv[k] = 0.0f; gc[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (in[k]) {
            const size_t t = (size_t)y * W + x;
            if (PACKED) gv[k] = __ldg(gi_packed + t); else gc[k] = __ldg(gi + t);
            dv[k] = __ldg(depth + t);
            nv[k] = __ldg(face_normal + t);
        }
    }
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        const int q = threadIdx.x + k * 256, lx = q % SW, ly = q / SW;
        if (q >= SW * SH) continue;
        float r = 0.0f, g = 0.0f, b = 0.0f, d = 1e18f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (in[k]) {
            if (PACKED) {
                const uint32_t v = gv[k];
                r = unpack_ufloat_fast<6>(v & 0x7FFu); g = unpack_ufloat_fast<6>((v >> 11) & 0x7FFu); b = unpack_ufloat_fast<5>(v >> 22);
            } else { r = gc[k].x; g = gc[k].y; b = gc[k].z; }
            d = __fdividef(-projB, projA + (2.0f * dv[k] - 1.0f));                          // reprojection.glsl:6-14 (2 ulp; contract 1e-4)
            const uint32_t n10 = nv[k];
            nx = float(n10 & 0x3FFu) * (2.0f / 1023.0f) - 1.0f; ny = float((n10 >> 10) & 0x3FFu) * (2.0f / 1023.0f) - 1.0f;
            nz = float((n10 >> 20) & 0x3FFu) * (2.0f / 1023.0f) - 1.0f;
        }
        s_c[0][ly][lx] = r; s_c[1][ly][lx] = g; s_c[2][ly][lx] = b; s_d[ly][lx] = d;
        s_n[0][ly][lx] = nx; s_n[1][ly][lx] = ny; s_n[2][ly][lx] = nz;
    }
    __syncthreads();
    float ar[RUN], ag[RUN], ab[RUN];
    {   // horizontal pass (DIRECTION (1,0), GIStage.cpp:427): lane = staged row, warp w -> tile columns 4w .. 4w+3
        const int row = lane, c0 = RUN * warp;                                             // window = staged columns c0 .. c0+9
        Texel w[WIN];
#pragma unroll
        for (int k = 0; k < WIN; k++)
            w[k] = Texel{s_c[0][row][c0 + k], s_c[1][row][c0 + k], s_c[2][row][c0 + k], s_d[row][c0 + k],
                         s_n[0][row][c0 + k], s_n[1][row][c0 + k], s_n[2][row][c0 + k]};
        run_pass(w, ar, ag, ab);
#pragma unroll
        for (int j = 0; j < RUN; j++) { s_t[0][row][c0 + j] = ar[j]; s_t[1][row][c0 + j] = ag[j]; s_t[2][row][c0 + j] = ab[j]; }
    }
    __syncthreads();
    {   // vertical pass (DIRECTION (0,1), GIStage.cpp:431): lane = tile column, warp w -> tile rows 4w .. 4w+3 (w = 6: two rows)
        const int col = lane, r0 = RUN * warp;                                             // window = staged rows r0 .. r0+9
        if (r0 >= TH) return;
        Texel w[WIN];
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            const int rr = r0 + k;
            if (rr < SH) w[k] = Texel{s_t[0][rr][col], s_t[1][rr][col], s_t[2][rr][col], s_d[rr][col + R],
                                      s_n[0][rr][col + R], s_n[1][rr][col + R], s_n[2][rr][col + R]};
            else w[k] = Texel{0.0f, 0.0f, 0.0f, 1e18f, 0.0f, 0.0f, 0.0f};                  // below the staged rows: only reached by rows that are not stored
        }
        run_pass(w, ar, ag, ab);
        const int x = x0 + R + col;
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const int y = y0 + R + r0 + j;
            if (r0 + j < TH && x < W && y < H) out[(size_t)y * W + x] = make_float4(ar[j], ag[j], ab[j], 0.0f);
        }
    }
}
int launch_gi_blur(const float4* gi, const uint32_t* gi_packed, const float* depth, const uint32_t* face_normal, int W, int H, float projA,
                   float projB, float4* out, cudaStream_t st) {
    dim3 grid((W + blur::TW - 1) / blur::TW, (H + blur::TH - 1) / blur::TH);
    if (gi_packed) gi_blur_fused_kernel<true><<<grid, 256, 0, st>>>(nullptr, gi_packed, depth, face_normal, W, H, projA, projB, out);
    else gi_blur_fused_kernel<false><<<grid, 256, 0, st>>>(gi, nullptr, depth, face_normal, W, H, projA, projB, out);
    return 1;
}

// =====================================================================================
// n4 (reduced to the term that depends on this path): deferredshading.frag:69-80 ambientTerm = gi * toLinear(diffuse) *
// occlusion, then tonemap (:38-43, Reinhard with exposure) and toSRGB (srgb_utils.glsl, gamma 2.2) -> RGBA8. The direct
// diffuse / specular terms need the VSM shadow map and are outside SURVEY section 8. HBM-bound: 16 + 4 (+4) B in, 4 B out.
// =====================================================================================
__global__ void composite_ambient_kernel(const float4* __restrict__ gi, const uint32_t* __restrict__ diffuse_rgba8, const float* __restrict__ occlusion,
                                         int n, float exposure, uint32_t* __restrict__ out_srgb8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 g = gi[i];
    const uint32_t d = diffuse_rgba8[i];
    const float occ = occlusion ? occlusion[i] : 1.0f;
    const float gc[3] = {g.x, g.y, g.z};
    uint32_t o = 0xFF000000u;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float dl = powf(float((d >> (8 * c)) & 0xFFu) / 255.0f, 2.2f);      // toLinear
        float v = gc[c] * dl * occ;                                               // ambientTerm
        v *= exposure; v = v / (1.0f + v);                                        // tonemap
        v = powf(v, 1.0f / 2.2f);                                                 // toSRGB
        o |= (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f) << (8 * c);
    }
    out_srgb8[i] = o;
}
int launch_composite_ambient(const float4* gi, const uint32_t* diffuse_rgba8, const float* occlusion, int n, float exposure, uint32_t* out_srgb8,
                             cudaStream_t st) {
    composite_ambient_kernel<<<(n + 255) / 256, 256, 0, st>>>(gi, diffuse_rgba8, occlusion, n, exposure, out_srgb8);
    return 1;
}

}  // namespace mlgi

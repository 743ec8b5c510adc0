// exact.cu -- the stages whose outputs are under a bit-exact contract with the CPU oracle.
// Compiled with -fmad=false (no FMA contraction); see common.cuh for the numeric rules.
//
//   a1  vpl_process_kernel      (vpl_processor.comp:41-72)
//   a2  point_count / point_generate / point_scatter   (ism.tesc/tese/geom, deterministic tessellator)
//   a3  splat_kernel            (ism.comp:54-139)  one warp per bucket, RED.MIN into the R32UI atlas
//   a4+a5 pullpush_kernel       (pull.comp x6 + push.comp x6) one 64x64 ISM per CTA, pyramid in smem
//   a6  cluster_mask / cluster_compact                 (clustering.comp:30-88)
//   a7  light_lists_kernel      (light_lists.comp:50-126) ballot + popc compaction, ascending ids
//   a8  cull_tiled_kernel       (tiled variant)
#include "common.cuh"
#include "pullpush_core.cuh"
#include "bulk_async.cuh"

#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace mlgi {

static constexpr unsigned FULL = 0xFFFFFFFFu;

// =====================================================================================
// a1: one thread per VPL. Also emits the hoisted record the gathering kernel consumes
// (lookAtRH basis + atlas tile origin are pure functions of the VPL, ism_utils.glsl:5-13,35-38).
// =====================================================================================
__global__ void vpl_process_kernel(const uint32_t* __restrict__ diffuse, const uint32_t* __restrict__ normal10,
                                   const float* __restrict__ depth, int rsm_w, int rsm_h, int N, Mat4 inv,
                                   float lightIntensity, int shuffle, const int32_t* __restrict__ shuffled, IsmMode mode, int compact,
                                   float* __restrict__ vpl, float* __restrict__ packed, VplRecord* __restrict__ records) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint32_t)N) return;
    const uint32_t ax = (uint32_t)ceilf(sqrtf(float(N)));            // vpl_processor.comp:38
    const uint32_t ay = (uint32_t)N / ax;                             // :46
    uint32_t tcx = (idx % ax) * (uint32_t)rsm_w / ax;                 // :50
    uint32_t tcy = (idx / ax) * (uint32_t)rsm_h / ay;
    float tfx = float(tcx) / float(rsm_w), tfy = float(tcy) / float(rsm_h);  // :51
    // compact: the three planes hold only the N sampled texels, in VPL order (host-input frames upload 12 B per VPL instead of the RSM)
    size_t t = compact ? (size_t)idx : (size_t)tcy * rsm_w + tcx;
    uint32_t d8 = diffuse[t];
    float3 dif = f3(float(d8 & 0xFFu) / 255.0f, float((d8 >> 8) & 0xFFu) / 255.0f, float((d8 >> 16) & 0xFFu) / 255.0f);
    float3 nrm = decode_rgb10_sn(normal10[t]);                        // :54
    float dep = depth[t];
    float4 wc = mat4_mul_vec4(inv, tfx, tfy, dep, 1.0f);              // :57
    float3 w = f3(wc.x / wc.w, wc.y / wc.w, wc.z / wc.w);             // :58
    float3 col = mul3s(dif, lightIntensity);                          // :60
    uint32_t slot = shuffle ? (uint32_t)shuffled[idx] : idx;          // :62-63
    float pn = pack3SN(nrm);                                          // :65
    float3 un = unpack3SN(pn);                                        // :68
    float* o = vpl + (size_t)slot * 12;
    o[0] = w.x; o[1] = w.y; o[2] = w.z; o[3] = 0.0f;
    o[4] = un.x; o[5] = un.y; o[6] = un.z; o[7] = 0.0f;
    o[8] = col.x; o[9] = col.y; o[10] = col.z; o[11] = 0.0f;
    float* p = packed + (size_t)slot * 4;
    p[0] = w.x; p[1] = w.y; p[2] = w.z; p[3] = pn;
    // hoisted record
    float3 up = f3(0.0f, 1.0f, 0.01f);
    float3 s = normalize3(cross3(un, up));
    float3 u = cross3(s, un);
    // ISM tile of this VPL: ism_utils.glsl:35-36 with ismIndex = slot - (scaleISMs ? vplStartIndex : 0) (ism.comp:124,
    // final_gathering.comp:44,120) in a layout of S2 x S2 ISMs; the int() truncations are the shader's (a slot outside the
    // scaled range lands on a negative or out-of-atlas tile, as it does in the reference)
    const float ismIndex = mode.scale ? float(slot) - float(mode.vpl_start) : float(slot);
    const float yy = float(f2i(ismIndex / float(mode.S2)));
    const float tx = float(f2i(ismIndex - yy * float(mode.S2))), ty = float(f2i(yy));
    const float h = mode.px * 0.5f;                                   // half the ISM edge: 32 by default (power of two: exact scaling)
    const float ox = tx * mode.px + h, oy = ty * mode.px + h;
    VplRecord r;
    r.r[0] = make_float4(-w.x, -w.y, -w.z, un.x);
    r.r[1] = make_float4(un.y, un.z, h * s.x, h * u.x);             // rows pre-scaled by half the ISM size (exact), paired by component
    r.r[2] = make_float4(h * s.y, h * u.y, h * s.z, h * u.z);
    r.r[3] = make_float4(col.x, col.y, col.z, 0.0f);
    r.r[4] = make_float4(12582912.0f + ox, 12582912.0f + oy, ox, oy);
    records[slot] = r;
}

int launch_vpl_process(const mlgi_rsm& rsm, int rsm_w, int rsm_h, int N, const Mat4& inv, float lightIntensity,
                       int shuffle, const int32_t* shuffled, const IsmMode& mode, int compact, float* vpl, float* packed, VplRecord* records,
                       cudaStream_t st) {
    vpl_process_kernel<<<(N + 63) / 64, 64, 0, st>>>(rsm.diffuse, rsm.face_normal, rsm.depth, rsm_w, rsm_h, N, inv,
                                                    lightIntensity, shuffle, shuffled, mode, compact, vpl, packed, records);
    return 1;
}

// =====================================================================================
// device-wide exclusive scan of uint32 (3 small kernels; n up to a few million)
// =====================================================================================
static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_ITEMS = 2;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0u;
        uint32_t y = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t z = __shfl_up_sync(FULL, y, o);
            if (lane >= o) y += z;
        }
        warp_sums[lane] = y - ws;           // exclusive
        if (lane == 31) *total = y;
    }
    __syncthreads();
    uint32_t r = warp_sums[warp] + x - v;
    __syncthreads();
    return r;
}

// =====================================================================================
// a2: points. Deterministic tessellator:
//   n = ceil(clamp(max edge * tessLevelFactor, 1, 64)), n^2 sub-triangles, sub id k enumerates
//   rows j = 0..n-1, inside a row i = 0..n-1-j: up(i,j) then down(i,j).
// =====================================================================================
__device__ __forceinline__ float3 load_vertex(const float* __restrict__ v, uint32_t i) { return f3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

__device__ __forceinline__ int tess_n(float3 v0, float3 v1, float3 v2, float tf) {
    float l0 = len3(sub3(v1, v2)), l1 = len3(sub3(v2, v0)), l2 = len3(sub3(v0, v1));  // ism.tesc:13-15
    float o0 = l0 * tf, o1 = l1 * tf, o2 = l2 * tf;                                    // :16-18
    float inner = gmax(gmax(o0, o1), o2);                                              // :20
    return f2i(ceilf(gclamp(inner, 1.0f, 64.0f)));
}

// random.glsl:34-46,50-59,65
__device__ __forceinline__ uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
__device__ __forceinline__ float random3(float3 v) {
    uint32_t h = hash1(__float_as_uint(v.x) ^ hash1(__float_as_uint(v.y)) ^ hash1(__float_as_uint(v.z)));
    return __uint_as_float((h & 0x007FFFFFu) | 0x3F800000u) - 1.0f;
}

// One warp takes 32 triangles at a time: per-lane sub-triangle counts, a warp scan, ONE atomicAdd per warp for the output
// range, then the lanes walk the group's points (owner lane found by a 5-step shuffle search over the scan). There is no
// global scan and no sort: the order of the point list is arbitrary (it depends on which warp allocates first), which is
// immaterial -- the splat is an order-independent min, and the tests compare point lists as sets. Each point carries its
// bucket id (ism.geom:57) in point_bucket; the bucket-grouped view of the reference's pointBuffer is only built on request
// (launch_group_points, used by the inspection API).
__global__ void __launch_bounds__(256) point_generate_kernel(const float* __restrict__ verts, const float* __restrict__ normals,
                                                             const uint32_t* __restrict__ indices, const uint32_t* __restrict__ prim_id,
                                                             int ntri, int N /* sampledVplCount: buckets */, int ismCount, float tf, int max_points,
                                                             float4* __restrict__ points, uint32_t* __restrict__ point_bucket,
                                                             uint32_t* __restrict__ generated) {
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    const float sqrtN = sqrtf(float(ismCount));                                           // ism.geom:63
    for (int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; group * 32 < ntri; group += warps_total) {
        const int t = group * 32 + lane;
        int n_mine = 0;
        if (t < ntri) {
            float3 v0 = load_vertex(verts, indices[3 * t]), v1 = load_vertex(verts, indices[3 * t + 1]), v2 = load_vertex(verts, indices[3 * t + 2]);
            n_mine = tess_n(v0, v1, v2, tf);
        }
        const uint32_t cnt = (uint32_t)(n_mine * n_mine);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        const uint32_t excl = incl - cnt;
        const uint32_t T = __shfl_sync(FULL, incl, 31);
        uint32_t base = 0;
        if (lane == 0 && T) base = atomicAdd(generated, T);
        base = __shfl_sync(FULL, base, 0);
        for (uint32_t k0 = 0; k0 < T; k0 += 32) {              // warp-uniform trip count
            const uint32_t kk = k0 + lane;
            int owner = 0;                                      // largest lane with excl <= kk (zero-count lanes can never win)
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t e = __shfl_sync(FULL, excl, (owner + step) & 31);
                if (e <= kk) owner += step;
            }
            const uint32_t start = __shfl_sync(FULL, excl, owner);
            const int n = __shfl_sync(FULL, n_mine, owner);
            const uint32_t out = base + kk;
            if (kk >= T || out >= (uint32_t)max_points) continue;
            const int tt = group * 32 + owner;
            const int k = (int)(kk - start);
            const int nn = n * n;
            // row j: largest j with 2nj - j^2 <= k
            int j = (int)(float(n) - sqrtf(float(nn - k)));
            if (j < 0) j = 0;
            if (j > n - 1) j = n - 1;
            while (j > 0 && 2 * n * j - j * j > k) j--;
            while (j < n - 1 && 2 * n * (j + 1) - (j + 1) * (j + 1) <= k) j++;
            const int r = k - (2 * n * j - j * j);
            const int i = r >> 1, down = r & 1;
            uint32_t i0 = indices[3 * tt], i1 = indices[3 * tt + 1], i2 = indices[3 * tt + 2];
            float3 v0 = load_vertex(verts, i0), v1 = load_vertex(verts, i1), v2 = load_vertex(verts, i2);
            float3 te_normal = load_vertex(normals, i0);                        // ism.tese:22
            int gi[3], gj[3];
            if (!down) { gi[0] = i; gj[0] = j; gi[1] = i + 1; gj[1] = j; gi[2] = i; gj[2] = j + 1; }
            else       { gi[0] = i + 1; gj[0] = j; gi[1] = i + 1; gj[1] = j + 1; gi[2] = i; gj[2] = j + 1; }
            const float fn = float(n);
            float3 tc[3], pp[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                tc[q] = f3(float(n - gi[q] - gj[q]) / fn, float(gi[q]) / fn, float(gj[q]) / fn);
                pp[q] = add3(add3(mul3s(v0, tc[q].x), mul3s(v1, tc[q].y)), mul3s(v2, tc[q].z));   // ism.tese:9-12
            }
            float3 position = div3s(add3(add3(pp[0], pp[1]), pp[2]), 3.0f);       // ism.geom:51
            float3 bary = div3s(add3(add3(tc[0], tc[1]), tc[2]), 3.0f);           // :52
            float maxdist = gmax(gmax(len3(sub3(position, pp[0])), len3(sub3(position, pp[1]))), len3(sub3(position, pp[2])));  // :54
            float sd = float(prim_id[tt] % 4096u) / 4096.0f;                      // :56
            float3 seed = f3(bary.x + sd, bary.y + sd, bary.z + sd);
            int bkt = f2i(random3(seed) * float(N));                              // :57
            if (bkt >= N) bkt = N - 1;
            float radius = maxdist * sqrtN;                                       // :63
            uint32_t nr = pack_unorm8(te_normal.x * 0.5f + 0.5f) | (pack_unorm8(te_normal.y * 0.5f + 0.5f) << 8) |
                          (pack_unorm8(te_normal.z * 0.5f + 0.5f) << 16) | (pack_unorm8(radius / 25.0f) << 24);  // :64
            points[out] = make_float4(position.x, position.y, position.z, __uint_as_float(nr));
            point_bucket[out] = (uint32_t)bkt;
        }
    }
}

int launch_point_generate(const mlgi_mesh& mesh, const IsmMode& mode, float tf, int max_points, float4* points, uint32_t* point_bucket,
                          uint32_t* generated, cudaStream_t st) {
    cudaMemsetAsync(generated, 0, sizeof(uint32_t), st);
    const int warps_needed = (mesh.num_triangles + 31) / 32;
    const int grid = std::max(1, std::min(148 * 8, (warps_needed + 7) / 8));
    point_generate_kernel<<<grid, 256, 0, st>>>(mesh.vertices, mesh.normals, mesh.indices, mesh.prim_id, mesh.num_triangles, mode.sampled,
                                                mode.ism_count, tf, max_points, points, point_bucket, generated);
    return 1;
}

// ---- bucket-grouped view of the point list (reference layout: pointBuffer[bucket][slot], here compacted with
// bucket_start[N+1]); built only when the inspection API asks for it ----
__global__ void __launch_bounds__(256) bucket_hist_kernel(const uint32_t* __restrict__ point_bucket, const uint32_t* __restrict__ generated,
                                                          int max_points, uint32_t* __restrict__ hist) {
    const uint32_t P = min(generated[0], (uint32_t)max_points);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) atomicAdd(&hist[point_bucket[p]], 1u);
}
__global__ void __launch_bounds__(256) point_scatter_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ point_bucket,
                                                            const uint32_t* __restrict__ generated, int max_points,
                                                            const uint32_t* __restrict__ bucket_start, uint32_t* __restrict__ bucket_cursor,
                                                            float4* __restrict__ points_grouped) {
    const uint32_t P = min(generated[0], (uint32_t)max_points);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        uint32_t b = point_bucket[p];
        uint32_t slot = bucket_start[b] + atomicAdd(&bucket_cursor[b], 1u);
        points_grouped[slot] = points[p];
    }
}
// the reverse (mlgi_ism_set_points: a grouped list injected by the caller): bucket id of every point
__global__ void __launch_bounds__(256) bucket_expand_kernel(const uint32_t* __restrict__ bucket_start, int N, uint32_t* __restrict__ point_bucket) {
    for (int b = blockIdx.x; b < N; b += gridDim.x)
        for (uint32_t p = bucket_start[b] + threadIdx.x; p < bucket_start[b + 1]; p += blockDim.x) point_bucket[p] = (uint32_t)b;
}

// single-block exclusive scan for the bucket histogram (N <= 16384): out has N+1 entries
__global__ void __launch_bounds__(SCAN_THREADS) bucket_scan_kernel(const uint32_t* __restrict__ hist, int N, uint32_t* __restrict__ out) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (int base = 0; base < N; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        uint32_t v = (i < N) ? hist[i] : 0u;
        uint32_t e = block_exclusive_scan(v, &total, ws);
        if (i < N) out[i] = carry + e;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[N] = carry;
}

int launch_group_points(const float4* points, const uint32_t* point_bucket, const uint32_t* generated, int max_points, int N,
                        uint32_t* bucket_hist, uint32_t* bucket_cursor, uint32_t* bucket_start, float4* points_grouped, cudaStream_t st) {
    cudaMemsetAsync(bucket_hist, 0, sizeof(uint32_t) * N, st);
    cudaMemsetAsync(bucket_cursor, 0, sizeof(uint32_t) * N, st);
    const int grid = 148 * 8;
    bucket_hist_kernel<<<grid, 256, 0, st>>>(point_bucket, generated, max_points, bucket_hist);
    bucket_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(bucket_hist, N, bucket_start);
    point_scatter_kernel<<<grid, 256, 0, st>>>(points, point_bucket, generated, max_points, bucket_start, bucket_cursor, points_grouped);
    return 3;
}
int launch_bucket_expand(const uint32_t* bucket_start, int N, uint32_t* point_bucket, cudaStream_t st) {
    bucket_expand_kernel<<<std::min(N, 148 * 8), 256, 0, st>>>(bucket_start, N, point_bucket);
    return 1;
}

// =====================================================================================
// a3: splat (ism.comp). Point-parallel: a quad of lanes owns one point of the (unordered) point list. The point's bucket b
// selects the 16-VPL window b..b+15 (ism.comp:60-69); each lane of the quad runs 4 of the 16 hemisphere tests (:94-107)
// against the hoisted VPL records, the quad ORs its masks, and lane q projects the q-th survivor (the reference keeps the
// first <= 4, :112) and writes it with atomicMin on (depth24<<8 | radius8). The window's lookAtRH rows come from the VPL
// record (32*s, 32*u: the same expressions the reference evaluates per work group, scaled by an exact power of two).
// No bucket sort and no shared memory: the kernel is latency-bound (long IEEE division / square-root chains), so it is laid
// out for parallelism -- every (point, survivor) projection is its own lane.
// =====================================================================================
__global__ void __launch_bounds__(256) splat_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ point_bucket,
                                                    const uint32_t* __restrict__ generated, int max_points,
                                                    const VplRecord* __restrict__ records, IsmMode mode, int S, float zFar,
                                                    uint32_t* __restrict__ atlas, int vpl_lo, int vpl_hi) {
    const uint32_t P = min(generated[0], (uint32_t)max_points);
    const int lane = threadIdx.x & 31, q = lane & 3;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const int N = mode.N;
    const float fA = float(64 * S);                            // atlas edge (fixed by totalVplCount)
    const float fS2 = float(mode.S2);
    const float invS = 1.0f / fS2;                             // S2 is a power of two: x * invS == x / S2 exactly
    const float inv_h = 2.0f / mode.px;                        // the record rows carry half the ISM edge (a power of two)
    const int A = 64 * S;
    const bool sharded = vpl_hi - vpl_lo < N;
    for (uint32_t w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u; w0 < P; w0 += warps_total * 8u) {   // warp-uniform
        const uint32_t pid = w0 + (uint32_t)(lane >> 2);
        bool have = pid < P;
        int b = 0;
        if (have) {
            b = (int)point_bucket[pid];
            if (sharded) {                                         // does the window [b, b+15] (mod N) reach this rank's ISMs?
                const int d0 = ((vpl_lo - b) % N + N) % N;         // distance from b to the first owned slot, going up
                have = d0 < 16 || (b >= vpl_lo && b < vpl_hi);
            }
        }
        float3 position = f3(0, 0, 0), pointNormal = f3(0, 0, 0);
        float pointRadius = 0.0f;
        uint32_t survivors = 0;   // bit i set: window VPL i passed both hemisphere tests
        if (have) {
            float4 rd = points[pid];
            position = f3(rd.x, rd.y, rd.z);
            uint32_t nr = __float_as_uint(rd.w);                   // ism.comp:85-87
            pointNormal = f3(float(nr & 0xFFu) / 255.0f * 2.0f - 1.0f, float((nr >> 8) & 0xFFu) / 255.0f * 2.0f - 1.0f,
                             float((nr >> 16) & 0xFFu) / 255.0f * 2.0f - 1.0f);
            pointRadius = float((nr >> 24) & 0xFFu) / 255.0f * 25.0f;
#pragma unroll
            for (int k = 0; k < 4; k++) {                          // :94-107, this lane's quarter of the window
                const int i = 4 * q + k;
                int id = (b + i) % N;                                                            // ism.comp:61-62
                if (mode.points_only) id = id % mode.vpl_count + mode.vpl_start;                  // :63-66
                const float4 r0 = __ldg(&records[id].r[0]);
                const float2 r1 = __ldg(reinterpret_cast<const float2*>(&records[id].r[1]));
                float3 rel = sub3(position, f3(-r0.x, -r0.y, -r0.z));
                bool cull = dot3(f3(r0.w, r1.x, r1.y), rel) < 0.0f || dot3(pointNormal, neg3(rel)) < 0.0f;
                if (!cull) survivors |= 1u << i;
            }
        }
        survivors |= __shfl_xor_sync(FULL, survivors, 1);
        survivors |= __shfl_xor_sync(FULL, survivors, 2);
        for (int k = 0; k < q; k++) survivors &= survivors - 1;   // lane q takes the q-th survivor in window order (:112-137)
        if (survivors != 0) {
            const int wi = __ffs(survivors) - 1;
            int gid = (b + wi) % N;
            if (mode.points_only) gid = gid % mode.vpl_count + mode.vpl_start;
            const bool owned = gid >= vpl_lo && gid < vpl_hi;   // multi-GPU shard: only this rank's ISMs
            if (!owned) continue;
            const float4 r0 = __ldg(&records[gid].r[0]), r1 = __ldg(&records[gid].r[1]), r2 = __ldg(&records[gid].r[2]);
            const float3 wpos = f3(-r0.x, -r0.y, -r0.z), wnrm = f3(r0.w, r1.x, r1.y);
            const float3 ws = f3(r1.z * inv_h, r2.x * inv_h, r2.z * inv_h);   // s = normalize(cross(nrm, up))  ism_utils.glsl:5-13
            const float3 wu = f3(r1.w * inv_h, r2.y * inv_h, r2.w * inv_h);   // u = cross(s, nrm)
            const float ismIndex = mode.scale ? float(gid) - float(mode.vpl_start) : float(gid);   // ism.comp:124
            const float yy = float(f2i(ismIndex / fS2));           // ism_utils.glsl:35-36
            const float tile_x = float(f2i(ismIndex - yy * fS2)), tile_y = float(f2i(yy));
            float3 rel = sub3(position, wpos);
            float dist = len3(rel);
            // paraboloid_project(rel, dist, nrm, zFar, id, S, true)  ism_utils.glsl:15-40
            float vx = (ws.x * rel.x + ws.y * rel.y) + ws.z * rel.z;
            float vy = (wu.x * rel.x + wu.y * rel.y) + wu.z * rel.z;
            float vz = -dot3(wnrm, rel);
            float sg = (vz > 0.0f) ? 1.0f : ((vz < 0.0f) ? -1.0f : 0.0f);
            vx = vx / dist; vy = vy / dist; vz = vz / dist;
            vz = 1.0f - vz;
            vx = vx / vz; vy = vy / vz;
            vz = dist / zFar;
            vz = vz * -sg;
            vx += 1.0f; vy += 1.0f;
            vx *= 0.5f; vy *= 0.5f;
            vx += tile_x; vy += tile_y;
            vx *= invS; vy *= invS;
            vx *= fA; vy *= fA;                                // ism.comp:131
            vz *= 16777216.0f;                                 // :132
            uint32_t depth24 = f2u(vz);
            if (depth24 > 0xFFFFFFu) depth24 = 0xFFFFFFu;      // Q10
            const uint32_t val = (depth24 << 8) | (f2u(pointRadius * 10.0f * 0.5f) & 0xFFu);  // :134-135  (/ sqrt(4.0))
            const int tx = f2i(vx), ty = f2i(vy);
            // imageAtomicMin :136. No warp aggregation: two lanes of a warp almost never hit the same texel, and ncu showed
            // match.any + redux costing 43% of the first version's stall samples; a plain RED.MIN per splat is what the L2
            // handles best here.
            if (tx >= 0 && ty >= 0 && tx < A && ty < A) atomicMin(&atlas[(uint32_t)ty * (uint32_t)A + (uint32_t)tx], val);
        }
    }
}

// ISMs [vpl_lo, vpl_hi) live in atlas tile rows [vpl_lo/S, ceil(vpl_hi/S)); only those rows are cleared.
void launch_atlas_clear(uint32_t* atlas, int S, int vpl_lo, int vpl_hi, cudaStream_t st) {
    const size_t A = 64 * (size_t)S;
    const int ty0 = vpl_lo / S, ty1 = (vpl_hi + S - 1) / S;
    cudaMemsetAsync(atlas + (size_t)ty0 * 64 * A, 0xFF, (size_t)(ty1 - ty0) * 64 * A * sizeof(uint32_t), st);   // ImperfectShadowmap.cpp:212
}
int launch_splat(const float4* points, const uint32_t* point_bucket, const uint32_t* generated, int max_points, const VplRecord* records,
                 const IsmMode& mode, int S, float zFar, uint32_t* atlas, int vpl_lo, int vpl_hi, bool clear, cudaStream_t st) {
    if (clear) launch_atlas_clear(atlas, S, vpl_lo, vpl_hi, st);
    splat_kernel<<<148 * 8, 256, 0, st>>>(points, point_bucket, generated, max_points, records, mode, S, zFar, atlas, vpl_lo, vpl_hi);
    return 1;
}

// =====================================================================================
// a4+a5: pull-push. One CTA per 64x64 ISM; the whole 6-level pyramid lives in shared memory
// (16 KB level 0 + 21.3 KB RGBA32F levels 1..6); HBM traffic = 4 B read + 2 B written per
// atlas texel instead of the reference's 12 full-atlas passes. Per-ISM independence: pull
// 2x2 blocks never straddle a 64-px tile; push masks every coarse tap of another ISM
// (push.comp:77-86). The arithmetic lives in pullpush_core.cuh as per-thread phases (also
// compiled for the host by tests/hostemu, which checks it against the oracle without a GPU);
// this kernel is the tile load, the phases with CTA barriers in between, and the tile store.
// ~42 KB of shared memory and <= 48 registers: 5 CTAs (40 warps) per SM.
// =====================================================================================
#ifndef PP_MINB
#define PP_MINB 5
#endif

__global__ void __launch_bounds__(pp::kThreads, PP_MINB) pullpush_kernel(const uint32_t* __restrict__ atlas, int S, float zFar,
                                                                         float ismPixelSize, uint16_t* __restrict__ ism, int tile_y0) {
    __shared__ __align__(16) pp::Shared s;
    const int A = 64 * S;
    const int gx0 = blockIdx.x * 64, gy0 = (blockIdx.y + tile_y0) * 64;
    const int tid = threadIdx.x;
    // ismPixelSize = A / ismIndices1d(ismCount) (ImperfectShadowmap.cpp:187-191): 64 unless scaleISMs. The 64x64 tile stays the
    // unit either way: pull blocks never straddle it, and push.comp:83 masks coarse taps across 64-texel borders whatever
    // the ISM size (an upstream inconsistency of the scaled mode, reproduced as it is).

    // ---- load the 64x64 R32UI tile: 1024 x uint4, coalesced 256 B rows ----
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = tid + k * pp::kThreads;  // uint4 index 0..1023
        v[k] = __ldg(reinterpret_cast<const uint4*>(atlas + (size_t)(gy0 + (q >> 4)) * A + gx0 + (q & 15) * 4));
    }
    pp::phase_tables(s, tid, zFar, ismPixelSize, gx0, gy0);      // tables + the all-empty block, while the loads are in flight
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = tid + k * pp::kThreads;
        *reinterpret_cast<uint4*>(&s.L0[(q >> 4) * 64 + (q & 15) * 4]) = v[k];
    }
    __syncthreads();
    pp::phase_pull1_classify(s, tid);
    __syncthreads();
    pp::phase_pull1_dense(s, tid, zFar, ismPixelSize, gx0, gy0);
    __syncthreads();
#pragma unroll 1
    for (int level = 2; level <= 6; level++) {
        pp::phase_pull(s, tid, level, gx0, gy0);
        __syncthreads();
    }
    // push levels 5..0, in place: push[i] overwrites pull[i]; only the owning thread reads the texel it replaces
#pragma unroll 1
    for (int level = 5; level >= 2; level--) {
        pp::phase_push_top(s, tid, level);
        __syncthreads();
    }
#ifdef PP_PUSH1_PER_PIXEL
    pp::phase_push_top(s, tid, 1);
#else
    pp::phase_push1(s, tid);
#endif
    __syncthreads();
    pp::phase_push0(s, tid);
    __syncthreads();
    // ---- store the R16 tile: 64 rows x 128 B, coalesced uint4 ----
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int q = tid + k * pp::kThreads;  // 0..511, 8 texels each
        const int row = q >> 3, c8 = q & 7;
        const uint4 a = *reinterpret_cast<const uint4*>(&s.L0[row * 64 + c8 * 8]);
        const uint4 b = *reinterpret_cast<const uint4*>(&s.L0[row * 64 + c8 * 8 + 4]);
        uint4 o;
        o.x = a.x | (a.y << 16); o.y = a.z | (a.w << 16); o.z = b.x | (b.y << 16); o.w = b.z | (b.w << 16);
        *reinterpret_cast<uint4*>(ism + (size_t)(gy0 + row) * A + gx0 + c8 * 8) = o;
    }
}

// TMA variant (default): the 64x64 R32UI tile arrives as ONE 2-D tensor-map copy (cp.async.bulk.tensor, 16 KB) issued by one
// thread and completing on an mbarrier while all threads fill the tables; the R16 result leaves as one tensor-map store of
// the packed 64x64x2 B tile. No register staging, no per-thread address arithmetic for the tile.
__global__ void __launch_bounds__(pp::kThreads, PP_MINB) pullpush_tma_kernel(const __grid_constant__ CUtensorMap tm_atlas,
                                                                             const __grid_constant__ CUtensorMap tm_ism, float zFar,
                                                                             float ismPixelSize, int tile_y0) {
    __shared__ __align__(128) pp::Shared s;
    __shared__ __align__(8) uint64_t bar;
    const int gx0 = blockIdx.x * 64, gy0 = (blockIdx.y + tile_y0) * 64;
    const int tid = threadIdx.x;
    if (tid == 0) { tma::mbar_init(&bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&bar, 64 * 64 * 4);
        tma::tile_g2s(s.L0, &tm_atlas, gx0, gy0, &bar);
    }
    pp::phase_tables(s, tid, zFar, ismPixelSize, gx0, gy0);      // tables + the all-empty block, while the tile is in flight
    tma::mbar_wait(&bar, 0);
    __syncthreads();
    pp::phase_pull1_classify(s, tid);
    __syncthreads();
    pp::phase_pull1_dense(s, tid, zFar, ismPixelSize, gx0, gy0);
    __syncthreads();
#pragma unroll 1
    for (int level = 2; level <= 6; level++) {
        pp::phase_pull(s, tid, level, gx0, gy0);
        __syncthreads();
    }
#pragma unroll 1
    for (int level = 5; level >= 2; level--) {
        pp::phase_push_top(s, tid, level);
        __syncthreads();
    }
#ifdef PP_PUSH1_PER_PIXEL
    pp::phase_push_top(s, tid, 1);
#else
    pp::phase_push1(s, tid);
#endif
    __syncthreads();
    pp::phase_push0(s, tid);
    __syncthreads();
    // pack the R16 codes (one per 32-bit word of L0) into a dense 64x64 uint16 tile in the pyramid's memory (dead by now)
    uint16_t* packed = reinterpret_cast<uint16_t*>(s.PY);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int q = tid + k * pp::kThreads;  // 0..511, 8 texels each
        const uint4 a = *reinterpret_cast<const uint4*>(&s.L0[q * 8]);
        const uint4 b = *reinterpret_cast<const uint4*>(&s.L0[q * 8 + 4]);
        uint4 o;
        o.x = a.x | (a.y << 16); o.y = a.z | (a.w << 16); o.z = b.x | (b.y << 16); o.w = b.z | (b.w << 16);
        *reinterpret_cast<uint4*>(packed + q * 8) = o;
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        tma::tile_s2g(&tm_ism, gx0, gy0, packed);
        tma::bulk_commit();
        tma::bulk_wait_read0();                // shared memory must stay valid until the copy engine has read it
    }
}

// test hook: the push phase's square-root elimination, evaluated on the device for caller-chosen radii
__global__ void sqrt_threshold_kernel(const float* __restrict__ rs, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pp::sqrt_gt_threshold(rs[i]);
}
int launch_sqrt_thresholds(const float* rs, int n, float* out, cudaStream_t st) {
    sqrt_threshold_kernel<<<(n + 255) / 256, 256, 0, st>>>(rs, n, out);
    return 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libmlgi does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
static bool make_tile_map(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, int A) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)A, (cuuint64_t)A}, strides[1] = {(cuuint64_t)A * elem_bytes};
    const cuuint32_t box[2] = {64, 64}, estr[2] = {1, 1};
    return fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_pullpush(const uint32_t* atlas, int S, float zFar, float ism_pixel_size, uint16_t* ism, int vpl_lo, int vpl_hi, cudaStream_t st) {
    const int ty0 = vpl_lo / S, ty1 = (vpl_hi + S - 1) / S;
    dim3 grid(S, ty1 - ty0);
    static const bool want_tma = [] { const char* e = getenv("MLGI_PP_TMA"); return !e || atoi(e) != 0; }();   // A/B switch: TMA tile copies (default) | LDG/STG
    CUtensorMap tm_atlas, tm_ism;
    if (want_tma && make_tile_map(&tm_atlas, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, atlas, 64 * S) &&
        make_tile_map(&tm_ism, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, ism, 64 * S)) {
        pullpush_tma_kernel<<<grid, pp::kThreads, 0, st>>>(tm_atlas, tm_ism, zFar, ism_pixel_size, ty0);
        return 1;
    }
    pullpush_kernel<<<grid, pp::kThreads, 0, st>>>(atlas, S, zFar, ism_pixel_size, ism, ty0);
    return 1;
}

// =====================================================================================
// a6: clustering. slice = #{s in 1..15 : -z >= 2^((s+3)/sf)} (table form of clustering.comp:50)
// =====================================================================================
__device__ __forceinline__ int slice_of(float zView, const SliceTable& tab) {
    float d = -zView;
    int s = 0;
#pragma unroll
    for (int i = 1; i <= 15; i++) if (d >= tab.b[i]) s = i;
    return s;
}

__global__ void __launch_bounds__(256) cluster_mask_kernel(const float* __restrict__ depth, int W, int H, float projA, float projB,
                                                           SliceTable tab, uint32_t* __restrict__ tile_masks) {
    __shared__ uint32_t smask;
    if (threadIdx.x == 0) smask = 0;
    __syncthreads();
    const int cx = blockIdx.x, cy = blockIdx.y;
    uint32_t m = 0;
    for (int p = threadIdx.x; p < 128 * 128; p += 256) {
        int fx = cx * 128 + (p & 127), fy = cy * 128 + (p >> 7);
        if (fx < W && fy < H) {                                   // clustering.comp:52
            float z = linear_depth(depth[(size_t)fy * W + fx], projA, projB);   // :48
            m |= 1u << slice_of(z, tab);                          // :50,54
        }
    }
    m = __reduce_or_sync(FULL, m);
    if ((threadIdx.x & 31) == 0) atomicOr(&smask, m);
    __syncthreads();
    if (threadIdx.x == 0) tile_masks[cy * gridDim.x + cx] = smask;
}

// canonical compact ids: ascending (tileY, tileX, slice)  (Q5)
__global__ void __launch_bounds__(SCAN_THREADS) cluster_compact_kernel(const uint32_t* __restrict__ tile_masks, int X, int Y,
                                                                        uint32_t* __restrict__ compact_ids, uint16_t* __restrict__ light_list_ids,
                                                                        uint32_t* __restrict__ num_used) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    uint32_t carry = 0;
    const int nt = X * Y;
    for (int base = 0; base < nt; base += SCAN_THREADS) {
        int t = base + threadIdx.x;
        uint32_t m = (t < nt) ? tile_masks[t] : 0u;
        uint32_t e = carry + block_exclusive_scan(__popc(m), &total, ws);
        if (t < nt) {
            int cx = t % X, cy = t / X;
#pragma unroll
            for (int s = 0; s < 16; s++) {
                uint16_t id = 0xFFFF;
                if (m & (1u << s)) {
                    compact_ids[e] = (uint32_t)cx | ((uint32_t)cy << 8) | ((uint32_t)s << 16);   // clustering.comp:83
                    id = (uint16_t)e;
                    e++;
                }
                light_list_ids[((size_t)s * Y + cy) * X + cx] = id;                                // :86
            }
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *num_used = carry;
}

int launch_clustering(const float* depth, int W, int H, float projA, float projB, const SliceTable& tab, uint32_t* tile_masks,
                      uint32_t* compact_ids, uint16_t* light_list_ids, uint32_t* num_used, cudaStream_t st) {
    const int X = (W + 127) / 128, Y = (H + 127) / 128;
    cluster_mask_kernel<<<dim3(X, Y), 256, 0, st>>>(depth, W, H, projA, projB, tab, tile_masks);
    cluster_compact_kernel<<<1, SCAN_THREADS, 0, st>>>(tile_masks, X, Y, compact_ids, light_list_ids, num_used);
    return 2;
}

// =====================================================================================
// a7 / a8 shared: one warp builds one (volume, class) sub-list. Lanes 0..7 unproject one
// corner each (light_lists.comp:85-90), every lane then tests VPLs lane, lane+32, ... of the
// class against the 8 corners; __ballot_sync + __popc give ascending, deterministic slots.
// =====================================================================================
__device__ __forceinline__ float3 unproject_corner(const Mat4& inv, float x, float y, float z) {
    float4 v = mat4_mul_vec4(inv, x * 2.0f - 1.0f, y * 2.0f - 1.0f, z * 2.0f - 1.0f, 1.0f * 2.0f - 1.0f);
    return f3(v.x / v.w, v.y / v.w, v.z / v.w);
}

__device__ __forceinline__ void build_sublist(float3 my_corner, const VplRecord* __restrict__ records, int L, int k,
                                              uint16_t* __restrict__ sub) {
    const int lane = threadIdx.x & 31;
    float3 corners[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        corners[j].x = __shfl_sync(FULL, my_corner.x, j);
        corners[j].y = __shfl_sync(FULL, my_corner.y, j);
        corners[j].z = __shfl_sync(FULL, my_corner.z, j);
    }
    uint32_t counter = 1;                                          // light_lists.comp:99
    for (int base = 0; base < L; base += 32) {
        const int t = base + lane;
        const uint32_t vplID = (uint32_t)(k * L + t);              // :104-106
        bool found = t >= L;                                       // lanes past the end do not hold up the early exit
        if (t < L) {
            // the record holds -pos and unpack3SN(packed normal), the very values light_lists.comp:108-110 derives
            // per test (corner - p == corner + (-p) exactly in IEEE arithmetic)
            const float4 r0 = __ldg(&records[vplID].r[0]);
            const float2 r1 = __ldg(reinterpret_cast<const float2*>(&records[vplID].r[1]));
            const float3 np = f3(r0.x, r0.y, r0.z), vn = f3(r0.w, r1.x, r1.y);
#pragma unroll
            for (int j = 0; j < 8; j++) {                          // :109-114 `found = found || ...` is order independent
                found = found || (dot3(add3(corners[j], np), vn) >= 0.0f);
            }
        }
        found = found && t < L;
        const unsigned bal = __ballot_sync(FULL, found);
        if (found) sub[counter + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)vplID;                // :116-119
        counter += __popc(bal);
    }
    if (lane == 0) sub[0] = (uint16_t)counter;                     // :124-125
}

__global__ void __launch_bounds__(128) light_lists_kernel(const VplRecord* __restrict__ records, int N, const uint32_t* __restrict__ compact_ids,
                                                          const uint32_t* __restrict__ num_used, Mat4 inv_vp, float projA, float projB,
                                                          int W, int H, SliceTable tab, int class_lo, int class_hi,
                                                          uint16_t* __restrict__ lists) {
    const int L = N / 16;
    const int ncls = class_hi - class_lo;
    const uint32_t pairs = num_used[0] * (uint32_t)ncls;
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    for (uint32_t pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pair < pairs; pair += warps_total) {
        const uint32_t id = pair / ncls;
        const int k = class_lo + (int)(pair % ncls);
        const uint32_t clusterID = compact_ids[id];
        const uint32_t clusterX = clusterID & 0xFFu, clusterY = (clusterID >> 8) & 0xFFu, clusterZ = (clusterID >> 16) & 0xFFu;
        // light_lists.comp:42-47 (Q4: slice 0 front = +nearPlane), :64-73
        float zFront = (clusterZ == 0) ? 0.05f : -tab.b[clusterZ];
        float zBack = -tab.b[clusterZ + 1];
        float ndcL = float(clusterX) * 128.0f / float(W);
        float ndcR = float(clusterX + 1) * 128.0f / float(W);
        float ndcB = float(clusterY) * 128.0f / float(H);
        float ndcT = float(clusterY + 1) * 128.0f / float(H);
        float ndcFront = log_depth(zFront, projA, projB);
        float ndcBack = log_depth(zBack, projA, projB);
        // corners 0..7 (:76-83). Q4: corners 6,7 repeat 4,5 (ndcL, ndcT).
        const int j = lane & 7;
        float cx = (j == 2 || j == 3) ? ndcR : ndcL;
        float cy = (j >= 4) ? ndcT : ndcB;
        float cz = (j & 1) ? ndcBack : ndcFront;
        float3 corner = unproject_corner(inv_vp, cx, cy, cz);
        build_sublist(corner, records, L, k, lists + ((size_t)id * 16 + k) * (L + 1));
    }
}

int launch_light_lists(const VplRecord* records, int N, const uint32_t* compact_ids, const uint32_t* num_used,
                       const Mat4& inv_vp, float projA, float projB, int W, int H, const SliceTable& tab, int class_lo,
                       int class_hi, uint16_t* lists, cudaStream_t st) {
    light_lists_kernel<<<148 * 8, 128, 0, st>>>(records, N, compact_ids, num_used, inv_vp, projA, projB, W, H, tab, class_lo,
                                                class_hi, lists);
    return 1;
}

// a8, two kernels. tile_bounds_kernel: one CTA per 64x64 tile unprojects the tile's pixels exactly as the gathering prologue
// does and reduces them to 14 numbers (position AABB, normal box, emin, emag) -- a streaming pass over the G-buffer (8 B per
// pixel). tile_lists_kernel: one warp per (tile, class) tests the class's VPLs against the 14 numbers (ordered ballot
// compaction). The rule is stated in DESIGN.md section 4 / oracle.cpp tile_keeps_vpl. Split in two because (1) the phases
// want different shapes (256 threads x 16 pixels at 40 registers vs one warp per sub-list), and (2) across GPUs the bounds are
// the exchange unit: every rank bounds only ITS tile rows and publishes 64 B per tile to the peers (bounds_peers), instead of
// every rank streaming the whole G-buffer.
#ifndef CULL_MINB
#define CULL_MINB 5
#endif
struct TileBounds { float pmin[3], pmax[3], nmin[3], nmax[3], emin, emag; };

__global__ void __launch_bounds__(256, CULL_MINB) tile_bounds_kernel(const float* __restrict__ depth, const uint32_t* __restrict__ face_normal, int W, int H,
                                                                    Mat4 inv_vp, int tile_y0, int ystride, BoundsOut bo) {
    __shared__ float sred[8][14];
    __shared__ float s_n10[1024];     // 10-bit channel -> v / 1023 * 2 - 1 (the IEEE division once per code instead of three per pixel)
    __shared__ float4 s_rowy[kTile];  // tile row -> column 1 of inv_vp times ndc.y = y / H * 2 - 1
    const int bx = blockIdx.x, by = tile_y0 + blockIdx.y * ystride;
    for (int i = threadIdx.x; i < 1024; i += 256) s_n10[i] = float(i) / 1023.0f * 2.0f - 1.0f;                 // final_gathering.comp:79
    if (threadIdx.x < kTile) {
        const float ny = float(by * kTile + threadIdx.x) / float(H) * 2.0f - 1.0f;                              // :70,74
        s_rowy[threadIdx.x] = make_float4(inv_vp.m[4] * ny, inv_vp.m[5] * ny, inv_vp.m[6] * ny, inv_vp.m[7] * ny);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float INF = __int_as_float(0x7f800000);
    float r[14];   // 0-2 pmin, 3-5 pmax, 6-8 nmin, 9-11 nmax, 12 emin, 13 emag
#pragma unroll
    for (int k = 0; k < 3; k++) { r[k] = INF; r[3 + k] = -INF; r[6 + k] = INF; r[9 + k] = -INF; }
    r[12] = INF; r[13] = 0.0f;
    {
        // A thread owns one column of the tile (x is fixed) and rows row0, row0 + 4, ...: everything of the unprojection
        // (:75, (m0*x + m1*y) + (m2*z + m3*1) per row of the matrix) that depends only on x is evaluated once per thread,
        // everything that depends only on y once per CTA (s_rowy); the per-pixel rest is 4 multiplies and 12 additions.
        // Same operations in the same order as the oracle (no FMA in this translation unit).
        constexpr int Q = kTile * kTile / 256, QB = 8;        // 16 pixels per thread, loads issued 8 at a time
        const int col = threadIdx.x & (kTile - 1), row0 = threadIdx.x / kTile;
        const int x = bx * kTile + col;
        const bool full = bx * kTile + kTile <= W && by * kTile + kTile <= H;   // CTA-uniform: no per-pixel bounds checks
        const float* m = inv_vp.m;
        const float nx = float(x) / float(W) * 2.0f - 1.0f;                    // final_gathering.comp:70,74
        const float mx0 = m[0] * nx, mx1 = m[1] * nx, mx2 = m[2] * nx, mx3 = m[3] * nx;
        const uint32_t off0 = (uint32_t)(by * kTile + row0) * (uint32_t)W + (uint32_t)x, step = 4u * (uint32_t)W;
#pragma unroll
        for (int qb = 0; qb < Q; qb += QB) {
            float dv[QB]; uint32_t nv[QB];
#pragma unroll
            for (int q = 0; q < QB; q++) {         // the batch's loads in flight before the first use
                const bool in = full || (x < W && by * kTile + row0 + 4 * (qb + q) < H);
                dv[q] = in ? __ldg(depth + off0 + (qb + q) * step) : 0.0f;
                nv[q] = in ? __ldg(face_normal + off0 + (qb + q) * step) : 0u;
            }
#pragma unroll
            for (int q = 0; q < QB; q++) {
                if (full || (x < W && by * kTile + row0 + 4 * (qb + q) < H)) {
                    const float4 ry = s_rowy[row0 + 4 * (qb + q)];
                    const float nz = dv[q] * 2.0f - 1.0f;
                    const float w0 = (mx0 + ry.x) + (m[8] * nz + m[12]), w1 = (mx1 + ry.y) + (m[9] * nz + m[13]);    // :75
                    const float w2 = (mx2 + ry.z) + (m[10] * nz + m[14]), w3 = (mx3 + ry.w) + (m[11] * nz + m[15]);
                    // bounds only: one IEEE reciprocal and three multiplies instead of the gathering pass's three divisions
                    // (:76). The position differs from the gathering pass's by <= 1.5 ulp, which the slack of the tile test
                    // covers many times over; the oracle evaluates the same expression.
                    float P[3], n[3];
                    const float rw = 1.0f / w3;
                    P[0] = w0 * rw; P[1] = w1 * rw; P[2] = w2 * rw;
                    n[0] = s_n10[nv[q] & 0x3FFu]; n[1] = s_n10[(nv[q] >> 10) & 0x3FFu]; n[2] = s_n10[(nv[q] >> 20) & 0x3FFu];   // :79
#pragma unroll
                    for (int c = 0; c < 3; c++) {   // FMNMX: equals the oracle's ternary min/max on the finite values seen here
                        r[c] = fminf(r[c], P[c]); r[3 + c] = fmaxf(r[3 + c], P[c]);
                        r[6 + c] = fminf(r[6 + c], n[c]); r[9 + c] = fmaxf(r[9 + c], n[c]);
                    }
                    const float ex = n[0] * P[0], ey = n[1] * P[1], ez = n[2] * P[2];
                    r[12] = fminf(r[12], (ex + ey) + ez);
                    r[13] = fmaxf(r[13], (fabsf(ex) + fabsf(ey)) + fabsf(ez));
                }
            }
        }
    }
    // min / max are exact and order-independent, so the reduction tree does not have to match the oracle's loop
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const bool is_min = i < 3 || (i >= 6 && i < 9) || i == 12;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float other = __shfl_xor_sync(FULL, r[i], o);
            r[i] = is_min ? fminf(r[i], other) : fmaxf(r[i], other);
        }
        if (lane == 0) sred[warp][i] = r[i];
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        const int i = threadIdx.x;
        const bool is_min = i < 3 || (i >= 6 && i < 9) || i == 12;
        float v = sred[0][i];
#pragma unroll
        for (int w = 1; w < 8; w++) v = is_min ? fminf(v, sred[w][i]) : fmaxf(v, sred[w][i]);
        const size_t e = ((size_t)by * gridDim.x + bx) * 16 + i;
        const size_t par = bo.frame_ctr ? (size_t)((*bo.frame_ctr + 1u) & 1u) * bo.parity_stride : 0;   // floats
        bo.local[par + e] = v;
#pragma unroll
        for (int g = 0; g < 8; g++) if (g < bo.num_peers) bo.peers[g][par + e] = v;      // multi-GPU: 64 B per tile to every other rank
    }
}

__global__ void __launch_bounds__(512) tile_lists_kernel(const float* __restrict__ bounds, size_t parity_stride, const uint32_t* __restrict__ frame_ctr,
                                                         const VplRecord* __restrict__ records, int N, int class_lo,
                                                         int class_hi, int tile_y0, int ystride, uint16_t* __restrict__ lists,
                                                         uint32_t* __restrict__ tile_weight) {
    __shared__ uint32_t s_weight;
    if (threadIdx.x == 0) s_weight = 0;
    __syncthreads();
    const int bx = blockIdx.x, by = tile_y0 + blockIdx.y * ystride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t tile = (size_t)by * gridDim.x + bx;
    TileBounds tb;
    {
        // (plain load: across GPUs the buffer was just written by the peers; the frame counter selects the parity buffer)
        const size_t par = frame_ctr ? (size_t)((*frame_ctr + 1u) & 1u) * parity_stride : 0;
        const float v = lane < 14 ? bounds[par + tile * 16 + lane] : 0.0f;
        float* t = reinterpret_cast<float*>(&tb);
#pragma unroll
        for (int i = 0; i < 14; i++) t[i] = __shfl_sync(FULL, v, i);
    }
    const int L = N / 16;
    for (int k = class_lo + warp; k < class_hi; k += nwarps) {
        uint16_t* sub = lists + (tile * 16 + k) * (L + 1);
        uint32_t counter = 1;
        constexpr int G = 4;                                  // chunks of 32 VPLs in flight per warp (ILP over the record loads)
        for (int base = 0; base < L; base += 32 * G) {
            float4 r0[G]; float2 r1[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int t = base + 32 * g + lane;
                const uint32_t id = (uint32_t)(k * L + (t < L ? t : 0));
                r0[g] = __ldg(&records[id].r[0]);                                         // (-pos, nrm.x)
                r1[g] = __ldg(reinterpret_cast<const float2*>(&records[id].r[1]));        // (nrm.y, nrm.z)
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int t = base + 32 * g + lane;
                const float vv[3] = {-r0[g].x, -r0[g].y, -r0[g].z}, mm[3] = {r0[g].w, r1[g].x, r1[g].y};
                float S1 = 0.0f, mA = 0.0f, tt = 0.0f, S2 = 0.0f, mB = 0.0f;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float lo = mm[c] * tb.pmin[c], hi = mm[c] * tb.pmax[c];
                    S1 = S1 + fmaxf(lo, hi);                              // finite operands: FMNMX == the oracle's ternary max
                    mA = mA + fmaxf(fabsf(lo), fabsf(hi));
                    const float tk = mm[c] * vv[c];
                    tt = tt + tk;
                    mA = mA + fabsf(tk);
                    const float l2 = tb.nmin[c] * vv[c], h2 = tb.nmax[c] * vv[c];
                    S2 = S2 + fmaxf(l2, h2);
                    mB = mB + fmaxf(fabsf(l2), fabsf(h2));
                }
                const bool keepA = (S1 - tt) > -(1e-5f * mA);
                const bool keepB = (S2 - tb.emin) > -(1e-5f * (mB + tb.emag));
                const bool found = t < L && keepA && keepB;
                const unsigned bal = __ballot_sync(FULL, found);
                if (found) sub[counter + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(k * L + t);
                counter += __popc(bal);
            }
        }
        if (lane == 0) { sub[0] = (uint16_t)counter; atomicAdd(&s_weight, counter - 1u); }
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_weight[tile] = s_weight;        // listed VPLs of this tile over the ctx's classes (scheduling weight)
}

// Gathering CTAs are scheduled in launch order, and a tile's cost is proportional to its list length: ordering the tiles
// by descending weight (longest first) removes the tail in which a few long-list tiles keep a handful of SMs busy. Counting
// sort into 128 weight bins by one CTA; the order inside a bin is arbitrary, which is immaterial (tiles are independent).
// The tiles to order are tile0 .. tile0 + ntiles - 1, or, with `ids`, the listed ones ids[tile0 .. tile0 + ntiles - 1] (a slab of
// interleaved tile rows of the host-input frame); the permutation is written to order[tile0 .. tile0 + ntiles - 1] either way.
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint32_t* __restrict__ tile_weight, const uint32_t* __restrict__ ids, int tile0, int ntiles,
                                                          int wmax, uint32_t* __restrict__ order) {
    __shared__ uint32_t hist[128], start[128];
    if (threadIdx.x < 128) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t div = (uint32_t)wmax / 128u + 1u;
    for (int t = threadIdx.x; t < ntiles; t += 1024) {
        const uint32_t id = ids ? ids[tile0 + t] : (uint32_t)(tile0 + t);
        atomicAdd(&hist[127u - min(tile_weight[id] / div, 127u)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t acc = 0; for (int b = 0; b < 128; b++) { start[b] = acc; acc += hist[b]; } }
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += 1024) {
        const uint32_t id = ids ? ids[tile0 + t] : (uint32_t)(tile0 + t);
        order[tile0 + atomicAdd(&start[127u - min(tile_weight[id] / div, 127u)], 1u)] = id;
    }
}

// tile rows ty0, ty0 + ystride, ... (nrows of them): the host-input frame culls each G-buffer slab as soon as it is uploaded
int launch_tile_bounds(const float* depth, const uint32_t* face_normal, int W, int H, const Mat4& inv_vp, int ty0, int ystride, int nrows,
                       const BoundsOut& bo, cudaStream_t st) {
    if (nrows <= 0) return 0;
    const int TX = (W + kTile - 1) / kTile;
    tile_bounds_kernel<<<dim3(TX, nrows), 256, 0, st>>>(depth, face_normal, W, H, inv_vp, ty0, ystride, bo);
    return 1;
}
int launch_tile_lists(const float* bounds, size_t parity_stride, const uint32_t* frame_ctr, int W, const VplRecord* records, int N, int class_lo, int class_hi, int ty0, int ystride, int nrows,
                      uint16_t* lists, uint32_t* tile_weight, cudaStream_t st) {
    if (nrows <= 0) return 0;
    const int TX = (W + kTile - 1) / kTile;
    const int warps = std::min(16, std::max(1, class_hi - class_lo));       // one warp per class of the tile
    tile_lists_kernel<<<dim3(TX, nrows), 32 * warps, 0, st>>>(bounds, parity_stride, frame_ctr, records, N, class_lo, class_hi, ty0, ystride, lists, tile_weight);
    return 1;
}
// the order, if requested, is a permutation of the tiles of rows [ty0, ty0 + nrows) (ystride 1), stored at their own positions
int launch_tile_order(const uint32_t* tile_weight, int W, int N, int class_lo, int class_hi, int ty0, int nrows, uint32_t* tile_order, cudaStream_t st) {
    const int TX = (W + kTile - 1) / kTile;
    tile_order_kernel<<<1, 1024, 0, st>>>(tile_weight, nullptr, ty0 * TX, nrows * TX, (class_hi - class_lo) * (N / 16), tile_order);
    return 1;
}
// the same for the tiles ids[first .. first + ntiles): order[first .. first + ntiles) = those tiles, longest list first
int launch_tile_order_ids(const uint32_t* tile_weight, const uint32_t* ids, int first, int ntiles, int N, int class_lo, int class_hi, uint32_t* tile_order, cudaStream_t st) {
    if (ntiles <= 0) return 0;
    tile_order_kernel<<<1, 1024, 0, st>>>(tile_weight, ids, first, ntiles, (class_hi - class_lo) * (N / 16), tile_order);
    return 1;
}

}  // namespace mlgi

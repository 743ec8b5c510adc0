// api.cu -- context, host orchestration and the C ABI declared in include/mlgi.h.
// Host-side matrix math follows glm's evaluation order in fp32 (the reference computes these
// on the host with glm: VPLProcessor.cpp:82,91, ClusteredShading.cpp:105, GIStage.cpp:289).
#include "common.cuh"

#include <algorithm>
#include <functional>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>

using namespace mlgi;

namespace {

std::string g_create_error;

// ---- glm-order fp32 matrix helpers (column-major) ----
void mat_mul(const float* a, const float* b, float* r) {        // r = a*b; column j = ((a0*b0j + a1*b1j) + a2*b2j) + a3*b3j
    float t[16];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++)
            t[4 * j + i] = ((a[i] * b[4 * j] + a[4 + i] * b[4 * j + 1]) + a[8 + i] * b[4 * j + 2]) + a[12 + i] * b[4 * j + 3];
    memcpy(r, t, sizeof(t));
}
void mat_inverse(const float* m, float* out) {                  // glm::inverse (cofactors)
#define M(c, r) m[4 * (c) + (r)]
    float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float F0[4] = {c00, c00, c02, c03}, F1[4] = {c04, c04, c06, c07}, F2[4] = {c08, c08, c10, c11};
    const float F3[4] = {c12, c12, c14, c15}, F4[4] = {c16, c16, c18, c19}, F5[4] = {c20, c20, c22, c23};
    const float V0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, V1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float V2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, V3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    const float SA[4] = {1, -1, 1, -1}, SB[4] = {-1, 1, -1, 1};
    float inv[16];
    for (int i = 0; i < 4; i++) {
        inv[0 + i] = ((V1[i] * F0[i] - V2[i] * F1[i]) + V3[i] * F2[i]) * SA[i];
        inv[4 + i] = ((V0[i] * F0[i] - V2[i] * F3[i]) + V3[i] * F4[i]) * SB[i];
        inv[8 + i] = ((V0[i] * F1[i] - V1[i] * F3[i]) + V3[i] * F5[i]) * SA[i];
        inv[12 + i] = ((V0[i] * F2[i] - V1[i] * F4[i]) + V2[i] * F5[i]) * SB[i];
    }
    const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
    const float det = (d0 + d1) + (d2 + d3);
    const float ood = 1.0f / det;
    for (int i = 0; i < 16; i++) out[i] = inv[i] * ood;
#undef M
}
Mat4 to_mat4(const float* p) { Mat4 r; memcpy(r.m, p, 64); return r; }

int ism_indices_1d(int n) {                                     // next even power of two (ism.comp:39)
    int l = 0;
    while ((1 << l) < n) l++;
    return 1 << ((l + 1) / 2);
}
SliceTable slice_table(float zFar) {                            // 2^((s+3)/sf), sf = 19/log2(zFar) (clustering.comp:22-24)
    SliceTable t;
    const double sf = 19.0 / std::log2((double)zFar);
    for (int s = 0; s <= 16; s++) t.b[s] = (float)std::exp2(double(s + 3) / sf);
    return t;
}

enum TimerId { T_VPLP = 0, T_POINTS, T_SPLAT, T_PULLPUSH, T_CLUSTERIDS, T_LIGHTLISTS, T_TILED, T_FG, T_BLUR, T_H2D, T_D2H, T_COUNT };
const char* kTimerNames[T_COUNT] = {"VPLP", "ISM render", "ISM CS", "PullPush", "ClusterIDs", "Light Lists", "Tiled cull", "FG", "GI blur", "H2D", "D2H"};

}  // namespace

struct mlgi_ctx {
    mlgi_config cfg;
    int S, A, L;                    // ISMs per atlas row, atlas edge, VPLs per class
    int ism_lo, ism_hi;             // ISMs this ctx builds: its classes' VPL slots widened to whole atlas tile rows plus one halo row
    int X, Y, TX, TY;               // clustered / tiled grids
    std::string err;
    unsigned long long launches = 0;
    // device buffers
    float* vpl = nullptr; float* packed = nullptr; VplRecord* records = nullptr; int32_t* shuffle = nullptr;
    float4* points_unsorted = nullptr; float4* points = nullptr; uint32_t* point_bucket = nullptr;
    uint32_t* bucket_hist = nullptr; uint32_t* bucket_cursor = nullptr; uint32_t* bucket_start = nullptr;
    uint32_t* counters = nullptr;   // [1] generated points (may exceed max_points: overflow), [2] num used clusters
    bool grouped_valid = false;     // points / bucket_start (bucket-grouped view of the point list) match points_unsorted / point_bucket
    uint32_t* atlas = nullptr; uint16_t* ism = nullptr;
    uint32_t* tile_masks = nullptr; uint32_t* compact_ids = nullptr; uint16_t* light_list_ids = nullptr;
    uint16_t* light_lists = nullptr; uint16_t* tile_lists = nullptr;
    uint32_t* tile_weight = nullptr; uint32_t* tile_order = nullptr;   // per-tile listed-VPL count and the longest-first tile permutation
    float* tile_bounds = nullptr;   // [tiles][16]: the 14 numbers the tile test uses (a8)
    uint32_t* slab_order = nullptr; // host-input frame: the tiles listed slab by slab (slab k = tile rows k, k + K, ...)
    int fg_sorted = 1; bool tile_order_valid = false;
    bool atlas_precleared = false;  // mlgi_frame clears the atlas on the side stream while the points are generated
    cudaEvent_t ev_clear = nullptr;
    uint32_t* gi_packed = nullptr;
    float4* gi = nullptr; float4* gi_planar = nullptr; float4* gi_blur = nullptr;
    unsigned long long* stats = nullptr;
    cudaTextureObject_t ism_tex = 0;      // ctx ISM atlas as an R16-unorm point-sampled texture
    // textures of caller-supplied ISM overrides, one per distinct pointer. A texture object may still be referenced by a
    // gathering kernel queued on some stream, so entries are only destroyed in mlgi_destroy (after a device synchronise)
    // or, when the cache is full, after an explicit cudaDeviceSynchronize.
    struct OverrideTex { const uint16_t* ptr; cudaTextureObject_t tex; };
    std::vector<OverrideTex> override_tex;
    int use_texture = 1, tex_float_coord = 0, fg_bulk = 1, xch_direct = 0;
    int fg_host_occ = 10;           // resident gathering CTAs per SM in the single-GPU host pipeline (A/B: MLGI_FG_HOST_OCC; 10 = no cap, the measured best)
    int cull_late = 0;              // start the culling after the splat instead of beside point generation (A/B: MLGI_CULL_LATE; measured 1.7595 vs 1.7529 ms: off)
    int pp_parts = 1;               // class groups of the pull-push / gathering pipeline in mlgi_frame (A/B: MLGI_PP_PARTS; measured: no gain, off)
    int xch_cull = 1;               // connected ranks bound 1/world of the tiles each and exchange the bounds (A/B: MLGI_XCH_CULL)
    IsmMode mode_vpl{}, mode_points{};   // the n3 knobs the VPL records / the point list were built with
    // multi-GPU exchange (mlgi_exchange_*): ONE IPC-exportable allocation per rank,
    //   [flags 256 B][class planes, parity 0][class planes, parity 1][G-buffer depth W*H*4][G-buffer normals W*H*4]
    //   [tile bounds, parity 0][tile bounds, parity 1]
    // flags[0..7]: "planes of frame f written" per rank; flags[16..23]: "tile bounds of frame f published" per rank;
    // flags[32..39] / flags[48..55]: the same two signals per PHASE of the host-input sharded frame (value 16 f + phase).
    struct Exchange {
        int rank = 0, world = 0;
        int push = 1;                         // 1: the gathering epilogue stores its staged 64-byte rows into every rank's buffer (default, measured)
                                              // 0: gathering writes its own buffer, the exchange kernel pulls over NVLink (MLGI_XCH=pull)
        char* base = nullptr;
        size_t buf_bytes = 0, gb_off = 0, bounds_off = 0, bounds_bytes = 0;
        // 64-pixel tile row t belongs to rank t mod world: the tiles it bounds, and in the host-input frame the G-buffer rows it
        // uploads and the image rows it returns (interleaved, so that every phase of that frame is a contiguous block of tile rows)
        char* peer_base[8] = {nullptr};       // peer-mapped allocations (own slot = base)
        uint32_t** d_peer_flags = nullptr;    // device array of world pointers (for the flag kernels)
        uint32_t* d_state = nullptr;          // [0] frame counter, [1] ticket, [2] error flag
        bool connected = false, local_peers = false;
    } xch;
    // CUDA graph of the frame (frame_graphed)
    int use_graph = 1, graph_topo = -1; bool graph_valid = false;
    cudaStream_t cap = nullptr; cudaGraphExec_t graph_exec = nullptr; std::vector<char> graph_key;
    unsigned long long graph_launches = 0, graph_replays = 0, graph_instantiations = 0;
    double last_enqueue_us = 0.0;
    // CUDA graph of the host-input frames (mlgi_frame_host*, mlgi_frame_sharded_host): uploads, stages and read-backs of one frame
    int host_graph = -1;            // -1: replay the sharded host frame only (measured at 4K on one GPU: the graph of mlgi_frame_host is 0.03-0.08 ms slower than
                                    // its direct launches, whose enqueueing overlaps the 2 ms of device work); 0 / 1: MLGI_HOST_GRAPH
    cudaGraphExec_t hgraph_exec = nullptr; std::vector<char> hgraph_key, hgraph_seen; bool hgraph_valid = false;
    unsigned long long hgraph_launches = 0, hgraph_replays = 0;
    float4* planar_ext = nullptr;   // caller-owned [16][PH][PW] float4 exchange buffer (mlgi_bind_planar_output)
    // staging for mlgi_frame_host
    float* st_depth = nullptr; uint32_t* st_normal = nullptr;
    uint32_t* st_rsm_samples = nullptr;     // device: [diffuse N | normal N | depth N] of the sampled RSM texels (host-input frames)
    uint32_t* h_rsm_samples = nullptr;      // pinned host staging of the same
    // stage bookkeeping
    bool have_vpl = false, have_points = false, have_atlas = false, have_ism = false, have_clustered = false, have_tiled = false, have_gi = false, have_blur = false;
    cudaStream_t side = nullptr;          // culling runs beside the ISM build (both only depend on the VPL stage)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_vpl = nullptr, ev_rsm = nullptr;
    // host-input frame (frame_host_impl): G-buffer slabs are uploaded (side), culled (aux), gathered (caller's stream) and
    // read back (aux) as a pipeline
    static constexpr int kSlabs = 16;
    cudaStream_t aux = nullptr, fg2 = nullptr;   // fg2: odd slabs' gathering launches, so that one slab's tail overlaps the next slab
    cudaStream_t rd = nullptr;                   // read-back of finished slabs (its own stream: the D2H copies must not queue behind the culling of later slabs)
    cudaEvent_t ev_ism = nullptr;
    cudaEvent_t ev_up[kSlabs] = {nullptr}, ev_cull[kSlabs] = {nullptr}, ev_fg[kSlabs] = {nullptr};
    // slab k of the single-GPU host pipeline = tile rows off, off + stride, ... (host_slab_plan)
    struct SlabPlan { int off, stride; };
    std::vector<SlabPlan> slab_plan;
    bool host_slabs_set = false;   // MLGI_HOST_SLABS given: also fixes the phase count of the sharded host frame (default there: by grid size)
    int host_slabs = 10;    // measured at 4K on PCIe 5 x16 (interleaved slabs, tiles sorted per slab): 4 slabs 2.16 ms, 6: 2.07, 8: 2.01, 10: 1.99, 12: 2.06, 16: 2.03
    cudaEvent_t ev[T_COUNT][2];
    bool ev_used[T_COUNT];
    bool events_created = false;
    int ev_launches[T_COUNT];
    std::vector<void*> allocs;
    // MLGI_HOST_TRACE=1: timeline of the host-input frame (timestamps of uploads, stages, read-backs) printed to stderr
    bool trace_on = false; size_t trace_n = 0;
    std::vector<std::pair<const char*, cudaEvent_t>> trace;
};

namespace {

int fail(mlgi_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
void trace_mark(mlgi_ctx* c, const char* name, cudaStream_t st) {
    if (!c->trace_on) return;
    if (c->trace_n == c->trace.size()) { cudaEvent_t e; cudaEventCreate(&e); c->trace.push_back({name, e}); }
    c->trace[c->trace_n].first = name;
    cudaEventRecord(c->trace[c->trace_n++].second, st);
}
void trace_dump(mlgi_ctx* c) {              // call after the frame's final synchronise
    if (!c->trace_on || c->trace_n == 0) return;
    fprintf(stderr, "[mlgi trace]");
    for (size_t i = 0; i < c->trace_n; i++) {
        float ms = 0.0f;
        cudaEventSynchronize(c->trace[i].second);
        cudaEventElapsedTime(&ms, c->trace[0].second, c->trace[i].second);
        fprintf(stderr, " %s=%.3f", c->trace[i].first, ms);
    }
    fprintf(stderr, "\n");
    c->trace_n = 0;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, MLGI_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// the R16 atlas as a texture: unnormalised coordinates, point sampling, clamp-to-edge, unorm -> float
// (what the reference's sampler does: ImperfectShadowmap.cpp:102-106, final_gathering.comp:122)
cudaError_t make_ism_texture(const uint16_t* ptr, int A, cudaTextureObject_t* out) {
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<uint16_t*>(ptr);
    rd.res.pitch2D.desc = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
    rd.res.pitch2D.width = A;
    rd.res.pitch2D.height = A;
    rd.res.pitch2D.pitchInBytes = (size_t)A * sizeof(uint16_t);
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    return cudaCreateTextureObject(out, &rd, &td, nullptr);
}

template <typename T>
cudaError_t dalloc(mlgi_ctx* c, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) c->allocs.push_back(*p);
    return e;
}

int check_params(mlgi_ctx* ctx, const mlgi_params* p) {
    if (!p) return fail(ctx, MLGI_E_INVALID, "params is NULL");
    if (p->vpl_start_index < 0 || p->vpl_end_index > ctx->cfg.vpl_count) return fail(ctx, MLGI_E_INVALID, "need 0 <= vpl_start_index < vpl_end_index <= vpl_count");
    if ((p->scale_isms || p->points_only_into_scaled_isms) && (ctx->cfg.class_lo != 0 || ctx->cfg.class_hi != 16))
        return fail(ctx, MLGI_E_UNSUPPORTED, "scaleISMs / pointsOnlyIntoScaledISMs are not available in a class-sharded ctx (the ISM rows of a class are no longer contiguous)");
    if (!p->use_push_pull) return fail(ctx, MLGI_E_UNSUPPORTED, "usePushPull=false (legacy point-sprite raster path) is out of scope");
    if (p->show_vpl_positions) return fail(ctx, MLGI_E_UNSUPPORTED, "SHOW_VPL_POSITIONS debug overlay is out of scope");
    if (!p->use_interleaving) return fail(ctx, MLGI_E_UNSUPPORTED, "USE_INTERLEAVING=false is broken upstream (SURVEY A.9) and out of scope");
    if (!(p->z_far > p->z_near) || !(p->z_near > 0.0f)) return fail(ctx, MLGI_E_INVALID, "need 0 < z_near < z_far");
    if (p->vpl_end_index <= p->vpl_start_index) return fail(ctx, MLGI_E_INVALID, "vpl_end_index must exceed vpl_start_index");
    return MLGI_OK;
}

// Every public entry point runs on the ctx device whatever device the calling thread had selected (a host may hold one ctx
// per GPU, or another library may have switched devices); the caller's device is restored on return.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(const mlgi_ctx* c) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != c->cfg.device) switched = cudaSetDevice(c->cfg.device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// n3: ism.geom:42-45 / ism.comp:36-39 / final_gathering.comp:42-46 / ImperfectShadowmap.cpp:187-191
IsmMode ism_mode(const mlgi_ctx* ctx, const mlgi_params* p) {
    IsmMode m;
    m.N = ctx->cfg.vpl_count;
    m.vpl_start = p->vpl_start_index;
    m.vpl_count = p->vpl_end_index - p->vpl_start_index;
    m.scale = p->scale_isms != 0; m.points_only = p->points_only_into_scaled_isms != 0;
    m.sampled = m.points_only ? m.vpl_count : m.N;
    m.ism_count = m.scale ? m.vpl_count : m.N;
    m.S2 = ism_indices_1d(m.ism_count);
    m.px = float(ctx->A / m.S2);
    return m;
}
bool same_mode(const IsmMode& a, const IsmMode& b) {
    return a.scale == b.scale && a.points_only == b.points_only && a.S2 == b.S2 && a.sampled == b.sampled && a.ism_count == b.ism_count &&
           (a.vpl_start == b.vpl_start || (!a.scale && !a.points_only));
}

struct StageTimer {
    mlgi_ctx* c; int id; cudaStream_t st; unsigned long long l0;
    StageTimer(mlgi_ctx* c_, int id_, cudaStream_t st_) : c(c_), id(id_), st(st_), l0(c_->launches) {
        if (c->cfg.enable_timers) cudaEventRecord(c->ev[id][0], st);
    }
    ~StageTimer() {
        if (c->cfg.enable_timers) { cudaEventRecord(c->ev[id][1], st); c->ev_used[id] = true; c->ev_launches[id] = (int)(c->launches - l0); }
    }
};

}  // namespace

// =====================================================================================
extern "C" {

MLGI_API int mlgi_version(void) { return 100; }

MLGI_API void mlgi_default_params(mlgi_params* p, int vpl_count) {   // GIStage.cpp:215-232
    memset(p, 0, sizeof(*p));
    p->light_intensity = 5.0f;
    p->gi_intensity_factor = 3000.0f;
    p->vpl_clamping_value = 0.001f;
    p->vpl_start_index = 0;
    p->vpl_end_index = vpl_count;
    p->tess_level_factor = 2.0f;
    p->use_push_pull = 1;
    p->enable_shadowing = 1;
    p->use_interleaving = 1;
    p->shuffle_lights = 1;
    p->z_near = 0.05f;
    p->z_far = 50.0f;
    p->cull_mode = MLGI_CULL_TILED;
}

MLGI_API const char* mlgi_last_error(const mlgi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

MLGI_API int mlgi_create(const mlgi_config* cfg, mlgi_ctx** out) {
    mlgi_ctx* ctx = nullptr;   // for fail()/CK before allocation
    if (!cfg || !out) return fail(nullptr, MLGI_E_INVALID, "cfg/out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MLGI_E_NO_DEVICE, "no CUDA device: libmlgi has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MLGI_E_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    const int N = cfg->vpl_count;
    if (N < 16 || N > 16384 || N % 16) return fail(nullptr, MLGI_E_INVALID, "vpl_count must be a multiple of 16 in [16,16384]");
    const int ax = (int)std::ceil(std::sqrt((float)N));
    if (N % ax) return fail(nullptr, MLGI_E_INVALID, "vpl_count must be a multiple of ceil(sqrt(N)) (vpl_processor.comp:38-50)");
    if (cfg->width < 1 || cfg->height < 1 || cfg->rsm_width < 1 || cfg->rsm_height < 1) return fail(nullptr, MLGI_E_INVALID, "bad image size");
    if ((cfg->width + 127) / 128 > 255 || (cfg->height + 127) / 128 > 255) return fail(nullptr, MLGI_E_INVALID, "viewport too large for 8-bit cluster coords (clustering.comp:83)");
    if (cfg->class_lo < 0 || cfg->class_hi > 16 || cfg->class_lo >= cfg->class_hi) return fail(nullptr, MLGI_E_INVALID, "need 0 <= class_lo < class_hi <= 16");
    if (cfg->max_points < 1 || cfg->max_triangles < 1) return fail(nullptr, MLGI_E_INVALID, "max_points / max_triangles must be positive");
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, MLGI_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    ctx = new mlgi_ctx();
    ctx->cfg = *cfg;
    ctx->S = ism_indices_1d(N);
    ctx->A = 64 * ctx->S;
    ctx->L = N / 16;
    {   // A shadow lookup can land one texel outside its own ISM (10-bit normals are not unit length), i.e. in the tile
        // row above or below; a class shard therefore also builds one halo row of ISMs on each side so that its pixels
        // are bit-identical to the unsharded frame.
        const int S = ctx->S, vlo = cfg->class_lo * ctx->L, vhi = cfg->class_hi * ctx->L;
        int r0 = vlo / S, r1 = (vhi + S - 1) / S;
        const int rows = (N + S - 1) / S;
        if (cfg->class_lo != 0 || cfg->class_hi != 16) { r0 = std::max(0, r0 - 1); r1 = std::min(rows, r1 + 1); }
        ctx->ism_lo = r0 * S;
        ctx->ism_hi = std::min(N, r1 * S);
    }
    ctx->X = (cfg->width + 127) / 128; ctx->Y = (cfg->height + 127) / 128;
    ctx->TX = (cfg->width + kTile - 1) / kTile; ctx->TY = (cfg->height + kTile - 1) / kTile;
    memset(ctx->ev_used, 0, sizeof(ctx->ev_used));
    memset(ctx->ev_launches, 0, sizeof(ctx->ev_launches));
    const size_t WH = (size_t)cfg->width * cfg->height, AA = (size_t)ctx->A * ctx->A, RS = (size_t)cfg->rsm_width * cfg->rsm_height;
    const size_t PW = (cfg->width + 3) / 4, PH = (cfg->height + 3) / 4;
    const size_t maxClusters = (size_t)ctx->X * ctx->Y * 16;
#define ALLOC(ptr, count)                                                                                     \
    do {                                                                                                      \
        cudaError_t e2 = dalloc(ctx, &ctx->ptr, (count));                                                     \
        if (e2 != cudaSuccess) {                                                                              \
            fail(nullptr, MLGI_E_CUDA, "cudaMalloc(" #ptr ", %zu): %s", (size_t)(count), cudaGetErrorString(e2)); \
            mlgi_destroy(ctx);                                                                                \
            return MLGI_E_CUDA;                                                                               \
        }                                                                                                     \
    } while (0)
    ALLOC(vpl, (size_t)N * 12); ALLOC(packed, (size_t)N * 4); ALLOC(records, (size_t)N); ALLOC(shuffle, (size_t)N);
    ALLOC(points_unsorted, (size_t)cfg->max_points); ALLOC(points, (size_t)cfg->max_points); ALLOC(point_bucket, (size_t)cfg->max_points);
    ALLOC(bucket_hist, (size_t)N); ALLOC(bucket_cursor, (size_t)N); ALLOC(bucket_start, (size_t)N + 1);
    ALLOC(counters, 4);
    ALLOC(atlas, AA); ALLOC(ism, AA);
    ALLOC(tile_masks, (size_t)ctx->X * ctx->Y); ALLOC(compact_ids, maxClusters); ALLOC(light_list_ids, maxClusters);
    // light_lists ([clusters][16][L+1] uint16: 1 GB at 8K / 16384 VPLs) is allocated by the first clustered cull (ensure_clustered)
    ALLOC(tile_lists, (size_t)ctx->TX * ctx->TY * 16 * (ctx->L + 1));
    ALLOC(tile_weight, (size_t)ctx->TX * ctx->TY); ALLOC(tile_order, (size_t)ctx->TX * ctx->TY); ALLOC(tile_bounds, (size_t)ctx->TX * ctx->TY * 16); ALLOC(slab_order, (size_t)ctx->TX * ctx->TY);
    ALLOC(gi, WH); ALLOC(gi_planar, (size_t)(cfg->class_hi - cfg->class_lo) * PW * PH); ALLOC(gi_blur, WH);
    ALLOC(stats, 2);
    if (cfg->gi_packed) ALLOC(gi_packed, WH);
    ALLOC(st_rsm_samples, (size_t)3 * N); ALLOC(st_depth, WH); ALLOC(st_normal, WH);
    if (cudaMallocHost((void**)&ctx->h_rsm_samples, (size_t)3 * N * sizeof(uint32_t)) != cudaSuccess) { fail(nullptr, MLGI_E_CUDA, "init: pinned RSM staging"); mlgi_destroy(ctx); return MLGI_E_CUDA; }
#undef ALLOC
    // VPLProcessor.cpp:58-65: fixed-seed shuffle table (libstdc++ std::shuffle; Q13)
    std::vector<int32_t> v(N);
    std::iota(v.begin(), v.end(), 0);
    std::mt19937 g(1979982);
    std::shuffle(v.begin(), v.end(), g);
    e = cudaMemcpy(ctx->shuffle, v.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(ctx->counters, 0, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->gi, 0, WH * sizeof(float4));
    // a class shard builds only ISMs [ism_lo, ism_hi); the discarded fetch of an unlit pixel in a pair may still touch
    // the other rows, so give them a defined value ("no occluder") once
    if (e == cudaSuccess) e = cudaMemset(ctx->ism, 0xFF, AA * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->atlas, 0xFF, AA * sizeof(uint32_t));
    if (e != cudaSuccess) { fail(nullptr, MLGI_E_CUDA, "init: %s", cudaGetErrorString(e)); mlgi_destroy(ctx); return MLGI_E_CUDA; }
    for (int i = 0; i < T_COUNT; i++) { cudaEventCreate(&ctx->ev[i][0]); cudaEventCreate(&ctx->ev[i][1]); }
    ctx->events_created = true;
    // The ISM chain (points -> splat -> pull-push) is the critical path into gathering; culling, forked beside it, only has
    // to be done by the time the chain ends. In the captured graph the chain's kernels therefore carry the highest stream
    // priority and the culling branch the lowest (kernel-node priorities are captured from the streams): when both have CTAs
    // to place, the chain goes first and culling fills what is left.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char* e9 = getenv("MLGI_PRIO")) { if (atoi(e9) == 0) prio_hi = prio_lo; }          // A/B switch
    cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_clear, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_vpl, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_rsm, cudaEventDisableTiming);
    cudaStreamCreateWithPriority(&ctx->cap, cudaStreamNonBlocking, prio_hi);
    if (const char* e12 = getenv("MLGI_HOST_GRAPH")) ctx->host_graph = atoi(e12) != 0;          // A/B switch: host-input frames replayed as a graph (default) | direct launches
    if (const char* e7 = getenv("MLGI_GRAPH")) ctx->use_graph = atoi(e7) != 0;                  // A/B switch: graph replay (default) | direct launches
    // aux carries the slab culling of the host-input frame: highest priority, so that its (small) kernels are placed as soon as
    // a slab has arrived instead of waiting behind the resident gathering CTAs of earlier slabs
    cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi);
    cudaStreamCreateWithFlags(&ctx->fg2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->rd, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_ism, cudaEventDisableTiming);
    for (int k = 0; k < mlgi_ctx::kSlabs; k++) {
        cudaEventCreateWithFlags(&ctx->ev_up[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_cull[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_fg[k], cudaEventDisableTiming);
    }
    if (const char* e10 = getenv("MLGI_HOST_TRACE")) ctx->trace_on = atoi(e10) != 0;
    if (const char* e5 = getenv("MLGI_HOST_SLABS")) { if (atoi(e5) > 0) ctx->host_slabs_set = true, ctx->host_slabs = std::max(1, std::min((int)mlgi_ctx::kSlabs, atoi(e5))); }   // A/B switch
    if (const char* e3 = getenv("MLGI_XCH_DIRECT")) ctx->xch_direct = atoi(e3) != 0;
    if (const char* e13 = getenv("MLGI_FG_HOST_OCC")) ctx->fg_host_occ = std::max(1, std::min(10, atoi(e13)));
    if (const char* e14 = getenv("MLGI_CULL_LATE")) ctx->cull_late = atoi(e14) != 0;
    if (const char* e4 = getenv("MLGI_PP_PARTS")) ctx->pp_parts = std::max(1, std::min(16, atoi(e4)));
    if (const char* e2 = getenv("MLGI_ISM_FETCH")) { ctx->use_texture = strcmp(e2, "ldg") != 0; ctx->tex_float_coord = strcmp(e2, "texf") == 0 ? 1 : strcmp(e2, "texi") == 0 ? 2 : strcmp(e2, "none") == 0 ? 3 : 0; }   // A/B switch: tex (default) | texf | texi | ldg
    if (const char* e11 = getenv("MLGI_XCH_CULL")) ctx->xch_cull = atoi(e11) != 0;             // A/B switch: bounds exchanged between ranks (default) | every rank bounds all tiles
    if (const char* e8 = getenv("MLGI_FG_BULK")) ctx->fg_bulk = atoi(e8) != 0;                 // A/B switch: records staged by cp.async.bulk (default) | LDG + STS
    if (const char* e6 = getenv("MLGI_FG_SORTED")) ctx->fg_sorted = atoi(e6) != 0;             // A/B switch: longest-list tiles first (default) | row-major
    if (make_ism_texture(ctx->ism, ctx->A, &ctx->ism_tex) != cudaSuccess) { cudaGetLastError(); ctx->ism_tex = 0; }
    {   // Slabs of the host-input frame, and its gathering order: tiles slab by slab. K uniform interleaved slabs (rows k, k + K, ...;
        // default 10, MLGI_HOST_SLABS = K). MLGI_HOST_SLABS=shaped selects seven slabs with 1/8, 1/4, 1/4, 1/8, 1/8, 1/16, 1/16 of the
        // tile rows (each still an arithmetic progression of rows: a small first slab that is ready when the ISM chain ends, small
        // last slabs because the frame ends with the last slab's upload -> cull -> gathering -> read-back) -- measured, no gain
        // (2.03 vs 2.01 / 1.99 ms for 8 / 10 uniform slabs): what delays a slab's gathering is its culling kernels waiting for SM
        // slots under the running gathering, and capping the gathering's occupancy to make room (MLGI_FG_HOST_OCC) costs more
        // than it saves (2.05 ms at 8 CTAs per SM).
        const char* e5 = getenv("MLGI_HOST_SLABS");
        if (e5 && strcmp(e5, "shaped") == 0 && ctx->TY >= 16) {
            const int plan[7][2] = {{0, 8}, {1, 4}, {2, 4}, {3, 8}, {7, 8}, {4, 16}, {12, 16}};
            for (auto& q : plan) ctx->slab_plan.push_back({q[0], q[1]});
        } else {
            const int K = std::min(ctx->host_slabs, ctx->TY);
            for (int k = 0; k < K; k++) ctx->slab_plan.push_back({k, K});
        }
        std::vector<uint32_t> order;
        order.reserve((size_t)ctx->TX * ctx->TY);
        for (auto& sp : ctx->slab_plan)
            for (int ty = sp.off; ty < ctx->TY; ty += sp.stride)
                for (int tx = 0; tx < ctx->TX; tx++) order.push_back((uint32_t)(ty * ctx->TX + tx));
        if (cudaMemcpy(ctx->slab_order, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            fail(nullptr, MLGI_E_CUDA, "init: slab order"); mlgi_destroy(ctx); return MLGI_E_CUDA;
        }
    }
    *out = ctx;
    return MLGI_OK;
}

MLGI_API void mlgi_destroy(mlgi_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard dg_(ctx);
    cudaDeviceSynchronize();
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_clear) cudaEventDestroy(ctx->ev_clear);
    if (ctx->ev_vpl) cudaEventDestroy(ctx->ev_vpl);
    if (ctx->ev_rsm) cudaEventDestroy(ctx->ev_rsm);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->hgraph_exec) cudaGraphExecDestroy(ctx->hgraph_exec);
    if (ctx->h_rsm_samples) cudaFreeHost(ctx->h_rsm_samples);
    if (ctx->cap) cudaStreamDestroy(ctx->cap);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->fg2) cudaStreamDestroy(ctx->fg2);
    if (ctx->rd) cudaStreamDestroy(ctx->rd);
    if (ctx->ev_ism) cudaEventDestroy(ctx->ev_ism);
    for (int k = 0; k < mlgi_ctx::kSlabs; k++) {
        if (ctx->ev_up[k]) cudaEventDestroy(ctx->ev_up[k]);
        if (ctx->ev_cull[k]) cudaEventDestroy(ctx->ev_cull[k]);
        if (ctx->ev_fg[k]) cudaEventDestroy(ctx->ev_fg[k]);
    }
    for (int g = 0; g < ctx->xch.world && !ctx->xch.local_peers; g++)
        if (g != ctx->xch.rank && ctx->xch.peer_base[g]) cudaIpcCloseMemHandle(ctx->xch.peer_base[g]);
    if (ctx->xch.base) cudaFree(ctx->xch.base);
    if (ctx->xch.d_peer_flags) cudaFree(ctx->xch.d_peer_flags);
    if (ctx->xch.d_state) cudaFree(ctx->xch.d_state);
    if (ctx->ism_tex) cudaDestroyTextureObject(ctx->ism_tex);
    for (auto& o : ctx->override_tex) cudaDestroyTextureObject(o.tex);
    for (auto& t : ctx->trace) cudaEventDestroy(t.second);
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->events_created)
        for (int i = 0; i < T_COUNT; i++) { cudaEventDestroy(ctx->ev[i][0]); cudaEventDestroy(ctx->ev[i][1]); }
    cudaGetLastError();
    delete ctx;
}

// ------------------------------------------------------------------------------------- a1
static int vpl_process_impl(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_params* p, void* stream, int compact);
MLGI_API int mlgi_vpl_process(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    return vpl_process_impl(ctx, rsm, light, p, stream, 0);
}
// The VPL build samples the RSM at N fixed texels (vpl_processor.comp:46-54: a regular sub-grid). The host-input frames gather
// those texels on the host (12 B per VPL: 48 KB at 4096 VPLs) and upload them instead of the 3 MB RSM -- per rank, so with G
// GPUs the job saves G x 3 MB of PCIe traffic per frame, and the VPL build that heads the ISM chain starts 0.06 ms earlier.
// the host-input frames read the RSM planes with the CPU: a device pointer there is a caller error, not a crash
static bool is_device_pointer(const void* ptr) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return pa.type == cudaMemoryTypeDevice;
}
static void gather_rsm_samples(mlgi_ctx* ctx, const mlgi_rsm* hr) {
    const uint32_t N = (uint32_t)ctx->cfg.vpl_count, rw = (uint32_t)ctx->cfg.rsm_width, rh = (uint32_t)ctx->cfg.rsm_height;
    const uint32_t ax = (uint32_t)ceilf(sqrtf(float(N))), ay = N / ax;              // the kernel's own arithmetic
    uint32_t* dst = ctx->h_rsm_samples;
    const float* hdepth = hr->depth;
    for (uint32_t idx = 0; idx < N; idx++) {
        const uint32_t tcx = (idx % ax) * rw / ax, tcy = (idx / ax) * rh / ay;
        const size_t t = (size_t)tcy * rw + tcx;
        dst[idx] = hr->diffuse[t];
        dst[N + idx] = hr->face_normal[t];
        memcpy(&dst[2 * N + idx], &hdepth[t], 4);
    }
}
static int vpl_process_impl(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_params* p, void* stream, int compact) {
    if (!rsm || !light || !rsm->diffuse || !rsm->face_normal || !rsm->depth) return fail(ctx, MLGI_E_INVALID, "rsm/light is NULL");
    int rc = check_params(ctx, p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // VPLProcessor.cpp:76-82,91: inverse(bias * P_light * V_light)
    const float bias[16] = {0.5f, 0, 0, 0, 0, 0.5f, 0, 0, 0, 0, 0.5f, 0, 0.5f, 0.5f, 0.5f, 1.0f};
    float bp[16], bpv[16], inv[16];
    mat_mul(bias, light->proj, bp);
    mat_mul(bp, light->view, bpv);
    mat_inverse(bpv, inv);
    StageTimer t(ctx, T_VPLP, st);
    ctx->launches += launch_vpl_process(*rsm, ctx->cfg.rsm_width, ctx->cfg.rsm_height, ctx->cfg.vpl_count, to_mat4(inv), p->light_intensity,
                                        p->shuffle_lights, ctx->shuffle, ism_mode(ctx, p), compact, ctx->vpl, ctx->packed, ctx->records, st);
    CK(cudaGetLastError());
    ctx->have_vpl = true; ctx->mode_vpl = ism_mode(ctx, p);
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a2
MLGI_API int mlgi_ism_points(mlgi_ctx* ctx, const mlgi_mesh* mesh, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!mesh) return fail(ctx, MLGI_E_INVALID, "mesh is NULL");
    if (mesh->num_triangles == 0) {            // no drawables (ImperfectShadowmap.cpp:225-240 draws nothing): no points
        cudaStream_t st0 = (cudaStream_t)stream;
        CK(cudaMemsetAsync(ctx->counters, 0, 2 * sizeof(uint32_t), st0));
        ctx->have_points = true; ctx->grouped_valid = false;
        return MLGI_OK;
    }
    if (!mesh->vertices || !mesh->normals || !mesh->indices || !mesh->prim_id) return fail(ctx, MLGI_E_INVALID, "mesh is NULL");
    if (mesh->num_triangles < 0) return fail(ctx, MLGI_E_INVALID, "negative triangle count");
    if (mesh->num_triangles > ctx->cfg.max_triangles) return fail(ctx, MLGI_E_CAPACITY, "mesh has %d triangles, ctx capacity %d", mesh->num_triangles, ctx->cfg.max_triangles);
    int rc = check_params(ctx, p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer t(ctx, T_POINTS, st);
    ctx->launches += launch_point_generate(*mesh, ism_mode(ctx, p), p->tess_level_factor, ctx->cfg.max_points, ctx->points_unsorted,
                                           ctx->point_bucket, ctx->counters + 1, st);
    CK(cudaGetLastError());
    ctx->have_points = true; ctx->grouped_valid = false; ctx->mode_points = ism_mode(ctx, p);
    return MLGI_OK;
}

// a caller-provided bucket-grouped point list (the reference's pointBuffer layout, compacted): becomes the ctx point list
MLGI_API int mlgi_ism_set_points(mlgi_ctx* ctx, const float* points, const uint32_t* bucket_start, int num_points, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!points || !bucket_start || num_points < 0) return fail(ctx, MLGI_E_INVALID, "points/bucket_start is NULL");
    if (num_points > ctx->cfg.max_points) return fail(ctx, MLGI_E_CAPACITY, "%d points exceed ctx capacity %d", num_points, ctx->cfg.max_points);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(ctx->points, points, sizeof(float4) * (size_t)num_points, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->points_unsorted, points, sizeof(float4) * (size_t)num_points, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->bucket_start, bucket_start, sizeof(uint32_t) * ((size_t)ctx->cfg.vpl_count + 1), cudaMemcpyDeviceToDevice, st));
    ctx->launches += launch_bucket_expand(ctx->bucket_start, ctx->cfg.vpl_count, ctx->point_bucket, st);
    const uint32_t np = (uint32_t)num_points;
    CK(cudaMemcpyAsync(ctx->counters + 1, &np, sizeof(np), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));              // np lives on this stack frame
    ctx->have_points = true; ctx->grouped_valid = true;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a3
MLGI_API int mlgi_ism_splat(mlgi_ctx* ctx, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    const bool clear = !ctx->atlas_precleared;      // consumed on every path, so a failed call cannot leave it set
    ctx->atlas_precleared = false;
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!ctx->have_vpl || !ctx->have_points) return fail(ctx, MLGI_E_ORDER, "mlgi_ism_splat needs mlgi_vpl_process and mlgi_ism_points first");
    const IsmMode mode = ism_mode(ctx, p);
    if (!same_mode(mode, ctx->mode_vpl)) return fail(ctx, MLGI_E_ORDER, "the VPL records were built with other scaleISMs / VPL-range settings: call mlgi_vpl_process with these params first");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer t(ctx, T_SPLAT, st);
    ctx->launches += launch_splat(ctx->points_unsorted, ctx->point_bucket, ctx->counters + 1, ctx->cfg.max_points, ctx->records,
                                  mode, ctx->S, p->z_far, ctx->atlas, ctx->ism_lo, ctx->ism_hi, clear, st);
    CK(cudaGetLastError());
    ctx->have_atlas = true;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a4+a5
MLGI_API int mlgi_ism_pullpush(mlgi_ctx* ctx, const uint32_t* atlas_override, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!atlas_override && !ctx->have_atlas) return fail(ctx, MLGI_E_ORDER, "mlgi_ism_pullpush needs mlgi_ism_splat first (or an atlas override)");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer t(ctx, T_PULLPUSH, st);
    ctx->launches += launch_pullpush(atlas_override ? atlas_override : ctx->atlas, ctx->S, p->z_far, ism_mode(ctx, p).px, ctx->ism,
                                     ctx->ism_lo, ctx->ism_hi, st);
    CK(cudaGetLastError());
    ctx->have_ism = true;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a6+a7
// Clustered culling stores compact cluster ids as uint16 with 0xFFFF = "unused" (the reference's R16UI lightListIds texture,
// ClusteredShading.cpp:122), so a viewport needs fewer than 65535 clusters; its worst-case list buffer is allocated on first
// use (tiled culling, the default, never touches it). Must run outside stream capture (cudaMalloc).
static int ensure_clustered(mlgi_ctx* ctx) {
    const size_t maxClusters = (size_t)ctx->X * ctx->Y * 16;
    if (maxClusters >= 65535) return fail(ctx, MLGI_E_UNSUPPORTED, "clustered culling needs X*Y*16 < 65535 clusters (16-bit list ids, ClusteredShading.cpp:122); this viewport has %zu -- use tiled culling", maxClusters);
    if (!ctx->light_lists) CK(dalloc(ctx, &ctx->light_lists, maxClusters * 16 * (ctx->L + 1)));
    return MLGI_OK;
}

MLGI_API int mlgi_cull_clustered(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!gb || !gb->depth || !cam) return fail(ctx, MLGI_E_INVALID, "gbuffer/camera is NULL");
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!ctx->have_vpl) return fail(ctx, MLGI_E_ORDER, "culling needs mlgi_vpl_process first");
    if ((rc = ensure_clustered(ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const SliceTable tab = slice_table(p->z_far);
    float pv[16], inv[16];
    mat_mul(cam->proj, cam->view, pv);          // ClusteredShading.cpp:105 inverse(projection * view)
    mat_inverse(pv, inv);
    {
        StageTimer t(ctx, T_CLUSTERIDS, st);
        ctx->launches += launch_clustering(gb->depth, ctx->cfg.width, ctx->cfg.height, cam->proj[10], cam->proj[14], tab, ctx->tile_masks,
                                           ctx->compact_ids, ctx->light_list_ids, ctx->counters + 2, st);
    }
    {
        StageTimer t(ctx, T_LIGHTLISTS, st);
        ctx->launches += launch_light_lists(ctx->records, ctx->cfg.vpl_count, ctx->compact_ids, ctx->counters + 2, to_mat4(inv), cam->proj[10], cam->proj[14], ctx->cfg.width, ctx->cfg.height,
                                            tab, ctx->cfg.class_lo, ctx->cfg.class_hi, ctx->light_lists, st);
    }
    CK(cudaGetLastError());
    ctx->have_clustered = true;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a8
// bounds + lists of tile rows ty0, ty0 + ystride, ... (nrows of them), single-GPU form (bounds stay in the ctx)
static int cull_tiled_rows(mlgi_ctx* ctx, const float* depth, const uint32_t* face_normal, const Mat4& inv, int ty0, int ystride, int nrows, cudaStream_t st) {
    BoundsOut bo;
    memset(&bo, 0, sizeof(bo));
    bo.local = ctx->tile_bounds;
    int n = launch_tile_bounds(depth, face_normal, ctx->cfg.width, ctx->cfg.height, inv, ty0, ystride, nrows, bo, st);
    n += launch_tile_lists(ctx->tile_bounds, 0, nullptr, ctx->cfg.width, ctx->records, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi, ty0, ystride, nrows,
                           ctx->tile_lists, ctx->tile_weight, st);
    return n;
}
MLGI_API int mlgi_cull_tiled(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!gb || !gb->depth || !gb->face_normal || !cam) return fail(ctx, MLGI_E_INVALID, "gbuffer/camera is NULL");
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!ctx->have_vpl) return fail(ctx, MLGI_E_ORDER, "culling needs mlgi_vpl_process first");
    cudaStream_t st = (cudaStream_t)stream;
    float vi[16], pi[16], inv[16];
    mat_inverse(cam->view, vi);                 // the gathering pass's own unprojection (GIStage.cpp:289), so both see the same positions
    mat_inverse(cam->proj, pi);
    mat_mul(vi, pi, inv);
    StageTimer t(ctx, T_TILED, st);
    ctx->launches += cull_tiled_rows(ctx, gb->depth, gb->face_normal, to_mat4(inv), 0, 1, ctx->TY, st);
    if (ctx->fg_sorted) ctx->launches += launch_tile_order(ctx->tile_weight, ctx->cfg.width, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi, 0, ctx->TY, ctx->tile_order, st);
    CK(cudaGetLastError());
    ctx->have_tiled = true; ctx->tile_order_valid = ctx->fg_sorted != 0;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a9
static int build_gather_args(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, const uint16_t* ism_override,
                             const mlgi_params* p, GatherArgs* out) {
    if (!gb || !gb->depth || !gb->face_normal || !cam) return fail(ctx, MLGI_E_INVALID, "gbuffer/camera is NULL");
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!ctx->have_vpl) return fail(ctx, MLGI_E_ORDER, "mlgi_final_gather needs mlgi_vpl_process first");
    if (p->enable_shadowing && !same_mode(ism_mode(ctx, p), ctx->mode_vpl))
        return fail(ctx, MLGI_E_ORDER, "the VPL records were built with other scaleISMs / VPL-range settings: call mlgi_vpl_process with these params first");
    if (p->enable_shadowing && !ism_override && !ctx->have_ism) return fail(ctx, MLGI_E_ORDER, "shadowing needs mlgi_ism_pullpush first (or an ISM override)");
    if (p->cull_mode == MLGI_CULL_CLUSTERED && !ctx->have_clustered) return fail(ctx, MLGI_E_ORDER, "clustered gathering needs mlgi_cull_clustered first");
    if (p->cull_mode == MLGI_CULL_TILED && !ctx->have_tiled) return fail(ctx, MLGI_E_ORDER, "tiled gathering needs mlgi_cull_tiled first");
    if (p->cull_mode < 0 || p->cull_mode > MLGI_CULL_NONE) return fail(ctx, MLGI_E_INVALID, "bad cull_mode");
    GatherArgs& a = *out;
    memset(&a, 0, sizeof(a));
    a.depth = gb->depth; a.face_normal = gb->face_normal; a.W = ctx->cfg.width; a.H = ctx->cfg.height;
    a.records = ctx->records; a.N = ctx->cfg.vpl_count; a.S = ctx->S;
    a.ism = ism_override ? ism_override : ctx->ism;
    a.ism_tex = 0;
    if (ctx->use_texture) {
        if (!ism_override) a.ism_tex = ctx->ism_tex;
        else {
            for (auto& o : ctx->override_tex) if (o.ptr == ism_override) a.ism_tex = o.tex;
            if (!a.ism_tex) {
                if (ctx->override_tex.size() >= 16) {       // bounded cache: nothing may still be sampling the old objects
                    cudaDeviceSynchronize();
                    for (auto& o : ctx->override_tex) cudaDestroyTextureObject(o.tex);
                    ctx->override_tex.clear();
                }
                cudaTextureObject_t t = 0;
                if (make_ism_texture(ism_override, ctx->A, &t) == cudaSuccess) { ctx->override_tex.push_back({ism_override, t}); a.ism_tex = t; }
                else cudaGetLastError();                    // no texture: the kernel falls back to plain loads of the override
            }
        }
    }
    a.cull_mode = p->cull_mode;
    a.lists.light_list_ids = ctx->light_list_ids;
    a.lists.lists = (p->cull_mode == MLGI_CULL_CLUSTERED) ? ctx->light_lists : ctx->tile_lists;
    a.lists.X = ctx->X; a.lists.Y = ctx->Y; a.lists.TX = ctx->TX;
    float vi[16], pi[16], vpi[16];
    mat_inverse(cam->view, vi);                 // GIStage.cpp:289: viewInverted * projectionInverted
    mat_inverse(cam->proj, pi);
    mat_mul(vi, pi, vpi);
    a.inv_vp = to_mat4(vpi);
    a.projA = cam->proj[10]; a.projB = cam->proj[14];
    a.tab = slice_table(p->z_far);
    a.zFar = p->z_far;
    a.giScale = p->gi_intensity_factor / float(p->vpl_end_index - p->vpl_start_index) * 16.0f;   // final_gathering.comp:132
    a.clampv = p->vpl_clamping_value;
    a.enable_shadowing = p->enable_shadowing;
    a.tile_begin = 0; a.tile_count = ctx->TX * ctx->TY;
    a.tex_float_coord = ctx->tex_float_coord;
    a.bulk_staging = ctx->fg_bulk;
    a.tile_order = (p->cull_mode == MLGI_CULL_TILED && ctx->fg_sorted && ctx->tile_order_valid) ? ctx->tile_order : nullptr;
    a.class_lo = ctx->cfg.class_lo; a.class_hi = ctx->cfg.class_hi;
    // output routing: a bound exchange buffer receives the class planes and the raster image is rebuilt by the
    // consumer (mlgi_planar_to_raster) after the all-gather; otherwise the raster image is written directly, plus
    // the ctx-owned planar sub-image when this ctx owns only some of the classes.
    const bool partial = ctx->cfg.class_lo != 0 || ctx->cfg.class_hi != 16;
    if (ctx->xch.connected) {
        a.out_raster = nullptr; a.planar_class0 = 0;
        a.peer_packed = ctx->cfg.gi_packed != 0;
        a.peer_parity_stride = ctx->xch.buf_bytes; a.frame_ctr = ctx->xch.d_state;
        if (ctx->xch.push) { a.num_peers = ctx->xch.world; for (int g = 0; g < ctx->xch.world; g++) a.peer_planar[g] = ctx->xch.peer_base[g] + 256; }
        else { a.num_peers = 1; a.peer_planar[0] = ctx->xch.base + 256; }
    } else if (ctx->planar_ext) {
        a.out_raster = nullptr; a.planar_class0 = 0;
        if (ctx->cfg.gi_packed) a.out_planar_packed = (uint32_t*)ctx->planar_ext; else a.out_planar = ctx->planar_ext;
    } else {
        a.out_raster = ctx->cfg.gi_packed == 2 ? nullptr : ctx->gi;      // gi_packed == 2: only the packed image is produced
        a.out_planar = (partial && ctx->cfg.gi_packed != 2) ? ctx->gi_planar : nullptr; a.planar_class0 = ctx->cfg.class_lo;
        a.out_packed = ctx->gi_packed;          // null unless cfg.gi_packed
    }
    a.stats = nullptr;
    return MLGI_OK;
}

MLGI_API int mlgi_final_gather(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, const uint16_t* ism_override,
                               const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    GatherArgs a;
    int rc = build_gather_args(ctx, gb, cam, ism_override, p, &a);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    {
        StageTimer t(ctx, T_FG, st);
        ctx->launches += launch_final_gather(a, st);
    }
    CK(cudaGetLastError());
    ctx->have_gi = true; ctx->have_blur = false;
    return MLGI_OK;
}

MLGI_API int mlgi_gather_stats(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, const uint16_t* ism_override,
                               const mlgi_params* p, unsigned long long* listed, unsigned long long* lit) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    GatherArgs a;
    int rc = build_gather_args(ctx, gb, cam, ism_override, p, &a);
    if (rc) return rc;
    a.stats = ctx->stats;
    a.out_raster = nullptr;     // counting variant: same loop, counters instead of stores
    a.out_planar = nullptr; a.out_packed = nullptr; a.out_planar_packed = nullptr; a.num_peers = 0;
    CK(cudaDeviceSynchronize());     // the legacy stream does not order against cudaStreamNonBlocking streams the stages ran on
    CK(cudaMemset(ctx->stats, 0, 2 * sizeof(unsigned long long)));
    ctx->launches += launch_final_gather(a, 0);
    unsigned long long h[2];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, ctx->stats, sizeof(h), cudaMemcpyDeviceToHost));
    if (listed) *listed = h[0];
    if (lit) *lit = h[1];
    return MLGI_OK;
}

MLGI_API int mlgi_bind_planar_output(mlgi_ctx* ctx, void* dev_planar16) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    ctx->planar_ext = (float4*)dev_planar16;
    ctx->graph_valid = false;
    return MLGI_OK;
}

MLGI_API int mlgi_planar_to_raster(mlgi_ctx* ctx, const void* dev_planar16, float* dev_raster, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!dev_planar16) return fail(ctx, MLGI_E_INVALID, "planar buffer is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->cfg.gi_packed) ctx->launches += launch_planar_packed_to_raster((const uint32_t*)dev_planar16, ctx->cfg.width, ctx->cfg.height,
                                                                           ctx->gi_packed, dev_raster ? (float4*)dev_raster : ctx->gi, st);
    else ctx->launches += launch_planar_to_raster((const float4*)dev_planar16, ctx->cfg.width, ctx->cfg.height, 0, 16,
                                             dev_raster ? (float4*)dev_raster : ctx->gi, st);
    CK(cudaGetLastError());
    ctx->have_gi = true; ctx->have_blur = false;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- multi-GPU exchange
MLGI_API int mlgi_exchange_create(mlgi_ctx* ctx, int rank, int world, void* out_handle64) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (world < 2 || world > 8 || 16 % world || rank < 0 || rank >= world || !out_handle64) return fail(ctx, MLGI_E_INVALID, "exchange needs world in {2,4,8} and 0 <= rank < world");
    if (ctx->cfg.class_lo != rank * (16 / world) || ctx->cfg.class_hi != (rank + 1) * (16 / world))
        return fail(ctx, MLGI_E_INVALID, "rank %d of %d must own interleave classes [%d,%d)", rank, world, rank * (16 / world), (rank + 1) * (16 / world));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    auto& x = ctx->xch;
    if (x.base) return fail(ctx, MLGI_E_INVALID, "exchange already created");
    const size_t PW = (ctx->cfg.width + 3) / 4, PH = (ctx->cfg.height + 3) / 4, WH = (size_t)ctx->cfg.width * ctx->cfg.height;
    x.buf_bytes = ((16 * PW * PH * (ctx->cfg.gi_packed ? 4 : 16)) + 255) & ~(size_t)255;
    x.gb_off = 256 + 2 * x.buf_bytes;
    x.bounds_off = x.gb_off + 2 * ((WH * 4 + 255) & ~(size_t)255);
    x.bounds_bytes = (((size_t)ctx->TX * ctx->TY * 16 * sizeof(float)) + 255) & ~(size_t)255;
    const size_t total = x.bounds_off + 2 * x.bounds_bytes;
    CK(cudaMalloc((void**)&x.base, total));
    CK(cudaMemset(x.base, 0, total));
    CK(cudaMalloc((void**)&x.d_peer_flags, 8 * sizeof(uint32_t*)));
    CK(cudaMalloc((void**)&x.d_state, 4 * sizeof(uint32_t)));
    CK(cudaMemset(x.d_state, 0, 4 * sizeof(uint32_t)));
    x.rank = rank; x.world = world;
    // measured at 4K/4096 (device-resident frame, profiles/r2_sweep_host_frame.txt): pull / push 0.386 / 0.358 ms on 8 GPUs, 0.565 / 0.540 on 4,
    // 0.965 / 0.945 on 2: a pull is an exposed tail after the gathering, pushed planes travel under it
    x.push = 1;
    if (const char* e = getenv("MLGI_XCH")) x.push = strcmp(e, "push") == 0;      // A/B switch
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, x.base));
    memcpy(out_handle64, &h, 64);
    return MLGI_OK;
}

MLGI_API int mlgi_exchange_connect(mlgi_ctx* ctx, const void* all_handles) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    auto& x = ctx->xch;
    if (!x.base || !all_handles) return fail(ctx, MLGI_E_ORDER, "mlgi_exchange_connect needs mlgi_exchange_create first");
    uint32_t* flags[8] = {nullptr};
    for (int g = 0; g < x.world; g++) {
        if (g == x.rank) x.peer_base[g] = x.base;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)all_handles + 64 * g, 64);
            void* p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            x.peer_base[g] = (char*)p;
        }
        flags[g] = (uint32_t*)x.peer_base[g];
    }
    CK(cudaMemcpy(x.d_peer_flags, flags, sizeof(flags), cudaMemcpyHostToDevice));
    x.connected = true;
    ctx->graph_valid = false;
    return MLGI_OK;
}

// same-process variant for a host that drives several GPUs from one process (e.g. the C++ GIStage mirror, or the
// single-GPU sequential-shard test): the peers' allocations are handed over as plain device pointers
MLGI_API int mlgi_exchange_connect_local(mlgi_ctx* ctx, mlgi_ctx* const* all_ctx) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    auto& x = ctx->xch;
    if (!x.base || !all_ctx) return fail(ctx, MLGI_E_ORDER, "mlgi_exchange_connect_local needs mlgi_exchange_create first");
    uint32_t* flags[8] = {nullptr};
    for (int g = 0; g < x.world; g++) {
        if (!all_ctx[g] || !all_ctx[g]->xch.base) return fail(ctx, MLGI_E_ORDER, "rank %d has no exchange allocation", g);
        if (all_ctx[g]->cfg.device != ctx->cfg.device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->cfg.device, all_ctx[g]->cfg.device));
            if (!can) return fail(ctx, MLGI_E_UNSUPPORTED, "device %d cannot access device %d", ctx->cfg.device, all_ctx[g]->cfg.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(all_ctx[g]->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, MLGI_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        x.peer_base[g] = all_ctx[g]->xch.base;
        flags[g] = (uint32_t*)x.peer_base[g];
    }
    for (int g = 0; g < x.world; g++) if (g != x.rank) x.peer_base[g] = all_ctx[g]->xch.base;
    CK(cudaMemcpy(x.d_peer_flags, flags, sizeof(flags), cudaMemcpyHostToDevice));
    x.connected = true; x.local_peers = true;
    ctx->graph_valid = false;
    return MLGI_OK;
}

static void build_exchange_args(mlgi_ctx* ctx, ExchangeArgs* out) {
    auto& x = ctx->xch;
    ExchangeArgs& e = *out;
    memset(&e, 0, sizeof(e));
    e.W = ctx->cfg.width; e.H = ctx->cfg.height; e.world = x.world; e.rank = x.rank; e.packed = ctx->cfg.gi_packed != 0;
    for (int g = 0; g < x.world; g++) {
        e.src[g] = (x.push ? x.base : x.peer_base[g]) + 256;
        e.peer_flags[g] = (uint32_t*)x.peer_base[g];
    }
    e.parity_stride = x.buf_bytes;
    e.flags = (const uint32_t*)x.base;
    e.frame_ctr = x.d_state; e.ticket = x.d_state + 1; e.error = x.d_state + 2;
    e.raster_packed = ctx->gi_packed;
    e.raster = ctx->cfg.gi_packed == 2 ? nullptr : ctx->gi;
    e.row_begin = 0; e.row_block = ctx->cfg.height; e.row_stride = ctx->cfg.height; e.row_blocks = 1;
    e.signal_wait = 1; e.advance = 1;
}

// Tiled culling of one rank of a connected job (inside mlgi_frame_sharded / mlgi_frame_sharded_host, which every rank calls
// once per frame): the rank bounds only ITS tile rows (t mod world == rank) -- 1/world of the G-buffer instead of all of it -- and stores the 14
// numbers per tile into every rank's bounds buffer (64 B per tile over NVLink, 130 KB per frame at 4K); flags order the
// exchange (bounded wait), then the rank builds the lists of its classes for all tiles. Same bounds, same lists as the
// single-GPU call.
static int cull_tiled_sharded(mlgi_ctx* ctx, const float* depth, const uint32_t* face_normal, const Mat4& inv, cudaStream_t st) {
    auto& x = ctx->xch;
    BoundsOut bo;
    memset(&bo, 0, sizeof(bo));
    bo.local = (float*)(x.base + x.bounds_off);
    for (int g = 0; g < x.world; g++) if (g != x.rank) bo.peers[bo.num_peers++] = (float*)(x.peer_base[g] + x.bounds_off);
    bo.parity_stride = x.bounds_bytes / sizeof(float);
    bo.frame_ctr = x.d_state;
    const int own = x.rank < ctx->TY ? (ctx->TY - x.rank + x.world - 1) / x.world : 0;      // tile rows rank, rank + world, ...
    int n = launch_tile_bounds(depth, face_normal, ctx->cfg.width, ctx->cfg.height, inv, x.rank, x.world, own, bo, st);
    n += launch_flag_signal(x.d_peer_flags, x.world, 16 + x.rank, x.d_state, 1u, 0u, st);
    n += launch_flag_wait((const uint32_t*)x.base + 16, x.world, x.d_state, x.d_state + 2, 1u, 0u, st);
    n += launch_tile_lists(bo.local, bo.parity_stride, x.d_state, ctx->cfg.width, ctx->records, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi,
                           0, 1, ctx->TY, ctx->tile_lists, ctx->tile_weight, st);
    if (ctx->fg_sorted) n += launch_tile_order(ctx->tile_weight, ctx->cfg.width, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi, 0, ctx->TY, ctx->tile_order, st);
    ctx->have_tiled = true; ctx->tile_order_valid = ctx->fg_sorted != 0;
    return n;
}

// signal the peers that this rank's planes of the current frame are written, wait for every rank's signal, then assemble
// the image: MLGI_BUF_GI_PACKED / MLGI_BUF_GI. One kernel; call after mlgi_frame on the same stream (mlgi_frame_sharded
// does both). Asynchronous; bounded wait.
MLGI_API int mlgi_exchange_finish(mlgi_ctx* ctx, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    auto& x = ctx->xch;
    if (!x.connected) return fail(ctx, MLGI_E_ORDER, "exchange is not connected");
    ExchangeArgs e;
    build_exchange_args(ctx, &e);
    ctx->launches += launch_exchange(e, (cudaStream_t)stream);
    CK(cudaGetLastError());
    ctx->have_gi = true; ctx->have_blur = false;
    return MLGI_OK;
}

MLGI_API int mlgi_exchange_status(mlgi_ctx* ctx) {      // synchronises; 0 = ok, MLGI_E_CUDA if a peer never signalled
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!ctx->xch.d_state) return MLGI_OK;
    uint32_t e = 0;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&e, ctx->xch.d_state + 2, 4, cudaMemcpyDeviceToHost));
    if (e) return fail(ctx, MLGI_E_CUDA, "exchange timed out waiting for rank %u", e - 1);
    return MLGI_OK;
}

MLGI_API int mlgi_set_timers(mlgi_ctx* ctx, int enable) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    ctx->cfg.enable_timers = enable ? 1 : 0;
    return MLGI_OK;
}

MLGI_API int mlgi_debug_sqrt_thresholds(mlgi_ctx* ctx, const float* dev_radii, int n, float* dev_thresholds, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!dev_radii || !dev_thresholds || n < 0) return fail(ctx, MLGI_E_INVALID, "radii/thresholds is NULL");
    ctx->launches += launch_sqrt_thresholds(dev_radii, n, dev_thresholds, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return MLGI_OK;
}

MLGI_API float mlgi_measure_fp32_peak_tflops(mlgi_ctx* ctx) {
    if (!ctx) return -1.0f;
    DeviceGuard dg_(ctx);
    return measure_fp32_peak_tflops(ctx->cfg.device, 0);
}

// ------------------------------------------------------------------------------------- n1
MLGI_API int mlgi_gi_blur(mlgi_ctx* ctx, const mlgi_gbuffer* gb, const mlgi_camera* cam, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!gb || !gb->depth || !gb->face_normal || !cam) return fail(ctx, MLGI_E_INVALID, "gbuffer/camera is NULL");
    if (!ctx->have_gi) return fail(ctx, MLGI_E_ORDER, "mlgi_gi_blur needs mlgi_final_gather first");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer t(ctx, T_BLUR, st);
    // packed-only contexts blur the GL_R11F_G11F_B10F image -- which is what the reference blurs (its giBuffer has that format)
    ctx->launches += launch_gi_blur(ctx->gi, ctx->cfg.gi_packed == 2 ? ctx->gi_packed : nullptr, gb->depth, gb->face_normal, ctx->cfg.width,
                                    ctx->cfg.height, cam->proj[10], cam->proj[14], ctx->gi_blur, st);
    CK(cudaGetLastError());
    ctx->have_blur = true;
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- n4 (ambient term only)
MLGI_API int mlgi_composite_ambient(mlgi_ctx* ctx, const uint32_t* diffuse_rgba8, const float* occlusion, float exposure, int use_blurred,
                                    uint32_t* out_srgb8, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!diffuse_rgba8 || !out_srgb8) return fail(ctx, MLGI_E_INVALID, "diffuse/out is NULL");
    if (!ctx->have_gi) return fail(ctx, MLGI_E_ORDER, "mlgi_composite_ambient needs mlgi_final_gather first");
    if (use_blurred && !ctx->have_blur) return fail(ctx, MLGI_E_ORDER, "mlgi_composite_ambient(use_blurred) needs mlgi_gi_blur first");
    ctx->launches += launch_composite_ambient(use_blurred ? ctx->gi_blur : ctx->gi, diffuse_rgba8, occlusion, ctx->cfg.width * ctx->cfg.height,
                                              exposure, out_srgb8, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- a10
// The frame as direct launches on `st` (also what gets captured into the CUDA graph).
static int frame_body(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                      const mlgi_camera* cam, const mlgi_params* p, bool with_exchange, cudaStream_t st) {
    int rc;
    void* stream = (void*)st;
    // The reference runs VPL build, ISM build and culling back to back on one GL queue (GIStage.cpp:373-403). Their real
    // dependencies are: point generation needs nothing from this frame, the atlas clear nothing at all, splat needs points +
    // VPL records + the cleared atlas, culling needs the VPL records, gathering needs the ISM and the lists. So the frame
    // forks three ways after the caller's stream position -- [points] on the caller's stream, [VPL build -> culling] on the
    // side stream, [atlas clear] on the aux stream -- and joins before the splat and before gathering.
    const bool fork = ctx->side != nullptr && ctx->aux != nullptr && p->cull_mode != MLGI_CULL_NONE;
    if (fork) {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
        launch_atlas_clear(ctx->atlas, ctx->S, ctx->ism_lo, ctx->ism_hi, ctx->aux);
        CK(cudaEventRecord(ctx->ev_clear, ctx->aux));
        void* cs = (void*)ctx->side;
        if ((rc = mlgi_vpl_process(ctx, rsm, light, p, cs))) return rc;           // GIStage.cpp:373-376
        CK(cudaEventRecord(ctx->ev_vpl, ctx->side));
    } else {
        if ((rc = mlgi_vpl_process(ctx, rsm, light, p, stream))) return rc;
    }
    // culling on the side stream (fork only): right after the VPL build, or (cull_late) only once the splat is done, so that it
    // runs beside the throughput-bound pull-push instead of stretching the latency-bound point generation and splat
    auto issue_cull = [&]() -> int {
        void* cs = (void*)ctx->side;
        int r;
        if (p->cull_mode == MLGI_CULL_CLUSTERED) { if ((r = mlgi_cull_clustered(ctx, gb, cam, p, cs))) return r; }   // :392-403
        else if (with_exchange && ctx->xch_cull) {          // every rank bounds 1/world of the tiles; the bounds are exchanged
            if (!gb || !gb->depth || !gb->face_normal || !cam) return fail(ctx, MLGI_E_INVALID, "gbuffer/camera is NULL");
            float vi[16], pi[16], inv[16];
            mat_inverse(cam->view, vi); mat_inverse(cam->proj, pi); mat_mul(vi, pi, inv);
            StageTimer t(ctx, T_TILED, ctx->side);
            ctx->launches += cull_tiled_sharded(ctx, gb->depth, gb->face_normal, to_mat4(inv), ctx->side);
            CK(cudaGetLastError());
        }
        else { if ((r = mlgi_cull_tiled(ctx, gb, cam, p, cs))) return r; }
        CK(cudaEventRecord(ctx->ev_join, ctx->side));
        return MLGI_OK;
    };
    const bool cull_late = fork && ctx->cull_late && !with_exchange;      // (sharded frames: the bounds exchange wants an early start)
    if (fork && !cull_late && (rc = issue_cull())) return rc;
    if ((rc = mlgi_ism_points(ctx, mesh, p, stream))) return rc;                  // :378-389 ISM render
    if (fork) {
        CK(cudaStreamWaitEvent(st, ctx->ev_vpl, 0));
        CK(cudaStreamWaitEvent(st, ctx->ev_clear, 0));
        ctx->atlas_precleared = true;
    }
    if ((rc = mlgi_ism_splat(ctx, p, stream))) return rc;                         //          ISM CS
    if (cull_late) {
        CK(cudaEventRecord(ctx->ev_rsm, st));                                     // (ev_rsm is free outside the host-input frames)
        CK(cudaStreamWaitEvent(ctx->side, ctx->ev_rsm, 0));
        if ((rc = issue_cull())) return rc;
    }
    // Optional (MLGI_PP_PARTS = 2..16, default off): pull-push and gathering as a pipeline over class groups. Gathering for classes
    // [c0, c1) only reads the ISMs of those classes' VPLs, i.e. a band of atlas rows (plus one halo row), so the pull-push of the
    // NEXT band (ALU-pipe bound) can run beside the gathering of this one (FMA-pipe bound) instead of in front of all of it
    // (stream events only: bands in order on the caller's stream, gathering groups in order on fg2). Measured at 4K/4096: 1.772
    // ms off, 1.769 / 1.782 / 1.934 ms with 2 / 4 / 8 groups -- the two kernels do not fill each other's idle issue slots, the
    // frame is work-conserving as it is. Kept as an A/B switch. Not with per-stage timers, the n3 ISM layouts or unshadowed gathering.
    const IsmMode im = ism_mode(ctx, p);
    const int ncls = ctx->cfg.class_hi - ctx->cfg.class_lo;
    const int parts = std::min(ctx->pp_parts, ncls);
    const bool pipelined = fork && parts > 1 && ctx->fg2 && !ctx->cfg.enable_timers && p->enable_shadowing && !im.scale && !im.points_only &&
                           im.vpl_start == 0 && im.vpl_count == ctx->cfg.vpl_count;
    if (!pipelined) {
        if ((rc = mlgi_ism_pullpush(ctx, nullptr, p, stream))) return rc;         //          pull / push
        if (fork) CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        if ((rc = mlgi_final_gather(ctx, gb, cam, nullptr, p, stream))) return rc;    // :405-408
    } else {
        if ((rc = check_params(ctx, p))) return rc;
        if (!ctx->have_atlas) return fail(ctx, MLGI_E_ORDER, "pull-push needs mlgi_ism_splat first");
        ctx->have_ism = true;
        GatherArgs a;
        if ((rc = build_gather_args(ctx, gb, cam, nullptr, p, &a))) return rc;
        const int S = ctx->S, L = ctx->L, band_hi = (ctx->ism_hi + S - 1) / S;
        int done_row = ctx->ism_lo / S;
        CK(cudaStreamWaitEvent(ctx->fg2, ctx->ev_join, 0));                       // the lists
        for (int j = 0; j < parts; j++) {
            const int c0 = ctx->cfg.class_lo + j * ncls / parts, c1 = ctx->cfg.class_lo + (j + 1) * ncls / parts;
            int r1 = j == parts - 1 ? band_hi : std::min(band_hi, (c1 * L + S - 1) / S + 1);       // the group's ISM rows + the halo row
            if (r1 > done_row) {
                ctx->launches += launch_pullpush(ctx->atlas, S, p->z_far, im.px, ctx->ism, done_row * S, std::min(ctx->cfg.vpl_count, r1 * S), st);
                done_row = r1;
            }
            CK(cudaEventRecord(ctx->ev_up[j], st));
            CK(cudaStreamWaitEvent(ctx->fg2, ctx->ev_up[j], 0));
            a.class_lo = c0; a.class_hi = c1;
            ctx->launches += launch_final_gather(a, ctx->fg2);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev_ism, ctx->fg2));
        CK(cudaStreamWaitEvent(st, ctx->ev_ism, 0));
        ctx->have_gi = true; ctx->have_blur = false;
    }
    if (with_exchange) rc = mlgi_exchange_finish(ctx, stream);
    return rc;
}

// mlgi_frame / mlgi_frame_sharded replay a CUDA graph of frame_body: one cudaGraphLaunch instead of 8 launches, 2 memsets
// and 6 event operations per frame. The graph is captured on an internal stream (the caller's may be the legacy stream,
// which cannot be captured) and keyed on everything the launches bake in -- pointers, matrices, parameters; when only
// values change (an animated light moves every frame) the re-captured graph updates the executable in place
// (cudaGraphExecUpdate), a structural change re-instantiates it. Per-stage timers need the direct path.
struct FrameKey {
    mlgi_rsm rsm; mlgi_camera light; mlgi_mesh mesh; mlgi_gbuffer gb; mlgi_camera cam; mlgi_params p;
    int with_exchange, push, fg_sorted, use_texture, tex_float_coord, fg_bulk, xch_cull, pp_parts, cull_late; const void* planar_ext;
};

// capture (if the key changed) and instantiate / update the frame's graph; no launch
static int ensure_graph(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                        const mlgi_camera* cam, const mlgi_params* p, bool with_exchange) {
    int rc = MLGI_OK;
    FrameKey key;
    memset(&key, 0, sizeof(key));
    key.rsm = *rsm; key.light = *light; key.mesh = *mesh; key.gb = *gb; key.cam = *cam; key.p = *p;
    key.with_exchange = with_exchange; key.push = ctx->xch.push; key.fg_sorted = ctx->fg_sorted; key.use_texture = ctx->use_texture;
    key.tex_float_coord = ctx->tex_float_coord; key.fg_bulk = ctx->fg_bulk; key.xch_cull = ctx->xch_cull; key.pp_parts = ctx->pp_parts; key.cull_late = ctx->cull_late; key.planar_ext = ctx->planar_ext;
    if (ctx->graph_valid && memcmp(&key, ctx->graph_key.data(), sizeof(key)) == 0) return MLGI_OK;
    if (p->cull_mode == MLGI_CULL_CLUSTERED && (rc = ensure_clustered(ctx))) return rc;   // cudaMalloc must not run inside a capture
    const unsigned long long l0 = ctx->launches;
    CK(cudaStreamBeginCapture(ctx->cap, cudaStreamCaptureModeThreadLocal));
    rc = frame_body(ctx, rsm, light, mesh, gb, cam, p, with_exchange, ctx->cap);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->cap, &g);
    ctx->graph_launches = ctx->launches - l0;
    ctx->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); cudaGetLastError(); ctx->graph_valid = false; return rc; }
    if (ce != cudaSuccess) { ctx->graph_valid = false; return fail(ctx, MLGI_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce)); }
    bool updated = false;
    // value-only changes (matrices, RSM pointers of an animated light ...) update the executable in place; a change of
    // the launch structure (culling mode, exchange, empty mesh) needs a new one
    const int topo = p->cull_mode | (with_exchange ? 8 : 0) | (mesh->num_triangles == 0 ? 16 : 0) | (ctx->xch.push ? 32 : 0) | (ctx->planar_ext ? 64 : 0);
    if (ctx->graph_exec && topo != ctx->graph_topo) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    ctx->graph_topo = topo;
    if (ctx->graph_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(ctx->graph_exec, g, &info) == cudaSuccess) updated = true;
        else { cudaGetLastError(); cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    }
    if (!updated) {
        const cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
        if (ie != cudaSuccess) { cudaGraphDestroy(g); ctx->graph_exec = nullptr; ctx->graph_valid = false; return fail(ctx, MLGI_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
        ctx->graph_instantiations++;
    }
    cudaGraphDestroy(g);
    ctx->graph_key.assign((const char*)&key, (const char*)&key + sizeof(key));
    ctx->graph_valid = true;
    return MLGI_OK;
}

static int frame_graphed(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                         const mlgi_camera* cam, const mlgi_params* p, bool with_exchange, void* stream) {
    if (!rsm || !light || !mesh || !gb || !cam || !p) return fail(ctx, MLGI_E_INVALID, "mlgi_frame: NULL argument");
    if (with_exchange && !ctx->xch.connected) return fail(ctx, MLGI_E_ORDER, "mlgi_frame_sharded needs a connected exchange");
    cudaStream_t st = (cudaStream_t)stream;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = MLGI_OK;
    if (!ctx->use_graph || ctx->cfg.enable_timers || !ctx->cap) {
        rc = frame_body(ctx, rsm, light, mesh, gb, cam, p, with_exchange, st);
    } else {
        if ((rc = ensure_graph(ctx, rsm, light, mesh, gb, cam, p, with_exchange))) return rc;
        CK(cudaGraphLaunch(ctx->graph_exec, st));
        ctx->launches += ctx->graph_launches;
        ctx->graph_replays++;
        // host-side stage bookkeeping of a replayed frame (what the stage calls set when they are issued directly)
        ctx->have_vpl = ctx->have_points = ctx->have_atlas = ctx->have_ism = ctx->have_gi = true;
        ctx->mode_vpl = ctx->mode_points = ism_mode(ctx, p);
        ctx->grouped_valid = false; ctx->have_blur = false; ctx->atlas_precleared = false;
        if (p->cull_mode == MLGI_CULL_CLUSTERED) ctx->have_clustered = true;
        if (p->cull_mode == MLGI_CULL_TILED) { ctx->have_tiled = true; ctx->tile_order_valid = ctx->fg_sorted != 0; }
    }
    ctx->last_enqueue_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// Builds (captures, instantiates, uploads) the graph the next mlgi_frame / mlgi_frame_sharded call with the same arguments
// will replay, without launching anything: takes the first-frame cost (capture + instantiation, ~1 ms of host time) out of the
// first frame, and -- for several ranks driven by one process on ONE GPU -- keeps a rank's instantiation from queueing behind
// another rank's kernel that is already waiting for it.
MLGI_API int mlgi_frame_prepare(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                                const mlgi_camera* cam, const mlgi_params* p, int sharded, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!rsm || !light || !mesh || !gb || !cam || !p) return fail(ctx, MLGI_E_INVALID, "mlgi_frame_prepare: NULL argument");
    if (sharded && !ctx->xch.connected) return fail(ctx, MLGI_E_ORDER, "mlgi_frame_prepare(sharded) needs a connected exchange");
    if (!ctx->use_graph || ctx->cfg.enable_timers || !ctx->cap) return MLGI_OK;       // direct launches: nothing to build
    int rc = ensure_graph(ctx, rsm, light, mesh, gb, cam, p, sharded != 0);
    if (rc) return rc;
    CK(cudaGraphUpload(ctx->graph_exec, (cudaStream_t)stream));
    return MLGI_OK;
}

MLGI_API int mlgi_frame(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                        const mlgi_camera* cam, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    return frame_graphed(ctx, rsm, light, mesh, gb, cam, p, false, stream);
}

// the frame of one rank of a class-sharded job = mlgi_frame + the exchange, as one graph launch
MLGI_API int mlgi_frame_sharded(mlgi_ctx* ctx, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* gb,
                                const mlgi_camera* cam, const mlgi_params* p, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    return frame_graphed(ctx, rsm, light, mesh, gb, cam, p, true, stream);
}

MLGI_API int mlgi_set_graph(mlgi_ctx* ctx, int enable) {
    if (!ctx) return MLGI_E_INVALID;
    ctx->use_graph = enable ? 1 : 0;
    return MLGI_OK;
}

MLGI_API int mlgi_get_frame_stats(mlgi_ctx* ctx, mlgi_frame_stats* out) {
    if (!ctx || !out) return MLGI_E_INVALID;
    out->host_enqueue_us = ctx->last_enqueue_us;
    out->graph_replays = ctx->graph_replays;
    out->graph_instantiations = ctx->graph_instantiations;
    out->graph_enabled = ctx->use_graph && !ctx->cfg.enable_timers;
    return MLGI_OK;
}

static const char* const kTrUp[16] = {"up0","up1","up2","up3","up4","up5","up6","up7","up8","up9","up10","up11","up12","up13","up14","up15"};
static const char* const kTrCull[16] = {"cull0","cull1","cull2","cull3","cull4","cull5","cull6","cull7","cull8","cull9","cull10","cull11","cull12","cull13","cull14","cull15"};
static const char* const kTrFg[16] = {"fg0","fg1","fg2","fg3","fg4","fg5","fg6","fg7","fg8","fg9","fg10","fg11","fg12","fg13","fg14","fg15"};
static const char* const kTrRd[16] = {"rd0","rd1","rd2","rd3","rd4","rd5","rd6","rd7","rd8","rd9","rd10","rd11","rd12","rd13","rd14","rd15"};

// The host-input frames as a CUDA graph. Each of them issues 40-80 API calls per frame (strided uploads, stage launches, event
// edges between five streams, read-backs); with several GPUs the device work per frame drops below a millisecond and those
// calls -- and the jitter of eight host processes issuing them -- become the frame time. A frame whose arguments (host pointers
// included) equal the previous frame's is captured once and replayed with one launch; arguments that change every frame (an
// animated light) keep the direct path, which needs no capture. Only pinned host memory is captured (pageable copies are
// staged synchronously by the runtime); per-stage timers and the timeline trace use the direct path.
struct HostFrameKey {
    mlgi_rsm hr; mlgi_camera light; mlgi_mesh mesh; mlgi_gbuffer hg; mlgi_camera cam; mlgi_params p;
    const void* host_out; int kind, packed, fg_sorted, use_texture, tex_float_coord, fg_bulk, host_slabs, xch_direct;
};
static bool is_pinned_host(const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return pa.type == cudaMemoryTypeHost;
}
static void mark_host_frame_done(mlgi_ctx* ctx, const mlgi_params* p) {
    ctx->have_vpl = ctx->have_points = ctx->have_atlas = ctx->have_ism = ctx->have_gi = true;
    ctx->mode_vpl = ctx->mode_points = ism_mode(ctx, p);
    ctx->grouped_valid = false; ctx->have_blur = false; ctx->atlas_precleared = false;
    if (p->cull_mode == MLGI_CULL_CLUSTERED) ctx->have_clustered = true;
    if (p->cull_mode == MLGI_CULL_TILED) { ctx->have_tiled = true; ctx->tile_order_valid = false; }
}
// body(stream) issues the frame without its final synchronise. Returns < 0 on error, 0 if the caller has to run the direct path,
// 1 if the frame was launched as a graph.
static int host_frame_graph(mlgi_ctx* ctx, const HostFrameKey& key, const mlgi_params* p, cudaStream_t st, const std::function<int(cudaStream_t)>& body) {
    const bool want = ctx->host_graph < 0 ? key.kind == 1 : ctx->host_graph != 0;
    if (!ctx->use_graph || !want || ctx->cfg.enable_timers || ctx->trace_on || !ctx->cap || key.xch_direct) return 0;
    const char* kb = (const char*)&key;
    const bool hit = ctx->hgraph_valid && memcmp(kb, ctx->hgraph_key.data(), sizeof(key)) == 0;
    if (!hit) {
        const bool seen = ctx->hgraph_seen.size() == sizeof(key) && memcmp(kb, ctx->hgraph_seen.data(), sizeof(key)) == 0;
        ctx->hgraph_seen.assign(kb, kb + sizeof(key));
        if (!seen) return 0;                         // first frame with these arguments: direct; capture when they repeat
        if (!is_pinned_host(key.hg.depth) || !is_pinned_host(key.hg.face_normal) || !is_pinned_host(key.host_out)) return 0;   // (the RSM is read by the CPU)
        if (p->cull_mode == MLGI_CULL_CLUSTERED) { int rc = ensure_clustered(ctx); if (rc) return rc; }
        const unsigned long long l0 = ctx->launches;
        CK(cudaStreamBeginCapture(ctx->cap, cudaStreamCaptureModeThreadLocal));
        const int rc = body(ctx->cap);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ctx->cap, &g);
        ctx->hgraph_launches = ctx->launches - l0;
        ctx->launches = l0;
        ctx->hgraph_valid = false;
        if (rc) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
        if (ce != cudaSuccess) { cudaGetLastError(); return fail(ctx, MLGI_E_CUDA, "host frame: cudaStreamEndCapture: %s", cudaGetErrorString(ce)); }
        bool updated = false;
        if (ctx->hgraph_exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(ctx->hgraph_exec, g, &info) == cudaSuccess) updated = true;
            else { cudaGetLastError(); cudaGraphExecDestroy(ctx->hgraph_exec); ctx->hgraph_exec = nullptr; }
        }
        if (!updated) {
            const cudaError_t ie = cudaGraphInstantiate(&ctx->hgraph_exec, g, 0);
            if (ie != cudaSuccess) { cudaGraphDestroy(g); ctx->hgraph_exec = nullptr; return fail(ctx, MLGI_E_CUDA, "host frame: cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
            ctx->graph_instantiations++;
        }
        cudaGraphDestroy(g);
        ctx->hgraph_key.assign(kb, kb + sizeof(key));
        ctx->hgraph_valid = true;
    }
    CK(cudaGraphLaunch(ctx->hgraph_exec, st));
    ctx->launches += ctx->hgraph_launches;
    ctx->hgraph_replays++; ctx->graph_replays++;
    mark_host_frame_done(ctx, p);
    return 1;
}

static int frame_host_body(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                           const mlgi_camera* cam, const mlgi_params* p, void* host_out, bool packed, cudaStream_t st);

static int frame_host_impl(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                           const mlgi_camera* cam, const mlgi_params* p, void* host_out, bool packed, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!hr || !hg || !light || !mesh || !cam || !p || !hr->diffuse || !hr->face_normal || !hr->depth || !hg->depth || !hg->face_normal) return fail(ctx, MLGI_E_INVALID, "host rsm/gbuffer is NULL");
    if (packed && !ctx->gi_packed) return fail(ctx, MLGI_E_INVALID, "mlgi_frame_host_packed needs cfg.gi_packed");
    cudaStream_t st = (cudaStream_t)stream;
    HostFrameKey key;
    memset(&key, 0, sizeof(key));
    key.hr = *hr; key.light = *light; key.mesh = *mesh; key.hg = *hg; key.cam = *cam; key.p = *p; key.host_out = host_out; key.kind = 0; key.packed = packed;
    key.fg_sorted = ctx->fg_sorted; key.use_texture = ctx->use_texture; key.tex_float_coord = ctx->tex_float_coord; key.fg_bulk = ctx->fg_bulk;
    key.host_slabs = ctx->host_slabs;
    if (is_device_pointer(hr->diffuse) || is_device_pointer(hr->face_normal) || is_device_pointer(hr->depth))
        return fail(ctx, MLGI_E_INVALID, "mlgi_frame_host: the RSM planes must be HOST pointers (the VPL samples are gathered by the CPU)");
    gather_rsm_samples(ctx, hr);
    int rc = host_frame_graph(ctx, key, p, st, [&](cudaStream_t cs) { return frame_host_body(ctx, hr, light, mesh, hg, cam, p, host_out, packed, cs); });
    if (rc < 0) return rc;
    if (rc == 0 && (rc = frame_host_body(ctx, hr, light, mesh, hg, cam, p, host_out, packed, st))) return rc;
    CK(cudaStreamSynchronize(st));
    trace_dump(ctx);
    return MLGI_OK;
}

static int frame_host_body(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                           const mlgi_camera* cam, const mlgi_params* p, void* host_out, bool packed, cudaStream_t st) {
    void* stream = (void*)st;
    const size_t WH = (size_t)ctx->cfg.width * ctx->cfg.height, NV = (size_t)ctx->cfg.vpl_count;
    // Three streams start together (after whatever the caller queued on `st` before): the side stream carries every upload
    // (RSM first, 3 MB, then the camera G-buffer, 8 B/pixel); the aux stream clears the atlas, builds the VPLs as soon as the RSM
    // has arrived and later culls the G-buffer slabs; the caller's stream generates the points (which need neither) and then
    // runs splat, pull-push and the gathering.
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
    trace_mark(ctx, "start", st);
    {
        StageTimer t(ctx, T_H2D, ctx->side);
        // the N sampled RSM texels, gathered on the host by the caller of this body (gather_rsm_samples)
        CK(cudaMemcpyAsync(ctx->st_rsm_samples, ctx->h_rsm_samples, (size_t)3 * NV * 4, cudaMemcpyHostToDevice, ctx->side));
    }
    CK(cudaEventRecord(ctx->ev_rsm, ctx->side));
    mlgi_rsm dr = {ctx->st_rsm_samples, ctx->st_rsm_samples + NV, reinterpret_cast<const float*>(ctx->st_rsm_samples + 2 * NV)};
    mlgi_gbuffer dg = {ctx->st_depth, ctx->st_normal};
    int rc;
    launch_atlas_clear(ctx->atlas, ctx->S, ctx->ism_lo, ctx->ism_hi, ctx->aux);
    CK(cudaEventRecord(ctx->ev_clear, ctx->aux));
    CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_rsm, 0));
    if ((rc = vpl_process_impl(ctx, &dr, light, p, (void*)ctx->aux, 1))) return rc;
    CK(cudaEventRecord(ctx->ev_vpl, ctx->aux));
    trace_mark(ctx, "vpl", ctx->aux);
    // points -> [VPL records, cleared atlas] -> splat -> pull-push on the caller's stream
    auto ism_chain = [&]() -> int {
        int r;
        if ((r = mlgi_ism_points(ctx, mesh, p, stream))) return r;
        CK(cudaStreamWaitEvent(st, ctx->ev_vpl, 0));
        CK(cudaStreamWaitEvent(st, ctx->ev_clear, 0));
        ctx->atlas_precleared = true;
        if ((r = mlgi_ism_splat(ctx, p, stream))) return r;
        return mlgi_ism_pullpush(ctx, nullptr, p, stream);
    };
    const bool whole_image = !ctx->xch.connected && !ctx->planar_ext;
    if (p->cull_mode == MLGI_CULL_TILED && whole_image && ctx->aux && ctx->fg2 && ctx->rd && ctx->slab_plan.size() > 1) {
        // Slab pipeline: the image is cut into K interleaved slabs of whole tile rows -- slab k = tile rows k, k + K, k + 2K, ... --
        // so that every slab is a representative sample of the image: gathering cost per tile varies by an order of magnitude
        // over the image (sky vs. floor), and with contiguous slabs the cheap slabs are gathered faster than the next one
        // uploads while the expensive ones pile up at the end (measured: GPU idle early, 0.9 of 1.9 ms in the last two of six
        // slabs). Slab k is uploaded on the side stream (one strided 2-D copy per G-buffer plane), culled on the aux stream as
        // soon as it has arrived, gathered on the caller's stream (after the ISM build) and read back on its own stream while
        // the next slabs are gathered, so PCIe traffic in both directions hides behind the ISM build and the gathering.
        const int K = (int)ctx->slab_plan.size(), W = ctx->cfg.width, H = ctx->cfg.height, TY = ctx->TY;
        auto slab_off = [&](int k) { return ctx->slab_plan[k].off; };
        auto slab_stride = [&](int k) { return ctx->slab_plan[k].stride; };
        auto slab_rows = [&](int k) { return slab_off(k) < TY ? (TY - slab_off(k) + slab_stride(k) - 1) / slab_stride(k) : 0; };
        float vi[16], pi[16], inv[16];
        mat_inverse(cam->view, vi); mat_inverse(cam->proj, pi); mat_mul(vi, pi, inv);
        const size_t row_px = (size_t)kTile * W;                       // pixels of one full tile row
        const int last_rows = H - (TY - 1) * kTile;                    // image rows of the last tile row (1..64)
        // strided copy of slab k between two images with 4-byte pixels (tile rows k, k + K, ...; the last tile row may be partial)
        auto copy_slab = [&](void* dst, const void* src, int k, cudaMemcpyKind kind, cudaStream_t cs) -> int {
            const int nrows = slab_rows(k);
            if (nrows <= 0) return MLGI_OK;
            const bool has_last = (TY - 1 - slab_off(k)) % slab_stride(k) == 0;
            const int full_rows = (has_last && last_rows < kTile) ? nrows - 1 : nrows;
            const size_t off = (size_t)slab_off(k) * row_px * 4, pitch = (size_t)slab_stride(k) * row_px * 4;
            if (full_rows > 0) CK(cudaMemcpy2DAsync((char*)dst + off, pitch, (const char*)src + off, pitch, row_px * 4, full_rows, kind, cs));
            if (full_rows < nrows) {
                const size_t o2 = (size_t)(TY - 1) * row_px * 4;
                CK(cudaMemcpyAsync((char*)dst + o2, (const char*)src + o2, (size_t)last_rows * W * 4, kind, cs));
            }
            return MLGI_OK;
        };
        // Enqueue order matters: the host issues ~10 calls per slab, so the first slab's upload and the ISM build are
        // queued before the remaining slabs -- both the copy engine and the SMs have work while the host is still enqueueing.
        // The uploads need no ordering against earlier work: this call returns only after every reader of the staging
        // buffers has finished. Culling follows the VPL build on the aux stream and waits for its slab (ev_up).
        auto upload = [&](int k) -> int {
            int r;
            if ((r = copy_slab(ctx->st_depth, hg->depth, k, cudaMemcpyHostToDevice, ctx->side))) return r;
            if ((r = copy_slab(ctx->st_normal, hg->face_normal, k, cudaMemcpyHostToDevice, ctx->side))) return r;
            CK(cudaEventRecord(ctx->ev_up[k], ctx->side));
            trace_mark(ctx, kTrUp[k], ctx->side);
            return MLGI_OK;
        };
        if ((rc = upload(0))) return rc;
        if ((rc = ism_chain())) return rc;
        trace_mark(ctx, "ism", st);
        int slab_tile0 = 0;
        for (int k = 0; k < K; k++) {
            if (k > 0 && (rc = upload(k))) return rc;
            CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_up[k], 0));
            ctx->launches += cull_tiled_rows(ctx, ctx->st_depth, ctx->st_normal, to_mat4(inv), slab_off(k), slab_stride(k), slab_rows(k), ctx->aux);
            if (ctx->fg_sorted) {       // the slab's tiles, longest list first (a slab launch has a tail like the whole-frame launch)
                ctx->launches += launch_tile_order_ids(ctx->tile_weight, ctx->slab_order, slab_tile0, slab_rows(k) * ctx->TX, ctx->cfg.vpl_count,
                                                       ctx->cfg.class_lo, ctx->cfg.class_hi, ctx->tile_order, ctx->aux);
                slab_tile0 += slab_rows(k) * ctx->TX;
            }
            CK(cudaEventRecord(ctx->ev_cull[k], ctx->aux));
            trace_mark(ctx, kTrCull[k], ctx->aux);
        }
        ctx->have_tiled = true; ctx->tile_order_valid = false;
        GatherArgs a;
        if ((rc = build_gather_args(ctx, &dg, cam, nullptr, p, &a))) return rc;
        a.tile_order = ctx->fg_sorted ? ctx->tile_order : ctx->slab_order;   // the tiles slab by slab, inside a slab longest list first
        a.max_ctas_per_sm = ctx->fg_host_occ;
        {
            // consecutive slabs alternate between the caller's stream and fg2: the gathering kernel's tail (a few long-list
            // tiles keeping a handful of SMs busy) then overlaps the next slab instead of idling the GPU once per slab
            StageTimer t(ctx, T_FG, st);
            CK(cudaEventRecord(ctx->ev_ism, st));
            CK(cudaStreamWaitEvent(ctx->fg2, ctx->ev_ism, 0));
            int tile0 = 0;
            for (int k = 0; k < K; k++) {
                cudaStream_t fs = (k & 1) ? ctx->fg2 : st;
                CK(cudaStreamWaitEvent(fs, ctx->ev_cull[k], 0));
                a.tile_begin = tile0; a.tile_count = slab_rows(k) * ctx->TX;
                tile0 += a.tile_count;
                ctx->launches += launch_final_gather(a, fs);
                CK(cudaEventRecord(ctx->ev_fg[k], fs));
                trace_mark(ctx, kTrFg[k], fs);
                CK(cudaStreamWaitEvent(ctx->rd, ctx->ev_fg[k], 0));
                if (host_out) {
                    if (packed) { if ((rc = copy_slab(host_out, ctx->gi_packed, k, cudaMemcpyDeviceToHost, ctx->rd))) return rc; }
                    else {       // fp32 RGBA: 16-byte pixels = four 4-byte "pixels" per pixel, the same strides scaled by 4
                        const int nrows = slab_rows(k);
                        const bool has_last = nrows > 0 && (TY - 1 - slab_off(k)) % slab_stride(k) == 0;
                        const int full_rows = (has_last && last_rows < kTile) ? nrows - 1 : nrows;
                        const size_t off = (size_t)slab_off(k) * row_px * 16, pitch = (size_t)slab_stride(k) * row_px * 16;
                        if (full_rows > 0) CK(cudaMemcpy2DAsync((char*)host_out + off, pitch, (const char*)ctx->gi + off, pitch, row_px * 16, full_rows, cudaMemcpyDeviceToHost, ctx->rd));
                        if (full_rows < nrows) {
                            const size_t o2 = (size_t)(TY - 1) * row_px * 16;
                            CK(cudaMemcpyAsync((char*)host_out + o2, (const char*)ctx->gi + o2, (size_t)last_rows * W * 16, cudaMemcpyDeviceToHost, ctx->rd));
                        }
                    }
                    trace_mark(ctx, kTrRd[k], ctx->rd);
                }
            }
            CK(cudaEventRecord(ctx->ev_join, ctx->rd));
            CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        }
        CK(cudaGetLastError());
        ctx->have_gi = true; ctx->have_blur = false;
        trace_mark(ctx, "end", st);
        return MLGI_OK;
    }
    CK(cudaMemcpyAsync(ctx->st_depth, hg->depth, WH * 4, cudaMemcpyHostToDevice, ctx->side));
    CK(cudaMemcpyAsync(ctx->st_normal, hg->face_normal, WH * 4, cudaMemcpyHostToDevice, ctx->side));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_vpl, 0));         // culling needs the VPL records
    void* cs = (void*)ctx->side;
    if (p->cull_mode == MLGI_CULL_CLUSTERED) { if ((rc = mlgi_cull_clustered(ctx, &dg, cam, p, cs))) return rc; }
    else if (p->cull_mode == MLGI_CULL_TILED) { if ((rc = mlgi_cull_tiled(ctx, &dg, cam, p, cs))) return rc; }
    CK(cudaEventRecord(ctx->ev_join, ctx->side));
    if ((rc = ism_chain())) return rc;
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    if ((rc = mlgi_final_gather(ctx, &dg, cam, nullptr, p, stream))) return rc;
    if (host_out) {
        StageTimer t(ctx, T_D2H, st);
        if (packed) CK(cudaMemcpyAsync(host_out, ctx->gi_packed, WH * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        else CK(cudaMemcpyAsync(host_out, ctx->gi, WH * sizeof(float4), cudaMemcpyDeviceToHost, st));
    }
    return MLGI_OK;
}

// ------------------------------------------------------------------------------------- sharded host-input frame
// One rank of a class-sharded job with HOST inputs (the N-GPU form of mlgi_frame_host_packed). The job's inputs
// cross PCIe once in total and the camera G-buffer is never replicated: rank r uploads the (small) RSM and ITS rows of the
// G-buffer -- the tile rows t mod world == r (mlgi_sharded_rows), 1/world of the image -- into its exchange allocation; it
// bounds the tiles of those rows and publishes the bounds (64 B per tile) to every rank; the gathering CTAs of every rank then
// read a tile's pixels from the rank that holds them, over NVLink (a rank's classes touch 1/world of a tile's pixels, once);
// finally the rank assembles its rows of the image from every rank's class planes and copies them to the caller's buffer.
// All of this runs as a pipeline over K phases (blocks of tile rows, see below), so that the PCIe traffic in both directions
// hides behind the gathering. host_gbuffer / host_gi_packed are full-size host arrays; only this rank's rows are read /
// written. Synchronises the stream before returning. Every rank of the job must make this call for every frame (each phase
// contains two bounded waits on the peers: tile bounds, class planes).
MLGI_API int mlgi_sharded_rows(mlgi_ctx* ctx, int* row_begin, int* row_block, int* row_stride) {
    if (!ctx) return MLGI_E_INVALID;
    auto& x = ctx->xch;
    if (!x.base) return fail(ctx, MLGI_E_ORDER, "mlgi_sharded_rows needs mlgi_exchange_create first");
    if (row_begin) *row_begin = x.rank * kTile;
    if (row_block) *row_block = kTile;
    if (row_stride) *row_stride = x.world * kTile;
    return MLGI_OK;
}

static int frame_sharded_host_body(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                                   const mlgi_camera* cam, const mlgi_params* p, uint32_t* host_gi_packed, cudaStream_t st);

MLGI_API int mlgi_frame_sharded_host(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                                     const mlgi_camera* cam, const mlgi_params* p, uint32_t* host_gi_packed, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    auto& x = ctx->xch;
    if (!x.connected) return fail(ctx, MLGI_E_ORDER, "mlgi_frame_sharded_host needs a connected exchange");
    if (!hr || !hg || !cam || !light || !mesh || !p || !hr->diffuse || !hr->face_normal || !hr->depth || !hg->depth || !hg->face_normal) return fail(ctx, MLGI_E_INVALID, "host rsm/gbuffer/camera is NULL");
    if (!ctx->gi_packed) return fail(ctx, MLGI_E_INVALID, "mlgi_frame_sharded_host needs cfg.gi_packed");
    if (ctx->cfg.enable_timers) return fail(ctx, MLGI_E_INVALID, "per-stage timers are not available on the sharded host path");
    if (p->cull_mode != MLGI_CULL_TILED) return fail(ctx, MLGI_E_UNSUPPORTED, "mlgi_frame_sharded_host runs tiled culling (the G-buffer is distributed by tile rows)");
    if (!ctx->aux || !ctx->side || !ctx->rd) return fail(ctx, MLGI_E_CUDA, "internal streams missing");
    cudaStream_t st = (cudaStream_t)stream;
    HostFrameKey key;
    memset(&key, 0, sizeof(key));
    key.hr = *hr; key.light = *light; key.mesh = *mesh; key.hg = *hg; key.cam = *cam; key.p = *p; key.host_out = host_gi_packed; key.kind = 1; key.packed = 1;
    key.fg_sorted = ctx->fg_sorted; key.use_texture = ctx->use_texture; key.tex_float_coord = ctx->tex_float_coord; key.fg_bulk = ctx->fg_bulk;
    key.host_slabs = ctx->host_slabs_set ? ctx->host_slabs : 0; key.xch_direct = ctx->xch_direct;
    if (is_device_pointer(hr->diffuse) || is_device_pointer(hr->face_normal) || is_device_pointer(hr->depth))
        return fail(ctx, MLGI_E_INVALID, "mlgi_frame_sharded_host: the RSM planes must be HOST pointers (the VPL samples are gathered by the CPU)");
    gather_rsm_samples(ctx, hr);
    int rc = host_frame_graph(ctx, key, p, st, [&](cudaStream_t cs) { return frame_sharded_host_body(ctx, hr, light, mesh, hg, cam, p, host_gi_packed, cs); });
    if (rc < 0) return rc;
    if (rc == 0 && (rc = frame_sharded_host_body(ctx, hr, light, mesh, hg, cam, p, host_gi_packed, st))) return rc;
    CK(cudaStreamSynchronize(st));
    trace_dump(ctx);
    return MLGI_OK;
}

static int frame_sharded_host_body(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                                   const mlgi_camera* cam, const mlgi_params* p, uint32_t* host_gi_packed, cudaStream_t st) {
    auto& x = ctx->xch;
    void* stream = (void*)st;
    const int W = ctx->cfg.width, H = ctx->cfg.height, TX = ctx->TX, TY = ctx->TY, G = x.world, g = x.rank;
    const size_t WH = (size_t)W * H, NV = (size_t)ctx->cfg.vpl_count;
    const size_t gb_plane = (WH * 4 + 255) & ~(size_t)255;
    int rc;
    // Phases: the image is cut into K blocks of whole tile rows, each a multiple of `world` rows high, so every rank owns the
    // same share (rows t mod world == rank) of every phase. Phase k's rows are uploaded, bounded and published while phase
    // k - 1 is gathered and phase k - 2 is assembled and read back: PCIe in both directions hides behind the gathering, as in
    // the single-GPU slab pipeline. K is sized so that a phase's gathering launch still fills the GPU a few times over
    // (MLGI_HOST_SLABS overrides it).
    // Measured at 4K (profiles/r2_sweep_host_frame.txt, end to end in ms by phase count): 2 GPUs 1.65 / 1.59 / 1.61 (2 / 4 / 6 phases),
    // 4 GPUs 1.53 / 1.19 / 1.28 (1 / 2 / 3), 8 GPUs 1.06 / 0.90 / 0.89 (1 / 2 / 3; the uploads of eight ranks share the host's
    // memory bandwidth, so finer phases pay there). Small images (less than about one wave of gathering CTAs per phase) use one phase.
    const long long ctas = (long long)TX * TY * (ctx->cfg.class_hi - ctx->cfg.class_lo);
    int K = ctx->host_slabs_set ? ctx->host_slabs : ctas < 3000 ? 1 : G == 2 ? 4 : G == 4 ? 2 : 3;
    K = std::max(1, std::min(K, std::min((int)mlgi_ctx::kSlabs, (TY + G - 1) / G)));
    const int rows_per_phase = G * ((TY + K * G - 1) / (K * G));
    K = (TY + rows_per_phase - 1) / rows_per_phase;
    const size_t row_px = (size_t)kTile * W;                      // pixels of one full tile row
    const int last_rows = H - (TY - 1) * kTile;                   // image rows of the last tile row (1..64)
    // this rank's tile rows inside tile rows [a, b): first, count
    auto own = [&](int a, int b, int* first, int* count) { *first = a + g; *count = *first < b ? (b - *first + G - 1) / G : 0; };
    // copy of those rows between two images with 4-byte pixels (the last tile row of the image may be partial)
    auto copy_rows = [&](void* dst, const void* src, int first, int count, cudaMemcpyKind kind, cudaStream_t cs) -> int {
        if (count <= 0) return MLGI_OK;
        const bool has_last = first + (count - 1) * G == TY - 1 && last_rows < kTile;
        const int full = has_last ? count - 1 : count;
        const size_t off = (size_t)first * row_px * 4, pitch = (size_t)G * row_px * 4;
        if (full > 0) CK(cudaMemcpy2DAsync((char*)dst + off, pitch, (const char*)src + off, pitch, row_px * 4, full, kind, cs));
        if (has_last) {
            const size_t o2 = (size_t)(TY - 1) * row_px * 4;
            CK(cudaMemcpyAsync((char*)dst + o2, (const char*)src + o2, (size_t)last_rows * W * 4, kind, cs));
        }
        return MLGI_OK;
    };
    // side: the RSM (3 MB; the ISM chain it feeds then runs under the G-buffer uploads), then this rank's rows phase by phase;
    // aux: atlas clear, VPL build, then per phase bounds -> publish -> wait -> lists -> order;
    // caller's stream: points, splat, pull-push, then per phase gathering + "planes written" signal;
    // rd: per phase wait for every rank's planes, assemble this rank's rows, copy them to the host.
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
    CK(cudaStreamWaitEvent(ctx->rd, ctx->ev_fork, 0));
    trace_mark(ctx, "start", st);
    float* own_depth = (float*)(x.base + x.gb_off);
    uint32_t* own_normal = (uint32_t*)(x.base + x.gb_off + gb_plane);
    CK(cudaMemcpyAsync(ctx->st_rsm_samples, ctx->h_rsm_samples, (size_t)3 * NV * 4, cudaMemcpyHostToDevice, ctx->side));   // the sampled RSM texels (gather_rsm_samples)
    CK(cudaEventRecord(ctx->ev_rsm, ctx->side));
    auto upload = [&](int k) -> int {
        int first, count, r;
        own(k * rows_per_phase, std::min(TY, (k + 1) * rows_per_phase), &first, &count);
        if ((r = copy_rows(own_depth, hg->depth, first, count, cudaMemcpyHostToDevice, ctx->side))) return r;
        if ((r = copy_rows(own_normal, hg->face_normal, first, count, cudaMemcpyHostToDevice, ctx->side))) return r;
        CK(cudaEventRecord(ctx->ev_up[k], ctx->side));
        trace_mark(ctx, kTrUp[k], ctx->side);
        return MLGI_OK;
    };
    if ((rc = upload(0))) return rc;
    mlgi_rsm dr = {ctx->st_rsm_samples, ctx->st_rsm_samples + NV, reinterpret_cast<const float*>(ctx->st_rsm_samples + 2 * NV)};
    mlgi_gbuffer dg = {own_depth, own_normal};
    launch_atlas_clear(ctx->atlas, ctx->S, ctx->ism_lo, ctx->ism_hi, ctx->aux);
    CK(cudaEventRecord(ctx->ev_clear, ctx->aux));
    CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_rsm, 0));
    if ((rc = vpl_process_impl(ctx, &dr, light, p, (void*)ctx->aux, 1))) return rc;
    CK(cudaEventRecord(ctx->ev_vpl, ctx->aux));
    trace_mark(ctx, "vpl", ctx->aux);
    if ((rc = mlgi_ism_points(ctx, mesh, p, stream))) return rc;
    CK(cudaStreamWaitEvent(st, ctx->ev_vpl, 0));
    CK(cudaStreamWaitEvent(st, ctx->ev_clear, 0));
    ctx->atlas_precleared = true;
    if ((rc = mlgi_ism_splat(ctx, p, stream))) return rc;
    if ((rc = mlgi_ism_pullpush(ctx, nullptr, p, stream))) return rc;
    trace_mark(ctx, "ism", st);
    // culling, phase by phase (aux): own bounds as soon as the rows are here, lists once every rank has published the phase
    float vi[16], pi[16], inv[16];
    mat_inverse(cam->view, vi); mat_inverse(cam->proj, pi); mat_mul(vi, pi, inv);
    BoundsOut bo;
    memset(&bo, 0, sizeof(bo));
    bo.local = (float*)(x.base + x.bounds_off);
    for (int q = 0; q < G; q++) if (q != g) bo.peers[bo.num_peers++] = (float*)(x.peer_base[q] + x.bounds_off);
    bo.parity_stride = x.bounds_bytes / sizeof(float);
    bo.frame_ctr = x.d_state;
    for (int k = 0; k < K; k++) {
        const int a = k * rows_per_phase, b = std::min(TY, (k + 1) * rows_per_phase);
        int first, count;
        own(a, b, &first, &count);
        if (k > 0 && (rc = upload(k))) return rc;
        CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_up[k], 0));
        ctx->launches += launch_tile_bounds(own_depth, own_normal, W, H, to_mat4(inv), first, G, count, bo, ctx->aux);
        ctx->launches += launch_flag_signal(x.d_peer_flags, G, 32 + g, x.d_state, 16u, (uint32_t)k, ctx->aux);
        ctx->launches += launch_flag_wait((const uint32_t*)x.base + 32, G, x.d_state, x.d_state + 2, 16u, (uint32_t)k, ctx->aux);
        ctx->launches += launch_tile_lists(bo.local, bo.parity_stride, x.d_state, W, ctx->records, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi,
                                           a, 1, b - a, ctx->tile_lists, ctx->tile_weight, ctx->aux);
        if (ctx->fg_sorted) ctx->launches += launch_tile_order(ctx->tile_weight, W, ctx->cfg.vpl_count, ctx->cfg.class_lo, ctx->cfg.class_hi, a, b - a, ctx->tile_order, ctx->aux);
        CK(cudaEventRecord(ctx->ev_cull[k], ctx->aux));
        trace_mark(ctx, kTrCull[k], ctx->aux);
    }
    CK(cudaGetLastError());
    ctx->have_tiled = true; ctx->tile_order_valid = ctx->fg_sorted != 0;
    // gathering, phase by phase (caller's stream); assembly + read-back of the finished phase on rd
    GatherArgs ga;
    if ((rc = build_gather_args(ctx, &dg, cam, nullptr, p, &ga))) return rc;
    ga.gb_world = G;
    for (int q = 0; q < G; q++) {
        ga.gb_depth[q] = (const float*)(x.peer_base[q] + x.gb_off);
        ga.gb_normal[q] = (const uint32_t*)(x.peer_base[q] + x.gb_off + gb_plane);
    }
    ExchangeArgs e;
    build_exchange_args(ctx, &e);
    e.raster = nullptr;
    e.signal_wait = 0;
    bool direct = false;
    // (measured at 2 GPUs, 4K: the exchange kernel storing 16.6 MB straight into mapped pinned memory takes 1.3 ms -- 13 GB/s of
    // SM stores over PCIe -- against 0.05 ms into HBM plus 0.3 ms for the copy engine; MLGI_XCH_DIRECT=1 keeps the direct form for A/B)
    if (host_gi_packed && ctx->xch_direct) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, host_gi_packed) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
            e.raster_packed = (uint32_t*)pa.devicePointer;
            direct = true;
        } else cudaGetLastError();
    }
    for (int k = 0; k < K; k++) {
        const int a = k * rows_per_phase, b = std::min(TY, (k + 1) * rows_per_phase);
        int first, count;
        own(a, b, &first, &count);
        CK(cudaStreamWaitEvent(st, ctx->ev_cull[k], 0));
        ga.tile_begin = a * TX; ga.tile_count = (b - a) * TX;
        ctx->launches += launch_final_gather(ga, st);
        ctx->launches += launch_flag_signal(x.d_peer_flags, G, 48 + g, x.d_state, 16u, (uint32_t)k, st);
        CK(cudaEventRecord(ctx->ev_fg[k], st));
        trace_mark(ctx, kTrFg[k], st);
        CK(cudaStreamWaitEvent(ctx->rd, ctx->ev_fg[k], 0));
        ctx->launches += launch_flag_wait((const uint32_t*)x.base + 48, G, x.d_state, x.d_state + 2, 16u, (uint32_t)k, ctx->rd);
        e.row_begin = first * kTile; e.row_block = kTile; e.row_stride = G * kTile; e.row_blocks = count;
        e.advance = k == K - 1;
        if (count > 0 || e.advance) ctx->launches += launch_exchange(e, ctx->rd);
        if (host_gi_packed && !direct && (rc = copy_rows(host_gi_packed, ctx->gi_packed, first, count, cudaMemcpyDeviceToHost, ctx->rd))) return rc;
        trace_mark(ctx, kTrRd[k], ctx->rd);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_join, ctx->rd));
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    ctx->have_gi = true; ctx->have_blur = false;
    trace_mark(ctx, "end", st);
    return MLGI_OK;
}

MLGI_API int mlgi_frame_host(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                             const mlgi_camera* cam, const mlgi_params* p, float* host_gi, void* stream) {
    return frame_host_impl(ctx, hr, light, mesh, hg, cam, p, host_gi, false, stream);
}

MLGI_API int mlgi_frame_host_packed(mlgi_ctx* ctx, const mlgi_rsm* hr, const mlgi_camera* light, const mlgi_mesh* mesh, const mlgi_gbuffer* hg,
                                    const mlgi_camera* cam, const mlgi_params* p, uint32_t* host_gi_packed, void* stream) {
    return frame_host_impl(ctx, hr, light, mesh, hg, cam, p, host_gi_packed, true, stream);
}

// ------------------------------------------------------------------------------------- inspection
MLGI_API int mlgi_get_buffer(mlgi_ctx* ctx, int id, void** dev_ptr, size_t* bytes) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    if (!dev_ptr || !bytes) return fail(ctx, MLGI_E_INVALID, "dev_ptr/bytes is NULL");
    const size_t N = ctx->cfg.vpl_count, WH = (size_t)ctx->cfg.width * ctx->cfg.height, AA = (size_t)ctx->A * ctx->A;
    const size_t maxClusters = (size_t)ctx->X * ctx->Y * 16, L1 = ctx->L + 1;
    const size_t PW = (ctx->cfg.width + 3) / 4, PH = (ctx->cfg.height + 3) / 4;
    switch (id) {
        case MLGI_BUF_VPL: *dev_ptr = ctx->vpl; *bytes = N * 48; break;
        case MLGI_BUF_PACKED_VPL: *dev_ptr = ctx->packed; *bytes = N * 16; break;
        case MLGI_BUF_SHUFFLE: *dev_ptr = ctx->shuffle; *bytes = N * 4; break;
        case MLGI_BUF_POINTS:
        case MLGI_BUF_BUCKET_START:
            // the bucket-grouped view (reference layout of pointBuffer) is not needed by the frame; build it on first request
            if (ctx->have_points && !ctx->grouped_valid) {
                CK(cudaDeviceSynchronize());
                ctx->launches += launch_group_points(ctx->points_unsorted, ctx->point_bucket, ctx->counters + 1, ctx->cfg.max_points, ctx->cfg.vpl_count,
                                                     ctx->bucket_hist, ctx->bucket_cursor, ctx->bucket_start, ctx->points, 0);
                CK(cudaDeviceSynchronize());
                ctx->grouped_valid = true;
            }
            if (id == MLGI_BUF_POINTS) { *dev_ptr = ctx->points; *bytes = (size_t)ctx->cfg.max_points * 16; }
            else { *dev_ptr = ctx->bucket_start; *bytes = (N + 1) * 4; }
            break;
        case MLGI_BUF_ATLAS: *dev_ptr = ctx->atlas; *bytes = AA * 4; break;
        case MLGI_BUF_ISM: *dev_ptr = ctx->ism; *bytes = AA * 2; break;
        case MLGI_BUF_COMPACT_CLUSTER_IDS: *dev_ptr = ctx->compact_ids; *bytes = maxClusters * 4; break;
        case MLGI_BUF_LIGHT_LIST_IDS: *dev_ptr = ctx->light_list_ids; *bytes = maxClusters * 2; break;
        case MLGI_BUF_LIGHT_LISTS:      // allocated on first use; a viewport that cannot use clustered culling has none
            if (ensure_clustered(ctx) != MLGI_OK) { *dev_ptr = nullptr; *bytes = 0; break; }
            *dev_ptr = ctx->light_lists; *bytes = maxClusters * 16 * L1 * 2; break;
        case MLGI_BUF_TILE_LISTS: *dev_ptr = ctx->tile_lists; *bytes = (size_t)ctx->TX * ctx->TY * 16 * L1 * 2; break;
        case MLGI_BUF_NUM_USED_CLUSTERS: *dev_ptr = ctx->counters + 2; *bytes = 4; break;
        case MLGI_BUF_GI: *dev_ptr = ctx->gi; *bytes = WH * 16; break;
        case MLGI_BUF_GI_PLANAR: *dev_ptr = ctx->gi_planar; *bytes = (size_t)(ctx->cfg.class_hi - ctx->cfg.class_lo) * PW * PH * 16; break;
        case MLGI_BUF_GI_BLUR: *dev_ptr = ctx->gi_blur; *bytes = WH * 16; break;
        case MLGI_BUF_GI_PACKED:
            if (!ctx->gi_packed) return fail(ctx, MLGI_E_INVALID, "MLGI_BUF_GI_PACKED needs cfg.gi_packed");
            *dev_ptr = ctx->gi_packed; *bytes = WH * 4; break;
        default: return fail(ctx, MLGI_E_INVALID, "unknown buffer id %d", id);
    }
    return MLGI_OK;
}

MLGI_API int mlgi_read_buffer(mlgi_ctx* ctx, int id, void* host_dst, size_t bytes, size_t offset) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    void* p = nullptr;
    size_t n = 0;
    int rc = mlgi_get_buffer(ctx, id, &p, &n);
    if (rc) return rc;
    if (!host_dst || bytes > n || offset > n - bytes) return fail(ctx, MLGI_E_INVALID, "read of %zu bytes at %zu exceeds buffer %d (%zu bytes)", bytes, offset, id, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_dst, (const char*)p + offset, bytes, cudaMemcpyDeviceToHost));
    return MLGI_OK;
}

MLGI_API int mlgi_read_buffer_async(mlgi_ctx* ctx, int id, void* host_dst, size_t bytes, size_t offset, void* stream) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    void* p = nullptr;
    size_t n = 0;
    int rc = mlgi_get_buffer(ctx, id, &p, &n);
    if (rc) return rc;
    if (!host_dst || bytes > n || offset > n - bytes) return fail(ctx, MLGI_E_INVALID, "read of %zu bytes at %zu exceeds buffer %d (%zu bytes)", bytes, offset, id, n);
    CK(cudaMemcpyAsync(host_dst, (const char*)p + offset, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MLGI_OK;
}

MLGI_API int mlgi_get_counts(mlgi_ctx* ctx, int* num_points, int* num_used_clusters) {
    if (!ctx) return MLGI_E_INVALID;
    DeviceGuard dg_(ctx);
    uint32_t h[4];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, ctx->counters, sizeof(h), cudaMemcpyDeviceToHost));
    if (num_points) *num_points = (int)std::min<uint32_t>(h[1], (uint32_t)ctx->cfg.max_points);
    if (num_used_clusters) *num_used_clusters = (int)h[2];
    if (h[1] > (uint32_t)ctx->cfg.max_points) return fail(ctx, MLGI_E_CAPACITY, "point buffer overflow: %u points generated, capacity %d", h[1], ctx->cfg.max_points);
    return MLGI_OK;
}

MLGI_API int mlgi_get_timings(mlgi_ctx* ctx, mlgi_timing* out, int cap) {
    if (!ctx || !out) return 0;
    if (!ctx->cfg.enable_timers) return 0;
    DeviceGuard dg_(ctx);
    cudaDeviceSynchronize();
    int n = 0;
    for (int i = 0; i < T_COUNT && n < cap; i++) {
        if (!ctx->ev_used[i]) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i][0], ctx->ev[i][1]) != cudaSuccess) { cudaGetLastError(); continue; }
        snprintf(out[n].name, sizeof(out[n].name), "%s", kTimerNames[i]);
        out[n].ms = ms;
        out[n].launches = ctx->ev_launches[i];
        n++;
    }
    return n;
}

MLGI_API unsigned long long mlgi_launch_count(const mlgi_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

}  // extern "C"

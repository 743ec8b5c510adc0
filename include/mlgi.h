/* mlgi.h -- C ABI of the B200-native GI frame pipeline (libmlgi.so).
 *
 * Drop-in boundary for the hot path of karyon/many-lights-gi: the stage objects that
 * MultiFramePainter wires together (source/mfs-painters/multiframepainter/):
 *     VPLProcessor::process          VPLProcessor.h:22        -> mlgi_vpl_process
 *     ImperfectShadowmap::process    ImperfectShadowmap.h:30  -> mlgi_ism_points + mlgi_ism_splat + mlgi_ism_pullpush
 *     ClusteredShading::process      ClusteredShading.h:30    -> mlgi_cull_clustered   (tiled variant: mlgi_cull_tiled)
 *     GIStage::compute_final_gathering GIStage.cpp:261        -> mlgi_final_gather
 *     GIStage::blur                  GIStage.cpp:308          -> mlgi_gi_blur
 *     GIStage::process               GIStage.cpp:346          -> mlgi_frame / mlgi_frame_host
 * The reference has no FFI of its own (it is a gloperate plugin, plugin.cpp:9-17); these
 * entry points are what a maintainer would bind from the C++ stage classes, see
 * INTEGRATION.md and many-lights-gi_b200/host/mlgi_stages.hpp.
 *
 * Conventions
 *  - plain C, pointers + sizes only. All matrices are column-major float[16] (GLSL/glm).
 *  - every function returns 0 on success or a negative mlgi_status; mlgi_last_error(ctx)
 *    returns a message. There is NO CPU fallback: without a CUDA device mlgi_create fails.
 *  - unless stated otherwise pointers are DEVICE pointers; inputs are caller-owned and must
 *    stay valid until the stream reaches the call; outputs are owned by the ctx and
 *    exposed through mlgi_get_buffer. All calls are asynchronous on `stream`
 *    (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *  - one ctx per GPU, not thread-safe (mirrors the reference's single GL thread).
 */
#ifndef MLGI_H
#define MLGI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MLGI_API __attribute__((visibility("default")))
#else
#define MLGI_API
#endif

typedef struct mlgi_ctx mlgi_ctx;

typedef enum {
    MLGI_OK = 0,
    MLGI_E_INVALID = -1,      /* bad argument / configuration */
    MLGI_E_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
    MLGI_E_NO_DEVICE = -3,    /* no usable CUDA device: the product has no CPU path */
    MLGI_E_UNSUPPORTED = -4,  /* reference feature outside the hot path (non push-pull raster path, debug overlay ...) */
    MLGI_E_CAPACITY = -5,     /* a ctx buffer is too small for the input (e.g. max_points) */
    MLGI_E_ORDER = -6         /* a stage was called before the stage that produces its input */
} mlgi_status;

typedef enum { MLGI_CULL_CLUSTERED = 0, MLGI_CULL_TILED = 1, MLGI_CULL_NONE = 2 } mlgi_cull_mode;

/* Sizes fixed at creation (the reference hard-codes them: SURVEY.md Appendix D). */
typedef struct {
    int device;          /* CUDA device ordinal */
    int width, height;   /* camera viewport (GIStage::resizeTexture, GIStage.cpp:468) */
    int vpl_count;       /* N = totalVplCount; multiple of 16 and of ceil(sqrt(N)); 16..16384 */
    int rsm_width, rsm_height; /* RSM size (1024x256 in GIStage.cpp:236) */
    int max_triangles;   /* capacity of the scene mesh handed to mlgi_ism_points */
    int max_points;      /* capacity of the point buffer (reference: 2^23, ImperfectShadowmap.cpp:95) */
    int class_lo, class_hi; /* interleave classes [lo,hi) owned by this ctx; 0,16 = all (multi-GPU shard) */
    int enable_timers;   /* record per-stage cudaEvents (PerfCounter.cpp:50-76 scope names) */
    int gi_packed;       /* 1: gathering also emits the reference's giBuffer format GL_R11F_G11F_B10F (GIStage.cpp:470)
                            into MLGI_BUF_GI_PACKED, and the multi-GPU exchange buffer is packed uint32;
                            2: ONLY the packed image is produced (MLGI_BUF_GI is not written: 4 B instead of 20 B stored per pixel) */
} mlgi_config;

/* Tunables; names and defaults mirror GIStage (GIStage.h:84-100, GIStage.cpp:215-232). */
typedef struct {
    float light_intensity;        /* 5.0   */
    float gi_intensity_factor;    /* 3000  */
    float vpl_clamping_value;     /* 0.001 */
    int   vpl_start_index;        /* 0     */
    int   vpl_end_index;          /* N     (by itself only affects the normalisation: Q12) */
    int   scale_isms;             /* 0     n3: vplCount = end - start larger ISMs fill the atlas (ISM of VPL i at index i - start;
                                           ism.comp:38-39,124, final_gathering.comp:43-46 as rewritten by GIStage.cpp:430) */
    int   points_only_into_scaled_isms; /* 0  n3: point buckets and their 16-VPL windows wrap inside [start, end) (ism.geom:43,57,
                                           ism.comp:56-66). Both flags need an unsharded ctx (class_lo = 0, class_hi = 16). The
                                           VPL records carry the ISM layout: pass the same params to every stage of a frame. */
    float tess_level_factor;      /* 2.0   */
    int   use_push_pull;          /* 1     (0 = legacy point-sprite raster path -> MLGI_E_UNSUPPORTED) */
    int   enable_shadowing;       /* 1     */
    int   show_vpl_positions;     /* 0     (debug overlay; 1 -> MLGI_E_UNSUPPORTED) */
    int   use_interleaving;       /* 1     (0 is broken upstream, SURVEY A.9 -> MLGI_E_UNSUPPORTED) */
    int   shuffle_lights;         /* 1     */
    float z_near, z_far;          /* camera AND light projection planes (GIStage.cpp:240-241); 0.05 / 50 */
    int   cull_mode;              /* mlgi_cull_mode used by mlgi_frame / mlgi_final_gather */
} mlgi_params;

typedef struct {
    float view[16];
    float proj[16];
} mlgi_camera;

/* RSM G-buffer in (RasterizationStage.cpp:139-149): */
typedef struct {
    const uint32_t* diffuse;      /* RGBA8, r in byte 0 (GL_RGB8)           rsm_width*rsm_height */
    const uint32_t* face_normal;  /* RGB10_A2 unorm of n*0.5+0.5, r in bits 0-9 */
    const float*    depth;        /* window-space depth [0,1]               */
} mlgi_rsm;

/* camera G-buffer in (MultiFramePainter.cpp:115-119): */
typedef struct {
    const float*    depth;        /* width*height fp32 window-space depth */
    const uint32_t* face_normal;  /* width*height RGB10_A2 */
} mlgi_gbuffer;

/* scene triangles in (ModelLoadingStage drawables, ImperfectShadowmap.cpp:225-240) */
typedef struct {
    const float*    vertices;     /* 3 floats per vertex */
    const float*    normals;      /* 3 floats per vertex */
    const uint32_t* indices;      /* 3 per triangle */
    const uint32_t* prim_id;      /* gl_PrimitiveIDIn of each triangle (index inside its draw call) */
    int num_vertices, num_triangles;
} mlgi_mesh;

/* Buffers a caller may inspect (the reference's BlitStage list, MultiFramePainter.cpp:153-176). */
typedef enum {
    MLGI_BUF_VPL = 0,              /* N x 12 float {pos,_,nrm,_,col,_}          VPLProcessor.cpp:34-41   */
    MLGI_BUF_PACKED_VPL,           /* N x 4 float (pos, pack3SN(nrm))           VPLProcessor.cpp:29-32   */
    MLGI_BUF_SHUFFLE,              /* N x int32 permutation                     VPLProcessor.cpp:58-65   */
    MLGI_BUF_POINTS,               /* P x 4 float (pos, unorm4x8(nrm, r/25)), grouped by bucket. The frame keeps an unordered
                                      list; this view (and BUCKET_START) is built when first requested (synchronises) */
    MLGI_BUF_BUCKET_START,         /* (N+1) x uint32 prefix of bucket sizes (replaces atomicCounter[N]) */
    MLGI_BUF_ATLAS,                /* A x A uint32 depth24<<8|radius8           ImperfectShadowmap.cpp:76-80 */
    MLGI_BUF_ISM,                  /* A x A uint16 R16 unorm push-pull result   ImperfectShadowmap.cpp:102-106 */
    MLGI_BUF_COMPACT_CLUSTER_IDS,  /* numUsed x uint32 x|y<<8|slice<<16         ClusteredShading.cpp:121 */
    MLGI_BUF_LIGHT_LIST_IDS,       /* [16][Y][X] uint16, 0xFFFF = unused        ClusteredShading.cpp:122 */
    MLGI_BUF_LIGHT_LISTS,          /* [numUsed][16][L+1] uint16                 ClusteredShading.cpp:124 */
    MLGI_BUF_TILE_LISTS,           /* [BY*BX][16][L+1] uint16 (tiled variant)   */
    MLGI_BUF_NUM_USED_CLUSTERS,    /* 1 x uint32                                */
    MLGI_BUF_GI,                   /* width*height x 4 float RGBA raster        GIStage giBuffer */
    MLGI_BUF_GI_PLANAR,            /* [class_hi-class_lo][ceil(H/4)][ceil(W/4)] x 4 float, written when the ctx owns a strict
                                      subset of the classes and no exchange buffer is bound (multi-GPU layout) */
    MLGI_BUF_GI_BLUR,              /* width*height x 4 float                    GIStage giBlurFinalBuffer */
    MLGI_BUF_GI_PACKED,            /* width*height x uint32 R11F_G11F_B10F (r in bits 0-10), only with cfg.gi_packed */
    MLGI_BUF_COUNT
} mlgi_buffer_id;

MLGI_API int  mlgi_version(void);
MLGI_API void mlgi_default_params(mlgi_params* p, int vpl_count);
MLGI_API int  mlgi_create(const mlgi_config* cfg, mlgi_ctx** out);
MLGI_API void mlgi_destroy(mlgi_ctx* ctx);
MLGI_API const char* mlgi_last_error(const mlgi_ctx* ctx);   /* ctx may be NULL: last create error */

/* a1  VPLProcessor::process (VPLProcessor.cpp:73-101, vpl_processor.comp:41-72) */
MLGI_API int mlgi_vpl_process(mlgi_ctx*, const mlgi_rsm* rsm, const mlgi_camera* light, const mlgi_params*, void* stream);

/* a2  scene -> points (ism.tesc/tese/geom, ImperfectShadowmap.cpp:225-240); deterministic tessellator */
MLGI_API int mlgi_ism_points(mlgi_ctx*, const mlgi_mesh* mesh, const mlgi_params*, void* stream);
/* caller-supplied point list instead (parity interface "given the same point list"):
 * points grouped by bucket, bucket_start has N+1 entries; both DEVICE pointers, copied into the ctx */
MLGI_API int mlgi_ism_set_points(mlgi_ctx*, const float* points, const uint32_t* bucket_start, int num_points, void* stream);

/* a3  ism.comp:54-139 (ImperfectShadowmap.cpp:244-258 + clear :212) */
MLGI_API int mlgi_ism_splat(mlgi_ctx*, const mlgi_params*, void* stream);
/* a4+a5  pull.comp x6 + push.comp x6 (ImperfectShadowmap.cpp:121-182) as one shared-memory pyramid per ISM.
 * atlas_override (device, A*A uint32) may replace the ctx atlas as input; NULL = ctx atlas. */
MLGI_API int mlgi_ism_pullpush(mlgi_ctx*, const uint32_t* atlas_override, const mlgi_params*, void* stream);

/* a6+a7  clustering.comp + light_lists.comp (ClusteredShading.cpp:70-112) */
MLGI_API int mlgi_cull_clustered(mlgi_ctx*, const mlgi_gbuffer*, const mlgi_camera*, const mlgi_params*, void* stream);
/* a8  tiled variant (upstream branch tiled_shading, README.md:10; spec in DESIGN.md) */
MLGI_API int mlgi_cull_tiled(mlgi_ctx*, const mlgi_gbuffer*, const mlgi_camera*, const mlgi_params*, void* stream);

/* a9  final_gathering.comp:63-135 (GIStage.cpp:261-306). ism_override (device A*A uint16) may replace the ctx ISM. */
MLGI_API int mlgi_final_gather(mlgi_ctx*, const mlgi_gbuffer*, const mlgi_camera*, const uint16_t* ism_override,
                               const mlgi_params*, void* stream);
/* n1  gi_blur.frag x2 (GIStage.cpp:308-344) */
MLGI_API int mlgi_gi_blur(mlgi_ctx*, const mlgi_gbuffer*, const mlgi_camera*, void* stream);

/* n4, ambient term of the deferred composite (deferredshading.frag:69-80 ambientTerm, tonemap :38-43, toSRGB
 * srgb_utils.glsl): out = sRGB8(tonemap(gi * toLinear(diffuse) * occlusion)). diffuse_rgba8: camera diffuse G-buffer
 * (device, width*height, r in bits 0-7); occlusion: device fp32 or NULL (= 1, SSAO is out of scope); use_blurred: read
 * giBlurFinalBuffer (needs mlgi_gi_blur) instead of giBuffer; out_srgb8: device, width*height RGBA8 (alpha 255). */
MLGI_API int mlgi_composite_ambient(mlgi_ctx*, const uint32_t* diffuse_rgba8, const float* occlusion, float exposure, int use_blurred,
                                    uint32_t* out_srgb8, void* stream);

/* a10 GIStage::process (GIStage.cpp:346-418) minus the RSM raster: a1..a9 in order, device inputs. */
MLGI_API int mlgi_frame(mlgi_ctx*, const mlgi_rsm*, const mlgi_camera* light, const mlgi_mesh*,
                        const mlgi_gbuffer*, const mlgi_camera* cam, const mlgi_params*, void* stream);

/* The frame entry points replay a CUDA graph of the stage launches (re-captured only when pointers or values change; an
 * animated light updates the instantiated graph in place). mlgi_set_graph(ctx, 0) or MLGI_GRAPH=0 selects direct launches;
 * per-stage timers always use direct launches. */
MLGI_API int mlgi_set_graph(mlgi_ctx*, int enable);
/* Builds (captures, instantiates, uploads) the graph that the next mlgi_frame (sharded = 0) or mlgi_frame_sharded (sharded = 1)
 * call with the same arguments replays, without launching it: takes the first-frame cost out of the first frame. Optional. */
MLGI_API int mlgi_frame_prepare(mlgi_ctx*, const mlgi_rsm*, const mlgi_camera* light, const mlgi_mesh*, const mlgi_gbuffer*,
                                const mlgi_camera* cam, const mlgi_params*, int sharded, void* stream);
typedef struct {
    double host_enqueue_us;                 /* host time spent inside the last mlgi_frame / mlgi_frame_sharded call */
    unsigned long long graph_replays;       /* frames issued as one cudaGraphLaunch */
    unsigned long long graph_instantiations;/* captures that needed a new executable (value-only changes update in place) */
    int graph_enabled;
} mlgi_frame_stats;
MLGI_API int mlgi_get_frame_stats(mlgi_ctx*, mlgi_frame_stats* out);

/* Same, but the per-frame inputs (RSM + camera G-buffer) are HOST pointers (pinned for
 * full speed) and the GI image is copied back to host_gi (width*height*4 float, may be
 * NULL to skip). The mesh stays resident on the device (it is static scene data in the
 * reference too). Of the RSM only the N texels the VPL build samples are read (by the CPU) and uploaded, 12 B per VPL; the
 * RSM planes may be pageable memory. Synchronises the stream before returning. This is the end-to-end call. */
MLGI_API int mlgi_frame_host(mlgi_ctx*, const mlgi_rsm* host_rsm, const mlgi_camera* light, const mlgi_mesh* dev_mesh,
                             const mlgi_gbuffer* host_gbuffer, const mlgi_camera* cam, const mlgi_params*,
                             float* host_gi, void* stream);

/* Same as mlgi_frame_host but the image comes back in the reference's own giBuffer format (GL_R11F_G11F_B10F,
 * GIStage.cpp:470): width*height uint32. Needs cfg.gi_packed. A quarter of the device->host bytes. */
MLGI_API int mlgi_frame_host_packed(mlgi_ctx*, const mlgi_rsm* host_rsm, const mlgi_camera* light, const mlgi_mesh* dev_mesh,
                                    const mlgi_gbuffer* host_gbuffer, const mlgi_camera* cam, const mlgi_params*,
                                    uint32_t* host_gi_packed, void* stream);

/* inspection */
MLGI_API int mlgi_get_buffer(mlgi_ctx*, int buffer_id, void** dev_ptr, size_t* bytes);
/* synchronous device->host copy of `bytes` bytes of a ctx buffer starting at byte `offset` (test/debug path) */
MLGI_API int mlgi_read_buffer(mlgi_ctx*, int buffer_id, void* host_dst, size_t bytes, size_t offset);
/* asynchronous variant on `stream` (host_dst should be pinned); no synchronisation */
MLGI_API int mlgi_read_buffer_async(mlgi_ctx*, int buffer_id, void* host_dst, size_t bytes, size_t offset, void* stream);
MLGI_API int mlgi_get_counts(mlgi_ctx*, int* num_points, int* num_used_clusters);  /* synchronises */
/* exact work counters of final gathering for the roofline accounting: runs the counting variant of the
 * kernel on the same inputs (no image is written) and returns (pixel,VPL) pairs listed / lit. Synchronises. */
MLGI_API int mlgi_gather_stats(mlgi_ctx*, const mlgi_gbuffer*, const mlgi_camera*, const uint16_t* ism_override,
                               const mlgi_params*, unsigned long long* pairs_listed, unsigned long long* pairs_lit);
/* FFMA micro-benchmark on the ctx device: the FP32-pipe roofline denominator in TFLOP/s. Synchronises. */
MLGI_API float mlgi_measure_fp32_peak_tflops(mlgi_ctx*);
/* test hook for the pull-push kernel's exact square-root elimination: thresholds[i] = T(radii[i]) with
 * sqrtf(x) > r  <=>  x > T(r) for every float x. Device pointers. */
MLGI_API int mlgi_debug_sqrt_thresholds(mlgi_ctx*, const float* dev_radii, int n, float* dev_thresholds, void* stream);

/* multi-GPU exchange: the gathering kernel writes the class-planar planes of the classes this
 * ctx owns straight into a caller-owned [16][ceil(H/4)][ceil(W/4)] x 4 float device buffer (plane c at
 * c*PH*PW), so one in-place all-gather over ranks assembles the image; NULL unbinds. With cfg.gi_packed the
 * buffer holds one uint32 (R11F_G11F_B10F) per pixel instead of 4 floats. */
MLGI_API int mlgi_bind_planar_output(mlgi_ctx*, void* dev_planar16);
/* Peer-memory exchange (B200/NVSwitch path, replaces the all-gather; SURVEY 8(e) mlgi_dist_init / mlgi_allgather): every rank
 * allocates an exchange region, the 64-byte CUDA IPC handles are swapped by the caller (any host transport), and from then
 * on ONE kernel per frame signals "my class planes are written" in every peer's flag array, waits for all ranks, and
 * assembles the image while pulling the other ranks' planes over NVLink with 16-byte loads (MLGI_XCH=push: the gathering
 * epilogue stores into every rank's region instead, the round-1 scheme). The frame counter lives in device memory, so the
 * whole frame is graph-replayable. One process per GPU (or one process, mlgi_exchange_connect_local), one node, world in
 * {2,4,8}; rank r must own classes [16r/world, 16(r+1)/world).
 *   mlgi_exchange_create : allocate, return this rank's handle (out_handle64 = 64 bytes)
 *   mlgi_exchange_connect: all_handles = world x 64 bytes in rank order; maps the peers
 *   mlgi_exchange_finish : after mlgi_frame on the same stream: signal, wait, pull + de-interleave into
 *                          MLGI_BUF_GI_PACKED / MLGI_BUF_GI. Asynchronous; bounded wait (no GPU hang if a peer dies)
 *   mlgi_exchange_status : synchronises and reports a timed-out wait */
MLGI_API int mlgi_exchange_create(mlgi_ctx*, int rank, int world, void* out_handle64);
MLGI_API int mlgi_exchange_connect(mlgi_ctx*, const void* all_handles);
/* same-process form (a host driving several GPUs from one thread, like the reference's single painter thread): all_ctx =
 * world ctx pointers in rank order, every one after mlgi_exchange_create; peer access is enabled as needed */
MLGI_API int mlgi_exchange_connect_local(mlgi_ctx*, mlgi_ctx* const* all_ctx);
MLGI_API int mlgi_exchange_finish(mlgi_ctx*, void* stream);
MLGI_API int mlgi_exchange_status(mlgi_ctx*);
/* SURVEY 8(e) mlgi_frame_sharded: one rank's frame of a class-sharded job = mlgi_frame + mlgi_exchange_finish as ONE graph
 * launch; afterwards MLGI_BUF_GI_PACKED / MLGI_BUF_GI hold the complete image on every rank. With tiled culling the ranks also
 * split the per-pixel half of the culling: each rank bounds the tiles of ITS tile rows (mlgi_sharded_rows) and publishes the
 * 14 numbers per tile to the peers' exchange regions (a second bounded wait inside the frame). A collective: every rank of
 * the job calls it once per frame. */
MLGI_API int mlgi_frame_sharded(mlgi_ctx*, const mlgi_rsm*, const mlgi_camera* light, const mlgi_mesh*,
                                const mlgi_gbuffer*, const mlgi_camera* cam, const mlgi_params*, void* stream);
/* The image rows this rank is responsible for in the sharded frames: the 64-pixel tile rows t with t mod world == rank, i.e. the
 * blocks of *row_block (= 64) rows starting at *row_begin + j * *row_stride, j = 0, 1, ..., clipped to the image height. These are
 * the tiles it bounds and, in mlgi_frame_sharded_host, the G-buffer rows it uploads and the image rows it returns (interleaved
 * so that every rank owns the same share of every phase of that frame). Valid after mlgi_exchange_create. */
MLGI_API int mlgi_sharded_rows(mlgi_ctx*, int* row_begin, int* row_block, int* row_stride);
/* The same with HOST inputs (the N-GPU form of mlgi_frame_host_packed; tiled culling). host_gbuffer and host_gi_packed are
 * FULL-size host arrays; this rank uploads only ITS rows (mlgi_sharded_rows) of the G-buffer and the G-buffer is never
 * replicated: the ranks exchange tile bounds (64 B per tile), and every rank's gathering CTAs read a tile's pixels from the
 * rank that uploaded them, over NVLink. The frame runs in phases (blocks of tile rows): while one phase is gathered the next
 * one's rows are uploaded, bounded and published, and the previous one's rows of the GL_R11F_G11F_B10F image are assembled
 * from every rank's class planes and copied back -- only this rank's rows, into host_gi_packed (pinned for full speed). Every
 * input byte crosses PCIe once in the whole job (the 3 MB RSM once per rank). MLGI_BUF_GI_PACKED then holds this rank's rows
 * only. Needs cfg.gi_packed. Synchronises the stream. */
MLGI_API int mlgi_frame_sharded_host(mlgi_ctx*, const mlgi_rsm* host_rsm, const mlgi_camera* light, const mlgi_mesh* dev_mesh,
                                     const mlgi_gbuffer* host_gbuffer, const mlgi_camera* cam, const mlgi_params*,
                                     uint32_t* host_gi_packed, void* stream);
/* consumer-side de-interleave of a complete planar image into raster order (dev_raster NULL = ctx GI buffer) */
MLGI_API int mlgi_planar_to_raster(mlgi_ctx*, const void* dev_planar16, float* dev_raster, void* stream);

/* per-stage timers (names of PerfCounter scopes, GIStage.cpp:368-411): fills up to `cap`
 * entries, returns the count. Synchronises. */
typedef struct { char name[24]; float ms; int launches; } mlgi_timing;
MLGI_API int mlgi_get_timings(mlgi_ctx*, mlgi_timing* out, int cap);
MLGI_API int mlgi_set_timers(mlgi_ctx*, int enable);   /* toggle cfg.enable_timers at run time */
/* number of kernels this library launched since creation (bench.py's gpu_launches) */
MLGI_API unsigned long long mlgi_launch_count(const mlgi_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* MLGI_H */

// mlgi_stages.hpp -- C++ host-side mirror of the reference's stage classes for the GI hot path.
//
// Same class names, method names, argument meaning and public outputs as the reference
// (source/mfs-painters/multiframepainter/), with GL objects replaced by device pointers and every
// call forwarded to the C ABI in include/mlgi.h:
//
//   reference (file:line)                              here
//   VPLProcessor::process          VPLProcessor.h:22   mlgi::VPLProcessor::process
//   ImperfectShadowmap::process    ImperfectShadowmap.h:30  mlgi::ImperfectShadowmap::process
//   ClusteredShading::process      ClusteredShading.h:30    mlgi::ClusteredShading::process   (+ TiledShading)
//   GIStage::process               GIStage.h:40        mlgi::GIStage::process
//
// Error behaviour: the reference's stages return void and print to stdout on failure; here every
// failing C-ABI call throws mlgi::Error carrying the status code and mlgi_last_error().
// Header-only; link against libmlgi.so.
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/mlgi.h"

namespace mlgi {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("mlgi error " + std::to_string(c) + ": " + m), code(c) {}
};

// Owns the mlgi_ctx shared by the stages of one painter (the reference's stages share one GL context).
class Context {
public:
    explicit Context(const mlgi_config& cfg) {
        int rc = mlgi_create(&cfg, &m_ctx);
        if (rc) throw Error(rc, mlgi_last_error(nullptr));
    }
    ~Context() { mlgi_destroy(m_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mlgi_ctx* get() const { return m_ctx; }
    void check(int rc) const { if (rc) throw Error(rc, mlgi_last_error(m_ctx)); }
    template <typename T> T* buffer(int id) const {
        void* p = nullptr; size_t n = 0;
        check(mlgi_get_buffer(m_ctx, id, &p, &n));
        return static_cast<T*>(p);
    }
private:
    mlgi_ctx* m_ctx = nullptr;
};

// The RSM renderer's outputs + light camera (RasterizationStage members read by VPLProcessor.cpp:76-88)
struct RsmInputs {
    mlgi_rsm buffers;     // diffuseBuffer, faceNormalBuffer, depthBuffer
    mlgi_camera light;    // camera->view(), projection->projection()
};

class VPLProcessor {
public:
    explicit VPLProcessor(Context& ctx) : m_ctx(ctx) {
        vplBuffer = ctx.buffer<float>(MLGI_BUF_VPL);                 // VPLProcessor.h:24
        packedVplBuffer = ctx.buffer<float>(MLGI_BUF_PACKED_VPL);    // VPLProcessor.h:25
    }
    // VPLProcessor.h:22  process(const RasterizationStage& rsmRenderer, float lightIntensity, bool shuffleLights)
    void process(const RsmInputs& rsmRenderer, float lightIntensity, bool shuffleLights, mlgi_params params, void* stream = nullptr) {
        params.light_intensity = lightIntensity;
        params.shuffle_lights = shuffleLights ? 1 : 0;
        m_ctx.check(mlgi_vpl_process(m_ctx.get(), &rsmRenderer.buffers, &rsmRenderer.light, &params, stream));
    }
    float* vplBuffer;
    float* packedVplBuffer;
private:
    Context& m_ctx;
};

class ImperfectShadowmap {
public:
    explicit ImperfectShadowmap(Context& ctx) : m_ctx(ctx) {
        softrenderBuffer = ctx.buffer<uint32_t>(MLGI_BUF_ATLAS);         // ImperfectShadowmap.h:42
        pointBuffer = ctx.buffer<float>(MLGI_BUF_POINTS);                // :45
        pushPullResultBuffer = ctx.buffer<uint16_t>(MLGI_BUF_ISM);       // :46
    }
    // ImperfectShadowmap.h:30-39  process(drawablesMap, vplProcessor, vplStartIndex, vplEndIndex, scaleISMs,
    //                                      pointsOnlyIntoScaledISMs, tessLevelFactor, usePushPull, zFar)
    void process(const mlgi_mesh& drawablesMap, const VPLProcessor&, int vplStartIndex, int vplEndIndex, bool scaleISMs,
                 bool pointsOnlyIntoScaledISMs, float tessLevelFactor, bool usePushPull, float zFar, mlgi_params params,
                 void* stream = nullptr) const {
        params.vpl_start_index = vplStartIndex; params.vpl_end_index = vplEndIndex;
        params.scale_isms = scaleISMs; params.points_only_into_scaled_isms = pointsOnlyIntoScaledISMs;
        params.tess_level_factor = tessLevelFactor; params.use_push_pull = usePushPull; params.z_far = zFar;
        render(drawablesMap, params, stream);       // ImperfectShadowmap.cpp:186
        pullpush(params, stream);                   // ImperfectShadowmap.cpp:191
    }
    uint32_t* softrenderBuffer;
    float* pointBuffer;
    uint16_t* pushPullResultBuffer;
protected:
    void render(const mlgi_mesh& mesh, const mlgi_params& p, void* stream) const {     // ImperfectShadowmap.cpp:194-261
        m_ctx.check(mlgi_ism_points(m_ctx.get(), &mesh, &p, stream));                  //   "ISM render"
        m_ctx.check(mlgi_ism_splat(m_ctx.get(), &p, stream));                          //   "ISM CS"
    }
    void pullpush(const mlgi_params& p, void* stream) const {                           // ImperfectShadowmap.cpp:121-182
        m_ctx.check(mlgi_ism_pullpush(m_ctx.get(), nullptr, &p, stream));
    }
private:
    Context& m_ctx;
};

class ClusteredShading {
public:
    explicit ClusteredShading(Context& ctx) : m_ctx(ctx) {
        compactUsedClusterIDs = ctx.buffer<uint32_t>(MLGI_BUF_COMPACT_CLUSTER_IDS);   // ClusteredShading.h:43
        lightListIds = ctx.buffer<uint16_t>(MLGI_BUF_LIGHT_LIST_IDS);                 // :44
        lightLists = ctx.buffer<uint16_t>(MLGI_BUF_LIGHT_LISTS);                      // :46
    }
    // ClusteredShading.h:30-39  process(vplProcessor, view, projection, viewport, zFar, vplStartIndex, vplEndIndex, depthBuffer, vplBuffer)
    void process(const VPLProcessor&, const float view[16], const float projection[16], float zFar, int vplStartIndex, int vplEndIndex,
                 const float* depthBuffer, mlgi_params params, void* stream = nullptr) {
        mlgi_camera cam;
        for (int i = 0; i < 16; i++) { cam.view[i] = view[i]; cam.proj[i] = projection[i]; }
        mlgi_gbuffer gb = {depthBuffer, nullptr};
        params.z_far = zFar; params.vpl_start_index = vplStartIndex; params.vpl_end_index = vplEndIndex;
        m_ctx.check(mlgi_cull_clustered(m_ctx.get(), &gb, &cam, &params, stream));
    }
    uint32_t* compactUsedClusterIDs;
    uint16_t* lightListIds;
    uint16_t* lightLists;
private:
    Context& m_ctx;
};

// tiled variant (upstream branch tiled_shading, README.md:10): same call shape, per-tile lists
class TiledShading {
public:
    explicit TiledShading(Context& ctx) : m_ctx(ctx) { tileLists = ctx.buffer<uint16_t>(MLGI_BUF_TILE_LISTS); }
    // the tile test also bounds the receiver side, so it reads the face normals as well
    void process(const VPLProcessor&, const float view[16], const float projection[16], float zFar, const float* depthBuffer,
                 const uint32_t* faceNormalBuffer, mlgi_params params, void* stream = nullptr) {
        mlgi_camera cam;
        for (int i = 0; i < 16; i++) { cam.view[i] = view[i]; cam.proj[i] = projection[i]; }
        mlgi_gbuffer gb = {depthBuffer, faceNormalBuffer};
        params.z_far = zFar;
        m_ctx.check(mlgi_cull_tiled(m_ctx.get(), &gb, &cam, &params, stream));
    }
    uint16_t* tileLists;
private:
    Context& m_ctx;
};

// GIStage (GIStage.h:35-100): owns the sub-stages, holds the tunables as members with the reference's
// names/defaults (GIStage.cpp:215-232) and runs them in the reference's order (GIStage.cpp:346-418).
class GIStage {
    Context m_ctx;          // declared first: the sub-stages below hold references to it
    mlgi_params m_params;
public:
    GIStage(const mlgi_config& cfg)
        : m_ctx(cfg), vplProcessor(m_ctx), ism(m_ctx), clusteredShading(m_ctx), tiledShading(m_ctx) {
        mlgi_default_params(&m_params, cfg.vpl_count);
        giBuffer = m_ctx.buffer<float>(MLGI_BUF_GI);                 // GIStage.h:45
        giBlurFinalBuffer = m_ctx.buffer<float>(MLGI_BUF_GI_BLUR);   // GIStage.h:47
        lightIntensity = m_params.light_intensity; giIntensityFactor = m_params.gi_intensity_factor;
        vplClampingValue = m_params.vpl_clamping_value; vplStartIndex = 0; vplEndIndex = cfg.vpl_count;
        scaleISMs = false; pointsOnlyIntoScaledISMs = false; tessLevelFactor = m_params.tess_level_factor;
        usePushPull = true; enableShadowing = true; showVPLPositions = false; useInterleaving = true; shuffleLights = true;
        useTiledShading = true;
    }
    // inputs wired by the painter (MultiFramePainter.cpp:115-119): public members in the reference too
    const float* depthBuffer = nullptr;           // GIStage.h:43
    const uint32_t* faceNormalBuffer = nullptr;   // GIStage.h:42
    mlgi_camera camera{};                         // camera->view() / projection->projection()
    float zNear = 0.05f, zFar = 50.0f;
    RsmInputs rsmRenderer{};                      // GIStage.h:56
    mlgi_mesh drawablesMap{};                     // modelLoadingStage.getDrawablesMap()

    // outputs
    float* giBuffer;
    float* giBlurFinalBuffer;

    // tunables (GIStage.h:84-100)
    float lightIntensity, giIntensityFactor, vplClampingValue, tessLevelFactor;
    int vplStartIndex, vplEndIndex;
    bool scaleISMs, pointsOnlyIntoScaledISMs, usePushPull, enableShadowing, showVPLPositions, useInterleaving, shuffleLights;
    bool useTiledShading;   // README.md:10: tiled instead of clustered culling

    Context& context() { return m_ctx; }

    // GIStage.cpp:346-418 (the RSM raster :367-371 is the caller's: rsmRenderer holds its outputs). One call = mlgi_frame: the
    // stages in the reference's order with culling and the VPL build forked beside point generation, replayed as a CUDA graph
    // (direct launches while per-stage timers are enabled).
    void process(void* stream = nullptr, bool blurPass = true) {
        mlgi_params p = params();
        mlgi_gbuffer gb = {depthBuffer, faceNormalBuffer};
        m_ctx.check(mlgi_frame(m_ctx.get(), &rsmRenderer.buffers, &rsmRenderer.light, &drawablesMap, &gb, &camera, &p, stream));
        if (blurPass) blur(stream);                                                                                    // :410-417
    }
    // the same frame issued stage by stage on one stream, literally as GIStage.cpp:373-417 does on its GL queue
    void processStaged(void* stream = nullptr, bool blurPass = true) {
        mlgi_params p = params();
        vplProcessor.process(rsmRenderer, lightIntensity, shuffleLights, p, stream);                                   // :373-376
        ism.process(drawablesMap, vplProcessor, vplStartIndex, vplEndIndex, scaleISMs, pointsOnlyIntoScaledISMs,
                    tessLevelFactor, usePushPull, zFar, p, stream);                                                    // :378-389
        if (useTiledShading) tiledShading.process(vplProcessor, camera.view, camera.proj, zFar, depthBuffer, faceNormalBuffer, p, stream);
        else clusteredShading.process(vplProcessor, camera.view, camera.proj, zFar, vplStartIndex, vplEndIndex, depthBuffer, p, stream);   // :392-403
        compute_final_gathering(stream);                                                                               // :405-408
        if (blurPass) blur(stream);                                                                                    // :410-417
    }
    // Multi-GPU (no counterpart in the reference): one GIStage per GPU, each owning the interleave classes
    // [16 r / world, 16 (r+1) / world) (cfg.class_lo / class_hi). exchangeCreate + exchangeConnect (one process per GPU: the
    // 64-byte handles travel over any host transport) or connectLocal (one process driving all GPUs), then processSharded on
    // every rank per frame: mlgi_frame + the peer-memory exchange as one graph launch; giPackedBuffer / giBuffer then hold the
    // whole image on every rank.
    void exchangeCreate(int rank, int world, void* outHandle64) { m_ctx.check(mlgi_exchange_create(m_ctx.get(), rank, world, outHandle64)); }
    void exchangeConnect(const void* allHandles) { m_ctx.check(mlgi_exchange_connect(m_ctx.get(), allHandles)); }
    static void connectLocal(GIStage* const* ranks, int world) {
        mlgi_ctx* all[8] = {nullptr};
        unsigned char handle[64];
        for (int r = 0; r < world; r++) { ranks[r]->exchangeCreate(r, world, handle); all[r] = ranks[r]->m_ctx.get(); }
        for (int r = 0; r < world; r++) ranks[r]->m_ctx.check(mlgi_exchange_connect_local(all[r], all));
    }
    void processSharded(void* stream = nullptr) {
        mlgi_params p = params();
        mlgi_gbuffer gb = {depthBuffer, faceNormalBuffer};
        m_ctx.check(mlgi_frame_sharded(m_ctx.get(), &rsmRenderer.buffers, &rsmRenderer.light, &drawablesMap, &gb, &camera, &p, stream));
    }
    // host-input forms: pinned host RSM / G-buffer in, GL_R11F_G11F_B10F image out (GIStage.cpp:470's giBuffer format)
    void processHost(const mlgi_rsm& hostRsm, const mlgi_gbuffer& hostGbuffer, uint32_t* hostGiPacked, void* stream = nullptr) {
        mlgi_params p = params();
        m_ctx.check(mlgi_frame_host_packed(m_ctx.get(), &hostRsm, &rsmRenderer.light, &drawablesMap, &hostGbuffer, &camera, &p, hostGiPacked, stream));
    }
    void processShardedHost(const mlgi_rsm& hostRsm, const mlgi_gbuffer& hostGbuffer, uint32_t* hostGiPacked, void* stream = nullptr) {
        mlgi_params p = params();
        m_ctx.check(mlgi_frame_sharded_host(m_ctx.get(), &hostRsm, &rsmRenderer.light, &drawablesMap, &hostGbuffer, &camera, &p, hostGiPacked, stream));
    }

    VPLProcessor vplProcessor;
    ImperfectShadowmap ism;
    ClusteredShading clusteredShading;
    TiledShading tiledShading;

protected:
    mlgi_params params() const {
        mlgi_params p = m_params;
        p.light_intensity = lightIntensity; p.gi_intensity_factor = giIntensityFactor; p.vpl_clamping_value = vplClampingValue;
        p.vpl_start_index = vplStartIndex; p.vpl_end_index = vplEndIndex; p.scale_isms = scaleISMs;
        p.points_only_into_scaled_isms = pointsOnlyIntoScaledISMs; p.tess_level_factor = tessLevelFactor;
        p.use_push_pull = usePushPull; p.enable_shadowing = enableShadowing; p.show_vpl_positions = showVPLPositions;
        p.use_interleaving = useInterleaving; p.shuffle_lights = shuffleLights; p.z_near = zNear; p.z_far = zFar;
        p.cull_mode = useTiledShading ? MLGI_CULL_TILED : MLGI_CULL_CLUSTERED;
        return p;
    }
    void compute_final_gathering(void* stream) {                      // GIStage.cpp:261-306
        mlgi_params p = params();
        mlgi_gbuffer gb = {depthBuffer, faceNormalBuffer};
        m_ctx.check(mlgi_final_gather(m_ctx.get(), &gb, &camera, nullptr, &p, stream));
    }
    void blur(void* stream) {                                         // GIStage.cpp:308-344
        mlgi_gbuffer gb = {depthBuffer, faceNormalBuffer};
        m_ctx.check(mlgi_gi_blur(m_ctx.get(), &gb, &camera, stream));
    }
};

}  // namespace mlgi

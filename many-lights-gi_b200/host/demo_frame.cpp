// demo_frame.cpp -- drives one GI frame through the C++ stage classes (mlgi_stages.hpp), the way
// MultiFramePainter::onPaint drives GIStage::process (MultiFramePainter.cpp:182-201), on the synthetic
// atrium fixture. Prints per-stage timers (PerfCounter scope names) and a checksum of the GI image.
//   usage: demo_frame [width height vpls [gi_image.raw]]   (the raw file receives the fp32 RGBA GI image of the graph-replayed frame)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "mlgi_stages.hpp"

extern "C" {
void* mlgi_scene_create(uint64_t seed, float cell_size, int clutter_boxes);
void mlgi_scene_destroy(void*);
int mlgi_scene_num_vertices(void*);
int mlgi_scene_num_triangles(void*);
void mlgi_scene_get_mesh(void*, float*, float*, uint32_t*, uint32_t*);
void mlgi_scene_look_at(const float*, const float*, const float*, float*);
void mlgi_scene_perspective(float, float, float, float, float*);
void mlgi_scene_ortho(float, float, float, float, float, float, float*);
void mlgi_scene_sun_eye(float, const float*, float*);
long long mlgi_scene_render(void*, const float*, const float*, int, int, float*, uint32_t*, uint32_t*);
}

template <typename T> static T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char** argv) {
    const int W = argc > 3 ? atoi(argv[1]) : 1920, H = argc > 3 ? atoi(argv[2]) : 1080, N = argc > 3 ? atoi(argv[3]) : 1024;
    const int RW = 1024, RH = 256;
    void* sc = mlgi_scene_create(0x5EED, 0.155f, 24);
    const int nv = mlgi_scene_num_vertices(sc), nt = mlgi_scene_num_triangles(sc);
    std::vector<float> verts(3 * (size_t)nv), normals(3 * (size_t)nv);
    std::vector<uint32_t> indices(3 * (size_t)nt), prim(nt);
    mlgi_scene_get_mesh(sc, verts.data(), normals.data(), indices.data(), prim.data());

    const float eye[3] = {-13.0f, 2.5f, -0.23f}, center[3] = {0.009f, -0.019f, -0.021f}, up[3] = {0, 1, 0};   // CrytekSponza preset
    const float lcenter[3] = {0.0f, 7.5f, 0.0f};
    float leye[3];
    mlgi_camera cam, light;
    mlgi_scene_look_at(eye, center, up, cam.view);
    mlgi_scene_perspective(60.0f * 3.14159265f / 180.0f, float(W) / float(H), 0.05f, 50.0f, cam.proj);
    mlgi_scene_sun_eye(266.0f, lcenter, leye);                                    // GIStage.cpp:228,357-361
    mlgi_scene_look_at(leye, lcenter, up, light.view);
    mlgi_scene_ortho(-10.0f, 10.0f, -2.5f, 2.5f, 0.05f, 50.0f, light.proj);       // height 5 -> width 20 (GIStage.cpp:236-238)

    std::vector<float> depth((size_t)W * H), rdepth((size_t)RW * RH);
    std::vector<uint32_t> normal((size_t)W * H), rnormal((size_t)RW * RH), rdiffuse((size_t)RW * RH);
    if (mlgi_scene_render(sc, cam.view, cam.proj, W, H, depth.data(), normal.data(), nullptr) != 0 ||
        mlgi_scene_render(sc, light.view, light.proj, RW, RH, rdepth.data(), rnormal.data(), rdiffuse.data()) != 0) {
        fprintf(stderr, "scene is not closed\n");
        return 2;
    }

    mlgi_config cfg = {};
    cfg.device = 0; cfg.width = W; cfg.height = H; cfg.vpl_count = N; cfg.rsm_width = RW; cfg.rsm_height = RH;
    cfg.max_triangles = nt; cfg.max_points = 1 << 23; cfg.class_lo = 0; cfg.class_hi = 16; cfg.enable_timers = 1; cfg.gi_packed = 0;
    try {
        mlgi::GIStage gi(cfg);
        gi.depthBuffer = upload(depth); gi.faceNormalBuffer = upload(normal);
        gi.camera = cam;
        gi.rsmRenderer.buffers = {upload(rdiffuse), upload(rnormal), upload(rdepth)};
        gi.rsmRenderer.light = light;
        float* dv = upload(verts); float* dn = upload(normals); uint32_t* di = upload(indices); uint32_t* dp = upload(prim);
        gi.drawablesMap = {dv, dn, di, dp, nv, nt};
        for (int i = 0; i < 3; i++) gi.process();      // frames (direct launches: the per-stage timers are on); the last one is reported
        cudaDeviceSynchronize();
        mlgi_timing t[16];
        const int n = mlgi_get_timings(gi.context().get(), t, 16);
        for (int i = 0; i < n; i++) printf("%-12s %8.3f ms  (%d launches)\n", t[i].name, t[i].ms, t[i].launches);
        std::vector<float> img((size_t)W * H * 4), img2((size_t)W * H * 4);
        gi.context().check(mlgi_read_buffer(gi.context().get(), MLGI_BUF_GI, img.data(), img.size() * 4, 0));
        // the same frame as one CUDA-graph launch (timers off), and stage by stage on one stream: all three must agree to the bit
        gi.context().check(mlgi_set_timers(gi.context().get(), 0));
        gi.process(nullptr, false);
        gi.context().check(mlgi_read_buffer(gi.context().get(), MLGI_BUF_GI, img2.data(), img2.size() * 4, 0));
        const bool same_graph = memcmp(img.data(), img2.data(), img.size() * 4) == 0;
        gi.processStaged(nullptr, false);
        gi.context().check(mlgi_read_buffer(gi.context().get(), MLGI_BUF_GI, img2.data(), img2.size() * 4, 0));
        const bool same_staged = memcmp(img.data(), img2.data(), img.size() * 4) == 0;
        mlgi_frame_stats fs;
        mlgi_get_frame_stats(gi.context().get(), &fs);
        printf("graph replay == direct launches: %s, staged == direct launches: %s (graph replays %llu)\n", same_graph ? "yes" : "NO",
               same_staged ? "yes" : "NO", (unsigned long long)fs.graph_replays);
        if (argc > 4) {
            FILE* f = fopen(argv[4], "wb");
            if (!f || fwrite(img2.data(), 4, img2.size(), f) != img2.size()) { fprintf(stderr, "cannot write %s\n", argv[4]); return 2; }
            fclose(f);
        }
        double sum = 0;
        for (size_t i = 0; i < img.size(); i += 4) sum += img[i] + img[i + 1] + img[i + 2];
        printf("mean radiance %.6f  (%dx%d, %d VPLs, %d triangles)\n", sum / (3.0 * W * H), W, H, N, nt);
        if (!same_graph || !same_staged) return 3;
    } catch (const mlgi::Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    mlgi_scene_destroy(sc);
    return 0;
}

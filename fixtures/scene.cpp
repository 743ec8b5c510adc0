// scene.cpp -- synthetic-input FIXTURE (untimed; not part of the hot path).
//
// The reference's hot path starts at G-buffers that its GL rasteriser produced
// (RasterizationStage.cpp:139-149, model.frag:92-93) from assets that are not shipped
// (.gitignore:1-7). This fixture stands in for that out-of-scope rasteriser: a closed
// procedural "atrium" of oriented boxes, its triangle mesh (input of
// the ISM point stage), and an analytic ray caster that renders the camera G-buffer
// (depth fp32 window space + RGB10_A2 face normal) and the RSM (RGBA8 diffuse + RGB10_A2
// normal + depth) from the same geometry, so that shadows are self-consistent.
//
// Exposes a plain C ABI (used through ctypes by tests/ and bench.py). CPU only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SCN_API extern "C" __attribute__((visibility("default")))

namespace {

struct d3 { double x, y, z; };
inline d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline d3 operator*(d3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline d3 cross(d3 a, d3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline d3 normalize(d3 a) { double l = std::sqrt(dot(a, a)); return a * (1.0 / l); }

struct Box {
    d3 c, h;          // centre, half extents
    double cs, sn;    // yaw (rotation about +y): local -> world  x' = cs*x + sn*z, z' = -sn*x + cs*z
    uint32_t rgba;    // diffuse colour, RGBA8 (r in byte 0)
    bool inward;      // room shell: normals point inside
};

inline d3 to_local(const Box& b, d3 v) { return {b.cs * v.x - b.sn * v.z, v.y, b.sn * v.x + b.cs * v.z}; }
inline d3 to_world(const Box& b, d3 v) { return {b.cs * v.x + b.sn * v.z, v.y, -b.sn * v.x + b.cs * v.z}; }

uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double urand(uint64_t& s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

struct Scene {
    std::vector<Box> boxes;
    std::vector<float> verts, normals;     // 3 floats per vertex
    std::vector<uint32_t> indices;         // 3 per triangle
    std::vector<uint32_t> prim_id;         // triangle index within its mesh (one mesh per box)
    std::vector<uint32_t> tri_box;         // owning box of each triangle
};

uint32_t object_colour(uint64_t seed, int objectId) {
    uint64_t s = seed ^ (uint64_t)objectId * 0x100000001B3ull;
    uint32_t r = 64 + (uint32_t)(splitmix64(s) % 176), g = 64 + (uint32_t)(splitmix64(s) % 176), b = 64 + (uint32_t)(splitmix64(s) % 176);
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}

void add_box(Scene& sc, uint64_t seed, d3 c, d3 h, double yaw, bool inward = false) {
    Box b; b.c = c; b.h = h; b.cs = std::cos(yaw); b.sn = std::sin(yaw); b.inward = inward;
    b.rgba = object_colour(seed, (int)sc.boxes.size());
    sc.boxes.push_back(b);
}

// one grid-tessellated quad per box face; every face owns its vertices (flat normals)
void tessellate(Scene& sc, double cell) {
    for (size_t bi = 0; bi < sc.boxes.size(); bi++) {
        const Box& b = sc.boxes[bi];
        uint32_t prim = 0;
        for (int axis = 0; axis < 3; axis++)
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                int ua = (axis + 1) % 3, va = (axis + 2) % 3;
                const double hh[3] = {b.h.x, b.h.y, b.h.z};
                int nu = std::max(1, (int)std::ceil(2 * hh[ua] / cell)), nv = std::max(1, (int)std::ceil(2 * hh[va] / cell));
                double nl[3] = {0, 0, 0};
                nl[axis] = b.inward ? -sgn : sgn;
                d3 nw = to_world(b, {nl[0], nl[1], nl[2]});
                uint32_t base = (uint32_t)(sc.verts.size() / 3);
                for (int j = 0; j <= nv; j++)
                    for (int i = 0; i <= nu; i++) {
                        double p[3];
                        p[axis] = sgn * hh[axis];
                        p[ua] = -hh[ua] + 2 * hh[ua] * i / nu;
                        p[va] = -hh[va] + 2 * hh[va] * j / nv;
                        d3 w = to_world(b, {p[0], p[1], p[2]}) + b.c;
                        sc.verts.push_back((float)w.x); sc.verts.push_back((float)w.y); sc.verts.push_back((float)w.z);
                        sc.normals.push_back((float)nw.x); sc.normals.push_back((float)nw.y); sc.normals.push_back((float)nw.z);
                    }
                for (int j = 0; j < nv; j++)
                    for (int i = 0; i < nu; i++) {
                        uint32_t v00 = base + j * (nu + 1) + i, v10 = v00 + 1, v01 = v00 + (nu + 1), v11 = v01 + 1;
                        const uint32_t t[6] = {v00, v10, v11, v00, v11, v01};
                        for (int k = 0; k < 6; k++) sc.indices.push_back(t[k]);
                        for (int k = 0; k < 2; k++) { sc.prim_id.push_back(prim++); sc.tri_box.push_back((uint32_t)bi); }
                    }
            }
    }
}

// ray vs oriented box, returns nearest t > tmin and the world normal of the hit face
bool hit_box(const Box& b, d3 o, d3 d, double tmin, double& tbest, d3& nbest) {
    d3 lo = to_local(b, o - b.c), ld = to_local(b, d);
    const double O[3] = {lo.x, lo.y, lo.z}, D[3] = {ld.x, ld.y, ld.z}, Hh[3] = {b.h.x, b.h.y, b.h.z};
    double t0 = -1e300, t1 = 1e300; int a0 = -1, a1 = -1; double s0 = 0, s1 = 0;
    for (int a = 0; a < 3; a++) {
        if (std::fabs(D[a]) < 1e-300) { if (std::fabs(O[a]) > Hh[a]) return false; continue; }
        double inv = 1.0 / D[a];
        double ta = (-Hh[a] - O[a]) * inv, tb = (Hh[a] - O[a]) * inv;
        double sa = -1, sb = 1;
        if (ta > tb) { std::swap(ta, tb); std::swap(sa, sb); }
        if (ta > t0) { t0 = ta; a0 = a; s0 = sa; }
        if (tb < t1) { t1 = tb; a1 = a; s1 = sb; }
    }
    if (t0 > t1) return false;
    double t; int a; double s;
    if (!b.inward) { t = t0; a = a0; s = s0; }   // entering a solid box
    else { t = t1; a = a1; s = -s1; }            // leaving the room shell: hit its inner wall
    if (a < 0 || t <= tmin || t >= tbest) return false;
    double nl[3] = {0, 0, 0};
    nl[a] = s;
    tbest = t;
    nbest = to_world(b, {nl[0], nl[1], nl[2]});
    return true;
}

struct M4 { double m[16]; };  // column-major
M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int c = 0; c < 4; c++) for (int rr = 0; rr < 4; rr++) {
        double s = 0;
        for (int k = 0; k < 4; k++) s += a.m[k * 4 + rr] * b.m[c * 4 + k];
        r.m[c * 4 + rr] = s;
    }
    return r;
}
bool invert(const M4& a, M4& out) {  // Gauss-Jordan, double
    double m[4][8];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { m[r][c] = a.m[c * 4 + r]; m[r][4 + c] = (r == c); }
    for (int i = 0; i < 4; i++) {
        int p = i;
        for (int r = i + 1; r < 4; r++) if (std::fabs(m[r][i]) > std::fabs(m[p][i])) p = r;
        if (std::fabs(m[p][i]) < 1e-300) return false;
        for (int c = 0; c < 8; c++) std::swap(m[i][c], m[p][c]);
        double inv = 1.0 / m[i][i];
        for (int c = 0; c < 8; c++) m[i][c] *= inv;
        for (int r = 0; r < 4; r++) if (r != i) { double f = m[r][i]; for (int c = 0; c < 8; c++) m[r][c] -= f * m[i][c]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out.m[c * 4 + r] = m[r][4 + c];
    return true;
}
void xform(const M4& a, const double v[4], double o[4]) {
    for (int r = 0; r < 4; r++) o[r] = a.m[r] * v[0] + a.m[4 + r] * v[1] + a.m[8 + r] * v[2] + a.m[12 + r] * v[3];
}

uint32_t pack_rgb10(d3 n) {  // RGB10_A2 unorm of n*0.5+0.5 (RasterizationStage.cpp:143, model.frag:92-93)
    auto q = [](double v) { double u = std::min(1.0, std::max(0.0, v * 0.5 + 0.5)); return (uint32_t)std::lrint(u * 1023.0); };
    return q(n.x) | (q(n.y) << 10) | (q(n.z) << 20);
}

}  // namespace

// ---------------------------------------------------------------------------------------
SCN_API void* mlgi_scene_create(uint64_t seed, float cell_size, int clutter_boxes) {
    Scene* sc = new Scene();
    // room shell 30 x 12 x 18 (like Sponza at vertexScale 0.01, ModelLoadingStage.cpp:180)
    add_box(*sc, seed, {0, 6, 0}, {15, 6, 9}, 0.0, true);
    // two rows of columns and a gallery slab on each side
    for (int i = 0; i < 6; i++)
        for (int s = -1; s <= 1; s += 2)
            add_box(*sc, seed, {-12.5 + 5.0 * i, 3.5, 4.5 * s}, {0.4, 3.5, 0.4}, 0.0);
    for (int s = -1; s <= 1; s += 2) add_box(*sc, seed, {0, 7.2, 6.75 * s}, {15, 0.2, 2.25}, 0.0);
    // clutter in the nave: rotated boxes on the floor
    uint64_t st = seed ^ 0xC0FFEEull;
    for (int i = 0; i < clutter_boxes; i++) {
        double hx = 0.3 + 0.7 * urand(st), hy = 0.25 + 0.9 * urand(st), hz = 0.3 + 0.7 * urand(st);
        double cx = -11.0 + 22.0 * urand(st), cz = -3.0 + 6.0 * urand(st), yaw = 3.14159265358979 * urand(st);
        add_box(*sc, seed, {cx, hy, cz}, {hx, hy, hz}, yaw);
    }
    // two hanging banners
    add_box(*sc, seed, {-4, 8.5, 0.0}, {0.05, 1.5, 1.2}, 0.3);
    add_box(*sc, seed, {5, 8.8, 0.5}, {0.05, 1.2, 1.0}, -0.4);
    tessellate(*sc, cell_size);
    return sc;
}
SCN_API void mlgi_scene_destroy(void* s) { delete (Scene*)s; }
SCN_API int mlgi_scene_num_vertices(void* s) { return (int)(((Scene*)s)->verts.size() / 3); }
SCN_API int mlgi_scene_num_triangles(void* s) { return (int)(((Scene*)s)->indices.size() / 3); }
SCN_API int mlgi_scene_num_boxes(void* s) { return (int)((Scene*)s)->boxes.size(); }
SCN_API void mlgi_scene_get_mesh(void* s, float* verts, float* normals, uint32_t* indices, uint32_t* prim_id) {
    Scene* sc = (Scene*)s;
    memcpy(verts, sc->verts.data(), sc->verts.size() * 4);
    memcpy(normals, sc->normals.data(), sc->normals.size() * 4);
    memcpy(indices, sc->indices.data(), sc->indices.size() * 4);
    memcpy(prim_id, sc->prim_id.data(), sc->prim_id.size() * 4);
}

// glm-style column-major fp32 matrices --------------------------------------------------
SCN_API void mlgi_scene_look_at(const float* eye, const float* center, const float* up, float* m) {
    d3 e{eye[0], eye[1], eye[2]}, c{center[0], center[1], center[2]}, u0{up[0], up[1], up[2]};
    d3 f = normalize(c - e), s = normalize(cross(f, u0)), u = cross(s, f);
    const double r[16] = {s.x, u.x, -f.x, 0, s.y, u.y, -f.y, 0, s.z, u.z, -f.z, 0, -dot(s, e), -dot(u, e), dot(f, e), 1};
    for (int i = 0; i < 16; i++) m[i] = (float)r[i];
}
SCN_API void mlgi_scene_perspective(float fovy, float aspect, float zn, float zf, float* m) {
    double t = std::tan((double)fovy / 2);
    double r[16] = {0};
    r[0] = 1.0 / (aspect * t); r[5] = 1.0 / t; r[10] = -((double)zf + zn) / ((double)zf - zn); r[11] = -1.0;
    r[14] = -(2.0 * zf * zn) / ((double)zf - zn);
    for (int i = 0; i < 16; i++) m[i] = (float)r[i];
}
SCN_API void mlgi_scene_ortho(float l, float r_, float b, float t, float zn, float zf, float* m) {
    double r[16] = {0};
    r[0] = 2.0 / ((double)r_ - l); r[5] = 2.0 / ((double)t - b); r[10] = -2.0 / ((double)zf - zn);
    r[12] = -((double)r_ + l) / ((double)r_ - l); r[13] = -((double)t + b) / ((double)t - b); r[14] = -((double)zf + zn) / ((double)zf - zn); r[15] = 1.0;
    for (int i = 0; i < 16; i++) m[i] = (float)r[i];
}
// GIStage.cpp:357-361: sun direction from sunCyclePosition; eye = center - dir*4
SCN_API void mlgi_scene_sun_eye(float sunCyclePosition, const float* center, float* eye) {
    const double degreeSpan = 80.0;
    double m = std::fmod((double)sunCyclePosition, degreeSpan * 2);
    if (m < 0) m += degreeSpan * 2;
    double degree = std::fabs(m - degreeSpan) + (180.0 - degreeSpan) / 2;
    double rad = degree * 3.14159265358979323846 / 180.0;
    d3 dir{0.0, -std::sin(rad), std::cos(rad)};
    eye[0] = (float)(center[0] - dir.x * 4.0); eye[1] = (float)(center[1] - dir.y * 4.0); eye[2] = (float)(center[2] - dir.z * 4.0);
}

// Renders depth (window space [0,1], fp32), RGB10_A2 face normal and RGBA8 diffuse through
// the given camera by casting one ray per pixel centre. Any output pointer may be null.
// Returns the number of pixels that hit nothing (must be 0 for a closed scene).
SCN_API long long mlgi_scene_render(void* s, const float* view, const float* proj, int W, int H,
                                    float* depth, uint32_t* normal10, uint32_t* diffuse8) {
    Scene* sc = (Scene*)s;
    M4 V, P;
    for (int i = 0; i < 16; i++) { V.m[i] = view[i]; P.m[i] = proj[i]; }
    M4 PV = mul(P, V), inv;
    if (!invert(PV, inv)) return -1;
    long long misses = 0;
#ifdef _OPENMP
    omp_set_num_threads(omp_get_num_procs());   // torchrun exports OMP_NUM_THREADS=1; this fixture is untimed
#endif
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : misses)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double nx = 2.0 * (x + 0.5) / W - 1.0, ny = 2.0 * (y + 0.5) / H - 1.0;
            double a[4] = {nx, ny, -1.0, 1.0}, b[4] = {nx, ny, 1.0, 1.0}, pa[4], pb[4];
            xform(inv, a, pa); xform(inv, b, pb);
            d3 o{pa[0] / pa[3], pa[1] / pa[3], pa[2] / pa[3]}, e{pb[0] / pb[3], pb[1] / pb[3], pb[2] / pb[3]};
            d3 d = e - o;  // t in [0,1] spans near..far plane
            double tbest = 1.0; d3 nbest{0, 0, 0}; int hit = -1;
            for (size_t bi = 0; bi < sc->boxes.size(); bi++)
                if (hit_box(sc->boxes[bi], o, d, 0.0, tbest, nbest)) hit = (int)bi;
            size_t idx = (size_t)y * W + x;
            if (hit < 0) {
                misses++;
                if (depth) depth[idx] = 1.0f;
                if (normal10) normal10[idx] = 0;
                if (diffuse8) diffuse8[idx] = 0;
                continue;
            }
            d3 p = o + d * tbest;
            double pw[4] = {p.x, p.y, p.z, 1.0}, clip[4];
            xform(PV, pw, clip);
            double zwin = 0.5 * (clip[2] / clip[3]) + 0.5;
            if (depth) depth[idx] = (float)std::min(std::max(zwin, 0.0), 0.99999);
            if (normal10) normal10[idx] = pack_rgb10(nbest);
            if (diffuse8) diffuse8[idx] = sc->boxes[hit].rgba;
        }
    return misses;
}

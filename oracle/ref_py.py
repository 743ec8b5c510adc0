"""ctypes wrapper of oracle/_ref/libref_glsl.so: the reference's own GLSL sources compiled as C++ (oracle/refglsl/).
TEST INFRASTRUCTURE ONLY -- used by tests/test_reference_glsl.py and tests/golden/make_ref_golden.py to pin the oracle."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_glsl.so")
_lib = None
P = C.c_void_p


_libs = {}
_vpls = 1024          # which build the wrappers below call into (the reference's own constant is 1024)


def lib_path(vpls=1024):
    return LIB_PATH if vpls == 1024 else os.path.join(HERE, "_ref", "libref_glsl_v%d.so" % vpls)


def available(vpls=1024):
    return os.path.exists(lib_path(vpls))


def use_vpls(vpls):
    """select the build whose shaders were compiled with totalVplCount = vpls (build_ref.py --vpls)"""
    global _vpls
    _vpls = vpls


def load():
    if _vpls not in _libs:
        _libs[_vpls] = C.CDLL(lib_path(_vpls))
    return _libs[_vpls]


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def vpl_process(rsm_diffuse, rsm_normal, rsm_depth, N, inv_biased_light_vp, light_intensity, shuffle, table):
    h, w = rsm_depth.shape
    vpl = np.zeros((N, 12), np.float32); packed = np.zeros((N, 4), np.float32)
    m = _f32(inv_biased_light_vp)
    load().ref_vplproc_run(_p(rsm_diffuse), _p(rsm_normal), _p(rsm_depth), w, h, N, _p(m), C.c_float(light_intensity), int(shuffle),
                           _p(np.ascontiguousarray(table, np.int32)), _p(vpl), _p(packed))
    return vpl, packed


def ism_splat(points_grouped, bucket_start, packed_vpl, N, z_far, A, vpl_start=0, vpl_end=None, scale_isms=False, points_only=False):
    """points in the reference's pointBuffer layout: bucket b at [b*stride, b*stride + count_b), stride = imageSize /
    sampledVplCount (ism.comp:81)"""
    counts = np.diff(bucket_start.astype(np.int64)).astype(np.uint32)
    stride = int(max(1, counts.max()))
    sampled = (N if vpl_end is None else vpl_end) - vpl_start if points_only else N
    cap = stride * sampled
    buf = np.zeros((cap, 4), np.float32)
    for b in range(sampled):
        n = int(counts[b])
        if n:
            buf[b * stride:b * stride + n] = points_grouped[int(bucket_start[b]):int(bucket_start[b]) + n]
    atlas = np.full((A, A), 0xFFFFFFFF, np.uint32)
    pv = _f32(packed_vpl)
    load().ref_splat_run2(_p(pv), N, _p(buf), cap, _p(counts), C.c_float(z_far), _p(atlas), A, vpl_start, N if vpl_end is None else vpl_end,
                          int(scale_isms), int(points_only))
    return atlas


def ism_pullpush(atlas, z_far, ism_pixel_size=64):
    """the 6 pull and 6 push passes of ImperfectShadowmap::pullpush (ImperfectShadowmap.cpp:121-182)"""
    lib = load()
    atlas = np.ascontiguousarray(atlas, np.uint32)
    A = atlas.shape[0]
    pull = {}
    out = np.zeros((A // 2, A // 2, 4), np.float32)
    lib.ref_pull0_run(_p(atlas), A, _p(out), 1, C.c_float(z_far), ism_pixel_size); pull[1] = out
    for l in range(2, 7):
        out = np.zeros((A >> l, A >> l, 4), np.float32)
        lib.ref_pull_run(_p(pull[l - 1]), A, _p(out), l, C.c_float(z_far), ism_pixel_size); pull[l] = out
    push = {6: pull[6]}
    for l in range(5, 0, -1):
        out = np.zeros((A >> l, A >> l, 4), np.float32)
        lib.ref_push_run(_p(pull[l]), _p(push[l + 1]), A, _p(out), l); push[l] = out
    r16 = np.zeros((A, A), np.uint16)
    lib.ref_push0_run(_p(atlas), _p(push[1]), A, _p(r16), 0)
    return r16


def clustering(depth, proj, z_far):
    H, W = depth.shape
    X, Y = (W + 127) // 128, (H + 127) // 128
    cap = X * Y * 16
    compact = np.zeros(cap, np.uint32); ids = np.full((16, Y, X), 0xFFFF, np.uint16); n = np.zeros(1, np.uint32)
    load().ref_clustering_run(_p(_f32(depth)), W, H, _p(_f32(proj)), C.c_float(z_far), _p(compact), cap, _p(ids), _p(n))
    return compact[:int(n[0])], ids, int(n[0])


def light_lists(packed_vpl, N, compact_ids, W, H, proj, inv_proj_view, z_far, rows=None):
    """returns the reference's lightLists image [slot][cluster] (r16ui; x = list id, y = class*64 + slot)"""
    n = len(compact_ids)
    rows = rows or N
    img = np.zeros((rows, max(n, 1)), np.uint16)
    load().ref_lightlists_run(_p(_f32(packed_vpl)), N, _p(np.ascontiguousarray(compact_ids, np.uint32)), n, W, H, _p(_f32(proj)),
                              _p(_f32(inv_proj_view)), C.c_float(z_far), _p(img), img.shape[1], img.shape[0])
    return img


def final_gather(depth, normal, vpl, N, ism, list_ids, lists_img, proj, view_proj_inverted, z_far, z_near, gi_intensity_factor=3000.0,
                 clamp=0.001, vpl_start=0, vpl_end=None, shadows=True, scale_isms=False):
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.float32)
    A = ism.shape[0]
    Y, X = list_ids.shape[1], list_ids.shape[2]
    fn = (load().ref_gather_scaled_run if scale_isms else load().ref_gather_run) if shadows else load().ref_gather_noshadow_run
    fn(_p(_f32(depth)), _p(np.ascontiguousarray(normal, np.uint32)), W, H, _p(_f32(vpl)), N, _p(np.ascontiguousarray(ism, np.uint16)), A,
       _p(np.ascontiguousarray(list_ids, np.uint16)), X, Y, _p(np.ascontiguousarray(lists_img, np.uint16)), lists_img.shape[1], lists_img.shape[0],
       _p(_f32(proj)), _p(_f32(view_proj_inverted)), C.c_float(z_far), C.c_float(z_near), C.c_float(gi_intensity_factor), C.c_float(clamp),
       vpl_start, N if vpl_end is None else vpl_end, _p(out))
    return out[:, :, :3]


def gi_blur(gi_rgb, depth, normal, proj):
    H, W = depth.shape
    rgba = np.zeros((H, W, 4), np.float32); rgba[:, :, :3] = gi_rgb
    tmp = np.zeros((H, W, 3), np.float32); out = np.zeros((H, W, 3), np.float32)
    load().ref_blur_h_run(_p(rgba), _p(_f32(depth)), _p(np.ascontiguousarray(normal, np.uint32)), W, H, _p(_f32(proj)), _p(tmp))
    rgba[:, :, :3] = tmp
    load().ref_blur_v_run(_p(rgba), _p(_f32(depth)), _p(np.ascontiguousarray(normal, np.uint32)), W, H, _p(_f32(proj)), _p(out))
    return out


def ism_points(vertices, normals, indices, prim_id, tess_level_factor, N, max_triangles=None, vpl_start=0, vpl_end=None, scale_isms=False,
               points_only=False):
    """programmable stages of the point pipeline (ism.tesc -> [tessellator] -> ism.tese -> ism.geom, usePushPull path).
    The fixed-function tessellator is replaced by the deterministic n^2 split of DESIGN.md section 3.4 (n = the
    equal_spacing rounding ceil(clamp(inner level, 1, 64)) of the GL spec). Returns (points[P,4], bucket[P]) grouped by
    bucket in pointBuffer order."""
    lib = load()
    assert N == _vpls, "ism.geom was compiled for totalVplCount = %d" % _vpls
    ntri = indices.shape[0] if max_triangles is None else min(indices.shape[0], max_triangles)
    V, Tc, Nr, Pid = [], [], [], []
    lv = np.zeros(4, np.float32); pos = np.zeros(3, np.float32)
    for t in range(ntri):
        tri = _f32(vertices[indices[t]])                      # 3 x 3
        n0 = _f32(normals[indices[t][0]])                     # ism.tese:22: vertex 0's normal
        lib.ref_tesc_run(_p(tri), C.c_float(tess_level_factor), _p(lv))
        n = int(np.ceil(np.float32(min(max(lv[0], np.float32(1.0)), np.float32(64.0)))))
        fn = np.float32(n)
        for j in range(n):
            for i in range(n - j):
                for down in (0, 1):
                    if down and i + j > n - 2:
                        continue
                    g = ((i, j), (i + 1, j), (i, j + 1)) if not down else ((i + 1, j), (i + 1, j + 1), (i, j + 1))
                    vs, tcs = [], []
                    for gi, gj in g:
                        tc = np.array([np.float32(n - gi - gj) / fn, np.float32(gi) / fn, np.float32(gj) / fn], np.float32)
                        lib.ref_tese_run(_p(tri), _p(n0), _p(tc), _p(pos))
                        vs.append(pos.copy()); tcs.append(tc)
                    V.append(vs); Tc.append(tcs); Nr.append(n0); Pid.append(int(prim_id[t]))
    T = len(V)
    V = _f32(np.array(V)); Tc = _f32(np.array(Tc)); Nr = _f32(np.array(Nr)); Pid = np.array(Pid, np.int32)
    stride = 4096
    buf = np.zeros((N * stride, 4), np.float32); counters = np.zeros(max(N, 1024), np.uint32)
    sampled = (N if vpl_end is None else vpl_end) - vpl_start if points_only else N
    # capacity = sampledVplCount * stride keeps ism.geom:67's `imageSize / sampledVplCount` equal to stride
    lib.ref_geom_run2(_p(V), _p(Tc), _p(Nr), _p(Pid), T, N, _p(buf), sampled * stride, _p(counters), vpl_start, N if vpl_end is None else vpl_end,
                      int(scale_isms), int(points_only))
    assert counters.max() <= stride
    pts = np.concatenate([buf[b * stride:b * stride + int(counters[b])] for b in range(N)])
    bkt = np.repeat(np.arange(N, dtype=np.int32), counters[:N].astype(np.int64))
    return pts, bkt


def composite_ambient(gi_rgb, diffuse_rgba8, occlusion=None, exposure=1.0):
    """deferredshading.frag with the direct light off: float sRGB [H,W,3]"""
    H, W = diffuse_rgba8.shape
    rgba = np.zeros((H, W, 4), np.float32); rgba[:, :, :3] = gi_rgb
    out = np.zeros((H, W, 3), np.float32)
    occ = None if occlusion is None else _f32(occlusion)
    load().ref_composite_run(_p(rgba), _p(np.ascontiguousarray(diffuse_rgba8, np.uint32)), _p(occ), W, H, C.c_float(exposure), _p(out))
    return out


def lists_image_from_canonical(lists, num_used, L):
    """the reference's lightLists image (x = list id, y = class * L + slot) holding the given canonical lists
    [cluster][16][L+1]; a full sub-list's last entry lands on the next class's counter and is overwritten by it, as in
    the reference (the overflow race)"""
    img = np.zeros((16 * L + L + 1, max(num_used, 1)), np.uint16)
    for k in range(16):
        for cl in range(num_used):
            n = int(lists[cl, k, 0])
            img[k * L + 1:k * L + n, cl] = lists[cl, k, 1:n]
            img[k * L, cl] = n
    return img

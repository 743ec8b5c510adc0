"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product path (many-lights-gi_b200/) never imports this module.
Parity: pinned against the reference's own shader sources compiled for the CPU (oracle/refglsl, see oracle.cpp header).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

P = C.c_void_p


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("liboracle.so not found -- run `make -C oracle` (or __graft_entry__.build())")
        lib = C.CDLL(LIB_PATH)
        lib.orc_hash.argtypes = [C.c_uint32]; lib.orc_hash.restype = C.c_uint32
        lib.orc_random3.argtypes = [C.c_float] * 3; lib.orc_random3.restype = C.c_float
        lib.orc_pack3SN.argtypes = [C.c_float] * 3; lib.orc_pack3SN.restype = C.c_uint32
        lib.orc_unpack3SN.argtypes = [C.c_uint32, P]
        lib.orc_packUnorm4x8.argtypes = [C.c_float] * 4; lib.orc_packUnorm4x8.restype = C.c_uint32
        lib.orc_packHalf2x16.argtypes = [C.c_float] * 2; lib.orc_packHalf2x16.restype = C.c_uint32
        lib.orc_unpackHalf2x16.argtypes = [C.c_uint32, P]
        lib.orc_mat4_inverse.argtypes = [P, P]
        lib.orc_mat4_mul.argtypes = [P, P, P]
        lib.orc_paraboloid_project.argtypes = [P, P, C.c_float, C.c_float, C.c_int, C.c_int, P]
        lib.orc_shuffle_table.argtypes = [C.c_int, P]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        lib.orc_ism_indices_1d.argtypes = [C.c_int]; lib.orc_ism_indices_1d.restype = C.c_int
        lib.orc_slice_boundaries.argtypes = [C.c_float, P]
        lib.orc_slice_log2.argtypes = [C.c_float, C.c_float]; lib.orc_slice_log2.restype = C.c_int
        lib.orc_slice_table.argtypes = [C.c_float, C.c_float]; lib.orc_slice_table.restype = C.c_int
        lib.orc_linear_depth.argtypes = [C.c_float, P]; lib.orc_linear_depth.restype = C.c_float
        lib.orc_vpl_process.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, P, P, C.c_float, C.c_int, P, P, P]
        lib.orc_ism_points.argtypes = [P, P, P, C.c_int, P, C.c_float, C.c_int, P, P, C.c_longlong]
        lib.orc_ism_points.restype = C.c_longlong
        lib.orc_ism_splat.argtypes = [P, P, P, C.c_int, C.c_float, P, C.c_int]
        lib.orc_ism_pullpush.argtypes = [P, C.c_int, C.c_float, P, P, P]
        lib.orc_ism_points_ex.argtypes = [P, P, P, C.c_int, P, C.c_float, C.c_int, C.c_int, P, P, C.c_longlong]
        lib.orc_ism_points_ex.restype = C.c_longlong
        lib.orc_ism_splat_ex.argtypes = [P, P, P, C.c_int, C.c_float, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_ism_pullpush_ex.argtypes = [P, C.c_int, C.c_int, C.c_float, P, P, P]
        lib.orc_final_gather_scale_isms.argtypes = [C.c_int]
        lib.orc_clustering.argtypes = [P, C.c_int, C.c_int, P, C.c_float, P, P]; lib.orc_clustering.restype = C.c_int
        lib.orc_light_lists.argtypes = [P, C.c_int, P, C.c_int, P, P, C.c_int, C.c_int, C.c_float, P]
        lib.orc_cull_tiled.argtypes = [P, P, C.c_int, C.c_int, P, C.c_int, P, P, C.c_int, P]
        lib.orc_final_gather.argtypes = [P, P, C.c_int, C.c_int, P, C.c_int, P, C.c_int, P, P, P, P, C.c_float,
                                         C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         P, P, P, C.c_float, P, P]
        lib.orc_pack_r11g11b10.argtypes = [P, C.c_longlong, C.c_int, P]
        lib.orc_unpack_r11g11b10.argtypes = [P, C.c_longlong, P]
        lib.orc_gi_blur_pass.argtypes = [P, P, P, C.c_int, C.c_int, P, C.c_int, P]
        lib.orc_composite_ambient.argtypes = [P, P, P, C.c_int, C.c_float, P, P]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def set_reference_math(on):
    """pow(dist, 4.0) through libm (bit-comparable with oracle/refglsl) instead of two multiplies"""
    load().orc_set_reference_math(int(on))


def num_threads():
    return load().orc_num_threads()


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; take every core this process may run on"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().orc_set_num_threads(n)
    return num_threads()


def ism_indices_1d(n):
    return load().orc_ism_indices_1d(n)


def shuffle_table(n):
    out = np.empty(n, np.int32)
    load().orc_shuffle_table(n, _p(out))
    return out


def slice_boundaries(z_far):
    out = np.empty(17, np.float32)
    load().orc_slice_boundaries(z_far, _p(out))
    return out


def mat4_inverse(m):
    m = _f32(m); out = np.empty(16, np.float32)
    load().orc_mat4_inverse(_p(m), _p(out))
    return out


def mat4_mul(a, b):
    a, b = _f32(a), _f32(b); out = np.empty(16, np.float32)
    load().orc_mat4_mul(_p(a), _p(b), _p(out))
    return out


def vpl_process(rsm_diffuse, rsm_normal, rsm_depth, N, light_view, light_proj, light_intensity=5.0, shuffle=True, table=None):
    h, w = rsm_depth.shape
    if table is None:
        table = shuffle_table(N)
    vpl = np.zeros((N, 12), np.float32)
    packed = np.zeros((N, 4), np.float32)
    lv, lp = _f32(light_view), _f32(light_proj)
    load().orc_vpl_process(_p(rsm_diffuse), _p(rsm_normal), _p(rsm_depth), w, h, N, _p(lv), _p(lp), light_intensity, int(shuffle),
                           _p(table), _p(vpl), _p(packed))
    return vpl, packed


def ism_mode(N, vpl_start=0, vpl_end=None, scale_isms=False, points_only=False):
    """n3 knobs -> (sampledVplCount, ismCount) of ism.geom:42-44 / ism.comp:36-38"""
    vpl_end = N if vpl_end is None else vpl_end
    count = vpl_end - vpl_start
    return (count if points_only else N), (count if scale_isms else N)


def ism_points(vertices, normals, indices, prim_id, tess_level_factor, N, sampled=None, ism_count=None):
    """returns (points[P,4] f32, bucket[P] i32) in canonical (triangle, sub id) order. sampled / ism_count: ism_mode()"""
    lib = load()
    ntri = indices.shape[0]
    sampled = N if sampled is None else sampled
    ism_count = N if ism_count is None else ism_count
    count = lib.orc_ism_points_ex(_p(vertices), _p(normals), _p(indices), ntri, _p(prim_id), tess_level_factor, sampled, ism_count, None, None, 0)
    pts = np.empty((count, 4), np.float32)
    bkt = np.empty(count, np.int32)
    c2 = lib.orc_ism_points_ex(_p(vertices), _p(normals), _p(indices), ntri, _p(prim_id), tess_level_factor, sampled, ism_count, _p(pts), _p(bkt), count)
    assert c2 == count
    return pts, bkt


def group_by_bucket(points, bucket, N):
    """stable sort by bucket -> (points grouped, bucket_start[N+1] u32): the canonical point list"""
    order = np.argsort(bucket, kind="stable")
    counts = np.bincount(bucket, minlength=N).astype(np.uint32)
    start = np.zeros(N + 1, np.uint32)
    start[1:] = np.cumsum(counts)
    return np.ascontiguousarray(points[order]), start


def ism_splat(points_grouped, bucket_start, packed_vpl, N, z_far, vpl_start=0, vpl_end=None, scale_isms=False, points_only=False):
    S = ism_indices_1d(N)
    A = 64 * S
    atlas = np.empty((A, A), np.uint32)
    load().orc_ism_splat_ex(_p(points_grouped), _p(bucket_start), _p(packed_vpl), N, z_far, _p(atlas), 1, vpl_start,
                            N if vpl_end is None else vpl_end, int(scale_isms), int(points_only))
    return atlas


def ism_pullpush(atlas, N, z_far, ism_count=None):
    A = atlas.shape[0]
    out = np.empty((A, A), np.uint16)
    load().orc_ism_pullpush_ex(_p(atlas), N, N if ism_count is None else ism_count, z_far, _p(out), None, None)
    return out


def clustering(depth, proj, z_far):
    H, W = depth.shape
    X, Y = (W + 127) // 128, (H + 127) // 128
    compact = np.zeros(X * Y * 16, np.uint32)
    ids = np.empty((16, Y, X), np.uint16)
    pr = _f32(proj)
    n = load().orc_clustering(_p(depth), W, H, _p(pr), z_far, _p(compact), _p(ids))
    return compact[:n].copy(), ids, n


def light_lists(packed_vpl, N, compact_ids, view, proj, W, H, z_far):
    n = compact_ids.shape[0]
    L = N // 16
    lists = np.zeros((max(n, 1), 16, L + 1), np.uint16)
    v, p = _f32(view), _f32(proj)
    load().orc_light_lists(_p(packed_vpl), N, _p(compact_ids), n, _p(v), _p(p), W, H, z_far, _p(lists))
    return lists[:n]


def cull_tiled(depth, normal, vpl, N, view, proj, tile=64):
    H, W = depth.shape
    BX, BY = (W + tile - 1) // tile, (H + tile - 1) // tile
    L = N // 16
    lists = np.zeros((BX * BY, 16, L + 1), np.uint16)
    v, p = _f32(view), _f32(proj)
    load().orc_cull_tiled(_p(depth), _p(normal), W, H, _p(vpl), N, _p(v), _p(p), tile, _p(lists))
    return lists


def final_gather(depth, normal, vpl, N, ism, cull_mode, light_list_ids, lists, view, proj, z_far,
                 gi_intensity_factor=3000.0, clamp=0.001, vpl_start=0, vpl_end=None, shadows=True,
                 class_lo=0, class_hi=16, tile=64, bounds=False, tie_eps=None, rows=None, scale_isms=False):
    """returns dict(rgb[H,W,3], lo, hi, f64, listed, lit); lo/hi/f64 only with bounds=True"""
    H, W = depth.shape
    if vpl_end is None:
        vpl_end = N
    out = np.zeros((H, W, 3), np.float32)
    lo = np.zeros((H, W, 3), np.float32) if bounds else None
    hi = np.zeros((H, W, 3), np.float32) if bounds else None
    f64 = np.zeros((H, W, 3), np.float32) if bounds else None
    stats = np.zeros(2, np.int64)
    if tie_eps is None:
        tie_eps = 64 * ism_indices_1d(N) * 1e-6   # a few fp32 ulps of the atlas coordinate, in texels
    r0, r1 = (0, H) if rows is None else rows
    v, p = _f32(view), _f32(proj)
    load().orc_final_gather_scale_isms(int(scale_isms))          # the SCALE_ISMS define (GIStage.cpp:430)
    try:
        load().orc_final_gather(_p(depth), _p(normal), W, H, _p(vpl), N, _p(ism), cull_mode, _p(light_list_ids), _p(lists),
                                _p(v), _p(p), z_far, gi_intensity_factor, clamp, vpl_start, vpl_end, int(shadows),
                                class_lo, class_hi, tile, r0, r1, _p(out), _p(lo), _p(hi), tie_eps, _p(stats), _p(f64))
    finally:
        load().orc_final_gather_scale_isms(0)
    return dict(rgb=out, lo=lo, hi=hi, f64=f64, listed=int(stats[0]), lit=int(stats[1]))


def gi_blur(gi_rgb, depth, normal, proj):
    H, W = depth.shape
    tmp = np.empty((H, W, 3), np.float32)
    out = np.empty((H, W, 3), np.float32)
    pr = _f32(proj)
    g = np.ascontiguousarray(gi_rgb, dtype=np.float32)
    load().orc_gi_blur_pass(_p(g), _p(depth), _p(normal), W, H, _p(pr), 0, _p(tmp))
    load().orc_gi_blur_pass(_p(tmp), _p(depth), _p(normal), W, H, _p(pr), 1, _p(out))
    return out


def composite_ambient(gi_rgb, diffuse_rgba8, occlusion=None, exposure=1.0):
    """returns (float sRGB [H,W,3], RGBA8 [H,W] uint32)"""
    H, W = diffuse_rgba8.shape
    g = _f32(gi_rgb); d = np.ascontiguousarray(diffuse_rgba8, dtype=np.uint32)
    occ = None if occlusion is None else _f32(occlusion)
    out_f = np.empty((H, W, 3), np.float32); out8 = np.empty((H, W), np.uint32)
    load().orc_composite_ambient(_p(g), _p(d), _p(occ) if occ is not None else None, W * H, exposure, _p(out_f), _p(out8))
    return out_f, out8


def pack_r11g11b10(rgb):
    """rgb [...,3 or 4] fp32 -> uint32 GL_R11F_G11F_B10F"""
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    n = a.size // a.shape[-1]
    out = np.empty(a.shape[:-1], np.uint32)
    load().orc_pack_r11g11b10(_p(a), n, a.shape[-1], _p(out))
    return out


def unpack_r11g11b10(packed):
    a = np.ascontiguousarray(packed, dtype=np.uint32)
    out = np.empty(a.shape + (3,), np.float32)
    load().orc_unpack_r11g11b10(_p(a), a.size, _p(out))
    return out

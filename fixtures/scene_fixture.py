"""Synthetic-input fixture binding (fixtures/libmlgi_scene.so, fixtures/scene.cpp): procedural atrium,
its triangle mesh, cameras and ray-cast G-buffers. Untimed; stands in for the reference's
out-of-scope rasteriser (RasterizationStage) and asset loader (ModelLoadingStage)."""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmlgi_scene.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmlgi_scene.so not found at %s -- run __graft_entry__.build()" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        P = C.c_void_p
        lib.mlgi_scene_create.argtypes = [C.c_uint64, C.c_float, C.c_int]
        lib.mlgi_scene_create.restype = P
        lib.mlgi_scene_destroy.argtypes = [P]
        lib.mlgi_scene_destroy.restype = None
        for f in ("mlgi_scene_num_vertices", "mlgi_scene_num_triangles", "mlgi_scene_num_boxes"):
            getattr(lib, f).argtypes = [P]
            getattr(lib, f).restype = C.c_int
        lib.mlgi_scene_get_mesh.argtypes = [P, P, P, P, P]
        lib.mlgi_scene_get_mesh.restype = None
        lib.mlgi_scene_look_at.argtypes = [P, P, P, P]
        lib.mlgi_scene_perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, P]
        lib.mlgi_scene_ortho.argtypes = [C.c_float] * 6 + [P]
        lib.mlgi_scene_sun_eye.argtypes = [C.c_float, P, P]
        lib.mlgi_scene_render.argtypes = [P, P, P, C.c_int, C.c_int, P, P, P]
        lib.mlgi_scene_render.restype = C.c_longlong
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Scene:
    """Closed procedural atrium."""

    def __init__(self, seed=0x5EED, cell_size=0.155, clutter_boxes=24):
        self.lib = load()
        self.h = self.lib.mlgi_scene_create(seed, cell_size, clutter_boxes)
        nv = self.lib.mlgi_scene_num_vertices(self.h)
        nt = self.lib.mlgi_scene_num_triangles(self.h)
        self.vertices = np.empty((nv, 3), np.float32)
        self.normals = np.empty((nv, 3), np.float32)
        self.indices = np.empty((nt, 3), np.uint32)
        self.prim_id = np.empty(nt, np.uint32)
        self.lib.mlgi_scene_get_mesh(self.h, self.vertices.ctypes.data, self.normals.ctypes.data, self.indices.ctypes.data,
                                     self.prim_id.ctypes.data)
        self.num_vertices, self.num_triangles = nv, nt

    def __del__(self):
        try:
            if self.h:
                self.lib.mlgi_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def look_at(self, eye, center, up=(0.0, 1.0, 0.0)):
        m = np.empty(16, np.float32)
        e, c, u = _f32(eye), _f32(center), _f32(up)
        self.lib.mlgi_scene_look_at(e.ctypes.data, c.ctypes.data, u.ctypes.data, m.ctypes.data)
        return m

    def perspective(self, fovy, aspect, zn, zf):
        m = np.empty(16, np.float32)
        self.lib.mlgi_scene_perspective(fovy, aspect, zn, zf, m.ctypes.data)
        return m

    def ortho(self, l, r, b, t, zn, zf):
        m = np.empty(16, np.float32)
        self.lib.mlgi_scene_ortho(l, r, b, t, zn, zf, m.ctypes.data)
        return m

    def sun_eye(self, sun_cycle_position, center):
        c = _f32(center)
        e = np.empty(3, np.float32)
        self.lib.mlgi_scene_sun_eye(sun_cycle_position, c.ctypes.data, e.ctypes.data)
        return e

    def render(self, view, proj, W, H, want_diffuse=False):
        depth = np.empty((H, W), np.float32)
        normal = np.empty((H, W), np.uint32)
        diffuse = np.empty((H, W), np.uint32) if want_diffuse else None
        v, p = _f32(view), _f32(proj)
        misses = self.lib.mlgi_scene_render(self.h, v.ctypes.data, p.ctypes.data, W, H, depth.ctypes.data, normal.ctypes.data,
                                            diffuse.ctypes.data if want_diffuse else None)
        if misses != 0:
            raise RuntimeError("scene not closed through this camera: %d pixels hit nothing" % misses)
        return depth, normal, diffuse


class Workload:
    """Everything one GI frame consumes, as numpy arrays (same bits go to the oracle and to CUDA).

    Camera: CrytekSponza preset eye/center (ModelLoadingStage.cpp:180), fovy 60 deg, zNear/zFar 0.05/50.
    Light : orthographic RSM, height 5 -> width 5*rsm_w/rsm_h (GIStage.cpp:236-238), eye = center - dir(sunCyclePosition)*4
            (GIStage.cpp:357-361), same zNear/zFar (GIStage.cpp:240-241).
    """

    def __init__(self, width, height, vpl_count, rsm_w=1024, rsm_h=256, cell_size=0.155, seed=0x5EED,
                 sun_cycle_position=266.0, clutter_boxes=24, scene=None, eye=(-13.0, 2.5, -0.23), center=(0.009, -0.019, -0.021)):
        self.W, self.H, self.N = width, height, vpl_count
        self.rsm_w, self.rsm_h = rsm_w, rsm_h
        self.z_near, self.z_far = 0.05, 50.0
        self.scene = scene if scene is not None else Scene(seed, cell_size, clutter_boxes)
        sc = self.scene
        self.cam_view = sc.look_at(eye, center)
        self.cam_proj = sc.perspective(math.radians(60.0), width / float(height), self.z_near, self.z_far)
        self.depth, self.normal, self.diffuse = sc.render(self.cam_view, self.cam_proj, width, height, want_diffuse=True)
        self.set_sun(sun_cycle_position)

    def set_sun(self, sun_cycle_position):
        sc = self.scene
        light_center = (0.0, 7.5, 0.0)
        eye = sc.sun_eye(sun_cycle_position, light_center)
        self.light_view = sc.look_at(eye, light_center)
        half_h = 2.5
        half_w = half_h * self.rsm_w / float(self.rsm_h)
        self.light_proj = sc.ortho(-half_w, half_w, -half_h, half_h, self.z_near, self.z_far)
        self.rsm_depth, self.rsm_normal, self.rsm_diffuse = sc.render(self.light_view, self.light_proj, self.rsm_w, self.rsm_h,
                                                                       want_diffuse=True)

"""ctypes binding of libmlgi.so (the C ABI in include/mlgi.h) for tests and bench.py.

This is plumbing only: device memory comes from torch tensors whose data_ptr() is handed to
the C ABI. The product is the CUDA library; there is NO CPU fallback -- if libmlgi.so is
missing or no CUDA device is present, every entry point fails loudly.

Reference interface mirrored by the C ABI: source/mfs-painters/multiframepainter/
{VPLProcessor,ImperfectShadowmap,ClusteredShading,GIStage}.h (see include/mlgi.h).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("MLGI_LIB_PATH") or os.path.join(PKG_ROOT, "lib", "libmlgi.so")   # override: A/B builds of the CUDA library

CULL_CLUSTERED, CULL_TILED, CULL_NONE = 0, 1, 2

(BUF_VPL, BUF_PACKED_VPL, BUF_SHUFFLE, BUF_POINTS, BUF_BUCKET_START, BUF_ATLAS, BUF_ISM,
 BUF_COMPACT_CLUSTER_IDS, BUF_LIGHT_LIST_IDS, BUF_LIGHT_LISTS, BUF_TILE_LISTS, BUF_NUM_USED_CLUSTERS,
 BUF_GI, BUF_GI_PLANAR, BUF_GI_BLUR, BUF_GI_PACKED, BUF_COUNT) = range(17)

E_INVALID, E_CUDA, E_NO_DEVICE, E_UNSUPPORTED, E_CAPACITY, E_ORDER = -1, -2, -3, -4, -5, -6

TILE = 64  # tiled-culling tile edge (csrc/common.cuh kTile)


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("vpl_count", C.c_int),
                ("rsm_width", C.c_int), ("rsm_height", C.c_int), ("max_triangles", C.c_int), ("max_points", C.c_int),
                ("class_lo", C.c_int), ("class_hi", C.c_int), ("enable_timers", C.c_int), ("gi_packed", C.c_int)]


class Params(C.Structure):
    _fields_ = [("light_intensity", C.c_float), ("gi_intensity_factor", C.c_float), ("vpl_clamping_value", C.c_float),
                ("vpl_start_index", C.c_int), ("vpl_end_index", C.c_int), ("scale_isms", C.c_int),
                ("points_only_into_scaled_isms", C.c_int), ("tess_level_factor", C.c_float), ("use_push_pull", C.c_int),
                ("enable_shadowing", C.c_int), ("show_vpl_positions", C.c_int), ("use_interleaving", C.c_int),
                ("shuffle_lights", C.c_int), ("z_near", C.c_float), ("z_far", C.c_float), ("cull_mode", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16)]

    @staticmethod
    def make(view, proj):
        cam = Camera()
        for i in range(16):
            cam.view[i] = float(view[i])
            cam.proj[i] = float(proj[i])
        return cam


class Rsm(C.Structure):
    _fields_ = [("diffuse", C.c_void_p), ("face_normal", C.c_void_p), ("depth", C.c_void_p)]


class GBuffer(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("face_normal", C.c_void_p)]


class Mesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("normals", C.c_void_p), ("indices", C.c_void_p), ("prim_id", C.c_void_p),
                ("num_vertices", C.c_int), ("num_triangles", C.c_int)]


class FrameStats(C.Structure):
    _fields_ = [("host_enqueue_us", C.c_double), ("graph_replays", C.c_ulonglong), ("graph_instantiations", C.c_ulonglong),
                ("graph_enabled", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("ms", C.c_float), ("launches", C.c_int)]


EXPORTS = [
    "mlgi_version", "mlgi_default_params", "mlgi_create", "mlgi_destroy", "mlgi_last_error",
    "mlgi_vpl_process", "mlgi_ism_points", "mlgi_ism_set_points", "mlgi_ism_splat", "mlgi_ism_pullpush",
    "mlgi_cull_clustered", "mlgi_cull_tiled", "mlgi_final_gather", "mlgi_gi_blur", "mlgi_composite_ambient", "mlgi_frame", "mlgi_frame_host", "mlgi_frame_host_packed",
    "mlgi_get_buffer", "mlgi_read_buffer", "mlgi_read_buffer_async", "mlgi_get_counts", "mlgi_gather_stats", "mlgi_measure_fp32_peak_tflops", "mlgi_debug_sqrt_thresholds", "mlgi_bind_planar_output", "mlgi_planar_to_raster", "mlgi_exchange_create", "mlgi_exchange_connect", "mlgi_exchange_connect_local", "mlgi_exchange_finish", "mlgi_exchange_status",
    "mlgi_frame_sharded", "mlgi_frame_sharded_host", "mlgi_frame_prepare", "mlgi_sharded_rows", "mlgi_set_graph", "mlgi_get_frame_stats", "mlgi_get_timings", "mlgi_set_timers", "mlgi_launch_count",
]

_lib = None


class MlgiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mlgi error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load libmlgi.so. Raises (never falls back) when the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmlgi.so not found at %s -- run __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.mlgi_version.restype = C.c_int
    lib.mlgi_default_params.argtypes = [C.POINTER(Params), C.c_int]
    lib.mlgi_default_params.restype = None
    lib.mlgi_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    lib.mlgi_destroy.argtypes = [P]
    lib.mlgi_destroy.restype = None
    lib.mlgi_last_error.argtypes = [P]
    lib.mlgi_last_error.restype = C.c_char_p
    lib.mlgi_vpl_process.argtypes = [P, C.POINTER(Rsm), C.POINTER(Camera), C.POINTER(Params), P]
    lib.mlgi_ism_points.argtypes = [P, C.POINTER(Mesh), C.POINTER(Params), P]
    lib.mlgi_ism_set_points.argtypes = [P, P, P, C.c_int, P]
    lib.mlgi_ism_splat.argtypes = [P, C.POINTER(Params), P]
    lib.mlgi_ism_pullpush.argtypes = [P, P, C.POINTER(Params), P]
    lib.mlgi_cull_clustered.argtypes = [P, C.POINTER(GBuffer), C.POINTER(Camera), C.POINTER(Params), P]
    lib.mlgi_cull_tiled.argtypes = [P, C.POINTER(GBuffer), C.POINTER(Camera), C.POINTER(Params), P]
    lib.mlgi_final_gather.argtypes = [P, C.POINTER(GBuffer), C.POINTER(Camera), P, C.POINTER(Params), P]
    lib.mlgi_gi_blur.argtypes = [P, C.POINTER(GBuffer), C.POINTER(Camera), P]
    lib.mlgi_composite_ambient.argtypes = [P, P, P, C.c_float, C.c_int, P, P]
    lib.mlgi_frame.argtypes = [P, C.POINTER(Rsm), C.POINTER(Camera), C.POINTER(Mesh), C.POINTER(GBuffer), C.POINTER(Camera),
                               C.POINTER(Params), P]
    lib.mlgi_frame_host.argtypes = [P, C.POINTER(Rsm), C.POINTER(Camera), C.POINTER(Mesh), C.POINTER(GBuffer), C.POINTER(Camera),
                                    C.POINTER(Params), P, P]
    lib.mlgi_frame_host_packed.argtypes = lib.mlgi_frame_host.argtypes
    lib.mlgi_get_buffer.argtypes = [P, C.c_int, C.POINTER(P), C.POINTER(C.c_size_t)]
    lib.mlgi_read_buffer.argtypes = [P, C.c_int, P, C.c_size_t, C.c_size_t]
    lib.mlgi_read_buffer_async.argtypes = [P, C.c_int, P, C.c_size_t, C.c_size_t, P]
    lib.mlgi_get_counts.argtypes = [P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mlgi_gather_stats.argtypes = [P, C.POINTER(GBuffer), C.POINTER(Camera), P, C.POINTER(Params),
                                      C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.mlgi_measure_fp32_peak_tflops.argtypes = [P]
    lib.mlgi_debug_sqrt_thresholds.argtypes = [P, P, C.c_int, P, P]
    lib.mlgi_measure_fp32_peak_tflops.restype = C.c_float
    lib.mlgi_bind_planar_output.argtypes = [P, P]
    lib.mlgi_planar_to_raster.argtypes = [P, P, P, P]
    lib.mlgi_exchange_create.argtypes = [P, C.c_int, C.c_int, P]
    lib.mlgi_exchange_connect.argtypes = [P, P]
    lib.mlgi_exchange_connect_local.argtypes = [P, C.POINTER(P)]
    lib.mlgi_frame_sharded.argtypes = lib.mlgi_frame.argtypes
    lib.mlgi_frame_prepare.argtypes = lib.mlgi_frame.argtypes[:-1] + [C.c_int, P]
    lib.mlgi_frame_sharded_host.argtypes = lib.mlgi_frame_host.argtypes
    lib.mlgi_sharded_rows.argtypes = [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mlgi_set_graph.argtypes = [P, C.c_int]
    lib.mlgi_get_frame_stats.argtypes = [P, C.POINTER(FrameStats)]
    lib.mlgi_exchange_finish.argtypes = [P, P]
    lib.mlgi_exchange_status.argtypes = [P]
    lib.mlgi_get_timings.argtypes = [P, C.POINTER(Timing), C.c_int]
    lib.mlgi_set_timers.argtypes = [P, C.c_int]
    lib.mlgi_launch_count.argtypes = [P]
    lib.mlgi_launch_count.restype = C.c_ulonglong
    _lib = lib
    return lib


def default_params(vpl_count):
    p = Params()
    load().mlgi_default_params(C.byref(p), vpl_count)
    return p


def _ptr(t):
    """device/host pointer of a torch tensor, numpy array or raw int."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def make_rsm(diffuse, face_normal, depth):
    return Rsm(_ptr(diffuse), _ptr(face_normal), _ptr(depth))


def make_gbuffer(depth, face_normal):
    return GBuffer(_ptr(depth), _ptr(face_normal))


def make_mesh(vertices, normals, indices, prim_id, num_vertices, num_triangles):
    return Mesh(_ptr(vertices), _ptr(normals), _ptr(indices), _ptr(prim_id), num_vertices, num_triangles)


class Context:
    """One mlgi_ctx (one GPU). Thin 1:1 wrapper; every method raises MlgiError on a non-zero status."""

    def __init__(self, width, height, vpl_count, rsm_width, rsm_height, max_triangles, max_points,
                 device=0, class_lo=0, class_hi=16, enable_timers=False, gi_packed=False):
        self.lib = load()
        self.cfg = Config(device, width, height, vpl_count, rsm_width, rsm_height, max_triangles, max_points,
                          class_lo, class_hi, int(enable_timers), int(gi_packed))      # gi_packed: False/True or 2 = packed image only
        self.h = C.c_void_p()
        rc = self.lib.mlgi_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise MlgiError(rc, self.lib.mlgi_last_error(None).decode())

    def close(self):
        if self.h:
            self.lib.mlgi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MlgiError(rc, self.lib.mlgi_last_error(self.h).decode())

    def vpl_process(self, rsm, light, params, stream=None):
        self._ck(self.lib.mlgi_vpl_process(self.h, C.byref(rsm), C.byref(light), C.byref(params), stream))

    def ism_points(self, mesh, params, stream=None):
        self._ck(self.lib.mlgi_ism_points(self.h, C.byref(mesh), C.byref(params), stream))

    def ism_set_points(self, points, bucket_start, num_points, stream=None):
        self._ck(self.lib.mlgi_ism_set_points(self.h, _ptr(points), _ptr(bucket_start), num_points, stream))

    def ism_splat(self, params, stream=None):
        self._ck(self.lib.mlgi_ism_splat(self.h, C.byref(params), stream))

    def ism_pullpush(self, params, atlas_override=None, stream=None):
        self._ck(self.lib.mlgi_ism_pullpush(self.h, _ptr(atlas_override), C.byref(params), stream))

    def cull_clustered(self, gb, cam, params, stream=None):
        self._ck(self.lib.mlgi_cull_clustered(self.h, C.byref(gb), C.byref(cam), C.byref(params), stream))

    def cull_tiled(self, gb, cam, params, stream=None):
        self._ck(self.lib.mlgi_cull_tiled(self.h, C.byref(gb), C.byref(cam), C.byref(params), stream))

    def final_gather(self, gb, cam, params, ism_override=None, stream=None):
        self._ck(self.lib.mlgi_final_gather(self.h, C.byref(gb), C.byref(cam), _ptr(ism_override), C.byref(params), stream))

    def gi_blur(self, gb, cam, stream=None):
        self._ck(self.lib.mlgi_gi_blur(self.h, C.byref(gb), C.byref(cam), stream))

    def composite_ambient(self, diffuse_rgba8, out_srgb8, occlusion=None, exposure=1.0, use_blurred=True, stream=None):
        """ambient term of the deferred composite (device tensors / pointers): sRGB8(tonemap(gi * toLinear(diffuse) * occlusion))"""
        self._ck(self.lib.mlgi_composite_ambient(self.h, _ptr(diffuse_rgba8), _ptr(occlusion), exposure, int(use_blurred), _ptr(out_srgb8), stream))

    def frame(self, rsm, light, mesh, gb, cam, params, stream=None):
        self._ck(self.lib.mlgi_frame(self.h, C.byref(rsm), C.byref(light), C.byref(mesh), C.byref(gb), C.byref(cam),
                                     C.byref(params), stream))

    def frame_host(self, host_rsm, light, dev_mesh, host_gb, cam, params, host_gi=None, stream=None):
        self._ck(self.lib.mlgi_frame_host(self.h, C.byref(host_rsm), C.byref(light), C.byref(dev_mesh), C.byref(host_gb),
                                          C.byref(cam), C.byref(params), _ptr(host_gi), stream))

    def frame_host_packed(self, host_rsm, light, dev_mesh, host_gb, cam, params, host_gi_packed=None, stream=None):
        self._ck(self.lib.mlgi_frame_host_packed(self.h, C.byref(host_rsm), C.byref(light), C.byref(dev_mesh), C.byref(host_gb),
                                                 C.byref(cam), C.byref(params), _ptr(host_gi_packed), stream))

    def frame_sharded(self, rsm, light, mesh, gb, cam, params, stream=None):
        self._ck(self.lib.mlgi_frame_sharded(self.h, C.byref(rsm), C.byref(light), C.byref(mesh), C.byref(gb), C.byref(cam),
                                             C.byref(params), stream))

    def frame_prepare(self, rsm, light, mesh, gb, cam, params, sharded=False, stream=None):
        """build the graph the next frame / frame_sharded call replays, without launching it"""
        self._ck(self.lib.mlgi_frame_prepare(self.h, C.byref(rsm), C.byref(light), C.byref(mesh), C.byref(gb), C.byref(cam),
                                             C.byref(params), int(bool(sharded)), stream))

    def frame_sharded_host(self, host_rsm, light, dev_mesh, host_gb, cam, params, host_gi_packed=None, stream=None):
        self._ck(self.lib.mlgi_frame_sharded_host(self.h, C.byref(host_rsm), C.byref(light), C.byref(dev_mesh), C.byref(host_gb),
                                                  C.byref(cam), C.byref(params), _ptr(host_gi_packed), stream))

    def sharded_rows(self):
        """(row_begin, row_block, row_stride): this rank uploads / bounds / returns the blocks of row_block image rows that start
        at row_begin + j * row_stride (its 64-pixel tile rows t with t mod world == rank)"""
        y0, blk, stride = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.mlgi_sharded_rows(self.h, C.byref(y0), C.byref(blk), C.byref(stride)))
        return int(y0.value), int(blk.value), int(stride.value)

    def sharded_row_mask(self):
        """boolean array over the image rows: True for the rows of sharded_rows()"""
        import numpy as np
        y0, blk, stride = self.sharded_rows()
        y = np.arange(self.cfg.height)
        return (y >= y0) & (((y - y0) % stride) < blk)

    def set_graph(self, enable):
        self._ck(self.lib.mlgi_set_graph(self.h, int(enable)))

    def frame_stats(self):
        st = FrameStats()
        self._ck(self.lib.mlgi_get_frame_stats(self.h, C.byref(st)))
        return {"host_enqueue_us": st.host_enqueue_us, "graph_replays": int(st.graph_replays),
                "graph_instantiations": int(st.graph_instantiations), "graph_enabled": bool(st.graph_enabled)}

    def buffer(self, buf_id):
        p = C.c_void_p()
        n = C.c_size_t()
        self._ck(self.lib.mlgi_get_buffer(self.h, buf_id, C.byref(p), C.byref(n)))
        return p.value, n.value

    def download(self, buf_id, dtype, count=None):
        """Copy a ctx buffer to a numpy array (synchronising)."""
        import numpy as np
        _, nbytes = self.buffer(buf_id)
        dt = np.dtype(dtype)
        n = nbytes // dt.itemsize if count is None else count
        out = np.empty(n, dtype=dt)
        self._ck(self.lib.mlgi_read_buffer(self.h, buf_id, out.ctypes.data, n * dt.itemsize, 0))
        return out

    def read_async(self, buf_id, host_tensor, nbytes, offset=0, stream=None):
        self._ck(self.lib.mlgi_read_buffer_async(self.h, buf_id, _ptr(host_tensor), nbytes, offset, stream))

    def counts(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.mlgi_get_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def gather_stats(self, gb, cam, params, ism_override=None):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self._ck(self.lib.mlgi_gather_stats(self.h, C.byref(gb), C.byref(cam), _ptr(ism_override), C.byref(params),
                                            C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_sqrt_thresholds(self, radii, out, stream=None):
        self._ck(self.lib.mlgi_debug_sqrt_thresholds(self.h, _ptr(radii), int(radii.numel()), _ptr(out), stream))

    def fp32_peak_tflops(self):
        return float(self.lib.mlgi_measure_fp32_peak_tflops(self.h))

    def bind_planar_output(self, planar16):
        self._ck(self.lib.mlgi_bind_planar_output(self.h, _ptr(planar16)))

    def planar_to_raster(self, planar16, raster=None, stream=None):
        self._ck(self.lib.mlgi_planar_to_raster(self.h, _ptr(planar16), _ptr(raster), stream))

    def timings(self):
        arr = (Timing * 16)()
        n = self.lib.mlgi_get_timings(self.h, arr, 16)
        return {arr[i].name.decode(): (arr[i].ms, arr[i].launches) for i in range(n)}

    def exchange_create(self, rank, world):
        """returns this rank's 64-byte IPC handle (bytes)"""
        buf = C.create_string_buffer(64)
        self._ck(self.lib.mlgi_exchange_create(self.h, rank, world, buf))
        return buf.raw

    def exchange_connect(self, handles):
        blob = b"".join(handles)
        self._ck(self.lib.mlgi_exchange_connect(self.h, C.c_char_p(blob)))

    def exchange_connect_local(self, contexts):
        """same-process peers: `contexts` = the world Context objects in rank order"""
        arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
        self._ck(self.lib.mlgi_exchange_connect_local(self.h, arr))

    def exchange_finish(self, stream=None):
        self._ck(self.lib.mlgi_exchange_finish(self.h, stream))

    def exchange_status(self):
        self._ck(self.lib.mlgi_exchange_status(self.h))

    def set_timers(self, enable):
        self._ck(self.lib.mlgi_set_timers(self.h, int(enable)))

    def launch_count(self):
        return int(self.lib.mlgi_launch_count(self.h))

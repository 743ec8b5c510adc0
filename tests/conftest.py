"""pytest configuration: `gpu` marker, import paths, shared small workloads.

-m "not gpu": oracle vs known-answer values / golden fixtures, host logic, C-ABI export check.
-m gpu      : parity tests proper -- CUDA path (through the C ABI) vs the CPU oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Some -m gpu tests emulate two ranks on ONE GPU: rank 0's exchange kernel spins until rank 1's frame (issued later, on other
# streams) has signalled. With the default 8 hardware work queues, unrelated streams can share a queue, and a kernel of rank 1
# queued behind rank 0's spinning kernel would never start (a false dependency that cannot occur with one process per GPU).
# More queues remove the aliasing; read by the driver when the CUDA context is created, i.e. after this line.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Same emulation, second hazard: with lazy module loading the FIRST launch of a kernel loads it, which can wait for running
# work -- i.e. for rank 0's spinning kernel, which waits for that very launch of rank 1 (measured: a 2 s bounded-wait timeout in
# the first frame only). Loading every kernel when the context is created removes it.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
for p in (os.path.join(ROOT, "many-lights-gi_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fixtures"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build the CPU-side libraries (oracle, scene fixture) if they are missing; the CUDA
    library is built by __graft_entry__.build() (nvcc cross-compiles without a GPU)."""
    import subprocess
    need = [os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "fixtures", "libmlgi_scene.so"),
            os.path.join(ROOT, "many-lights-gi_b200", "lib", "libmlgi.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT)


_ensure_built()


class Case:
    """A seeded synthetic workload plus lazily computed oracle results for every stage."""

    def __init__(self, W, H, N, rsm_w, rsm_h, cell, clutter=8, sun=266.0, seed=0x5EED, **camera):
        from scene_fixture import Workload
        self.wl = Workload(W, H, N, rsm_w=rsm_w, rsm_h=rsm_h, cell_size=cell, clutter_boxes=clutter, sun_cycle_position=sun, seed=seed, **camera)
        self.W, self.H, self.N = W, H, N
        self.z_far = 50.0
        self._c = {}

    def get(self, key):
        import oracle_py as O
        c, wl, N = self._c, self.wl, self.N
        if key in c:
            return c[key]
        if key in ("vpl", "packed"):
            c["vpl"], c["packed"] = O.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, N, wl.light_view, wl.light_proj)
        elif key in ("points_raw", "bucket_raw"):
            sc = wl.scene
            c["points_raw"], c["bucket_raw"] = O.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, N)
        elif key in ("points", "bucket_start"):
            c["points"], c["bucket_start"] = O.group_by_bucket(self.get("points_raw"), self.get("bucket_raw"), N)
        elif key == "atlas":
            c["atlas"] = O.ism_splat(self.get("points"), self.get("bucket_start"), self.get("packed"), N, self.z_far)
        elif key == "ism":
            c["ism"] = O.ism_pullpush(self.get("atlas"), N, self.z_far)
        elif key in ("compact", "list_ids", "num_used"):
            c["compact"], c["list_ids"], c["num_used"] = O.clustering(wl.depth, wl.cam_proj, self.z_far)
        elif key == "light_lists":
            c["light_lists"] = O.light_lists(self.get("packed"), N, self.get("compact"), wl.cam_view, wl.cam_proj, self.W, self.H, self.z_far)
        elif key == "tile_lists":
            c["tile_lists"] = O.cull_tiled(wl.depth, wl.normal, self.get("vpl"), N, wl.cam_view, wl.cam_proj, 64)
        else:
            raise KeyError(key)
        return c[key]

    def gather(self, cull_mode, shadows=True, bounds=True, class_lo=0, class_hi=16):
        import oracle_py as O
        key = ("fg", cull_mode, shadows, bounds, class_lo, class_hi)
        if key not in self._c:
            wl = self.wl
            ids = self.get("list_ids") if cull_mode == 0 else None
            lists = self.get("light_lists") if cull_mode == 0 else (self.get("tile_lists") if cull_mode == 1 else None)
            self._c[key] = O.final_gather(wl.depth, wl.normal, self.get("vpl"), self.N, self.get("ism") if shadows else None,
                                          cull_mode, ids, lists, wl.cam_view, wl.cam_proj, self.z_far, shadows=shadows,
                                          bounds=bounds, class_lo=class_lo, class_hi=class_hi)
        return self._c[key]


@pytest.fixture(scope="session")
def case_alt():
    """another scene seed, sun position and camera (looking across the atrium from the far side): guards against a path
    that only works for the default view"""
    return Case(256, 144, 256, 256, 64, cell=1.0, clutter=10, sun=250.0, seed=1234, eye=(6.0, 3.0, 1.5), center=(-6.0, 4.0, -1.0))


def make_case_mid():
    return Case(384, 216, 1024, 1024, 256, cell=0.5, clutter=12)


@pytest.fixture(scope="session")
def case_small():
    """ragged 160x100 viewport (not a multiple of 4, 64 or 128), 64 VPLs, coarse mesh (n>1 tessellation)"""
    return Case(160, 100, 64, 128, 128, cell=1.5, clutter=6)


@pytest.fixture(scope="session")
def case_c1():
    """BASELINE config 0: 256x256, 64 VPLs from a 128^2 RSM, 4x4 interleave (shadows optional)"""
    return Case(256, 256, 64, 128, 128, cell=0.8, clutter=8)


@pytest.fixture(scope="session")
def case_mid():
    """reference defaults at reduced resolution: 1024 VPLs, 1024x256 RSM, 2048^2 atlas"""
    return make_case_mid()

"""Pins the oracle (oracle/oracle.cpp) against the reference's OWN shader sources: oracle/refglsl/build_ref.py reads the
GLSL files where they lie under /root/reference/data/shaders, rewrites them syntactically into C++20 and compiles them
with g++ against a small GLSL runtime (oracle/refglsl/glsl_rt.h) into oracle/_ref/libref_glsl.so. Every stage of the
hot path (a1, a2's programmable stages, a3, a4+a5, a6, a7, a9, n1) is then run on the same inputs as the oracle and
compared BIT FOR BIT. (Only the GL fixed-function tessellator inside a2 has no source to compile.)

Runs wherever the library exists (it is built here, where /root/reference is mounted, and travels with the snapshot);
tests/test_golden.py::test_oracle_reproduces_reference_derived_golden covers the same ground from committed hashes."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import ref_py as R  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    if not R.available():
        if not os.path.isdir("/root/reference/data/shaders"):
            pytest.skip("oracle/_ref/libref_glsl.so is not built and /root/reference is not mounted")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "refglsl", "build_ref.py")])
    return R


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_a1_vpl_processor(ref, case_mid):
    c, wl = case_mid, case_mid.wl
    bias = np.array([0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.5, 0, 0.5, 0.5, 0.5, 1], np.float32)      # VPLProcessor.cpp:76-82
    inv = O.mat4_inverse(O.mat4_mul(O.mat4_mul(bias, wl.light_proj), wl.light_view))
    for shuffle in (True, False):
        vpl, packed = ref.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, c.N, inv, 5.0, shuffle, O.shuffle_table(c.N))
        o_vpl, o_packed = O.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, c.N, wl.light_view, wl.light_proj, 5.0, shuffle)
        assert np.array_equal(bits(vpl), bits(o_vpl)) and np.array_equal(bits(packed), bits(o_packed))


def canonical_points(points, bucket):
    a = np.concatenate([bits(points), np.asarray(bucket).astype(np.uint32)[:, None]], 1)
    return a[np.lexsort(a.T[::-1])]


def test_a2_programmable_point_stages(ref, case_mid):
    """ism.tesc (tessellation levels), ism.tese (interpolation) and ism.geom (centre, radius, bucket, pointBuffer write) run
    from source; only the fixed-function tessellator between them is replaced by the deterministic split of DESIGN.md 3.4"""
    sc = case_mid.wl.scene
    M = 6000
    pts, bkt = ref.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, case_mid.N, max_triangles=M)
    o_pts, o_bkt = O.ism_points(sc.vertices, sc.normals, sc.indices[:M], sc.prim_id[:M], 2.0, case_mid.N)
    assert pts.shape[0] > M            # some triangles are tessellated (n > 1)
    assert np.array_equal(canonical_points(pts, bkt), canonical_points(o_pts, o_bkt))


def test_a3_splat(ref, case_mid):
    c = case_mid
    atlas = ref.ism_splat(c.get("points"), c.get("bucket_start"), c.get("packed"), c.N, 50.0, c.get("atlas").shape[0])
    assert np.array_equal(atlas, c.get("atlas"))


def test_a45_pull_push_twelve_passes(ref, case_mid, case_small):
    for c in (case_mid, case_small):                      # 2048^2 and 512^2 atlases
        assert np.array_equal(ref.ism_pullpush(c.get("atlas"), 50.0), c.get("ism"))
    # synthetic atlases: empty, a single point, dense noise, zero depth (exercise the validity / radius branches)
    rng = np.random.default_rng(7)
    A = 512
    cases = [np.full((A, A), 0xFFFFFFFF, np.uint32)]
    one = cases[0].copy(); one[100, 200] = (123456 << 8) | 17; cases.append(one)
    cases.append(((rng.integers(0, 1 << 24, (A, A), dtype=np.uint64) << 8) | rng.integers(0, 256, (A, A), dtype=np.uint64)).astype(np.uint32))
    sparse = cases[0].copy(); m = rng.random((A, A)) < 0.03
    sparse[m] = ((rng.integers(1 << 20, 1 << 23, int(m.sum()), dtype=np.uint64) << 8) | rng.integers(1, 60, int(m.sum()), dtype=np.uint64)).astype(np.uint32)
    cases.append(sparse)
    zero = cases[0].copy(); zero[::7, ::5] = 5; cases.append(zero)
    for atlas in cases:
        assert np.array_equal(ref.ism_pullpush(atlas, 50.0), O.ism_pullpush(atlas, 64, 50.0))


def test_a6_a7_clustering_and_light_lists(ref, case_mid):
    c, wl = case_mid, case_mid.wl
    compact, ids, n = ref.clustering(wl.depth, wl.cam_proj, 50.0)
    oc, oi, on = c.get("compact"), c.get("list_ids"), c.get("num_used")
    assert n == on and np.array_equal(compact, oc[:on])                 # sequential dispatch == the oracle's canonical order (Q5)
    assert np.array_equal(ids != 0xFFFF, oi != 0xFFFF) and np.array_equal(ids[ids != 0xFFFF], oi[oi != 0xFFFF])
    inv_pv = O.mat4_inverse(O.mat4_mul(wl.cam_proj, wl.cam_view))        # ClusteredShading.cpp:105
    img = ref.light_lists(c.get("packed"), c.N, oc[:on], c.W, c.H, wl.cam_proj, inv_pv, 50.0, rows=c.N + 64)
    ol = c.get("light_lists")
    full = 0
    for cl in range(on):
        for k in range(16):
            n_or = int(ol[cl, k, 0])
            if n_or - 1 == 64:
                full += 1             # Q2: the reference's 65th slot is the next sub-list's counter (overflow race); the oracle has L+1 slots
            else:
                assert int(img[k * 64, cl]) == n_or
            m = min(n_or, 64)
            assert np.array_equal(img[k * 64 + 1:k * 64 + m, cl], ol[cl, k, 1:m])
    assert full < on * 16             # the comparison is not vacuous


def _full_sublist_pixels(c):
    wl = c.wl
    B = O.slice_boundaries(50.0)
    ys, xs = np.mgrid[0:c.H, 0:c.W]
    P = wl.cam_proj.reshape(4, 4).T
    dl = P[2, 3] / (P[2, 2] + (2.0 * wl.depth - 1.0))
    sl = np.zeros((c.H, c.W), np.int64)
    for i in range(1, 16):
        sl[dl >= B[i]] = i
    lid = c.get("list_ids")[sl, ys // 128, xs // 128].astype(np.int64)
    cls = (xs & 3) + 4 * (ys & 3)
    return c.get("light_lists")[lid, cls, 0].astype(np.int64) - 1 >= 64


@pytest.mark.parametrize("shadows", [True, False])
def test_a9_final_gathering(ref, case_mid, shadows):
    c, wl = case_mid, case_mid.wl
    oc, oi, on, ol = c.get("compact"), c.get("list_ids"), c.get("num_used"), c.get("light_lists")
    inv_pv = O.mat4_inverse(O.mat4_mul(wl.cam_proj, wl.cam_view))
    img = ref.light_lists(c.get("packed"), c.N, oc[:on], c.W, c.H, wl.cam_proj, inv_pv, 50.0, rows=c.N + 64)
    vpi = O.mat4_mul(O.mat4_inverse(wl.cam_view), O.mat4_inverse(wl.cam_proj))      # GIStage.cpp:289
    got = ref.final_gather(wl.depth, wl.normal, c.get("vpl"), c.N, c.get("ism"), oi, img, wl.cam_proj, vpi, 50.0, 0.05, shadows=shadows)
    ok_px = ~_full_sublist_pixels(c)          # Q2 overflow: the reference reads the next sub-list's counter as a VPL id there
    assert ok_px.mean() > 0.5
    O.set_reference_math(1)                   # pow(dist, 4.0) through libm, as the compiled GLSL evaluates it
    try:
        want = O.final_gather(wl.depth, wl.normal, c.get("vpl"), c.N, c.get("ism") if shadows else None, 0, oi, ol, wl.cam_view, wl.cam_proj,
                              50.0, shadows=shadows)["rgb"]
    finally:
        O.set_reference_math(0)
    assert np.array_equal(bits(got)[ok_px], bits(want)[ok_px])
    # the shipped default (two multiplies instead of pow) stays within 2 ulp of it
    fast = c.gather(0, shadows, bounds=False)["rgb"]
    assert np.all(np.abs(fast[ok_px] - got[ok_px]) <= 4e-7 * np.maximum(np.abs(got[ok_px]), 1e-3))


def test_n1_blur(ref, case_mid):
    c, wl = case_mid, case_mid.wl
    gi = c.gather(0, True, bounds=False)["rgb"]
    got, want = ref.gi_blur(gi, wl.depth, wl.normal, wl.cam_proj), O.gi_blur(gi, wl.depth, wl.normal, wl.cam_proj)
    inner = (slice(3, -3), slice(3, -3))      # texelFetch outside the image is undefined in GL; the oracle skips those taps
    assert np.array_equal(bits(got[inner]), bits(want[inner]))


def test_n4_ambient_term_of_the_composite(ref, case_mid):
    """deferredshading.frag from source with the direct light switched off (lightIntensity = 0, black specular buffer):
    what remains is toSRGB(tonemap(gi * toLinear(diffuse) * occlusion)), the term mlgi_composite_ambient implements"""
    c, wl = case_mid, case_mid.wl
    gi = c.gather(0, True, bounds=False)["rgb"]
    occ = np.linspace(0.25, 1.0, c.W * c.H, dtype=np.float32).reshape(c.H, c.W)
    for occlusion, exposure in ((None, 1.0), (occ, 2.5)):
        got = ref.composite_ambient(gi, wl.diffuse, occlusion, exposure)
        want, _ = O.composite_ambient(gi, wl.diffuse, occlusion, exposure)
        assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("vpls", [64, 4096])
def test_other_vpl_counts(vpls, case_small):
    """the reference hard-codes 1024 VPLs; with the literal substituted (build_ref.py --vpls) the same sources are run at
    64 VPLs (8x8 ISMs, 512^2 atlas) and 4096 VPLs (64x64 ISMs, 4096^2 atlas): a1, a3, a4+a5, a6 and a9 (light_lists.comp
    hard-codes 64-entry sub-lists, so the lists given to a9 are the oracle's, laid out as the reference would)"""
    if not R.available(vpls):
        if not os.path.isdir("/root/reference/data/shaders"):
            pytest.skip("reference build for %d VPLs not available" % vpls)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "refglsl", "build_ref.py"), "--vpls", str(vpls)])
    from conftest import Case
    c = case_small if vpls == 64 else Case(128, 72, 4096, 1024, 256, cell=0.5, clutter=12)
    wl, N, L = c.wl, c.N, c.N // 16
    R.use_vpls(vpls)
    try:
        bias = np.array([0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.5, 0, 0.5, 0.5, 0.5, 1], np.float32)
        inv = O.mat4_inverse(O.mat4_mul(O.mat4_mul(bias, wl.light_proj), wl.light_view))
        vpl, packed = R.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, N, inv, 5.0, True, O.shuffle_table(N))
        assert np.array_equal(bits(vpl), bits(c.get("vpl"))) and np.array_equal(bits(packed), bits(c.get("packed")))
        A = c.get("atlas").shape[0]
        assert np.array_equal(R.ism_splat(c.get("points"), c.get("bucket_start"), c.get("packed"), N, 50.0, A), c.get("atlas"))
        assert np.array_equal(R.ism_pullpush(c.get("atlas"), 50.0), c.get("ism"))
        compact, ids, n = R.clustering(wl.depth, wl.cam_proj, 50.0)
        oc, oi, on, ol = c.get("compact"), c.get("list_ids"), c.get("num_used"), c.get("light_lists")
        assert n == on and np.array_equal(compact, oc[:on]) and np.array_equal(ids[ids != 0xFFFF], oi[oi != 0xFFFF])
        img = R.lists_image_from_canonical(ol, on, L)
        vpi = O.mat4_mul(O.mat4_inverse(wl.cam_view), O.mat4_inverse(wl.cam_proj))
        got = R.final_gather(wl.depth, wl.normal, c.get("vpl"), N, c.get("ism"), oi, img, wl.cam_proj, vpi, 50.0, 0.05, shadows=True)
        B = O.slice_boundaries(50.0)
        ys, xs = np.mgrid[0:c.H, 0:c.W]
        P = wl.cam_proj.reshape(4, 4).T
        dl = P[2, 3] / (P[2, 2] + (2.0 * wl.depth - 1.0))
        sl = np.zeros((c.H, c.W), np.int64)
        for i in range(1, 16):
            sl[dl >= B[i]] = i
        full = ol[oi[sl, ys // 128, xs // 128].astype(np.int64), (xs & 3) + 4 * (ys & 3), 0].astype(np.int64) - 1 >= L
        O.set_reference_math(1)
        try:
            want = O.final_gather(wl.depth, wl.normal, c.get("vpl"), N, c.get("ism"), 0, oi, ol, wl.cam_view, wl.cam_proj, 50.0, shadows=True)["rgb"]
        finally:
            O.set_reference_math(0)
        assert (~full).sum() > 300       # at 64 VPLs (4 per class) most sub-lists are full and hit the overflow race (Q2)
        assert np.array_equal(bits(got)[~full], bits(want)[~full])
    finally:
        R.use_vpls(1024)


# n3: the scaled-ISM experiment knobs (vplStartIndex / vplEndIndex, scaleISMs, pointsOnlyIntoScaledISMs; ism.geom:37-46,
# ism.comp:31-40,56-66,124, final_gathering.comp:40-46 with GIStage.cpp:430's rewrite of SCALE_ISMS,
# ImperfectShadowmap.cpp:184-191 for the pull-push pixel size)
N3_MODES = [dict(vpl_start=256, vpl_end=512, scale_isms=True, points_only=True),      # 256 ISMs of 128^2, points only into them
            dict(vpl_start=256, vpl_end=512, scale_isms=True, points_only=False),     # scaled layout, every VPL still splatted
            dict(vpl_start=0, vpl_end=256, scale_isms=False, points_only=True),       # buckets wrap inside the range, 64^2 ISMs
            dict(vpl_start=100, vpl_end=1000, scale_isms=True, points_only=True)]     # 900 ISMs: ismIndices1d stays 32


@pytest.mark.parametrize("mode", N3_MODES, ids=lambda m: "s%d_e%d_scale%d_only%d" % (m["vpl_start"], m["vpl_end"], m["scale_isms"], m["points_only"]))
def test_n3_scaled_isms(ref, case_mid, mode):
    c, wl, sc, N = case_mid, case_mid.wl, case_mid.wl.scene, case_mid.N
    sampled, ism_count = O.ism_mode(N, **mode)
    M = 3000
    pts, bkt = ref.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, N, max_triangles=M, **mode)
    o_pts, o_bkt = O.ism_points(sc.vertices, sc.normals, sc.indices[:M], sc.prim_id[:M], 2.0, N, sampled, ism_count)
    assert np.array_equal(canonical_points(pts, bkt), canonical_points(o_pts, o_bkt))
    assert int(o_bkt.max()) < sampled
    # splat + 12 pull-push passes on the full point set of the mode
    a_pts, a_bkt = O.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, N, sampled, ism_count)
    pg, bs = O.group_by_bucket(a_pts, a_bkt, N)
    A = c.get("atlas").shape[0]
    atlas = O.ism_splat(pg, bs, c.get("packed"), N, 50.0, **mode)
    assert np.array_equal(ref.ism_splat(pg, bs, c.get("packed"), N, 50.0, A, **mode), atlas)
    assert not np.array_equal(atlas, c.get("atlas"))
    ism = O.ism_pullpush(atlas, N, 50.0, ism_count=ism_count)
    assert np.array_equal(ref.ism_pullpush(atlas, 50.0, ism_pixel_size=A // O.ism_indices_1d(ism_count)), ism)
    # final gathering with the matching SCALE_ISMS define
    oc, oi, on, ol = c.get("compact"), c.get("list_ids"), c.get("num_used"), c.get("light_lists")
    inv_pv = O.mat4_inverse(O.mat4_mul(wl.cam_proj, wl.cam_view))
    img = ref.light_lists(c.get("packed"), N, oc[:on], c.W, c.H, wl.cam_proj, inv_pv, 50.0, rows=N + 64)
    vpi = O.mat4_mul(O.mat4_inverse(wl.cam_view), O.mat4_inverse(wl.cam_proj))
    got = ref.final_gather(wl.depth, wl.normal, c.get("vpl"), N, ism, oi, img, wl.cam_proj, vpi, 50.0, 0.05, vpl_start=mode["vpl_start"],
                           vpl_end=mode["vpl_end"], scale_isms=mode["scale_isms"])
    ok_px = ~_full_sublist_pixels(c)
    O.set_reference_math(1)
    try:
        want = O.final_gather(wl.depth, wl.normal, c.get("vpl"), N, ism, 0, oi, ol, wl.cam_view, wl.cam_proj, 50.0, vpl_start=mode["vpl_start"],
                              vpl_end=mode["vpl_end"], scale_isms=mode["scale_isms"])["rgb"]
    finally:
        O.set_reference_math(0)
    assert np.array_equal(bits(got)[ok_px], bits(want)[ok_px])

"""Golden vectors DERIVED FROM THE REFERENCE: runs the reference's GLSL sources (compiled for the CPU by
oracle/refglsl/build_ref.py; needs /root/reference) on the `case_mid` fixture and commits SHA-256 digests of what they
produce. tests/test_golden.py then checks the oracle against these digests wherever the tests run (the GPU box has no
/root/reference). usage: python tests/golden/make_ref_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import conftest  # noqa: E402
import oracle_py as O  # noqa: E402
import ref_py as R  # noqa: E402
from test_reference_glsl import _full_sublist_pixels, bits  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    c = conftest.make_case_mid()
    wl = c.wl
    g = {"case": "case_mid: %dx%d, %d VPLs" % (c.W, c.H, c.N), "source": "reference GLSL compiled by oracle/refglsl (see DESIGN.md section 6)"}
    bias = np.array([0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.5, 0, 0.5, 0.5, 0.5, 1], np.float32)
    inv = O.mat4_inverse(O.mat4_mul(O.mat4_mul(bias, wl.light_proj), wl.light_view))
    vpl, packed = R.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, c.N, inv, 5.0, True, O.shuffle_table(c.N))
    g["a1_vpl"], g["a1_packed"] = sha(vpl), sha(packed)
    # a2: tesc / tese / geom from source on the first 6000 triangles (the GL tessellator itself has no source); the full point
    # list fed to the stages downstream is the oracle's
    from test_reference_glsl import canonical_points
    sc = wl.scene
    pts, bkt = R.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, c.N, max_triangles=6000)
    g["a2_points_first_6000_triangles"] = sha(canonical_points(pts, bkt))
    atlas = R.ism_splat(c.get("points"), c.get("bucket_start"), packed, c.N, 50.0, c.get("atlas").shape[0])
    g["a3_atlas"] = sha(atlas)
    ism = R.ism_pullpush(atlas, 50.0)
    g["a45_ism"] = sha(ism)
    compact, ids, n = R.clustering(wl.depth, wl.cam_proj, 50.0)
    g["a6_num_used"], g["a6_compact"], g["a6_used_mask"], g["a6_ids_used"] = n, sha(compact), sha(ids != 0xFFFF), sha(ids[ids != 0xFFFF])
    inv_pv = O.mat4_inverse(O.mat4_mul(wl.cam_proj, wl.cam_view))
    img = R.light_lists(packed, c.N, compact, c.W, c.H, wl.cam_proj, inv_pv, 50.0, rows=c.N + 64)
    # canonical digest of the lists: per (cluster, class) the count and the first min(count, 63) ids (Q2: slot 64 overflows)
    canon = []
    for cl in range(n):
        for k in range(16):
            cnt = int(img[k * 64, cl])
            canon.append(np.concatenate([[min(cnt, 65)], img[k * 64 + 1:k * 64 + min(cnt, 64), cl]]).astype(np.uint16))
    # a sub-list whose 64 VPLs all pass has its counter overwritten by the overflow race; record only the ids there
    g["a7_lists_ids"] = sha(np.concatenate([x[1:] for x in canon]))
    vpi = O.mat4_mul(O.mat4_inverse(wl.cam_view), O.mat4_inverse(wl.cam_proj))
    ok_px = ~_full_sublist_pixels(c)
    for shadows in (True, False):
        got = R.final_gather(wl.depth, wl.normal, vpl, c.N, ism, ids, img, wl.cam_proj, vpi, 50.0, 0.05, shadows=shadows)
        g["a9_radiance_%s" % ("shadowed" if shadows else "unshadowed")] = sha(bits(got)[ok_px])
    g["a9_pixels_compared"] = int(ok_px.sum())
    gi = c.gather(0, True, bounds=False)["rgb"]
    g["n1_blur_interior"] = sha(R.gi_blur(gi, wl.depth, wl.normal, wl.cam_proj)[3:-3, 3:-3])
    g["n4_composite_ambient"] = sha(R.composite_ambient(gi, wl.diffuse, None, 1.0))
    json.dump(g, open(os.path.join(HERE, "ref_mid.json"), "w"), indent=1, sort_keys=True)
    print("wrote ref_mid.json", {k: (v[:12] if isinstance(v, str) else v) for k, v in g.items()})



if __name__ == "__main__":
    main()

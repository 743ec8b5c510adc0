"""Generates tests/golden/golden_small.{json,npz}: outputs of the CPU oracle (oracle/liboracle.so) for the
`case_small` workload of tests/conftest.py (160x100, 64 VPLs, 128^2 RSM, coarse mesh).

The reference (GLSL + GL driver) cannot be executed in this environment, so these vectors pin OUR oracle
(and through it the CUDA path) against drift; they are not outputs of the reference itself.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402  (sets import paths, builds if needed)
from conftest import Case  # noqa: E402

TILES = (0, 9, 27, 63)   # ISM indices whose 64x64 atlas / R16 tiles are stored verbatim


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build():
    case = Case(160, 100, 64, 128, 128, cell=1.5, clutter=6)
    g = {}
    arrays = {}
    for key in ("vpl", "packed", "points", "bucket_start", "atlas", "ism", "compact", "list_ids", "light_lists", "tile_lists"):
        g["sha256_" + key] = sha(case.get(key))
    g["num_points"] = int(case.get("points").shape[0])
    g["num_used_clusters"] = int(case.get("num_used"))
    for mode in (0, 1, 2):
        r = case.gather(mode, True, bounds=False)
        g["sha256_radiance_mode%d" % mode] = sha(r["rgb"])
        g["pairs_listed_mode%d" % mode] = r["listed"]
        g["pairs_lit_mode%d" % mode] = r["lit"]
    atlas, ism = case.get("atlas"), case.get("ism")
    S = 8
    for t in TILES:
        ty, tx = divmod(t, S)
        arrays["atlas_tile_%d" % t] = atlas[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64].copy()
        arrays["ism_tile_%d" % t] = ism[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64].copy()
    arrays["vpl"] = case.get("vpl")
    arrays["packed"] = case.get("packed")
    arrays["tile_list_counts"] = case.get("tile_lists")[:, :, 0].copy()
    arrays["cluster_ids"] = case.get("compact")
    rgb = case.gather(1, True, bounds=False)["rgb"]
    arrays["radiance_rows_48_52"] = rgb[48:52].copy()
    return g, arrays


if __name__ == "__main__":
    g, arrays = build()
    json.dump(g, open(os.path.join(HERE, "golden_small.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **arrays)
    print("wrote golden_small.json / .npz:", {k: v for k, v in g.items() if not k.startswith("sha")})

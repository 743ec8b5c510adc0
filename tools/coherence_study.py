"""offline study: warp-level lit coherence of the gathering kernel under different pixel->lane mappings"""
import sys, os, numpy as np
sys.path.insert(0, 'many-lights-gi_b200'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'fixtures')
import oracle_py as O
from scene_fixture import Workload
W, H, N = 3840, 2160, 4096
wl = Workload(W, H, N)
vpl, packed = O.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, N, wl.light_view, wl.light_proj)
V = wl.cam_view.reshape(4, 4).T.astype(np.float64); P = wl.cam_proj.reshape(4, 4).T.astype(np.float64)
inv = np.linalg.inv(V) @ np.linalg.inv(P)
rng = np.random.default_rng(0)
L = N // 16
res = {k: [] for k in ("spatial8x4", "sorted_normal", "pixel", "rows16x2")}
TX, TY = W // 64, (H + 63) // 64
for _ in range(150):
    tx, ty, cls = rng.integers(0, TX), rng.integers(0, TY - 1), rng.integers(0, 16)
    cx, cy = cls & 3, cls >> 2
    xs = tx * 64 + 4 * np.arange(16) + cx; ys = ty * 64 + 4 * np.arange(16) + cy
    X, Y = np.meshgrid(xs, ys)   # [j,i]
    d = wl.depth[Y, X].astype(np.float64)
    ndc = np.stack([X / W * 2 - 1, Y / H * 2 - 1, d * 2 - 1, np.ones_like(d)], -1)
    w = ndc @ inv.T
    Pw = w[..., :3] / w[..., 3:4]
    n10 = wl.normal[Y, X]
    nrm = np.stack([(n10 & 1023) / 1023 * 2 - 1, ((n10 >> 10) & 1023) / 1023 * 2 - 1, ((n10 >> 20) & 1023) / 1023 * 2 - 1], -1)
    vp = vpl[cls * L:(cls + 1) * L, 0:3].astype(np.float64); vn = vpl[cls * L:(cls + 1) * L, 4:7].astype(np.float64)
    diff = Pw[:, :, None, :] - vp[None, None]          # [16,16,L,3]
    a = (diff * vn[None, None]).sum(-1); b = -(diff * nrm[:, :, None, :]).sum(-1)
    lit = (a > 0) & (b > 0)                             # [16,16,L]
    res["pixel"].append(lit.mean())
    # current mapping: warp w: i in 8*(w&1)+[0,8), j in 8*(w>>1)+{0..3} and +4
    anyw = []
    for wq in range(4):
        i0, j0 = 8 * (wq & 1), 8 * (wq >> 1)
        blk = lit[j0:j0 + 8, i0:i0 + 8]                 # 64 pixels of the warp
        anyw.append(blk.reshape(64, L).any(0))
    res["spatial8x4"].append(np.mean(anyw))
    # rows mapping: warp = 4 rows of 16
    res["rows16x2"].append(np.mean([lit[4 * wq:4 * wq + 4].reshape(64, L).any(0) for wq in range(4)]))
    # sorted by packed normal then position
    flat = lit.reshape(256, L); key = n10.reshape(256)
    order = np.argsort(key, kind="stable")
    res["sorted_normal"].append(np.mean([flat[order[64 * wq:64 * wq + 64]].any(0) for wq in range(4)]))
for k, v in res.items():
    print("%-14s %.3f" % (k, np.mean(v)))

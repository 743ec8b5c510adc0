"""Offline study (CPU, numpy): how many (pixel, VPL) pairs each conservative tile test would keep, against the pairs
that are actually lit (min(a,b) > 0). Used to choose the tiled-culling rule."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200")); sys.path.insert(0, os.path.join(ROOT, "fixtures")); sys.path.insert(0, ROOT)
from scene_fixture import Workload
from oracle import oracle_py as orc

W, H, N = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (3840, 2160, 4096)))
step = int(sys.argv[4]) if len(sys.argv) > 4 else 5
TILE = int(sys.argv[5]) if len(sys.argv) > 5 else 64
orc.use_all_cores()
wl = Workload(W, H, N)
vpl, packed = orc.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, N, wl.light_view, wl.light_proj)
vp = vpl[:, 0:3].astype(np.float64); vn = vpl[:, 4:7].astype(np.float64)
view = wl.cam_view.reshape(4, 4).T.astype(np.float64); proj = wl.cam_proj.reshape(4, 4).T.astype(np.float64)
inv = np.linalg.inv(proj @ view)
ys, xs = np.mgrid[0:H, 0:W]
ndc = np.stack([2 * (xs + 0.5) / W - 1, 2 * (ys + 0.5) / H - 1, 2 * wl.depth.astype(np.float64) - 1, np.ones((H, W))], -1)
Pw = ndc @ inv.T; Pw = Pw[..., :3] / Pw[..., 3:4]
n = wl.normal
Nw = np.stack([(n & 1023), (n >> 10) & 1023, (n >> 20) & 1023], -1).astype(np.float64) / 1023.0 * 2 - 1
L = N // 16
TX, TY = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
tot = dict(pairs=0, lit=0, sphere=0, a_exact=0, b_box=0, both_box=0, warp_lit=0, ab_exact=0, b_exact=0, b_cone=0)
tiles = [(tx, ty) for ty in range(TY) for tx in range(TX)][::step]
for tx, ty in tiles:
    x0, y0 = tx * TILE, ty * TILE
    Pt = Pw[y0:y0 + TILE, x0:x0 + TILE]; Nt = Nw[y0:y0 + TILE, x0:x0 + TILE]
    Pf = Pt.reshape(-1, 3); Nf = Nt.reshape(-1, 3)
    c = Pf.mean(0); R = np.sqrt(((Pf - c) ** 2).sum(1).max()) * 1.01
    e = (Nf * Pf).sum(1)
    nmin, nmax, emin = Nf.min(0), Nf.max(0), e.min()
    for cls in range(16):
        cx, cy = cls & 3, cls >> 2
        Pc = Pt[cy::4, cx::4].reshape(-1, 3); Nc = Nt[cy::4, cx::4].reshape(-1, 3)
        v = vp[cls * L:(cls + 1) * L]; nv = vn[cls * L:(cls + 1) * L]
        d = Pc[:, None, :] - v[None, :, :]
        a = (d * nv[None]).sum(-1); b = -(d * Nc[:, None, :]).sum(-1)
        lit = (a > 0) & (b > 0)
        sphere = ((c - v) * nv).sum(1) >= -R
        bbox = np.maximum(nmin * v, nmax * v).sum(1) - emin > 0
        tot["pairs"] += lit.size; tot["lit"] += lit.sum()
        tot["sphere"] += sphere.sum() * len(Pc); tot["b_box"] += bbox.sum() * len(Pc)
        tot["both_box"] += (sphere & bbox).sum() * len(Pc)
        tot["a_exact"] += (a > 0).any(0).sum() * len(Pc); tot["b_exact"] += (b > 0).any(0).sum() * len(Pc)
        tot["ab_exact"] += lit.any(0).sum() * len(Pc)
        # warp granularity: 8x4 class pixels
        hh, ww = Pt[cy::4, cx::4].shape[:2]
        l2 = lit.reshape(hh, ww, -1)
        for j in range(0, hh, 4):
            for i in range(0, ww, 8):
                blk = l2[j:j + 4, i:i + 8]
                tot["warp_lit"] += blk.any((0, 1)).sum() * blk.shape[0] * blk.shape[1]
for k, v in tot.items():
    print("%-10s %14d  %.3f" % (k, v, v / tot["pairs"]))

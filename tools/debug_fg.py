"""debug helper: worst radiance mismatches of the CUDA gathering vs the oracle for one test case"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import conftest  # noqa  (sets paths)
from conftest import Case
import mlgi
from test_gpu_parity import Gpu

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
case = {"small": lambda: Case(160, 100, 64, 128, 128, cell=1.5, clutter=6), "c1": lambda: Case(256, 256, 64, 128, 128, cell=0.8, clutter=8),
        "mid": lambda: Case(384, 216, 1024, 1024, 256, cell=0.5, clutter=12),
        "4k": lambda: Case(3840, 2160, 4096, 1024, 256, cell=0.155, clutter=24)}[which]()
g = Gpu(case, max_points=1 << 23)
p = mlgi.default_params(case.N)
p.cull_mode = mlgi.CULL_TILED
g.ctx.vpl_process(g.rsm, g.light, p)
g.ctx.cull_tiled(g.gb, g.cam, p)
g.ctx.final_gather(g.gb, g.cam, p, ism_override=g.up(case.get("ism")))
gi = g.ctx.download(mlgi.BUF_GI, np.float32).reshape(case.H, case.W, 4)[:, :, :3]
ref = case.gather(1, True)
rgb, lo, hi = ref["rgb"], ref["lo"], ref["hi"]
mean = rgb.mean()
rel = np.abs(gi - rgb) / (np.abs(rgb) + 1e-6 * mean)
notie = (lo == hi)
rel_nt = np.where(notie, rel, 0)
idx = np.argsort(rel_nt.ravel())[::-1][:12]
print("mean radiance", mean, "tie-free frac", notie.mean())
for i in idx:
    y, x, c = np.unravel_index(i, rel.shape)
    print("px(%d,%d) ch%d gpu %.8g ref %.8g rel %.3g  value/mean %.3g" % (x, y, c, gi[y, x, c], rgb[y, x, c], rel[y, x, c], rgb[y, x, c] / mean))
for thr in (1e-3, 1e-4, 1e-5):
    print("pixels with rel >", thr, ":", int((rel_nt > thr).sum()), "of", rel.size)



f64 = ref["f64"]
atol = 1e-5 * mean + 2.0 * np.abs(rgb - f64)
excess = np.maximum(lo - (1e-3 * np.abs(lo) + atol) - gi, gi - (hi + (1e-3 * np.abs(hi) + atol)))
ratio = np.abs(gi - rgb) / (1e-3 * np.abs(rgb) + atol)
bad = excess > 0
print("outside:", int(bad.sum()))
order = np.argsort(np.where(bad, ratio, 0).max(2).ravel())[::-1][:14]
for i in order:
    y, x = divmod(int(i), case.W)
    print("px(%d,%d) cls %d err/tol %.1f gpu %s ref32 %s ref64 %s lo==hi %s" % (x, y, (x & 3) + 4 * (y & 3), ratio[y, x].max(), gi[y, x], rgb[y, x], f64[y, x], np.array_equal(lo[y, x], hi[y, x])))

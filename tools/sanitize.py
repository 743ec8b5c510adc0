"""small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every stage, both cull modes,
blur, packed output, sharded context"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import conftest  # noqa
from conftest import Case
import mlgi
from test_gpu_parity import Gpu
case = Case(160, 100, 64, 128, 128, cell=1.5, clutter=6)
for kw in (dict(), dict(class_lo=4, class_hi=8)):
    g = Gpu(case, **kw)
    for mode in (mlgi.CULL_TILED, mlgi.CULL_CLUSTERED, mlgi.CULL_NONE):
        p = mlgi.default_params(case.N); p.cull_mode = mode
        g.ctx.frame(g.rsm, g.light, g.mesh, g.gb, g.cam, p)
        g.ctx.gi_blur(g.gb, g.cam)
    print("listed/lit", g.ctx.gather_stats(g.gb, g.cam, p))
    g.close()
print("sanitize run done")

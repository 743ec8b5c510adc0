"""small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every stage, all cull modes,
blur, packed-only output, graph replay, a class shard, and the peer-memory exchange (two shards of this process)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import conftest  # noqa
from conftest import Case
import mlgi
import torch
from test_gpu_parity import Gpu
case = Case(160, 100, 64, 128, 128, cell=1.5, clutter=6)
for kw in (dict(), dict(class_lo=4, class_hi=8)):
    g = Gpu(case, **kw)
    for graph in (False, True):
        g.ctx.set_graph(graph)
        for mode in (mlgi.CULL_TILED, mlgi.CULL_CLUSTERED, mlgi.CULL_NONE):
            p = mlgi.default_params(case.N); p.cull_mode = mode
            g.ctx.frame(g.rsm, g.light, g.mesh, g.gb, g.cam, p)
            g.ctx.frame(g.rsm, g.light, g.mesh, g.gb, g.cam, p)
            g.ctx.gi_blur(g.gb, g.cam)
    print("listed/lit", g.ctx.gather_stats(g.gb, g.cam, p))
    g.close()
wl, sc = case.wl, case.wl.scene
g = Gpu(case)
p = mlgi.default_params(case.N)
for packed in (2, 0):
    shards = [mlgi.Context(case.W, case.H, case.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 22, class_lo=8 * r, class_hi=8 * r + 8, gi_packed=packed) for r in range(2)]
    for r, s in enumerate(shards):
        s.exchange_create(r, 2)
    for s in shards:
        s.exchange_connect_local(shards)
    streams = [torch.cuda.Stream() for _ in range(2)]
    for frame in range(2):
        for r, s in enumerate(shards):
            s.frame_sharded(g.rsm, g.light, g.mesh, g.gb, g.cam, p, stream=streams[r].cuda_stream)
        torch.cuda.synchronize()
    for s in shards:
        s.exchange_status(); s.close()
g.close()
print("sanitize run done")

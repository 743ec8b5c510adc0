"""debug: two class-shard contexts on ONE GPU through mlgi_frame_sharded_host; reports which pixels of a rank's rows differ"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import conftest  # noqa
from conftest import Case
import mlgi, torch
from test_gpu_parity import Gpu
c = Case(384, 216, 1024, 1024, 256, cell=0.5, clutter=12)
wl, sc = c.wl, c.wl.scene
full = Gpu(c); full.ctx.close()
full.ctx = mlgi.Context(c.W, c.H, c.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 22, gi_packed=1)
p = mlgi.default_params(c.N)
full.ctx.frame(full.rsm, full.light, full.mesh, full.gb, full.cam, p)
ref = full.ctx.download(mlgi.BUF_GI_PACKED, np.uint32).reshape(c.H, c.W)
world = 2
shards = [mlgi.Context(c.W, c.H, c.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 22, class_lo=8 * r, class_hi=8 * r + 8, gi_packed=2) for r in range(world)]
for r, s in enumerate(shards): s.exchange_create(r, world)
for s in shards: s.exchange_connect_local(shards)
rows = [s.sharded_row_mask() for s in shards]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(wl.rsm_diffuse), pin(wl.rsm_normal), pin(wl.rsm_depth), pin(wl.depth), pin(wl.normal)]
hr, hg = mlgi.make_rsm(h[0], h[1], h[2]), mlgi.make_gbuffer(h[3], h[4])
outs = [torch.zeros((c.H, c.W), dtype=torch.int32).pin_memory() for _ in range(world)]
streams = [torch.cuda.Stream() for _ in range(world)]
errors = []
def run(r, out):
    try:
        shards[r].frame_sharded_host(hr, full.light, full.mesh, hg, full.cam, p, host_gi_packed=out, stream=streams[r].cuda_stream)
    except Exception as e:
        errors.append((r, repr(e)))
import time
for frame in range(4):
    for o in outs: o.zero_()
    t0 = time.time()
    th = [threading.Thread(target=run, args=(r, outs[r])) for r in range(world)]
    for t in th: t.start()
    for t in th: t.join(timeout=120)
    dt = time.time() - t0
    print("frame", frame, "errors", errors, "seconds %.3f" % dt)
    for r in range(world):
        got = outs[r].numpy().view(np.uint32)
        bad = (got != ref) & rows[r][:, None]
        ys, xs = np.nonzero(bad)
        cls = (xs & 3) + 4 * (ys & 3)
        print(" rank", r, "bad", int(bad.sum()), "bad rows", sorted(set(ys.tolist()))[:12], "bad classes", sorted(set(cls.tolist())),
              "zeros among bad", int((got[bad] == 0).sum()), "written outside own rows", int(got[~rows[r]].any()))
    for r, s in enumerate(shards):
        try: s.exchange_status(); print(" status rank", r, "ok")
        except Exception as e: print(" status rank", r, repr(e))

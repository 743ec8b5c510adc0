"""Sweep of mlgi_frame_sharded_host settings inside ONE torchrun launch (the workload is built once): for every setting
"SLABS[:GRAPH[:XCH]]" given on the command line the rank re-creates its shard context with MLGI_HOST_SLABS / MLGI_HOST_GRAPH / MLGI_XCH
set, connects the exchange and times the end-to-end call (wall clock, max over ranks) and the device-resident mlgi_frame_sharded.
usage: torchrun --nproc-per-node N tools/sweep_host_frame.py 1 2 3 4:0 1:1:push"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200")); sys.path.insert(0, os.path.join(ROOT, "fixtures")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import mlgi, bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = bench.build_workload("4k_4096")
sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
dev = torch.device("cuda", local)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
d = [up(a) for a in (wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, wl.depth, wl.normal, sc.vertices, sc.normals, sc.indices, sc.prim_id)]
h = [pin(a) for a in (wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, wl.depth, wl.normal)]
host_gi = torch.zeros((H, W), dtype=torch.int32).pin_memory()
rsm, gb = mlgi.make_rsm(d[0], d[1], d[2]), mlgi.make_gbuffer(d[3], d[4])
hr, hg = mlgi.make_rsm(h[0], h[1], h[2]), mlgi.make_gbuffer(h[3], h[4])
mesh = mlgi.make_mesh(d[5], d[6], d[7], d[8], sc.num_vertices, sc.num_triangles)
light, cam = mlgi.Camera.make(wl.light_view, wl.light_proj), mlgi.Camera.make(wl.cam_view, wl.cam_proj)
p = mlgi.default_params(N)
per = 16 // world
steps = 30
for setting in sys.argv[1:]:
    f = setting.split(":")
    for k in ("MLGI_HOST_SLABS", "MLGI_HOST_GRAPH", "MLGI_XCH"):
        os.environ.pop(k, None)
    if f[0] not in ("", "auto"): os.environ["MLGI_HOST_SLABS"] = f[0]
    if len(f) > 1 and f[1] != "": os.environ["MLGI_HOST_GRAPH"] = f[1]
    if len(f) > 2: os.environ["MLGI_XCH"] = f[2]
    ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, device=local, class_lo=rank * per, class_hi=(rank + 1) * per, gi_packed=2)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.exchange_create(rank, world))
    ctx.exchange_connect(handles)
    res = {}
    for name, call in (("e2e", lambda: ctx.frame_sharded_host(hr, light, mesh, hg, cam, p, host_gi_packed=host_gi)),
                       ("dev", lambda: ctx.frame_sharded(rsm, light, mesh, gb, cam, p))):
        for _ in range(4): call()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        for _ in range(steps): call()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.time() - t0) / steps * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[name] = float(dt.item())
        dist.barrier()
    ctx.exchange_status()
    if rank == 0:
        print("N=%d slabs:graph:xch=%-10s e2e %.4f ms   device-resident %.4f ms" % (world, setting, res["e2e"], res["dev"]), flush=True)
    dist.barrier()
    ctx.close()
dist.destroy_process_group()

#!/usr/bin/env python
"""What ONE rank of a G-GPU class-sharded job does, measured on a single GPU (no exchange): frame time and per-stage times of
a context that owns interleave classes [16r/G, 16(r+1)/G). Used to study the multi-GPU critical path without multi-GPU
box time.   python tools/shard_sim.py [--world 8] [--rank 0] [--workload 4k_4096] [--steps 50]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--workload", default="4k_4096")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    import numpy as np, torch, mlgi
    wl = bench.build_workload(a.workload)
    sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
    dev = torch.device("cuda:0")
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    t = [up(wl.rsm_diffuse), up(wl.rsm_normal), up(wl.rsm_depth), up(wl.depth), up(wl.normal), up(sc.vertices), up(sc.normals), up(sc.indices), up(sc.prim_id)]
    per = 16 // a.world
    ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, class_lo=a.rank * per, class_hi=(a.rank + 1) * per, gi_packed=2)
    p = mlgi.default_params(N)
    rsm, gb = mlgi.make_rsm(t[0], t[1], t[2]), mlgi.make_gbuffer(t[3], t[4])
    mesh = mlgi.make_mesh(t[5], t[6], t[7], t[8], sc.num_vertices, sc.num_triangles)
    light, cam = mlgi.Camera.make(wl.light_view, wl.light_proj), mlgi.Camera.make(wl.cam_view, wl.cam_proj)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        ctx.frame(rsm, light, mesh, gb, cam, p, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ctx.frame(rsm, light, mesh, gb, cam, p, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ctx.set_timers(True)
    acc = {}
    for _ in range(20):
        ctx.frame(rsm, light, mesh, gb, cam, p, stream=st)
        torch.cuda.synchronize()
        for k, (v, n) in ctx.timings().items():
            acc[k] = acc.get(k, 0.0) + v / 20
    print(json.dumps({"world": a.world, "rank": a.rank, "workload": a.workload, "frame_ms_no_exchange": ms,
                      "ideal_ms": None, "stage_ms": {k: round(v, 4) for k, v in acc.items()}}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE config 4: 64-frame animated-light batch at 4K, 4096 VPLs (the ISMs are rebuilt every frame as the sun
moves, GIStage.cpp:357-365). Frames are independent, so N GPUs take every N-th frame (replicas, no exchange) and, on
each GPU, --depth contexts on separate streams let the ISM build of frame i+1 run beside the gathering of frame i.
  python tools/bench_anim.py [--frames 64] [--depth 2]            (torchrun for N > 1)
Prints one JSON line on rank 0: frames/s and Gpixel*VPL/s over the batch."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200")); sys.path.insert(0, os.path.join(ROOT, "fixtures"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import numpy as np, torch, mlgi
    from scene_fixture import Workload
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W, H, N = 3840, 2160, 4096
    wl = Workload(W, H, N)
    sc = wl.scene
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mine = list(range(rank, args.frames, world))
    rsms, lights = [], []
    for i in mine:                                   # sunCyclePosition += sunCycleSpeed (0.1) per frame, GIStage.cpp:229,362-365
        wl.set_sun(266.0 + 0.1 * i)
        t = (up(wl.rsm_diffuse), up(wl.rsm_normal), up(wl.rsm_depth))
        rsms.append((t, mlgi.make_rsm(*t)))
        lights.append(mlgi.Camera.make(wl.light_view, wl.light_proj))
    d_gb = (up(wl.depth), up(wl.normal))
    gb = mlgi.make_gbuffer(*d_gb)
    d_mesh = (up(sc.vertices), up(sc.normals), up(sc.indices), up(sc.prim_id))
    mesh = mlgi.make_mesh(*d_mesh, sc.num_vertices, sc.num_triangles)
    cam = mlgi.Camera.make(wl.cam_view, wl.cam_proj)
    params = mlgi.default_params(N)
    ctxs = [mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, device=local) for _ in range(args.depth)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.depth)]

    def batch():
        for k, i in enumerate(mine):
            s = k % args.depth
            ctxs[s].frame(rsms[k][1], lights[k], mesh, gb, cam, params, stream=streams[s].cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch(); barrier()
    best = 1e9
    for _ in range(args.repeat):
        barrier(); t0 = time.time(); batch(); barrier(); best = min(best, time.time() - t0)
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "animated-light batch throughput (4K, 4096 VPLs, ISM rebuild every frame)", "frames": args.frames,
                          "n_gpus": world, "streams_per_gpu": args.depth, "seconds": best, "frames_per_s": args.frames / best,
                          "ms_per_frame": 1e3 * best / args.frames, "gpixel_vpl_per_s": args.frames * W * H * (N // 16) / best / 1e9,
                          "sharding": "by frame (replicas, no exchange)"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

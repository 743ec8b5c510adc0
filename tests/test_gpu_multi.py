"""Multi-GPU tests (need >= 2 GPUs; skipped otherwise -- the single-GPU box runs the same exchange kernels through
test_gpu_parity.py::test_exchange_kernels_with_two_sequential_shards_on_one_gpu): the peer-memory exchange (pull and push
modes, graph-replayed mlgi_frame_sharded), the host-input sharded frame and the NCCL all-gather path all reproduce the
single-GPU image bit for bit on every rank."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, packed, q):
    import sys
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    import conftest  # noqa: F401
    from conftest import Case
    import mlgi
    from mlgi import dist as mdist
    if mode in ("pull", "push"):
        os.environ["MLGI_XCH"] = mode
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if mode == "nccl" else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", rank)} if mode == "nccl" else {}))
    try:
        case = Case(384, 216, 1024, 1024, 256, cell=0.5, clutter=12)
        wl, sc = case.wl, case.wl.scene
        dev = torch.device("cuda", rank)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        t = [up(wl.rsm_diffuse), up(wl.rsm_normal), up(wl.rsm_depth), up(wl.depth), up(wl.normal), up(sc.vertices), up(sc.normals),
             up(sc.indices), up(sc.prim_id)]
        rsm, gb = mlgi.make_rsm(t[0], t[1], t[2]), mlgi.make_gbuffer(t[3], t[4])
        mesh = mlgi.make_mesh(t[5], t[6], t[7], t[8], sc.num_vertices, sc.num_triangles)
        light, cam = mlgi.Camera.make(wl.light_view, wl.light_proj), mlgi.Camera.make(wl.cam_view, wl.cam_proj)
        p = mlgi.default_params(case.N)
        # single-GPU reference image on this rank's GPU
        full = mlgi.Context(case.W, case.H, case.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 22, device=rank, gi_packed=packed)
        full.frame(rsm, light, mesh, gb, cam, p)
        ref = full.download(mlgi.BUF_GI_PACKED, np.uint32) if packed else full.download(mlgi.BUF_GI, np.float32)
        full.close()
        lo, hi = mdist.class_range(rank, world)
        ctx = mlgi.Context(case.W, case.H, case.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 22, device=rank, class_lo=lo, class_hi=hi,
                           gi_packed=packed)
        ok = True
        if mode in ("pull", "push", "host"):
            handles = [None] * world
            dist.all_gather_object(handles, ctx.exchange_create(rank, world))
            ctx.exchange_connect(handles)
        if mode in ("pull", "push"):
            for _ in range(3):                       # three frames: exercises the parity double-buffering
                ctx.frame_sharded(rsm, light, mesh, gb, cam, p)
            ctx.exchange_status()
        elif mode == "host":
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            h = [pin(wl.rsm_diffuse), pin(wl.rsm_normal), pin(wl.rsm_depth), pin(wl.depth), pin(wl.normal)]
            host_gi = torch.zeros((case.H, case.W), dtype=torch.int32).pin_memory()
            mine = ctx.sharded_row_mask()
            for _ in range(3):
                host_gi.zero_()
                ctx.frame_sharded_host(mlgi.make_rsm(h[0], h[1], h[2]), light, mesh, mlgi.make_gbuffer(h[3], h[4]), cam, p, host_gi_packed=host_gi)
                ok = ok and bool(np.array_equal(host_gi.numpy().view(np.uint32)[mine], ref.reshape(case.H, case.W)[mine]))
                ok = ok and not host_gi.numpy()[~mine].any()     # only this rank's rows are written
            ctx.exchange_status()
        else:
            PH, PW = (case.H + 3) // 4, (case.W + 3) // 4
            planar = torch.zeros((16, PH, PW) if packed else (16, PH, PW, 4), dtype=torch.int32 if packed else torch.float32, device=dev)
            ctx.bind_planar_output(planar)
            for _ in range(2):
                ctx.frame(rsm, light, mesh, gb, cam, p)
                mdist.allgather_planar(planar, rank, world)
                ctx.planar_to_raster(planar, None)
        torch.cuda.synchronize()
        got = ctx.download(mlgi.BUF_GI_PACKED, np.uint32) if packed else ctx.download(mlgi.BUF_GI, np.float32)
        if mode == "host":       # the host-input frame assembles only this rank's rows (checked above, in the host buffer)
            pass
        elif packed:
            ok = ok and bool(np.array_equal(got, ref))
        else:
            ok = ok and bool(np.array_equal(got.reshape(-1, 4)[:, :3], ref.reshape(-1, 4)[:, :3]))
        dist.barrier()
        ctx.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,packed", [("pull", True), ("pull", False), ("push", True), ("host", True), ("nccl", True)])
def test_two_gpu_exchange_is_bit_identical(mode, packed):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, packed, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {0: True, 1: True}

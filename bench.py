#!/usr/bin/env python
"""bench.py -- GI frame benchmark (BASELINE.json metric) for the B200-native pipeline.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement of the GLSL
                                                           # (the reference itself is GL-only and cannot run here)

A "step" is one GI frame of the hot path on one batch of synthetic input: VPL build, scene->points,
ISM splat, pull-push, tiled culling, final gathering (GIStage::process minus the RSM raster,
GIStage.cpp:346-418). Workload at every N: 3840x2160, 4096 VPLs / ISMs of 64^2 in a 4096^2 atlas,
~250k-triangle procedural atrium (BASELINE.json configs[2]); for N>1 the frame is sharded by
interleave class and assembled with one NCCL all-gather (strong scaling).

Prints ONE JSON line (rank 0). `value` = nominal interleaved (pixel,VPL) pairs per second over the
whole frame with inputs resident in HBM; `e2e` = the same through mlgi_frame_host (pinned host
G-buffers/RSM in, GI image out, copies inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200"))

WORKLOADS = {
    # name: (W, H, N, rsm_w, rsm_h, cell_size)
    "c1_256_64": (256, 256, 64, 128, 128, 0.155),
    "1080p_1024": (1920, 1080, 1024, 1024, 256, 0.155),
    "4k_4096": (3840, 2160, 4096, 1024, 256, 0.155),
    "8k_16384": (7680, 4320, 16384, 1024, 256, 0.155),
}
# FMA = 2, rsqrt/rcp = 1, min/max = 1
FLOP_LISTED, FLOP_LIT, FLOP_PIXEL = 26.0, 48.0, 60.0


PROFILE_FG, PROFILE_PP = "r1_final_gather_final", "r1_pullpush_final"   # committed ncu summaries the traffic figures come from


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        for ts, ln in self.lines:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name):
    from mlgi.scene import Workload
    W, H, N, rw, rh, cell = WORKLOADS[name]
    return Workload(W, H, N, rsm_w=rw, rsm_h=rh, cell_size=cell)


# ---------------------------------------------------------------------------------------
# CPU side: the oracle (a port of the reference GLSL) on the host cores. bench.py may touch oracle/
# ONLY here (cpu_baseline leg and --impl reference).
# ---------------------------------------------------------------------------------------
class CpuPort:
    def __init__(self, wl):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        self.O, self.wl = O, wl
        self.cores = O.use_all_cores()
        self.prepared = False

    def prepare(self):
        """all stages before final gathering, full size, timed once"""
        O, wl = self.O, self.wl
        sc, N = wl.scene, wl.N
        t = {}
        t0 = time.time(); self.vpl, self.packed = O.vpl_process(wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, N, wl.light_view, wl.light_proj); t["VPLP"] = time.time() - t0
        t0 = time.time(); pts, bkt = O.ism_points(sc.vertices, sc.normals, sc.indices, sc.prim_id, 2.0, N); pg, bs = O.group_by_bucket(pts, bkt, N); t["ISM render"] = time.time() - t0
        t0 = time.time(); atlas = O.ism_splat(pg, bs, self.packed, N, wl.z_far); t["ISM CS"] = time.time() - t0
        t0 = time.time(); self.ism = O.ism_pullpush(atlas, N, wl.z_far); t["PullPush"] = time.time() - t0
        t0 = time.time(); self.tiles = O.cull_tiled(wl.depth, wl.normal, self.vpl, N, wl.cam_view, wl.cam_proj, 64); t["Tiled cull"] = time.time() - t0
        self.t_other = sum(t.values())
        self.stage_s = t
        self.prepared = True

    def gather_rows(self, r0, r1):
        O, wl = self.O, self.wl
        t0 = time.time()
        O.final_gather(wl.depth, wl.normal, self.vpl, wl.N, self.ism, 1, None, self.tiles, wl.cam_view, wl.cam_proj, wl.z_far, rows=(r0, r1))
        return time.time() - t0

    def pick_rows(self, budget_s):
        """rows (multiple of 4 so every interleave class is hit) whose gathering fits budget_s"""
        H = self.wl.H
        mid = (H // 2) & ~3
        self.gather_rows(mid, mid + 4)                     # thread-pool warm-up
        probe = min(H - mid, 32) & ~3
        t4 = max(self.gather_rows(mid, mid + probe), 1e-4) * 4.0 / probe
        rows = int(max(4, min(H, (budget_s / t4) * 4))) & ~3
        r0 = max(0, ((H - rows) // 2) & ~3)
        return r0, min(H, r0 + rows)

    def frame_estimate(self, r0, r1, t_rows):
        wl = self.wl
        t_frame = self.t_other + t_rows * wl.H / float(r1 - r0)
        pairs = wl.W * wl.H * (wl.N // 16)
        return pairs / t_frame / 1e9, t_frame


def run_reference(args, rank, world):
    """reference arm: the reference's algorithm on the host CPU (oracle port; the GL original cannot run)."""
    if rank != 0:
        return
    wl = build_workload(args.workload)
    cpu = CpuPort(wl)
    cpu.prepare()
    per_step = max(0.5, 150.0 / float(args.steps + args.warmup))
    r0, r1 = cpu.pick_rows(per_step)
    for _ in range(args.warmup):
        cpu.gather_rows(r0, r1)
    ts = [cpu.gather_rows(r0, r1) for _ in range(args.steps)]
    t_rows = sum(ts) / len(ts)
    value, t_frame = cpu.frame_estimate(r0, r1, t_rows)
    sample = ("per step: final gathering of rows [%d,%d) of %d (all 16 classes), extrapolated to the frame; VPL/points/splat/"
              "pull-push/tiled-cull timed once at full size = %.2f s" % (r0, r1, wl.H, cpu.t_other))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_frame * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(args, wl, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


METRIC = "final-gather Gpixel*VPL/s over the whole GI frame (ISM build + pull-push + tiled culling + final gathering)"
UNIT = "Gpixel*VPL/s"


def bench_config(args, wl, world):
    return {"workload": "%s: %dx%d, %d VPLs, %d ISMs of 64^2 in a %d^2 atlas, RSM %dx%d, %d-triangle procedural atrium" % (
                args.workload, wl.W, wl.H, wl.N, wl.N, 64 * isqrt_pow2(wl.N), wl.rsm_w, wl.rsm_h, wl.scene.num_triangles),
            "cull": args.cull, "interleave": "4x4", "shadows": True,
            "parallelism": ("interleave-class shard x%d + %s" % (world, "fused peer-store exchange over NVLink" if args.exchange == "fused" else "NCCL all-gather")) if world > 1 else "single GPU",
            "l2": "no flush: per-step working set (G-buffer + atlas + ISM + tile lists + GI image) exceeds the 126 MB L2"}


def isqrt_pow2(n):
    l = 0
    while (1 << l) < n:
        l += 1
    return 1 << ((l + 1) // 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k_4096", choices=sorted(WORKLOADS))
    ap.add_argument("--cull", default="tiled", choices=["tiled", "clustered"])
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: gathering kernels store their planes into every peer over NVLink (fused) or one NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import mlgi
    from mlgi import dist as mdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    wl = build_workload(args.workload)
    sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = dict(rsm_diffuse=up(wl.rsm_diffuse), rsm_normal=up(wl.rsm_normal), rsm_depth=up(wl.rsm_depth), depth=up(wl.depth),
             normal=up(wl.normal), verts=up(sc.vertices), normals=up(sc.normals), indices=up(sc.indices), prim=up(sc.prim_id))
    per = 16 // world
    lo, hi = rank * per, (rank + 1) * per
    # gi_packed: besides the fp32 image the gathering kernel emits the reference's giBuffer format (GL_R11F_G11F_B10F,
    # GIStage.cpp:470); that is what crosses NVLink (multi-GPU) and PCIe (end-to-end call)
    ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, device=local, class_lo=lo, class_hi=hi, gi_packed=True)
    params = mlgi.default_params(N)
    params.cull_mode = mlgi.CULL_TILED if args.cull == "tiled" else mlgi.CULL_CLUSTERED
    rsm = mlgi.make_rsm(d["rsm_diffuse"], d["rsm_normal"], d["rsm_depth"])
    gb = mlgi.make_gbuffer(d["depth"], d["normal"])
    mesh = mlgi.make_mesh(d["verts"], d["normals"], d["indices"], d["prim"], sc.num_vertices, sc.num_triangles)
    light = mlgi.Camera.make(wl.light_view, wl.light_proj)
    cam = mlgi.Camera.make(wl.cam_view, wl.cam_proj)
    PH, PW = (H + 3) // 4, (W + 3) // 4
    planar = None
    fused = world > 1 and args.exchange == "fused"
    if fused:
        handles = [None] * world
        dist.all_gather_object(handles, ctx.exchange_create(rank, world))        # 64-byte CUDA IPC handles, once
        ctx.exchange_connect(handles)
    elif world > 1:
        planar = torch.zeros((16, PH, PW), dtype=torch.int32, device=dev)      # R11F_G11F_B10F class planes
        ctx.bind_planar_output(planar)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.frame(rsm, light, mesh, gb, cam, params, stream=stream)
        if fused:
            ctx.exchange_finish(stream=stream)           # flags in peer memory + de-interleave; the planes already crossed NVLink
        elif world > 1:
            mdist.allgather_planar(planar, rank, world)                                      # in place, NCCL over NVLink
            ctx.planar_to_raster(planar, None, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    pairs = W * H * (N // 16)
    value = pairs / (ms_step * 1e-3) / 1e9

    # ---- per-stage times + the dominant kernel's live duration (events on the launching stream) ----
    ctx.set_timers(True)
    acc = {}
    for _ in range(min(args.steps, 50)):
        step()
        torch.cuda.synchronize()
        for k, (ms, n) in ctx.timings().items():
            a = acc.setdefault(k, [0.0, 0, n])
            a[0] += ms; a[1] += 1
    ctx.set_timers(False)
    stages = {k: v[0] / v[1] for k, v in acc.items()}
    stage_launches = {k: v[2] for k, v in acc.items()}
    fg_ms = stages.get("FG", float("nan"))
    pp_ms = stages.get("PullPush", float("nan"))
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None

    # ---- end to end: host buffers through mlgi_frame_host ----
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h = [pin(wl.rsm_diffuse), pin(wl.rsm_normal), pin(wl.rsm_depth), pin(wl.depth), pin(wl.normal)]
        host_gi = torch.empty((H, W), dtype=torch.int32).pin_memory()          # R11F_G11F_B10F
        hr, hg = mlgi.make_rsm(h[0], h[1], h[2]), mlgi.make_gbuffer(h[3], h[4])
        n_e2e = max(5, min(args.steps, 50))
        if world == 1:
            def e2e_step():
                ctx.frame_host_packed(hr, light, mesh, hg, cam, params, host_gi_packed=host_gi, stream=stream)
        else:
            # N > 1: the job's inputs cross PCIe once in total. Every rank uploads the (small) RSM and ITS 1/N slab of the
            # camera G-buffer rows from pinned host memory, the slabs are all-gathered over NVLink (every rank needs the
            # whole G-buffer: its pixels are interleaved over the image), then frame + exchange, and every rank reads ITS
            # slab of the assembled image back -- the union over ranks is the full result on the host.
            assert H % world == 0, "e2e at N GPUs splits the image rows evenly"
            y0, y1 = rank * H // world, (rank + 1) * H // world
            gkeys = ("depth", "normal")
            # The G-buffer traffic (slab upload + all-gather) runs on a second stream beside the ISM build, which only needs
            # the RSM; culling and gathering start when both are done (the stage entry points of the C ABI, in the order
            # mlgi_frame runs them).
            s_io = torch.cuda.Stream(device=dev)
            def e2e_step():
                main = torch.cuda.current_stream()
                s_io.wait_stream(main)
                with torch.cuda.stream(s_io):
                    for k, src in zip(gkeys, h[3:]):
                        full = d[k].view(torch.uint8).view(H, -1)
                        full[y0:y1].copy_(src.view(torch.uint8).view(H, -1)[y0:y1], non_blocking=True)
                        dist.all_gather_into_tensor(full.view(-1), full[y0:y1].reshape(-1))
                for k, src in zip(("rsm_diffuse", "rsm_normal", "rsm_depth"), h[:3]):
                    d[k].copy_(src.view(d[k].dtype).reshape(d[k].shape), non_blocking=True)
                ctx.vpl_process(rsm, light, params, stream=stream)
                ctx.ism_points(mesh, params, stream=stream)
                ctx.ism_splat(params, stream=stream)
                ctx.ism_pullpush(params, stream=stream)
                main.wait_stream(s_io)
                if args.cull == "tiled":
                    ctx.cull_tiled(gb, cam, params, stream=stream)
                else:
                    ctx.cull_clustered(gb, cam, params, stream=stream)
                ctx.final_gather(gb, cam, params, stream=stream)
                if fused:
                    ctx.exchange_finish(stream=stream)
                else:
                    mdist.allgather_planar(planar, rank, world)
                    ctx.planar_to_raster(planar, None, stream=stream)
                ctx.read_async(mlgi.BUF_GI_PACKED, host_gi[y0:y1], (y1 - y0) * W * 4, offset=y0 * W * 4, stream=stream)
                torch.cuda.current_stream().synchronize()
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.time()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.time() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the end-to-end path must have produced the image of the device-resident path (same inputs): compare this rank's rows
        step()
        torch.cuda.synchronize()
        ref_img = ctx.download(mlgi.BUF_GI_PACKED, np.uint32).reshape(H, W)
        ra, rb = (0, H) if world == 1 else (y0, y1)
        e2e_ok = bool(np.array_equal(host_gi.numpy().view(np.uint32)[ra:rb], ref_img[ra:rb]))
        if world > 1:
            t = torch.tensor([1.0 if e2e_ok else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            e2e_ok = bool(t.item() > 0.5)
        if not e2e_ok:
            raise SystemExit("bench.py: the end-to-end path did not reproduce the device-resident frame")
        nb = lambda x: x.numel() * x.element_size()
        h2d = sum(nb(x) for x in h) if world == 1 else world * sum(nb(x) for x in h[:3]) + sum(nb(x) for x in h[3:])   # summed over ranks
        e2e = {"value": pairs / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_gi.numel() * 4,
               "ms_per_step": dt * 1e3, "steps": n_e2e, "verified_against_device_path": e2e_ok,
               "call": ("mlgi_frame_host_packed (pinned host RSM + G-buffer in, GL_R11F_G11F_B10F GI image out = the reference's giBuffer "
                        "format, stream sync)") if world == 1 else
                       "per rank: pinned-host upload of the RSM and of 1/N of the G-buffer rows + NVLink all-gather of the G-buffer (second stream, "
                       "beside the ISM build) + the stage calls of mlgi_frame + exchange + read-back of 1/N of the assembled "
                       "GL_R11F_G11F_B10F image rows (bytes summed over ranks)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (final gathering, FP32 pipe) and of pull-push (HBM) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    def ncu_traffic(name):
        """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` summary
        (profiles/<name>.metrics.csv). Only valid for the configuration the capture was taken on."""
        if args.workload != "4k_4096" or args.cull != "tiled" or world != 1:
            return None
        try:
            import csv
            m = {r[0]: (r[1], float(r[2])) for r in csv.reader(open(os.path.join(ROOT, "profiles", name + ".metrics.csv"))) if len(r) == 3 and r[0].startswith("dram__bytes")}
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return sum(v * scale[u] for u, v in m.values()) if len(m) == 2 else None
        except Exception:
            return None
    listed, lit = ctx.gather_stats(gb, cam, params)
    npx = (W * H * (hi - lo)) // 16
    fg_flops = FLOP_LISTED * listed + FLOP_LIT * lit + FLOP_PIXEL * npx
    fp32_peak = ctx.fp32_peak_tflops()
    fg_tflops = fg_flops / (fg_ms * 1e-3) / 1e12
    A = 64 * isqrt_pow2(N)
    pp_bytes = 6.0 * A * A * (hi - lo) / 16.0
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"kernel": "final_gather_kernel", "bound": "fp32", "achieved": fg_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": fg_tflops / fp32_peak if fp32_peak > 0 else None, "traffic": ncu_traffic(PROFILE_FG),
                "traffic_note": "DRAM bytes per launch (read + write) from profiles/%s.metrics.csv; the kernel is FP32-bound, its algorithmic bytes are 8 B/pixel in + 20 B/pixel out" % PROFILE_FG,
                "peak_source": "FFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP32-pipe figure; nominal 148*128*2*1.965 GHz = 74.4)",
                "algorithmic_flops_per_launch": fg_flops, "pairs_listed": listed, "pairs_lit": lit, "pixels": npx,
                "flop_model": "26*listed + 48*lit + 60*pixels (SURVEY 8d)", "kernel_ms": fg_ms,
                "share_of_step": fg_ms / ms_step}
    roofline_hbm = {"kernel": "pullpush_kernel", "bound": "hbm", "achieved": pp_bytes / (pp_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": pp_bytes / (pp_ms * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic(PROFILE_PP),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": pp_bytes, "kernel_ms": pp_ms}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu = CpuPort(wl)
        cpu.prepare()
        r0, r1 = cpu.pick_rows(12.0)
        t_rows = cpu.gather_rows(r0, r1)
        v, t_frame = cpu.frame_estimate(r0, r1, t_rows)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": "port",
                        "sample": "oracle (C++ port of the reference GLSL, OpenMP): VPL/points/splat/pull-push/tiled-cull at full size (%.2f s) + "
                                  "final gathering of rows [%d,%d) of %d (%.2f s), extrapolated to one frame = %.1f s" % (
                                      cpu.t_other, r0, r1, H, t_rows, t_frame),
                        "frame_s_estimate": t_frame, "stage_s": cpu.stage_s}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(args, wl, world), "gi_ms_per_frame": ms_step,
           "fg_gpixel_vpl_per_s": pairs * ((hi - lo) / 16.0) / (fg_ms * 1e-3) / 1e9 * world if world > 1 else pairs / (fg_ms * 1e-3) / 1e9,
           "stage_ms": stages, "stage_launches": stage_launches, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline}
    print(json.dumps(out), flush=True)
    sampler.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

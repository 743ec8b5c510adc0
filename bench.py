#!/usr/bin/env python
"""bench.py -- GI frame benchmark (BASELINE.json metric) for the B200-native pipeline.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement of the GLSL
                                                           # (the reference itself is GL-only and cannot run here)

A "step" is one GI frame of the hot path on one batch of synthetic input: VPL build, scene->points,
ISM splat, pull-push, tiled culling, final gathering (GIStage::process minus the RSM raster,
GIStage.cpp:346-418). Workload at every N: 3840x2160, 4096 VPLs / ISMs of 64^2 in a 4096^2 atlas,
~290k-triangle procedural atrium (BASELINE.json configs[2]); for N>1 the frame is sharded by
interleave class and assembled over NVLink peer memory (strong scaling).

Prints ONE JSON line (rank 0). `value` = nominal interleaved (pixel,VPL) pairs per second over the
whole frame with inputs resident in HBM; `e2e` = the same through the host-input entry point
(mlgi_frame_host_packed; mlgi_frame_sharded_host at N>1): pinned host G-buffers/RSM in, GI image out,
copies inside the timed region. `image_sha256` = digest of the assembled GL_R11F_G11F_B10F image; at N>1
the run FAILS unless it equals the image a 1-GPU context produces on the same rank.
Other workloads: --workload anim64 (BASELINE configs[4]), 8k_16384 (configs[3]), 1080p_1024 (configs[1]);
--cull clustered (configs[2] "tiled vs clustered").
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200"))
sys.path.insert(0, os.path.join(ROOT, "fixtures"))

WORKLOADS = {
    # name: (W, H, N, rsm_w, rsm_h, cell_size)
    "c1_256_64": (256, 256, 64, 128, 128, 0.155),
    "1080p_1024": (1920, 1080, 1024, 1024, 256, 0.155),
    "4k_4096": (3840, 2160, 4096, 1024, 256, 0.155),
    "8k_16384": (7680, 4320, 16384, 1024, 256, 0.155),
    "anim64": (3840, 2160, 4096, 1024, 256, 0.155),     # 64 sun positions, ISM rebuild every frame (GIStage.cpp:357-365)
}
ANIM_FRAMES = 64
# FMA = 2, rsqrt/rcp = 1, min/max = 1
FLOP_LISTED, FLOP_LIT, FLOP_PIXEL = 26.0, 48.0, 60.0

PROFILE_FG, PROFILE_PP = "r2_final_gather", "r2_pullpush"   # committed ncu summaries the traffic figures come from

METRIC = "final-gather Gpixel*VPL/s over the whole GI frame (ISM build + pull-push + tiled culling + final gathering)"
UNIT = "Gpixel*VPL/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML inside this process at ~1 kHz
    (a 20-step timed region lasts tens of milliseconds), nvidia-smi at 50 ms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.samples, self.stop_flag, self.source = device, None, [], False, None
        self.sm_max = None

    def _nvml_loop(self):
        import pynvml as nv
        h = self.handle
        R = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.time(), float(mhz), [k for k, b in R.items() if bits & b]))
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except (ValueError, IndexError):
                    idx = self.device
            self.handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.source = "nvml@1kHz"
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi@50ms"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz = float(f[1]); self.sm_max = max(self.sm_max or 0.0, float(f[2]))
            except ValueError:
                continue
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]) if v.lower().startswith("active")]
            self.samples.append((time.time(), mhz, reasons))

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, reasons = [], set()
        for ts, mhz, rs in list(self.samples):
            if ts < t0 or ts > t1:
                continue
            sm.append(mhz)
            reasons.update(rs)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(reasons), "samples": len(sm),
                "source": self.source}


def bind_to_gpu_numa_node(device):
    """N > 1, one process per GPU: run this rank on the CPU cores NVML reports as local to its GPU, BEFORE the pinned host buffers are
    allocated (first touch), so that the rank's uploads and read-backs do not cross the socket interconnect. Eight ranks uploading at
    once share the host's memory bandwidth; a host application would place its staging buffers the same way. MLGI_BENCH_BIND=0 disables.
    Returns the number of cores bound to, or None."""
    if os.environ.get("MLGI_BENCH_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device]) if vis else device
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 64
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cores = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1} & allowed
        if cores and cores != allowed:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        return None
    return None


def build_workload(name):
    from scene_fixture import Workload
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


def isqrt_pow2(n):
    l = 0
    while (1 << l) < n:
        l += 1
    return 1 << ((l + 1) // 2)


def bench_config(args, wl, world):
    xch = {"pull": "one exchange kernel pulling the class planes over NVLink peer memory",
           "push": "gathering epilogue storing the class planes into every peer over NVLink", "nccl": "NCCL all-gather"}[args.exchange]
    cfg = {"workload": "%s: %dx%d, %d VPLs, %d ISMs of 64^2 in a %d^2 atlas, RSM %dx%d, %d-triangle procedural atrium" % (
               args.workload, wl.W, wl.H, wl.N, wl.N, 64 * isqrt_pow2(wl.N), wl.rsm_w, wl.rsm_h, wl.scene.num_triangles),
           "cull": args.cull, "interleave": "4x4", "shadows": True,
           "parallelism": ("interleave-class shard x%d + %s" % (world, xch)) if world > 1 else "single GPU",
           "l2": "no flush: per-step working set (G-buffer + atlas + ISM + tile lists + GI image) exceeds the 126 MB L2"}
    if args.workload == "anim64":
        cfg["animation"] = ("%d sun positions (sunCyclePosition += 0.1 per frame, GIStage.cpp:229,362-365), step i renders frame i %% %d; "
                            "the RSMs are resident in HBM" % (ANIM_FRAMES, ANIM_FRAMES))
        if world > 1 and args.anim_shard == "frame":
            cfg["parallelism"] = "whole frames round-robin over %d replicas, no exchange" % world
    return cfg


class Job:
    """one workload on this rank: device inputs, ctx (class shard at world > 1), exchange, step functions"""

    def __init__(self, args, wl, rank, world, local, dist, replicas=False):
        import numpy as np
        import torch
        import mlgi
        from mlgi import dist as mdist
        self.np, self.torch, self.mlgi, self.mdist, self.dist = np, torch, mlgi, mdist, dist
        self.args, self.wl, self.rank, self.world, self.local = args, wl, rank, world, local
        self.dev = torch.device("cuda", local)
        sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
        up = self.up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.d = dict(rsm_diffuse=up(wl.rsm_diffuse), rsm_normal=up(wl.rsm_normal), rsm_depth=up(wl.rsm_depth), depth=up(wl.depth),
                      normal=up(wl.normal), verts=up(sc.vertices), normals=up(sc.normals), indices=up(sc.indices), prim=up(sc.prim_id))
        self.sharded = world > 1 and not replicas
        per = 16 // world if self.sharded else 16
        self.lo, self.hi = (rank * per, (rank + 1) * per) if self.sharded else (0, 16)
        # gi_packed = 2: the gathering kernel emits only the reference's giBuffer format (GL_R11F_G11F_B10F, GIStage.cpp:470);
        # that is what crosses NVLink (multi-GPU) and PCIe (end-to-end call)
        self.ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, device=local, class_lo=self.lo, class_hi=self.hi, gi_packed=2)
        self.params = mlgi.default_params(N)
        self.params.cull_mode = mlgi.CULL_TILED if args.cull == "tiled" else mlgi.CULL_CLUSTERED
        d = self.d
        self.rsm = mlgi.make_rsm(d["rsm_diffuse"], d["rsm_normal"], d["rsm_depth"])
        self.gb = mlgi.make_gbuffer(d["depth"], d["normal"])
        self.mesh = mlgi.make_mesh(d["verts"], d["normals"], d["indices"], d["prim"], sc.num_vertices, sc.num_triangles)
        self.light = mlgi.Camera.make(wl.light_view, wl.light_proj)
        self.cam = mlgi.Camera.make(wl.cam_view, wl.cam_proj)
        self.PH, self.PW = (H + 3) // 4, (W + 3) // 4
        self.planar = None
        self.peer = self.sharded and args.exchange in ("pull", "push")
        if self.peer:
            handles = [None] * world
            dist.all_gather_object(handles, self.ctx.exchange_create(rank, world))        # 64-byte CUDA IPC handles, once
            self.ctx.exchange_connect(handles)
        elif self.sharded:
            self.planar = torch.zeros((16, self.PH, self.PW), dtype=torch.int32, device=self.dev)      # R11F_G11F_B10F class planes
            self.ctx.bind_planar_output(self.planar)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.anim = None

    def prepare_animation(self, frames):
        """BASELINE configs[4]: one RSM + light camera per sun position, resident in HBM"""
        wl, mlgi = self.wl, self.mlgi
        self.anim = []
        for i in range(frames):
            wl.set_sun(266.0 + 0.1 * i)
            t = (self.up(wl.rsm_diffuse), self.up(wl.rsm_normal), self.up(wl.rsm_depth))
            self.anim.append((t, mlgi.make_rsm(*t), mlgi.Camera.make(wl.light_view, wl.light_proj)))
        wl.set_sun(266.0)

    def step(self, i=0):
        ctx = self.ctx
        rsm, light = (self.rsm, self.light) if self.anim is None else self.anim[i % len(self.anim)][1:]
        if self.peer:
            ctx.frame_sharded(rsm, light, self.mesh, self.gb, self.cam, self.params, stream=self.stream)   # one graph launch: stages + exchange kernel
        else:
            ctx.frame(rsm, light, self.mesh, self.gb, self.cam, self.params, stream=self.stream)
            if self.sharded:
                self.mdist.allgather_planar(self.planar, self.rank, self.world)                          # in place, NCCL over NVLink
                ctx.planar_to_raster(self.planar, None, stream=self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup, step=None):
        """W warm-up steps, then K steps between CUDA events on the launching stream, barrier + synchronize on both sides,
        max over ranks. Returns (ms per step, wall t0, wall t1, launches, mean host enqueue us)."""
        torch = self.torch
        step = step or self.step
        for i in range(warmup):
            step(i)
        self.barrier()
        l0 = self.ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        enq = 0.0
        t0 = time.time()
        e0.record()
        for i in range(steps):
            step(warmup + i)
            enq += self.ctx.frame_stats()["host_enqueue_us"]
        e1.record()
        self.barrier()
        t1 = time.time()
        ms_total = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total / steps, t0, t1, self.ctx.launch_count() - l0, enq / steps

    def image_digest(self):
        """sha256 of the assembled GL_R11F_G11F_B10F image this rank holds"""
        img = self.ctx.download(self.mlgi.BUF_GI_PACKED, self.np.uint32)
        return hashlib.sha256(img.tobytes()).hexdigest(), img

    def verify_against_single_gpu(self, img):
        """N > 1: the assembled image must equal, bit for bit, what an unsharded context produces on this rank's GPU for
        the same inputs (SURVEY 8e "bit-identical"); returns (ok on every rank, digest of the 1-GPU image)"""
        mlgi, np, torch, wl = self.mlgi, self.np, self.torch, self.wl
        sc = wl.scene
        full = mlgi.Context(wl.W, wl.H, wl.N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23, device=self.local, gi_packed=2)
        full.frame(self.rsm, self.light, self.mesh, self.gb, self.cam, self.params, stream=self.stream)
        torch.cuda.synchronize()
        ref = full.download(mlgi.BUF_GI_PACKED, np.uint32)
        full.close()
        ok = bool(np.array_equal(ref, img))
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5), hashlib.sha256(ref.tobytes()).hexdigest()

    def close(self):
        self.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k_4096", choices=sorted(WORKLOADS))
    ap.add_argument("--cull", default="tiled", choices=["tiled", "clustered"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "pull", "push", "nccl"],
                    help="N>1: one exchange kernel pulls the class planes over NVLink peer memory (pull), the gathering epilogue stores "
                         "them into every peer (push), the library's default (auto = push), or one NCCL all-gather")
    ap.add_argument("--anim-shard", default="class", choices=["class", "frame"],
                    help="anim64 at N>1: shard every frame by interleave class (default) or hand whole frames to replicas")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scale-8k", action="store_true", help="skip the extra 8K/16384 measurement (BASELINE configs[3]) added to the 4k_4096 line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.exchange == "auto":
        args.exchange = os.environ.get("MLGI_XCH") or "push"       # the library's own default
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import mlgi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if world > 1 and args.exchange in ("pull", "push"):
        os.environ["MLGI_XCH"] = args.exchange
    torch.cuda.set_device(local)
    cpu_bind = bind_to_gpu_numa_node(local) if world > 1 else None
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    wl = build_workload(args.workload)
    W, H, N = wl.W, wl.H, wl.N
    replicas = args.workload == "anim64" and args.anim_shard == "frame" and world > 1
    job = Job(args, wl, rank, world, local, dist, replicas=replicas)
    ctx, params, gb, cam = job.ctx, job.params, job.gb, job.cam
    lo, hi = job.lo, job.hi
    if args.workload == "anim64":
        job.prepare_animation(ANIM_FRAMES)
    step = job.step
    if replicas:                         # rank r renders frames r, r + world, ... : `world` frames per step in the whole job
        step = lambda i=0: job.step(i * world + rank)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, t_wall0, t_wall1, launches, enqueue_us = job.timed(args.steps, args.warmup, step)
    frames_per_step = world if replicas else 1
    pairs = W * H * (N // 16)
    value = frames_per_step * pairs / (ms_step * 1e-3) / 1e9
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
    frame_stats = ctx.frame_stats()

    # ---- the image: digest, and at N > 1 bit-identity with a 1-GPU context on this rank ----
    anim_saved, job.anim = job.anim, None                 # the static-light frame (frame 0 of the animation)
    job.step()
    torch.cuda.synchronize()
    digest, img = job.image_digest()
    single_gpu_ok, single_digest = True, digest
    if job.sharded:
        single_gpu_ok, single_digest = job.verify_against_single_gpu(img)
        if not single_gpu_ok:
            raise SystemExit("bench.py: the %d-GPU image differs from the 1-GPU image (sha256 %s vs %s on rank %d)" % (world, digest, single_digest, rank))
    job.anim = anim_saved

    # ---- end to end: host buffers through the host-input entry points ----
    e2e = None
    if not args.no_e2e and not replicas and (not job.sharded or job.peer):
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h = [pin(wl.rsm_diffuse), pin(wl.rsm_normal), pin(wl.rsm_depth), pin(wl.depth), pin(wl.normal)]
        host_gi = torch.zeros((H, W), dtype=torch.int32).pin_memory()          # R11F_G11F_B10F
        hr, hg = mlgi.make_rsm(h[0], h[1], h[2]), mlgi.make_gbuffer(h[3], h[4])
        n_e2e = max(5, min(args.steps, 50))
        rowmask = np.ones(H, dtype=bool)
        if not job.sharded:
            def e2e_step():
                ctx.frame_host_packed(hr, job.light, job.mesh, hg, cam, params, host_gi_packed=host_gi, stream=job.stream)
        else:
            # N > 1, ONE library call per rank and frame: upload of the RSM and of this rank's 1/N of the G-buffer rows (its tile
            # rows t mod N) from pinned host memory, tile-bounds exchange, gathering with remote G-buffer reads over NVLink, plane
            # exchange, read-back of this rank's 1/N of the assembled image rows, pipelined in phases -- the union over ranks is
            # the full result on the host
            rowmask = ctx.sharded_row_mask()
            def e2e_step():
                ctx.frame_sharded_host(hr, job.light, job.mesh, hg, cam, params, host_gi_packed=host_gi, stream=job.stream)
        for _ in range(3):
            e2e_step()
        job.barrier()
        t0 = time.time()
        for _ in range(n_e2e):
            e2e_step()
        job.barrier()
        dt = (time.time() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the end-to-end path must have produced the image of the device-resident path (same inputs): this rank's rows
        e2e_ok = bool(np.array_equal(host_gi.numpy().view(np.uint32)[rowmask], img.reshape(H, W)[rowmask]))
        if world > 1:
            t = torch.tensor([1.0 if e2e_ok else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            e2e_ok = bool(t.item() > 0.5)
        if not e2e_ok:
            raise SystemExit("bench.py: the end-to-end path did not reproduce the device-resident frame")
        nb = lambda x: x.numel() * x.element_size()
        # bytes that cross PCIe host->device per step, summed over ranks: the library uploads the camera G-buffer once in total and, per
        # rank, the N RSM texels the VPL build samples (12 B per VPL, gathered on the host from the caller's RSM planes)
        h2d = world * wl.N * 12 + sum(nb(x) for x in h[3:])
        e2e = {"value": pairs / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_gi.numel() * 4,
               "ms_per_step": dt * 1e3, "steps": n_e2e, "verified_against_device_path": e2e_ok,
               "call": ("mlgi_frame_host_packed (pinned host RSM + G-buffer in, GL_R11F_G11F_B10F GI image out = the reference's giBuffer "
                        "format, stream sync)") if world == 1 else
                       "mlgi_frame_sharded_host, one call per rank: pinned-host upload of the RSM and of 1/N of the G-buffer (whole tile rows), "
                       "tile bounds exchanged between ranks, gathering reads each tile's pixels from the rank holding them over NVLink, plane "
                       "exchange, 1/N of the assembled GL_R11F_G11F_B10F image rows copied back to pinned host memory; upload, culling, gathering and "
                       "read-back pipelined over phases of tile rows (bytes summed over ranks)"}
        if world > 1:
            e2e["host_cores_bound_to_gpu_numa_node"] = cpu_bind      # bind_to_gpu_numa_node (None: not bound)

    # ---- per-stage times + the dominant kernels' live durations (events on the launching stream; direct launches) ----
    ctx.set_timers(True)
    acc = {}
    for _ in range(min(args.steps, 50)):
        ctx.frame(job.rsm, job.light, job.mesh, gb, cam, params, stream=job.stream)
        ctx.gi_blur(gb, cam, stream=job.stream)          # n1, the consumer right after the path (not part of the GI frame metric)
        torch.cuda.synchronize()
        for k, (ms, n) in ctx.timings().items():
            a = acc.setdefault(k, [0.0, 0, n])
            a[0] += ms; a[1] += 1
    ctx.set_timers(False)
    stages = {k: v[0] / v[1] for k, v in acc.items()}
    stage_launches = {k: v[2] for k, v in acc.items()}
    fg_ms = stages.get("FG", float("nan"))
    pp_ms = stages.get("PullPush", float("nan"))
    listed, lit = ctx.gather_stats(gb, cam, params)       # exact work counters of this rank's gathering launch
    fp32_peak = ctx.fp32_peak_tflops() if rank == 0 else 0.0
    job.close()

    # ---- BASELINE configs[3] beside the headline: 8K / 16384 VPLs on the same ranks (the ">= 6x at 8 GPUs" configuration) ----
    scale_8k = None
    if args.workload == "4k_4096" and not args.no_scale_8k:
        a8 = argparse.Namespace(**vars(args))
        a8.workload = "8k_16384"
        wl8 = build_workload("8k_16384")
        job8 = Job(a8, wl8, rank, world, local, dist)
        k8 = max(3, min(args.steps, 10))
        ms8, _, _, _, _ = job8.timed(k8, 3)
        d8, img8 = job8.image_digest()
        ok8 = True if not job8.sharded else job8.verify_against_single_gpu(img8)[0]
        if not ok8:
            raise SystemExit("bench.py: the %d-GPU 8K image differs from the 1-GPU image" % world)
        p8 = wl8.W * wl8.H * (wl8.N // 16)
        scale_8k = {"workload": bench_config(a8, wl8, world)["workload"], "ms_per_frame": ms8, "value": p8 / (ms8 * 1e-3) / 1e9, "unit": UNIT,
                    "steps": k8, "warmup": 3, "image_sha256": d8, "matches_single_gpu_image": ok8}
        job8.close()

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

    npx = (W * H * (hi - lo)) // 16
    fg_flops = FLOP_LISTED * listed + FLOP_LIT * lit + FLOP_PIXEL * npx
    fg_tflops = fg_flops / (fg_ms * 1e-3) / 1e12
    A = 64 * isqrt_pow2(N)
    pp_bytes = 6.0 * A * A * (hi - lo) / 16.0
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"kernel": "final_gather_kernel", "bound": "fp32", "achieved": fg_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": fg_tflops / fp32_peak if fp32_peak > 0 else None, "traffic": ncu_traffic(PROFILE_FG),
                "traffic_note": "DRAM bytes per launch (read + write) from profiles/%s.metrics.csv; the kernel is FP32-bound, its algorithmic bytes are 8 B/pixel in + 4 B/pixel out (packed image only)" % PROFILE_FG,
                "peak_source": "FFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP32-pipe figure; nominal 148*128*2*1.965 GHz = 74.4)",
                "algorithmic_flops_per_launch": fg_flops, "pairs_listed": listed, "pairs_lit": lit, "pixels": npx,
                "flop_model": "26*listed + 48*lit + 60*pixels (SURVEY 8d)", "kernel_ms": fg_ms,
                "share_of_step": fg_ms / ms_step}
    roofline_hbm = {"kernel": "pullpush_kernel", "bound": "hbm", "achieved": pp_bytes / (pp_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": pp_bytes / (pp_ms * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic(PROFILE_PP),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": pp_bytes, "kernel_ms": pp_ms, "share_of_step": pp_ms / ms_step}

    blur_ms = stages.get("GI blur")
    blur_bytes = float(W) * H * (4 + 4 + 4 + 16)        # packed GI + depth + normal in, fp32 RGBA out
    roofline_blur = None if not blur_ms else {"kernel": "gi_blur_fused_kernel", "bound": "hbm", "achieved": blur_bytes / (blur_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                              "unit": "GB/s", "frac": blur_bytes / (blur_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                              "algorithmic_bytes_per_launch": blur_bytes, "kernel_ms": blur_ms,
                                              "note": "n1 (gi_blur.frag x2 fused); runs after the GI frame, not inside the timed region"}

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
           "config": bench_config(args, wl, world), "gi_ms_per_frame": ms_step / frames_per_step,
           "fg_gpixel_vpl_per_s": pairs / (fg_ms * 1e-3) / 1e9,       # all ranks gather concurrently: whole-job pairs over one rank's launch
           "image_sha256": digest, "single_gpu_image_sha256": single_digest, "matches_single_gpu_image": single_gpu_ok,
           "stage_ms": stages, "stage_launches": stage_launches, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
           "host_enqueue_us_per_frame": enqueue_us, "graph": frame_stats,
           "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_blur": roofline_blur, "cpu_baseline": cpu_baseline, "scale_8k": scale_8k}
    if args.workload == "anim64":
        out["frames_per_s"] = frames_per_step / (ms_step * 1e-3)
    print(json.dumps(out), flush=True)
    sampler.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

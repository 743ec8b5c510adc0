"""Does the texture unit floor an unnormalised point-sampled coordinate exactly? Renders the 4K frame with the exact-integer
coordinate (MLGI_ISM_FETCH=tex) and with the unfloored coordinate (texf) and counts the pixels whose bits differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200")); sys.path.insert(0, os.path.join(ROOT, "fixtures")); sys.path.insert(0, ROOT)
import torch, mlgi, bench
name = sys.argv[1] if len(sys.argv) > 1 else "4k_4096"
other = sys.argv[2] if len(sys.argv) > 2 else "texf"        # texf | texi | ldg
wl = bench.build_workload(name); sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = [up(a) for a in (wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, wl.depth, wl.normal, sc.vertices, sc.normals, sc.indices, sc.prim_id)]
imgs = {}
for mode in ("tex", other):
    os.environ["MLGI_ISM_FETCH"] = mode
    ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23)
    p = mlgi.default_params(N)
    args = (mlgi.make_rsm(t[0], t[1], t[2]), mlgi.Camera.make(wl.light_view, wl.light_proj), mlgi.make_mesh(t[5], t[6], t[7], t[8], sc.num_vertices, sc.num_triangles),
            mlgi.make_gbuffer(t[3], t[4]), mlgi.Camera.make(wl.cam_view, wl.cam_proj), p)
    for _ in range(5): ctx.frame(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ctx.frame(*args)
    e1.record(); torch.cuda.synchronize()
    imgs[mode] = ctx.download(mlgi.BUF_GI, np.float32).reshape(H, W, 4)[:, :, :3].copy()
    print(mode, "frame %.4f ms" % (e0.elapsed_time(e1) / 50))
    ctx.close()
a, b = imgs["tex"], imgs[other]
diff = (a.view(np.uint32) != b.view(np.uint32)).any(-1)
rel = np.abs(a - b).max(-1) / np.maximum(np.abs(a).max(-1), 1e-12)
print("pixels differing: %d of %d (%.4f %%), max relative difference %.3e, pixels above 1e-3: %d" % (diff.sum(), diff.size, 100.0 * diff.mean(), rel.max(), (rel > 1e-3).sum()))

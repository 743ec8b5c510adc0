"""Render one GI frame on the GPU and write viewable images (PIL): the raw interleaved GI image, the blurred one, the
ambient composite (deferredshading.frag ambient term) and a crop of the ISM atlas. usage: render_frame.py [workload] [outdir]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "many-lights-gi_b200")); sys.path.insert(0, os.path.join(ROOT, "fixtures")); sys.path.insert(0, ROOT)
import torch
from PIL import Image
import mlgi
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "1080p_1024"
outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
os.makedirs(outdir, exist_ok=True)
wl = bench.build_workload(name)
sc, W, H, N = wl.scene, wl.W, wl.H, wl.N
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = [up(a) for a in (wl.rsm_diffuse, wl.rsm_normal, wl.rsm_depth, wl.depth, wl.normal, sc.vertices, sc.normals, sc.indices, sc.prim_id, wl.diffuse)]
ctx = mlgi.Context(W, H, N, wl.rsm_w, wl.rsm_h, sc.num_triangles, 1 << 23)
p = mlgi.default_params(N)
gb = mlgi.make_gbuffer(t[3], t[4]); cam = mlgi.Camera.make(wl.cam_view, wl.cam_proj)
ctx.frame(mlgi.make_rsm(t[0], t[1], t[2]), mlgi.Camera.make(wl.light_view, wl.light_proj),
          mlgi.make_mesh(t[5], t[6], t[7], t[8], sc.num_vertices, sc.num_triangles), gb, cam, p)
ctx.gi_blur(gb, cam)
out = torch.zeros((H, W), dtype=torch.int32, device=dev)
def save(arr_u32, fn, scale=None):
    rgb = np.stack([(arr_u32 >> s) & 0xFF for s in (0, 8, 16)], -1).astype(np.uint8)
    im = Image.fromarray(rgb)
    if scale: im = im.resize((im.width // scale, im.height // scale), Image.BOX)
    im.save(os.path.join(outdir, fn)); print("wrote", fn, im.size)
shrink = max(1, W // 960)
for blurred, fn in ((True, "frame_%s_composite.png" % name), (False, "frame_%s_composite_unblurred.png" % name)):
    ctx.composite_ambient(t[9], out, exposure=1.0, use_blurred=blurred)
    torch.cuda.synchronize()
    save(out.cpu().numpy().view(np.uint32), fn, shrink)
# GI alone (white diffuse), tonemapped the same way
white = torch.full((H, W), 0x00FFFFFF, dtype=torch.int32, device=dev)
ctx.composite_ambient(white, out, exposure=1.0, use_blurred=True); torch.cuda.synchronize()
save(out.cpu().numpy().view(np.uint32), "frame_%s_gi.png" % name, shrink)
save(wl.diffuse, "frame_%s_diffuse.png" % name, shrink)
ism = ctx.download(mlgi.BUF_ISM, np.uint16)
A = int(round(len(ism) ** 0.5))
crop = (ism.reshape(A, A)[:512, :512] >> 8).astype(np.uint8)
Image.fromarray(crop).save(os.path.join(outdir, "frame_%s_ism_crop.png" % name)); print("wrote ism crop")
ctx.close()

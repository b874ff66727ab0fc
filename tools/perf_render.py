"""Inference (demo) throughput through the drop-in API: MC_Model(mode=1)(img_idx) renders whole 800x800 views in
`batch`-sized chunks (BASELINE configs[3]).  Reports rays/s per view incl. the final copy of the image to the host."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import synthetic as syn
from mc_nerf_b200.model import MC_Model

dev = "cuda:0"
img = int(sys.argv[1]) if len(sys.argv) > 1 else 800
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
views = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw = dict(n_cam=110, img_h=img, img_w=img, batch=batch, samples=64, scale=2, with_images=False)
sp = syn.make_sys_param(device=dev, **kw)
torch.manual_seed(0)
m = MC_Model(sp).to(dev)
with torch.no_grad():
    for k, v in syn.init_camera_weights(sp).items():
        getattr(m, k).copy_(v)
tmp = tempfile.mkdtemp()
m.nerf.weights_pth = tmp
m.nerf.save_model(m, 0)
sp2 = syn.make_sys_param(device=dev, mode=1, **kw)
sp2["demo_ckpt"] = m.nerf.file_path
demo = MC_Model(sp2).to(dev).eval()
with torch.no_grad():
    demo(torch.tensor([0]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for v in range(views):
        rgb, dep, opa = demo(torch.tensor([v + 1]))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / views
print(f"demo render {img}x{img}, chunk {batch}: {dt*1e3:.1f} ms/view  {img*img/dt/1e6:.2f} Mrays/s  "
      f"(110 views: {110*dt:.1f} s)")

from mc_nerf_b200._lib import lib
L = lib()
with torch.no_grad():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.profile_begin()
    demo(torch.tensor([5]))
    torch.cuda.synchronize()
    calls = [(n, e0.elapsed_time(e1)) for n, e0, e1 in L._prof if n == "mcnerf_mlp_tc_fwd"]
    prof = L.profile_end()
    wall = time.perf_counter() - t0
print("one profiled view: wall %.1f ms; kernel ms by entry: " % (wall * 1e3)
      + ", ".join(f"{k.replace('mcnerf_', '')} {v:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8])
      + f"; sum {sum(prof.values()):.1f}")
from mc_nerf_b200 import render
print("mlp_tc_fwd calls of the view (ms):", " ".join(f"{ms:.2f}" for _, ms in calls),
      "| fine rows selected in the last chunk:", int(render.LAST["n_rows_dev"].item()), "of capacity", render.LAST["n_rows"])

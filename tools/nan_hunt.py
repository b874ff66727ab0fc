"""Diagnostic: one small bf16 train render + backward; reports non-finite gradients (run under compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops, render, synthetic as syn
from mc_nerf_b200.model import MC_Model
DEV = "cuda"
kw = dict(n_cam=4, img_h=16, img_w=16, batch=96, samples=32, scale=2, coarse=(4, 256, (2,)), fine=(4, 256, (2,)))
sp = syn.make_sys_param(device=DEV, **kw)
sp["mlp_precision"] = "bf16"
sp["noise_sampler"] = "device"
torch.manual_seed(3)
m = MC_Model(sp).to(DEV)
B, Sc, Sf = 96, 32, 64
g = torch.Generator().manual_seed(1)
rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV).requires_grad_(True)
ro = (torch.randn(B, 3, generator=g) * 0.2).to(DEV).requires_grad_(True)
gt = torch.rand(B, 3, generator=g).to(DEV)
seed = torch.tensor([42, -7], dtype=torch.int64, device=DEV)
jitter = ops.philox_fill(seed, 4, B, normal=False, lo=0.0, hi=7.0 / Sc)
for rep in range(2):
    for p in list(m.nerf.parameters()) + [rd, ro]:
        p.grad = None
    rgb_c, rgb_f = m.nerf.render_rays_train(rd, ro, 25, 1.0, rng=dict(seed=seed, jitter=jitter))
    (((rgb_c - gt) ** 2).mean() + ((rgb_f - gt) ** 2).mean()).backward()
    torch.cuda.synchronize()
    print("rep", rep, "rgb finite", bool(torch.isfinite(rgb_c).all()), bool(torch.isfinite(rgb_f).all()), "n_rows_dev", render.LAST.get("n_rows_dev"))
    for (name, p) in list(m.nerf.named_parameters()) + [("rd", rd), ("ro", ro)]:
        bad = ~torch.isfinite(p.grad)
        if bad.any():
            idx = bad.nonzero()
            cols = sorted(set(idx[:, -1].tolist()))[:12] if idx.shape[1] > 1 else []
            rows = sorted(set(idx[:, 0].tolist()))[:8]
            print("  non-finite grad:", name, tuple(p.shape), "count", int(bad.sum()), "rows", rows, "cols", cols)

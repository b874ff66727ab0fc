"""Compositing kernels against the HBM roofline (north_star: warp-scan kernels with coalesced, vectorised HBM access).
Shapes: BASELINE configs[1] (4096 rays, 64+128 samples) and the configs[4] stress shape (65536 rays, 64+256 samples).
Algorithmic bytes per ray: forward out4 16 B/sample + noise 4 B/sample + outputs; backward reads out4,
noise, g_rgb and writes g_out4 16 B/sample."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
DEV = "cuda"
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {}
hbm = float(peaks.get("hbm_gbs", 6542.7))


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    tot = 0.0
    for _ in range(n):
        flush.zero_()                       # evict the 126 MB L2 between iterations
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


for B, S in ((4096, 192), (65536, 320), (262144, 192)):
    g = torch.Generator(device=DEV).manual_seed(0)
    out4 = torch.rand(B, S, 4, device=DEV, generator=g).requires_grad_(True)
    noise = torch.randn(B, S, device=DEV, generator=g)
    rays_d = torch.nn.functional.normalize(torch.randn(B, 3, device=DEV, generator=g), dim=-1)
    jitter = torch.rand(B, 1, device=DEV, generator=g) * 0.01
    res = {}
    def fwd():
        res["o"] = ops.CompositeFn.apply(out4, noise, rays_d, None, jitter, 1.0, 8.0, True)
    ms_f = timed(fwd)
    rgb = res["o"][0]
    grgb = torch.randn_like(rgb)
    def bwd():
        out4.grad = None
        res["o"] = ops.CompositeFn.apply(out4, noise, rays_d, None, jitter, 1.0, 8.0, True)
        res["o"][0].backward(grgb)
    ms_fb = timed(bwd)
    ms_b = ms_fb - ms_f
    bytes_f = B * (S * 20 + 12 + 4 + 20)          # out4 + noise in; rays_d, jitter in; rgb/depth/opacity out
    bytes_b = B * (S * 20 + S * 16 + 12 + 4 + 12)  # out4 + noise in, g_out4 out, rays_d, jitter, g_rgb
    print(f"B={B} S={S}: composite fwd {ms_f*1e3:.1f} us = {bytes_f/ms_f/1e6:.0f} GB/s ({bytes_f/ms_f/1e6/hbm*100:.0f}% of {hbm:.0f}); "
          f"bwd {ms_b*1e3:.1f} us = {bytes_b/ms_b/1e6:.0f} GB/s ({bytes_b/ms_b/1e6/hbm*100:.0f}%)")

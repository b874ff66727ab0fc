"""Inference forward kernel at demo-render sizes: dense sample grid vs an explicit sample index (the fine pass), to
separate the kernel's sustained rate from the cost of the gathered input stage.
usage: python tools/perf_mlp_infer_modes.py [rays] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from oracle import mcnerf_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
DEV = "cuda"
p = orc.init_mlp_params(8, 256, (4,), seed=3)
tensors = {k: p[k].to(DEV).contiguous() for k in ops.param_names(8)}
ps = ops.make_mlp_params(tensors, 8, 256, (4,))
tcw = ops.TcWeights().get(ps, tensors, False)
g = torch.Generator().manual_seed(0)
ro = (torch.randn(B, 3, generator=g) * 0.5).to(DEV)
rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)


def run(S, sel, label):
    smp = ops.make_sampling(1.0, 8.0, S, 10)
    n = B * S if sel is None else sel.shape[0]
    out = torch.empty(n, 4, device=DEV)
    tin = ops.make_tc_input_rays(ro, rd, None, smp, sel, n, None)
    for _ in range(2):
        ops.mlp_tc_fwd(ps, tcw, tin, out, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.mlp_tc_fwd(ps, tcw, tin, out, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{label}: rows={n} {ms:.3f} ms  {ms / n * 262144 * 1e3:.1f} us per 262144 rows  {2 * 629248 * n / ms / 1e9:.0f} TFLOP/s")


run(64, None, "dense coarse grid (S=64)          ")
run(128, None, "dense fine grid (S=128)           ")
full = torch.arange(B * 128, dtype=torch.int32, device=DEV)
run(128, full, "fine grid through an identity index")
keep = full[(torch.rand(B * 128, device=DEV) < 0.77)].contiguous()
run(128, keep, "fine grid, 77 % random selection    ")

# the same kernel on the rays of a real camera (one origin, a smooth fan of directions) and on the demo's weights
from mc_nerf_b200 import synthetic as syn
from mc_nerf_b200.model import MC_Model
sp = syn.make_sys_param(n_cam=110, img_h=800, img_w=800, batch=B, samples=64, scale=2, device=DEV, with_images=False)
torch.manual_seed(0)
m = MC_Model(sp).to(DEV)
with torch.no_grad():
    for k, v in syn.init_camera_weights(sp).items():
        getattr(m, k).copy_(v)
    rays_d, rays_o = m.get_rays(m.test_pose, torch.tensor([1]), m.intr_test_inv.to(DEV))
ro, rd = rays_o[:B].contiguous(), rays_d[:B].contiguous()
run(64, None, "camera rays, random-init oracle weights (S=64)")
tensors = {k: v.detach() for k, v in m.nerf.nerf_coarse.param_dict().items()}
ps = ops.make_mlp_params(tensors, 8, 256, (4,))
tcw = ops.TcWeights().get(ps, tensors, False)
run(64, None, "camera rays, nn.Linear-init weights (S=64)    ")
ro = (torch.randn(B, 3, generator=g) * 0.5).to(DEV)
rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
run(64, None, "random rays, nn.Linear-init weights (S=64)    ")

# context experiments: fresh output buffer per call; other kernels interleaved (as in the demo's chunk loop)
smp = ops.make_sampling(1.0, 8.0, 64, 10)
n = B * 64


def timed_calls(label, fresh, interleave):
    outs, evs = [], []
    for i in range(iters + 2):
        out = torch.empty(n, 4, device=DEV) if (fresh or i == 0) else outs[-1]
        outs.append(out)
        if interleave:
            noise = torch.randn(B, 256, device=DEV)
            noise2 = noise * 2 + 1
        tin = ops.make_tc_input_rays(ro, rd, None, smp, None, n, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mlp_tc_fwd(ps, tcw, tin, out, None)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs[2:]]
    print(f"{label}: per-call ms min {min(ms):.2f} median {sorted(ms)[len(ms)//2]:.2f} max {max(ms):.2f}")


timed_calls("reused output, back to back        ", False, False)
timed_calls("fresh output per call              ", True, False)
timed_calls("fresh output + interleaved kernels ", True, True)

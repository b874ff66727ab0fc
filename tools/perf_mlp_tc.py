"""Micro-benchmark of the tcgen05 MLP kernels alone (device-resident inputs, CUDA events).
usage: python tools/perf_mlp_tc.py [rows] [iters] [--profile]   (--profile: 1 warm-up + 1 measured pass, for ncu)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from mc_nerf_b200._lib import lib
from oracle import mcnerf_oracle as orc

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = int(args[0]) if len(args) > 0 else 4096 * 192
iters = int(args[1]) if len(args) > 1 else 10
profile = "--profile" in sys.argv
DEV = "cuda"
p = orc.init_mlp_params(8, 256, (4,), seed=3)
tensors = {k: p[k].to(DEV).contiguous() for k in ops.param_names(8)}
ps = ops.make_mlp_params(tensors, 8, 256, (4,))
tcw = ops.TcWeights().get(ps, tensors)
S = 192
B = rows // S
g = torch.Generator().manual_seed(0)
ro = (torch.randn(B, 3, generator=g) * 0.5).to(DEV)
rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
smp = ops.make_sampling(1.0, 8.0, S, 10)
out = torch.empty(B * S, 4, device=DEV)
gout = torch.randn(B * S, 4, device=DEV)
tin = ops.make_tc_input_rays(ro, rd, None, smp, None, B * S, None)
grads = {k: torch.zeros_like(v) for k, v in tensors.items()}
gs = ops.fill_mlp_struct(ops.MlpGrads(), grads, 8)
stash = ops.tc_stash(ps, B * S, DEV)
ws = ops.tc_bwd_workspace(ps, B * S, DEV)
g_o, g_d = torch.zeros(B, 3, device=DEV), torch.zeros(B, 3, device=DEV)
M = B * S
flop_f = 2 * 629248 * M


def timed(fn, n):
    for _ in range(1 if profile else 3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = 1 if profile else iters
if not profile or "--infer" in sys.argv:
    ms = timed(lambda: ops.mlp_tc_fwd(ps, tcw, tin, out, None), n)
    print(f"fwd inference   : rows={M} {ms:.3f} ms  {flop_f/ms/1e9:.1f} TFLOP/s")
ms = timed(lambda: ops.mlp_tc_fwd(ps, tcw, tin, out, stash), n)
print(f"fwd train(stash): rows={M} {ms:.3f} ms  {flop_f/ms/1e9:.1f} TFLOP/s   stash {stash.numel()/1e9:.2f} GB")
L = lib()
L.profile_begin()
ms = timed(lambda: ops.mlp_tc_bwd(ps, tcw, tin, out, gout, stash, ws, gs, g_rays_o=g_o, g_rays_d=g_d), n)
L.profile_end()
print(f"bwd (dgrad chain + wgrad + bias): {ms:.3f} ms  {2*flop_f/ms/1e9:.1f} TFLOP/s (2x forward flops)")

"""Fused backward (chain + weight gradients in one launch, mlp_tc_bwd_fused.cu) vs the two back-to-back kernels:
gradient agreement and device time over the split of the SMs.
usage: python tools/perf_bwd_fused.py [rows] [iters] [--sweep]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from oracle import mcnerf_oracle as orc

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = int(args[0]) if len(args) > 0 else 262144
iters = int(args[1]) if len(args) > 1 else 10
DEV = "cuda"
p = orc.init_mlp_params(8, 256, (4,), seed=3)
tensors = {k: p[k].to(DEV).contiguous() for k in ops.param_names(8)}
ps = ops.make_mlp_params(tensors, 8, 256, (4,))
tcw = ops.TcWeights().get(ps, tensors)
S = 64
B = rows // S
g = torch.Generator().manual_seed(0)
ro = (torch.randn(B, 3, generator=g) * 0.5).to(DEV)
rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
smp = ops.make_sampling(1.0, 8.0, S, 10)
out = torch.empty(B * S, 4, device=DEV)
gout = torch.randn(B * S, 4, device=DEV)
tin = ops.make_tc_input_rays(ro, rd, None, smp, None, B * S, None)
stash = ops.tc_stash(ps, B * S, DEV)
ws = ops.tc_bwd_workspace(ps, B * S, DEV)
ops.mlp_tc_fwd(ps, tcw, tin, out, stash)
M = B * S
flop = 4 * 629248 * M


def run(env, n=iters, check=None):
    for k in ("MCNERF_BWD_FUSED", "MCNERF_FUSED_CHAIN_PAIRS", "MCNERF_FUSED_DISCARD", "MCNERF_FUSED_BPC", "MCNERF_FUSED_PF", "MCNERF_FUSED_MODE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    grads = {k: torch.zeros_like(v) for k, v in tensors.items()}
    gs = ops.fill_mlp_struct(ops.MlpGrads(), grads, 8)
    g_o, g_d = torch.zeros(B, 3, device=DEV), torch.zeros(B, 3, device=DEV)
    ops.mlp_tc_bwd(ps, tcw, tin, out, gout, stash, ws, gs, g_rays_o=g_o, g_rays_d=g_d)
    torch.cuda.synchronize()
    res = {k: v.clone() for k, v in grads.items()}
    res["g_o"], res["g_d"] = g_o.clone(), g_d.clone()
    for _ in range(2):
        ops.mlp_tc_bwd(ps, tcw, tin, out, gout, stash, ws, gs, g_rays_o=g_o, g_rays_d=g_d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.mlp_tc_bwd(ps, tcw, tin, out, gout, stash, ws, gs, g_rays_o=g_o, g_rays_d=g_d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    worst = None
    if check is not None:
        worst = max(float((res[k] - check[k]).norm() / check[k].norm().clamp_min(1e-20)) for k in check)
    print(f"{env}: {ms:.3f} ms  {flop/ms/1e9:.0f} TFLOP/s" + (f"  max rel diff vs two-kernel path {worst:.2e}" if worst is not None else ""),
          flush=True)
    return res


ref = run({"MCNERF_BWD_FUSED": "0"})
run({"MCNERF_BWD_FUSED": "1"}, check=ref)
run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_DISCARD": "1"}, check=ref)
if "--sweep" in sys.argv:
    for pc in (36, 38, 40, 42, 44, 46, 48):
        run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_CHAIN_PAIRS": str(pc)}, check=ref)
        run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_CHAIN_PAIRS": str(pc), "MCNERF_FUSED_DISCARD": "1"}, check=ref)
if "--modes" in sys.argv:
    for pc in (74, 48):
        run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_CHAIN_PAIRS": str(pc), "MCNERF_FUSED_MODE": "1", "MCNERF_FUSED_PF": "0"})
    for bpc in (48, 1000):
        run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_MODE": "2", "MCNERF_FUSED_PF": "0", "MCNERF_FUSED_BPC": str(bpc)})
        run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_MODE": "2", "MCNERF_FUSED_PF": "0", "MCNERF_FUSED_BPC": str(bpc), "MCNERF_FUSED_DISCARD": "1"})
    os.environ["MCNERF_BWD_FUSED"] = "0"
    from mc_nerf_b200._lib import lib
    L = lib(); L.profile_begin()
    grads = {k: torch.zeros_like(v) for k, v in tensors.items()}
    gs = ops.fill_mlp_struct(ops.MlpGrads(), grads, 8)
    g_o, g_d = torch.zeros(B, 3, device=DEV), torch.zeros(B, 3, device=DEV)
    for _ in range(5):
        ops.mlp_tc_bwd(ps, tcw, tin, out, gout, stash, ws, gs, g_rays_o=g_o, g_rays_d=g_d)
    print({k: round(v / 5, 4) for k, v in L.profile_end().items()})
if "--pf" in sys.argv:
    for pc in (44, 48, 52):
        for pf in (0, 6, 12, 24):
            run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_CHAIN_PAIRS": str(pc), "MCNERF_FUSED_PF": str(pf)}, check=ref)
if "--stats" in sys.argv:
    os.environ["MCNERF_FUSED_STATS"] = "1"
    run({"MCNERF_BWD_FUSED": "1", "MCNERF_FUSED_CHAIN_PAIRS": "42"}, n=1)

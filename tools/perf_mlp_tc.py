"""Micro-benchmark of the tcgen05 MLP kernels alone (device-resident inputs, CUDA events).
usage: python tools/perf_mlp_tc.py [rows] [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from oracle import mcnerf_oracle as orc

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
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
tin = ops.make_tc_input_rays(ro, rd, None, smp, None, B * S, None)
for mode in ("inference", "train(stash)"):
    stash = ops.tc_stash(ps, B * S, DEV) if mode != "inference" else None
    for _ in range(3):
        ops.mlp_tc_fwd(ps, tcw, tin, out, stash)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.mlp_tc_fwd(ps, tcw, tin, out, stash)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2 * 629248 * B * S / (ms / 1e3) / 1e12
    print(f"fwd {mode}: rows={B*S} {ms:.3f} ms  {tf:.1f} TFLOP/s (algorithmic)  {B*S/ms*1e3/1e6:.1f} Msamples/s")

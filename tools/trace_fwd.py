"""MCNERF_TC_DEBUG=7 python tools/trace_fwd.py : clock64 trace of the forward kernel's barrier hand-offs (CTA 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from oracle import mcnerf_oracle as orc
DEV = "cuda"
p = orc.init_mlp_params(8, 256, (4,), seed=3)
tensors = {k: p[k].to(DEV).contiguous() for k in ops.param_names(8)}
ps = ops.make_mlp_params(tensors, 8, 256, (4,))
tcw = ops.TcWeights().get(ps, tensors)
B, S = 4096, 192
ro = torch.randn(B, 3, device=DEV) * 0.5
rd = torch.nn.functional.normalize(torch.randn(B, 3, device=DEV), dim=-1)
smp = ops.make_sampling(1.0, 8.0, S, 10)
out = torch.empty(B * S, 4, device=DEV)
tin = ops.make_tc_input_rays(ro, rd, None, smp, None, B * S, None)
train = "--train" in sys.argv
stash = ops.tc_stash(ps, B * S, DEV) if train else None
os.environ.pop("MCNERF_TC_DEBUG", None)
for _ in range(2):
    ops.mlp_tc_fwd(ps, tcw, tin, out, stash)
torch.cuda.synchronize()
os.environ["MCNERF_TC_DEBUG"] = "7"
ops.mlp_tc_fwd(ps, tcw, tin, out, stash)
torch.cuda.synchronize()

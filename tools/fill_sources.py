"""Which Python lines cause the small fill / elementwise kernels of one train step (torch.profiler with stacks)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
sp, model, loss_fn, opt, batch = bench.build_workload("cuda:0", 0, "bf16")
dev_batch = tuple(t.to("cuda:0") for t in batch)
def step():
    opt.zero_grad()
    loss_dict, _, _, _ = model(dev_batch, 25, bench.STAGE, bench.RATIO)
    loss = loss_fn(loss_dict, bench.STAGE)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
names = ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::index", "aten::mul", "aten::add", "aten::copy_",
         "aten::cat", "aten::sub", "aten::div", "aten::sum", "aten::clone", "aten::masked_fill_", "aten::arange")
cnt = collections.Counter()
for ev in prof.events():
    if ev.name in names and ev.device_time_total > 0:
        frames = [f for f in (ev.stack or []) if "mc_nerf_b200" in f or "bench.py" in f or "autograd" in f]
        where = frames[0].strip() if frames else "(no python frame: autograd engine)"
        cnt[(ev.name, str(ev.input_shapes)[:60], where[:110])] += 1
for (n, shp, where), c in sorted(cnt.items(), key=lambda kv: kv[0][2]):
    print(f"{c:2d}x {n:18s} {shp:60s} {where}")

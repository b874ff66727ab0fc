"""cProfile of the host side of one train step (which Python/torch calls cost the CPU time)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
pr = cProfile.Profile()
pr.enable()
for _ in range(20): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])

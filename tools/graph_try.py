import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mc_nerf_b200.graph import GraphedTrainStep
sp, model, loss_fn, opt, batch = bench.build_workload("cuda:0", 0, "bf16")
host_batch = tuple(t.pin_memory() for t in batch)
step = GraphedTrainStep(model, loss_fn)
for i in range(5):
    loss = step(host_batch, 25, bench.STAGE, bench.RATIO); opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(50):
    loss = step(host_batch, 25, bench.STAGE, bench.RATIO); opt.step()
th = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"graphed: {e0.elapsed_time(e1)/50:.3f} ms/step, host issue {th/50*1e3:.3f} ms/step, loss {loss.item():.5f}")
# compare the loss trajectory with eager from the same start

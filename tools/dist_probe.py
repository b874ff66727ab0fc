"""Minimal NCCL probe (run under torchrun): init, all-reduce, barrier; prints progress with timestamps."""
import os, sys, time
t0 = time.time()
def log(*a):
    print(f"[rank {os.environ.get('RANK')}] +{time.time()-t0:.1f}s", *a, flush=True)
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
log("cuda ok", torch.cuda.get_device_name(local))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
log("init ok")
x = torch.ones(1 << 20, device=f"cuda:{local}") * (local + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
log("allreduce ok", float(x[0]))
dist.barrier()
torch.cuda.synchronize()
log("barrier ok")
dist.destroy_process_group()
log("done")

"""Train-step time at the shapes of the reference's SHIPPED config (config/config.yaml: batch 7000, 128 coarse samples,
scale 5 -> 640 fine samples capped at 128 per ray, coarse net 4x128 skip [2], fine net 8x256 skip [4]) through the
drop-in API, eager launches or (--graph) CUDA-graph replay.   python tools/perf_default_config.py [steps] [--graph]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import synthetic as syn
from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
from mc_nerf_b200._lib import lib

graph = "--graph" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if args else 10
dev = "cuda:0"
sp = syn.make_sys_param(n_cam=110, img_h=800, img_w=800, batch=7000, samples=128, scale=5, device=dev, with_images=False,
                        coarse=(4, 128, (2,)), fine=(8, 256, (4,)), pixel_sampler="device")
torch.manual_seed(42)
m = MC_Model(sp).to(dev)
with torch.no_grad():
    for k, v in syn.init_camera_weights(sp).items():
        getattr(m, k).copy_(v)
loss_fn = MC_NeRF_Loss(sp)
opt = RAdam(list(m.parameters()), lr=5e-4, weight_decay=4e-4)
batch = tuple(t.to(dev) for t in syn.make_train_batch(sp, img_id=3))


def eager_step():
    opt.zero_grad()
    loss = loss_fn(m(batch, 25, "GLOBAL_OPTIM_EPOCH", 0.5)[0], "GLOBAL_OPTIM_EPOCH")
    loss.backward()
    opt.step()
    return loss


if graph:
    from mc_nerf_b200.graph import GraphedTrainStep
    gstep = GraphedTrainStep(m, loss_fn)

    def step():
        loss = gstep(batch, 25, "GLOBAL_OPTIM_EPOCH", 0.5)
        opt.step()
        return loss
else:
    step = eager_step


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    loss = step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / steps * 1e3
L = lib()
L.profile_begin()
eager_step()
prof = L.profile_end()
top = ", ".join(f"{k.replace('mcnerf_', '')} {v:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:6])
print(f"default config.yaml shapes (pad={os.environ.get('MCNERF_TC_PAD', '1')}, {'graph replay' if graph else 'eager'}): {ms:.2f} ms/step = {7000 / ms * 1e3 / 1e6:.3f} Mrays/s, "
      f"loss {loss.item():.4f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB; kernel ms: {top}")

"""BASELINE configs[4] "Room-style rig stress": 65536 rays per optimiser step, 64 coarse + 256 fine samples per ray (scale 4),
through the drop-in API.  One optimiser step = `micro` accumulation micro-steps of 65536/micro rays (the activation and
gradient stashes of all 65536 x 320 samples, ~230 GB, do not fit 180 GB at once; the reference would need the same
accumulation).  With scale 4 the reference's train-only cap (128 fine samples per ray on average; model/mc_nerf.py:630-632) is active,
so the fine network sees <= 128 x rays rows; the cap is drawn on the device (render.select_and_cap)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import synthetic as syn, render
from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam

dev = "cuda:0"
total = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
micro = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
scale = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rays = total // micro
sp = syn.make_sys_param(n_cam=110, img_h=800, img_w=800, batch=rays, samples=64, scale=scale, device=dev, with_images=False)
torch.manual_seed(1)
model = MC_Model(sp).to(dev)
with torch.no_grad():
    for k, v in syn.init_camera_weights(sp).items():
        getattr(model, k).copy_(v)
loss_fn = MC_NeRF_Loss(sp)
opt = RAdam(list(model.parameters()), lr=5e-4, eps=1e-8, weight_decay=4e-4)
batches = [tuple(t.to(dev) for t in syn.make_train_batch(sp, img_id=(3 + 7 * i) % 110, seed=11 + i)) for i in range(micro)]


def step(count=False):
    opt.zero_grad()
    evals = 0
    for b in batches:
        loss_dict, _, _, _ = model(b, 25, "GLOBAL_OPTIM_EPOCH", 0.8)
        (loss_fn(loss_dict, "GLOBAL_OPTIM_EPOCH") / micro).backward()
        if count:      # (one host sync per micro-step: only in the untimed warm-up)
            nd = render.LAST.get("n_rows_dev")
            evals += rays * 64 + (int(nd.item()) if nd is not None else int(render.LAST["n_rows"]))
    opt.step()
    return evals


for _ in range(2):
    evals = step(count=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flops = 6.0 * 629248 * evals
print(f"stress: {total} rays/step as {micro} x {rays}, 64+{64*scale} samples: {ms:.1f} ms/step = {total/ms*1e3/1e6:.3f} Mrays/s; "
      f"{evals/1e6:.2f} M MLP evaluations/step -> {flops/ms/1e9:.0f} TFLOP/s algorithmic over the whole step "
      f"(peak mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB)")

"""Two real GPUs: the ray-sharded train step (BASELINE configs[2]) with the gradient exchange of parallel.GradSync -
both transports, the hand-written NVLink all-reduce (p2p) and NCCL - against the SAME batch on one GPU.
Needs two visible GPUs (`gpurun --gpus 2`); with one it is skipped for that stated reason (the single-GPU suite covers
the all-reduce kernel with two virtual ranks, tests/test_kernels_gpu.py, and bench.py --gpus N reports ranks_identical)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MCNERF_ROOT"])
from mc_nerf_b200 import parallel, synthetic as syn
from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
transport, out = sys.argv[1], sys.argv[2]
dev = f"cuda:{local}"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(dev))


def main():
    B = 256
    kw = dict(n_cam=6, img_h=32, img_w=32, samples=32, scale=2, coarse=(4, 256, (2,)), fine=(4, 256, (2,)), device=dev)


    def build(batch):
        sp = syn.make_sys_param(batch=batch, **kw)
        sp["mlp_precision"] = "bf16"
        sp["pixel_sampler"] = "device"
        torch.manual_seed(42)
        m = MC_Model(sp).to(dev)
        with torch.no_grad():
            for k, v in syn.init_camera_weights(sp).items():
                getattr(m, k).copy_(v)
        return sp, m, MC_NeRF_Loss(sp)


    def step(m, loss_fn, batch, seed):
        torch.manual_seed(seed)                    # pixel choice and noise come from torch's CUDA generator
        loss_dict, _, _, _ = m(batch, 25, "GLOBAL_OPTIM_EPOCH", 0.5)
        loss = loss_fn(loss_dict, "GLOBAL_OPTIM_EPOCH")
        loss.backward()
        return float(loss)


    # reference first (no GradSync anywhere yet): every rank's shard gradients, summed by a plain all-reduce
    sp2, m2, loss_fn2 = build(B // world)
    parallel.broadcast_parameters(m2)
    batch = tuple(t.to(dev) for t in syn.make_train_batch(sp2, img_id=3, seed=11))
    step(m2, loss_fn2, batch, 1000 + rank)
    names = [k for k, _ in m2.named_parameters()]
    tot = {}
    for k, p in m2.named_parameters():
        t = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone()
        dist.all_reduce(t)
        tot[k] = t
    del m2

    # sharded: every rank renders B / world rays of the same camera (its own pixels), gradients summed by GradSync
    sp, m, loss_fn = build(B // world)
    parallel.broadcast_parameters(m)
    sync = parallel.GradSync(m, overlap=True, transport=transport).install()
    opt = RAdam(list(m.parameters()), lr=5e-4, eps=1e-8, weight_decay=4e-4)
    opt.grad_scale = 1.0 / world
    losses = []
    for it in range(3):
        opt.zero_grad()
        losses.append(step(m, loss_fn, batch, 1000 + 10 * it + rank))
        sync.finish()
        if it == 0:
            g0 = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        opt.step()
    torch.cuda.synchronize()
    assert parallel.parameters_identical(m), "ranks diverged"
    # every rank holds the SUM of the shard gradients: identical bits everywhere
    flat = torch.cat([g.reshape(-1) for g in g0.values()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "reduced gradients differ between ranks"
    # and the sum equals the plain all-reduce of the same shard gradients (per tensor, relative to the tensor's largest entry;
    # the two sums differ only in fp32 summation order and in the run-to-run order of the ray-gradient atomics)
    err, worst = 0.0, None
    for k in names:
        e = float((tot[k] - g0[k]).abs().max() / tot[k].abs().max().clamp_min(1e-30))
        if e > err:
            err, worst = e, k
    assert err < 1e-4, f"GradSync sum vs plain all-reduce of the same shard gradients: {err} at {worst}"
    # the same exchange inside a replayed CUDA graph (what bench.py times): fork / all-reduce kernels / join are graph
    # nodes.  A fresh model: autograd's AccumulateGrad nodes of a model that already ran eager backward passes live on the
    # default stream, which a capture on another stream may not touch.
    sync.uninstall()
    sp3, m3, loss_fn3 = build(B // world)
    parallel.broadcast_parameters(m3)
    sync3 = parallel.GradSync(m3, overlap=True, transport=transport).install()
    opt3 = RAdam(list(m3.parameters()), lr=5e-4, eps=1e-8, weight_decay=4e-4)
    opt3.grad_scale = 1.0 / world
    from mc_nerf_b200.graph import GraphedTrainStep
    gstep = GraphedTrainStep(m3, loss_fn3)
    gstep.after_backward.append(sync3.finish)
    for it in range(4):
        torch.manual_seed(2000 + 10 * it + rank)
        loss = gstep(batch, 25, "GLOBAL_OPTIM_EPOCH", 0.5)
        opt3.step()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss)), "graph-replayed loss"
    assert parallel.parameters_identical(m3), "ranks diverged under graph replay"
    gflat = torch.cat([p.grad.reshape(-1) for p in m3.parameters()])
    gref = gflat.clone()
    dist.broadcast(gref, 0)
    assert torch.equal(gflat, gref), "graph-replayed reduced gradients differ between ranks"
    sync3.uninstall()
    if rank == 0:
        torch.save(dict(losses=losses, err=err, n_collectives=sync.n_collectives, transport=sync.transport), out)


main()                       # every model, gradient buffer, captured graph and the GradSync pool die with main()'s frame,
import gc                    # i.e. before the process group goes away and before the interpreter clears modules in
gc.collect()                 # arbitrary order (a live torch.cuda.MemPool at that point aborts)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_gpu_sharded_step_gradsync(transport, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs (run under `gpurun --gpus 2`); one-GPU coverage: two-virtual-rank kernel test")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.pt"
    env = dict(os.environ, MCNERF_ROOT=ROOT)
    port = 29600 + (os.getpid() % 200) + (0 if transport == "p2p" else 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script), transport, str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = torch.load(out)
    assert res["transport"] == transport and res["err"] < 1e-4
    assert all(torch.isfinite(torch.tensor(res["losses"])))

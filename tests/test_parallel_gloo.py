"""N > 1 path on CPU: two gloo ranks each take half of a ray batch, run the oracle's forward/backward on their
shard, and the flat-buffer all-reduce (mc_nerf_b200/parallel.py) must reproduce the full-batch gradients
(mean of equal shard means = global mean: the reference's DDP semantic)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    sys.path.insert(0, ROOT)
    from mc_nerf_b200 import synthetic as syn
    from oracle import mcnerf_oracle as orc
    kw = dict(n_cam=4, img_h=8, img_w=8, batch=16, samples=8, scale=2, coarse=(2, 16, ()), fine=(2, 16, ()))
    sp = syn.make_sys_param(**kw)
    cfg = orc.cfg_from_sys_param(sp)
    pc = orc.init_mlp_params(*cfg["coarse"], seed=1)
    pf = orc.init_mlp_params(*cfg["fine"], seed=2)
    g = torch.Generator().manual_seed(3)
    B = 16
    rays_d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)
    rays_o = torch.randn(B, 3, generator=g) * 0.3
    gt = torch.rand(B, 3, generator=g)
    rng = dict(jitter=torch.rand(B, 1, generator=g) * 0.5, noise_c=torch.randn(B, 8, generator=g),
               noise_sel=torch.randn(B, 8, generator=g), noise_f=torch.randn(B, 16, generator=g))
    return orc, cfg, pc, pf, rays_d, rays_o, gt, rng


def _grads(orc, cfg, pc, pf, rays_d, rays_o, gt, rng, sl):
    pc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    r = {k: v[sl] for k, v in rng.items()}
    rgb_c, rgb_f = orc.render_rays(pc, pf, cfg, rays_d[sl], rays_o[sl], r, train=True)
    loss = torch.nn.functional.mse_loss(rgb_c, gt[sl]) + torch.nn.functional.mse_loss(rgb_f, gt[sl])
    loss.backward()
    return pc, pf


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    from mc_nerf_b200.parallel import FlatGradAllReduce, shard_rays
    orc, cfg, pc, pf, rays_d, rays_o, gt, rng = _problem()
    b, e = shard_rays(rays_d.shape[0])
    pc_r, pf_r = _grads(orc, cfg, pc, pf, rays_d, rays_o, gt, rng, slice(b, e))
    params = list(pc_r.values()) + list(pf_r.values())
    # one parameter is left untouched on rank 1 to exercise the "unused parameter" convention
    if rank == 1:
        params[0].grad = None
    FlatGradAllReduce(params)()
    if rank == 0:
        torch.save([p.grad.clone() for p in params], out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    orc, cfg, pc, pf, rays_d, rays_o, gt, rng = _problem()
    # selection threshold uses the batch-global max weight; with thresh 1e-3 << max both shards use 1e-3
    pc_f, pf_f = _grads(orc, cfg, pc, pf, rays_d, rays_o, gt, rng, slice(0, 16))
    full = [p.grad for p in list(pc_f.values()) + list(pf_f.values())]
    # rank 1 contributed zeros for parameter 0 -> compare that one against half of rank 0's shard gradient
    pc_0, _ = _grads(orc, cfg, pc, pf, rays_d, rays_o, gt, rng, slice(0, 8))
    torch.testing.assert_close(got[0], list(pc_0.values())[0].grad * 0.5, rtol=1e-5, atol=1e-7)
    for a, b in list(zip(got, full))[1:]:
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-7)


def test_shard_rays_partitions_the_batch():
    from mc_nerf_b200.parallel import shard_rays
    for n, w in ((4096, 8), (1000, 3), (7, 2)):
        cuts = [shard_rays(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


def _worker_flat(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mc_nerf_b200.parallel import FlatGradAllReduce
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(4, 3), (4,), (2, 4), (2,)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(5))]
    flat = torch.randn(sum(torch.Size(s).numel() for s in shapes), generator=g)      # one buffer, as render.py produces
    off = 0
    for p_, s_ in zip(params, shapes):
        k = torch.Size(s_).numel()
        p_.grad = flat[off:off + k].view(s_)
        off += k
    params[-1].grad = torch.randn(5, generator=g)
    ar = FlatGradAllReduce(params)
    ar()
    assert ar.n_collectives == 1
    if rank == 0:
        torch.save([p_.grad.clone() for p_ in params] + [flat.clone()], out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_buffer_views_are_reduced_in_place(tmp_path):
    out = str(tmp_path / "f.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_flat, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    shapes = [(4, 3), (4,), (2, 4), (2,)]
    flats, lasts = [], []
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        flats.append(torch.randn(sum(torch.Size(s).numel() for s in shapes), generator=g))
        lasts.append(torch.randn(5, generator=g))
    mean_flat = (flats[0] + flats[1]) / 2
    torch.testing.assert_close(got[-1], mean_flat)                     # the shared buffer itself was reduced
    torch.testing.assert_close(torch.cat([t.reshape(-1) for t in got[:4]]), mean_flat)
    torch.testing.assert_close(got[4], (lasts[0] + lasts[1]) / 2)


def _worker_gradsync(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mc_nerf_b200 import parallel

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(3, 2)

    class Nerf(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.nerf_coarse, self.nerf_fine = Net(), Net()

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weights_pose = torch.nn.Parameter(torch.zeros(4, 6))
            self.weights_fx = torch.nn.Parameter(torch.zeros(4))
            self.nerf = Nerf()

    torch.manual_seed(10 + rank)                   # ranks start different (ref: main.py:273-277 seeds 42 + rank)
    m = Model()
    with torch.no_grad():
        for p in m.parameters():
            p.normal_()
    assert not parallel.parameters_identical(m)
    parallel.broadcast_parameters(m)
    assert parallel.parameters_identical(m)
    sync = parallel.GradSync(m, overlap=False)
    g = torch.Generator().manual_seed(200 + rank)
    flats = {}
    for name, net in (("fine", m.nerf.nerf_fine), ("coarse", m.nerf.nerf_coarse)):     # the renderer's backward order
        flat = torch.randn(sum(p.numel() for p in net.parameters()), generator=g)
        off = 0
        for p in net.parameters():
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        flats[name] = flat
        if name == "fine" or rank >= 0:
            sync._on_ready(name, flat)              # what render.GRAD_HOOK calls inside RenderFn.backward
    m.weights_pose.grad = torch.randn(4, 6, generator=g)
    m.weights_fx.grad = None if rank == 1 else torch.randn(4, generator=g)      # untouched on rank 1 only
    sync.finish()
    assert sync.n_collectives == 3
    if rank == 0:
        torch.save(dict(fine=flats["fine"], coarse=flats["coarse"], pose=m.weights_pose.grad, fx=m.weights_fx.grad,
                        w=m.nerf.nerf_fine.a.weight.grad), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_sums_fixed_buffers_and_keeps_ranks_in_step(tmp_path):
    """GradSync (the overlapped path of bench.py, here on gloo / CPU without streams): per-network flat buffers are
    reduced IN PLACE when the renderer reports them, the camera tensors in finish(); a gradient that is None on one
    rank only contributes zeros and cannot desynchronise the collectives.  The result is the SUM (1/N goes into the
    optimiser's grad_scale)."""
    out = str(tmp_path / "s.pt")
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_gradsync, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    exp = {}
    for rank in range(2):
        g = torch.Generator().manual_seed(200 + rank)
        fine, coarse = torch.randn(8, generator=g), torch.randn(8, generator=g)
        pose = torch.randn(4, 6, generator=g)
        fx = torch.zeros(4) if rank == 1 else torch.randn(4, generator=g)
        for k, v in dict(fine=fine, coarse=coarse, pose=pose, fx=fx).items():
            exp[k] = exp.get(k, 0) + v
    for k in exp:
        torch.testing.assert_close(got[k], exp[k])
    torch.testing.assert_close(got["w"].reshape(-1), exp["fine"][:6])       # .grad views follow their flat buffer

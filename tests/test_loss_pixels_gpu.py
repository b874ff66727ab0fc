"""GPU tests of the two step-level kernels that replace strings of tiny launches: the fused train-stage loss
(ref: model/loss.py:15-58) against the reference's torch expression, and the pixel sampler
(ref: model/mc_nerf.py:327-345, randperm(H*W)[:batch]) bit-exact against its numpy oracle plus the statistical
properties of a random permutation head."""
import numpy as np
import pytest
import torch

from mc_nerf_b200 import synthetic as syn
from oracle import mcnerf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops():
    from mc_nerf_b200 import ops as _ops
    return _ops


def torch_loss(loss_mod, rgb_c, rgb_f, gt, px, px_gt):
    """the reference expression for a rendering stage (model/loss.py:15-31), plain torch ops"""
    total = 0.0
    if px is not None:
        l_intr = loss_mod.get_reproject_loss([px, px_gt])
        total = total + l_intr / (l_intr.detach() + 1e-8)
    return total + loss_mod.get_rgb_loss([rgb_c, rgb_f, gt])


@pytest.mark.parametrize("n_rays,with_fine,with_px", [(4096, True, True), (1000, True, True), (64, False, True),
                                                      (333, True, False), (1, True, True)])
def test_fused_train_loss_matches_torch_expression(n_rays, with_fine, with_px):
    from mc_nerf_b200.model import MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=110, img_h=600, img_w=800, with_images=False)
    loss_mod = MC_NeRF_Loss(sp)
    g = torch.Generator().manual_seed(n_rays)
    mk = lambda *shape, scale=1.0: (torch.rand(*shape, generator=g) * scale).to(DEV)
    rgb_c, rgb_f, gt = mk(n_rays, 3), (mk(n_rays, 3) if with_fine else None), mk(n_rays, 3)
    px, px_gt = (mk(1, 110, 5, 2, scale=800.0), mk(1, 110, 5, 2, scale=800.0)) if with_px else (None, None)
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (rgb_c, rgb_f, px)]
    ref = torch_loss(loss_mod, leaves[0], leaves[1], gt, leaves[2], px_gt)
    ref.backward()
    mine = [t.clone().requires_grad_(True) if t is not None else None for t in (rgb_c, rgb_f, px)]
    loss_dict = {"rgb": [mine[0], mine[1], gt]}
    if with_px:
        loss_dict["intr"] = [mine[2], px_gt]
    out = loss_mod(loss_dict, "GLOBAL_OPTIM_EPOCH")
    assert out.grad_fn is not None and "TrainLossFn" in type(out.grad_fn).__name__      # the fused kernel ran
    (out * 1.0).backward()
    torch.testing.assert_close(out.detach(), ref.detach(), rtol=2e-6, atol=1e-7)
    for a, b in zip(mine, leaves):
        if a is not None:
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-10)


def test_fused_train_loss_scales_with_upstream_gradient():
    from mc_nerf_b200.model import MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=6, img_h=16, img_w=16, with_images=False)
    rgb_c = torch.rand(50, 3, device=DEV, requires_grad=True)
    rgb_f = torch.rand(50, 3, device=DEV, requires_grad=True)
    gt = torch.rand(50, 3, device=DEV)
    out = MC_NeRF_Loss(sp)({"rgb": [rgb_c, rgb_f, gt]}, "FINE_TUNE_EPOCH")
    (out * 0.25).backward()
    torch.testing.assert_close(rgb_c.grad, 0.25 * 2 * (rgb_c.detach() - gt) / 150, rtol=1e-5, atol=1e-9)


def test_camera_stage_loss_keeps_the_torch_path():
    from mc_nerf_b200.model import MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=6, img_h=16, img_w=16, with_images=False)
    px = torch.rand(1, 6, 5, 2, device=DEV, requires_grad=True)
    gt = torch.rand(1, 6, 5, 2, device=DEV)
    out = MC_NeRF_Loss(sp)({"intr": [px, gt], "extr": [px, gt]}, "CAM_PARAM_EPOCH")
    assert "TrainLossFn" not in type(out.grad_fn).__name__


def draw(n, batch, s0, s1):
    seed = torch.tensor([s0, s1], dtype=torch.int64, device=DEV)
    ws = ops().sample_pixels_workspace(n, batch, DEV)
    assert ws is not None
    i64, i32 = ops().sample_pixels(n, batch, seed, ws)
    again64, _ = ops().sample_pixels(n, batch, seed, ws)            # workspace is left ready for the next call
    assert torch.equal(i64, again64) and torch.equal(i64, i32.to(torch.int64))
    assert int(ws[:4].view(torch.int32).item()) == 0
    return i64.cpu().numpy()


@pytest.mark.parametrize("n,batch", [(640000, 4096), (10000, 1024), (256, 64), (100, 256), (1, 1), (640000, 8192),
                                     (5000, 4999), (768, 256), (1 << 20, 17)])
def test_pixel_sampler_is_the_head_of_the_full_sort(n, batch):
    for s0, s1 in [(1, 2), (-(1 << 61) + 12345, (1 << 60) + 99)]:
        got = draw(n, batch, s0, s1)
        want = orc.sample_pixels(n, batch, s0, s1)
        assert got.shape == want.shape == (min(n, batch),)
        assert np.array_equal(got, want)                               # index work: bit-exact
        assert len(np.unique(got)) == len(got) and got.min() >= 0 and got.max() < n


def test_pixel_sampler_statistics_of_a_permutation_head():
    n, batch, reps = 640000, 4096, 24
    counts = np.zeros(64)
    first_half = 0
    for r in range(reps):
        idx = draw(n, batch, 1000 + r, 77)
        counts += np.bincount(idx * 64 // n, minlength=64)
        first_half += np.corrcoef(np.arange(batch), idx)[0, 1]
    expected = reps * batch / 64
    chi2 = ((counts - expected) ** 2 / expected).sum()
    assert chi2 < 130                       # 63 dof: mean 63, P[chi2 > 130] ~ 1e-6
    assert abs(first_half / reps) < 0.02    # position in the batch carries no information about the pixel index
    assert not np.array_equal(draw(n, batch, 1, 2), draw(n, batch, 1, 3))


def test_pixel_sampler_too_large_for_one_block_falls_back_to_randperm():
    assert ops().sample_pixels_workspace(640000, 65536, DEV) is None
    from mc_nerf_b200.model import MC_Model
    sp = syn.make_sys_param(n_cam=4, img_h=300, img_w=300, batch=65536, samples=8, scale=2, device=DEV,
                            with_images=False, pixel_sampler="device")
    m = MC_Model(sp).to(DEV)
    i64, i32 = m._choose_pixels(90000)
    assert i64.shape == (65536,) and len(torch.unique(i64)) == 65536


def test_model_train_step_with_device_sampler_follows_torch_seed():
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=6, img_h=24, img_w=32, batch=256, samples=16, scale=2, device=DEV, with_images=False,
                            pixel_sampler="device")
    m = MC_Model(sp).to(DEV)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(m, k).copy_(v)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=2, seed=3))
    losses = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        loss = MC_NeRF_Loss(sp)(m(batch, 25, "GLOBAL_OPTIM_EPOCH", 0.5)[0], "GLOBAL_OPTIM_EPOCH")
        losses.append(loss.item())
    assert losses[0] == losses[1] and losses[0] != losses[2]

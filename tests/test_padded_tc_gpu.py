"""Networks narrower than 256 (the reference's default coarse net is 4x128, config/config.yaml:76-78) run on the
tcgen05 path through a zero-padded 256-wide shadow (ops.PaddedNet).  Checked against the already-validated native
path: a 256-wide network whose parameters ARE the zero-padded ones must give bit-identical renders and the same
gradients on the valid blocks; and against the fp32 CUDA-core path within the bf16 path's stated tolerance."""
import pytest
import torch

from mc_nerf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, SC, SCALE = 300, 16, 2


def build(depth, width, skips, precision, seed=3):
    from mc_nerf_b200.model import MC_Model
    sp = syn.make_sys_param(n_cam=4, img_h=16, img_w=16, batch=B, samples=SC, scale=SCALE, device=DEV, with_images=False,
                            coarse=(depth, width, skips), fine=(depth, width, skips))
    sp["mlp_precision"] = precision
    torch.manual_seed(seed)
    return sp, MC_Model(sp).to(DEV)


def pad_into(narrow, wide):
    """zero the wide network and copy the narrow parameters into the top-left blocks"""
    with torch.no_grad():
        for (k, pn), (k2, pw) in zip(narrow.named_parameters(), wide.named_parameters()):
            assert k == k2
            pw.zero_()
            if pn.dim() == 2:
                pw[:pn.shape[0], :pn.shape[1]].copy_(pn)
            else:
                pw[:pn.shape[0]].copy_(pn)


def inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
    ro = (torch.randn(B, 3, generator=g) * 0.3).to(DEV)
    rng = dict(jitter=torch.rand(B, 1, generator=g) * (7.0 / SC), noise_c=torch.randn(B, SC, generator=g),
               noise_sel=torch.randn(B, SC, generator=g), noise_f=torch.randn(B, SC * SCALE, generator=g))
    return rd, ro, {k: v.to(DEV) for k, v in rng.items()}, torch.rand(B, 3, generator=g).to(DEV)


def step(nerf, rd, ro, rng, gt):
    for p in nerf.parameters():
        p.grad = None
    rgb_c, rgb_f = nerf.render_rays_train(rd, ro, 25, 0.5, rng=rng)
    (((rgb_c - gt) ** 2).mean() + ((rgb_f - gt) ** 2).mean()).backward()
    return rgb_c.detach(), rgb_f.detach(), {k: p.grad.clone() for k, p in nerf.named_parameters()}


@pytest.mark.parametrize("depth,width,skips", [(4, 128, (2,)), (2, 64, ()), (3, 40, (1,)), (8, 248, (4,))])
def test_narrow_network_equals_its_zero_padded_wide_twin(depth, width, skips):
    from mc_nerf_b200 import render
    _, narrow = build(depth, width, skips, "bf16")
    _, wide = build(depth, 256, skips, "bf16")
    pad_into(narrow.nerf, wide.nerf)
    assert render.use_tc(narrow.nerf.render_cfg, narrow.nerf.render_cfg.coarse)
    rd, ro, rng, gt = inputs()
    for round_ in range(2):
        cn, fn, gn = step(narrow.nerf, rd, ro, rng, gt)
        cw, fw, gw = step(wide.nerf, rd, ro, rng, gt)
        assert torch.equal(cn, cw) and torch.equal(fn, fw)
        for k, g in gn.items():
            blk = gw[k][:g.shape[0], :g.shape[1]] if g.dim() == 2 else gw[k][:g.shape[0]]
            assert g.is_contiguous()
            if k.endswith("weight"):
                assert torch.equal(g, blk), k
            else:          # bias gradients are accumulated with atomics (order varies run to run)
                assert float((g - blk).norm()) <= 1e-5 * float(blk.norm()) + 1e-12, k
            rest = gw[k].clone()
            (rest[:g.shape[0], :g.shape[1]] if g.dim() == 2 else rest[:g.shape[0]]).zero_()
            assert float(rest.abs().max()) == 0.0, k           # padded units carry no gradient
        with torch.no_grad():                                   # second round: the shadow must follow the parameters
            for p in narrow.nerf.parameters():
                p.mul_(1.25)
        pad_into(narrow.nerf, wide.nerf)
        if round_ == 0:
            first = cn
    assert not torch.equal(first, cn)


def test_narrow_network_bf16_path_close_to_fp32_path():
    _, m16 = build(4, 128, (2,), "bf16")
    _, m32 = build(4, 128, (2,), "fp32")
    rd, ro, rng, gt = inputs(1)
    c16, f16, g16 = step(m16.nerf, rd, ro, rng, gt)
    c32, f32, g32 = step(m32.nerf, rd, ro, rng, gt)
    assert float((c16 - c32).abs().max()) <= 1e-3 and float((f16 - f32).abs().quantile(0.99)) <= 1e-3
    for k in g32:
        if k.endswith("weight"):
            assert float((g16[k] - g32[k]).norm()) <= 0.1 * float(g32[k].norm()) + 1e-9, k


def test_optimiser_step_reaches_the_shadow():
    from mc_nerf_b200.model import RAdam
    _, m = build(4, 128, (2,), "bf16")
    rd, ro, rng, gt = inputs(2)
    opt = RAdam(list(m.nerf.parameters()), lr=1e-2)
    c0, _, _ = step(m.nerf, rd, ro, rng, gt)
    for _ in range(8):
        opt.step()
        c1, _, _ = step(m.nerf, rd, ro, rng, gt)
    assert float((c1 - c0).abs().max()) > 1e-4

"""Fused tails (mc_nerf_b200/csrc/tails.cu) and the device RNG (philox.cuh):
 * the coarse tail = composite_fwd + sigma2weights of the unfused kernels, bit for bit;
 * the fine tail on the COMPACTED rows = scatter into defaults + dense compositing (and its backward = dense backward
   + gather), bit for bit, over ragged selections incl. empty and full rays;
 * Philox streams bit-exact against the numpy oracle, N(0,1) / U(a,b) moments;
 * a render in device-RNG mode = the same render fed the explicit tensors of those streams (so the backward pass
   regenerates exactly the forward's noise)."""
import ctypes

import numpy as np
import pytest
import torch

from mc_nerf_b200 import synthetic as syn
from oracle import mcnerf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
    from mc_nerf_b200 import ops
    from mc_nerf_b200._lib import lib
    return ops, lib(), ops._p, ops._stream


@pytest.mark.parametrize("B,S", [(37, 64), (5, 8), (130, 128), (9, 200)])
def test_coarse_tail_equals_unfused_kernels(B, S):
    ops, L, _p, _stream = _lib()
    g = torch.Generator().manual_seed(B * S)
    out4 = torch.cat([torch.randn(B, S, 1, generator=g) * 3, torch.rand(B, S, 3, generator=g)], -1).to(DEV).contiguous()
    n1, n2 = torch.randn(B, S, generator=g).to(DEV), torch.randn(B, S, generator=g).to(DEV)
    jit = (torch.rand(B, generator=g) * 0.1).to(DEV)
    rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
    rgb, w_sel, w_max = ops.coarse_tail_fwd(out4, n1, n2, None, jit, B, 1.0, 8.0, S, True)
    cc = ops.make_composite_cfg(1.0, 8.0, S, True)
    rgb0 = torch.empty(B, 3, device=DEV)
    L.call("mcnerf_composite_fwd", _p(out4), _p(n1), _p(rd), _p(jit), None, B, ctypes.byref(cc), _p(rgb0), None, None,
           None, _stream())
    wm0 = torch.zeros(1, device=DEV)
    w0 = ops.sigma2weights(out4, n2, jitter=jit, near=1.0, far=8.0, sigma_stride=4, n_rays=B, S=S, w_max=wm0)
    assert torch.equal(rgb, rgb0) and torch.equal(w_sel, w0) and torch.equal(w_max, wm0)
    # backward of the colour path
    g_rgb = torch.randn(B, 3, generator=g).to(DEV)
    ga, gb = torch.empty_like(out4), torch.empty_like(out4)
    L.call("mcnerf_coarse_tail_bwd", _p(out4), _p(n1), None, _p(jit), B, ctypes.byref(cc), _p(g_rgb), _p(ga), _stream())
    L.call("mcnerf_composite_bwd", _p(out4), _p(n1), _p(jit), None, B, ctypes.byref(cc), _p(g_rgb), _p(gb), _stream())
    assert torch.equal(ga, gb)


@pytest.mark.parametrize("B,Sc,scale,density", [(41, 64, 2, 0.5), (7, 8, 2, 0.3), (33, 32, 4, 0.8), (16, 64, 2, 0.0),
                                                (16, 64, 2, 1.0), (19, 128, 2, 0.6)])
def test_fine_tail_on_compacted_rows_equals_dense_path(B, Sc, scale, density):
    ops, L, _p, _stream = _lib()
    Sf = Sc * scale
    g = torch.Generator().manual_seed(B + Sc)
    # selection weights: `density` of the coarse samples above the threshold; some rays empty, some full
    w_sel = torch.rand(B, Sc, generator=g)
    w_sel = torch.where(torch.rand(B, Sc, generator=g) < density, w_sel + 0.5, w_sel * 1e-4)
    if density not in (0.0, 1.0):
        w_sel[0] = 1e-5
        w_sel[-1] = 0.9
    w_sel = w_sel.to(DEV).contiguous()
    w_max = w_sel.max().reshape(1).contiguous()
    thresh = 1e-3
    sel_idx, offs, n_sel = ops.select_fine(w_sel, w_max, scale, thresh)
    n = int(n_sel.item())
    cap = B * Sf
    out_sel = torch.full((cap, 4), float("nan"), device=DEV)
    out_sel[:n] = torch.cat([torch.randn(n, 1, generator=g) * 3, torch.rand(n, 3, generator=g)], -1).to(DEV)
    noise = torch.randn(B, Sf, generator=g).to(DEV)
    jit = (torch.rand(B, generator=g) * 0.05).to(DEV)
    rd = (torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1) * 1.1).to(DEV).contiguous()
    cf = ops.make_composite_cfg(1.0, 8.0, Sf, True)
    rgb, dep, opa = (torch.empty(B, 3, device=DEV), torch.empty(B, 1, device=DEV), torch.empty(B, 1, device=DEV))
    L.call("mcnerf_fine_tail_fwd", _p(out_sel), _p(w_sel), _p(w_max), thresh, scale, _p(offs, torch.int32), _p(rd), _p(jit),
           _p(noise), None, B, ctypes.byref(cf), -20.0, _p(rgb), _p(dep), _p(opa), _stream())
    dense = torch.empty(cap, 4, device=DEV)
    L.call("mcnerf_scatter_fine", _p(out_sel), _p(sel_idx, torch.int32), cap, _p(n_sel, torch.int32), cap, -20.0, _p(dense),
           _stream())
    rgb0, dep0, opa0 = torch.empty_like(rgb), torch.empty_like(dep), torch.empty_like(opa)
    L.call("mcnerf_composite_fwd", _p(dense), _p(noise), _p(rd), _p(jit), None, B, ctypes.byref(cf), _p(rgb0), _p(dep0),
           _p(opa0), None, _stream())
    assert torch.equal(rgb, rgb0) and torch.equal(dep, dep0) and torch.equal(opa, opa0)
    g_rgb = torch.randn(B, 3, generator=g).to(DEV)
    g_sel = torch.zeros(cap, 4, device=DEV)
    L.call("mcnerf_fine_tail_bwd", _p(out_sel), _p(w_sel), _p(w_max), thresh, scale, _p(offs, torch.int32), _p(jit), _p(noise),
           None, B, ctypes.byref(cf), -20.0, _p(g_rgb), _p(g_sel), _stream())
    g_dense = torch.empty_like(dense)
    L.call("mcnerf_composite_bwd", _p(dense), _p(noise), _p(jit), None, B, ctypes.byref(cf), _p(g_rgb), _p(g_dense), _stream())
    g_sel0 = torch.zeros(cap, 4, device=DEV)
    L.call("mcnerf_gather_fine", _p(g_dense), _p(sel_idx, torch.int32), cap, _p(n_sel, torch.int32), _p(g_sel0), _stream())
    assert torch.equal(g_sel[:n], g_sel0[:n])
    assert float(g_sel[n:].abs().max()) == 0.0 if n < cap else True          # rows beyond the count are never written


def test_philox_streams_match_the_oracle_and_have_the_right_moments():
    ops, L, _p, _stream = _lib()
    seed = torch.tensor([0x1234567890ABCDEF - (1 << 63), 987654321], dtype=torch.int64, device=DEV)
    s0, s1 = int(seed[0]), int(seed[1])
    n = 100000
    for stream in (1, 2, 3):
        x = ops.philox_fill(seed, stream, n).cpu().numpy()
        ref = orc.philox_normal(n, s0, s1, stream)
        assert np.max(np.abs(x - ref)) < 2e-5
    u = ops.philox_fill(seed, 4, n, normal=False, lo=0.25, hi=0.75).cpu().numpy()
    assert np.max(np.abs(u - orc.philox_uniform(n, s0, s1, 4, 0.25, 0.75))) < 1e-7
    big = ops.philox_fill(seed, 1, 4_000_000)
    assert abs(float(big.mean())) < 3e-3 and abs(float(big.var()) - 1.0) < 5e-3
    assert abs(float((big ** 3).mean())) < 1e-2 and abs(float((big ** 4).mean()) - 3.0) < 3e-2     # skewness, kurtosis
    q = torch.tensor([0.001, 0.1587, 0.5, 0.8413, 0.999], device=DEV)
    emp = torch.quantile(big[:1_000_000], q).cpu()
    torch.testing.assert_close(emp, torch.tensor([-3.0902, -1.0, 0.0, 1.0, 3.0902]), rtol=0, atol=2e-2)
    assert 0.25 < u.min() and u.max() < 0.75 and abs(u.mean() - 0.5) < 2e-3
    other = ops.philox_fill(seed, 2, 1000)
    assert float((other - big[:1000]).abs().max()) > 0.1                      # streams are independent


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_device_rng_render_equals_render_with_explicit_streams(precision):
    """Device-RNG mode (noise generated in the kernels, regenerated in backward) against the SAME render fed explicit
    tensors of the same Philox streams: renders and every gradient bit-identical."""
    from mc_nerf_b200 import ops, render
    from mc_nerf_b200.model import MC_Model
    kw = dict(n_cam=4, img_h=16, img_w=16, batch=96, samples=32, scale=2, coarse=(4, 256, (2,)), fine=(4, 256, (2,)))
    sp = syn.make_sys_param(device=DEV, **kw)
    sp["mlp_precision"] = precision
    sp["noise_sampler"] = "device"
    torch.manual_seed(3)
    m = MC_Model(sp).to(DEV)
    cfg = m.nerf.render_cfg
    assert cfg.device_rng
    B, Sc, Sf = 96, 32, 64
    g = torch.Generator().manual_seed(1)
    rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV).requires_grad_(True)
    ro = (torch.randn(B, 3, generator=g) * 0.2).to(DEV).requires_grad_(True)
    gt = torch.rand(B, 3, generator=g).to(DEV)
    seed = torch.tensor([42, -7], dtype=torch.int64, device=DEV)

    def run(rng):
        for p in list(m.nerf.parameters()) + [rd, ro]:
            p.grad = None
        rgb_c, rgb_f = m.nerf.render_rays_train(rd, ro, 25, 1.0, rng=rng)
        (((rgb_c - gt) ** 2).mean() + ((rgb_f - gt) ** 2).mean()).backward()
        return rgb_c.detach(), rgb_f.detach(), [p.grad.clone() for p in list(m.nerf.parameters()) + [rd, ro]]

    jitter = ops.philox_fill(seed, 4, B, normal=False, lo=0.0, hi=7.0 / Sc)
    a = run(dict(seed=seed, jitter=jitter))
    b = run(dict(jitter=jitter, noise_c=ops.philox_fill(seed, 1, B * Sc).view(B, Sc),
                 noise_sel=ops.philox_fill(seed, 2, B * Sc).view(B, Sc), noise_f=ops.philox_fill(seed, 3, B * Sf).view(B, Sf)))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    n_par = len(list(m.nerf.parameters()))
    for i, (x, y) in enumerate(zip(a[2], b[2])):
        if precision == "bf16" and i < n_par:
            assert torch.equal(x, y)        # weight / bias gradients of the tensor-core path: fixed reduction order
        else:                               # ray gradients (and the fp32 path's split-K sums) use fp32 atomics: the
            torch.testing.assert_close(x, y, rtol=2e-5, atol=1e-9)      # run-to-run summation order is free
    # and the default path draws its own seed from torch's generator: reproducible under torch.manual_seed
    torch.manual_seed(11)
    c = m.nerf.render_rays_train(rd, ro, 25, 1.0)
    torch.manual_seed(11)
    d = m.nerf.render_rays_train(rd, ro, 25, 1.0)
    torch.manual_seed(12)
    e = m.nerf.render_rays_train(rd, ro, 25, 1.0)
    assert torch.equal(c[1], d[1]) and not torch.equal(c[1], e[1])


def test_fused_tails_switch_gives_identical_renders(monkeypatch):
    """MCNERF_FUSED_TAILS=0 (the unfused kernels: composite + sigma2weights + scatter / gather) = the fused tails."""
    from mc_nerf_b200.model import MC_Model
    kw = dict(n_cam=4, img_h=16, img_w=16, batch=64, samples=16, scale=2, coarse=(4, 64, (2,)), fine=(4, 64, (2,)))
    sp = syn.make_sys_param(device=DEV, **kw)
    sp["mlp_precision"] = "fp32"
    torch.manual_seed(5)
    m = MC_Model(sp).to(DEV)
    g = torch.Generator().manual_seed(2)
    B = 64
    rd = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
    ro = (torch.randn(B, 3, generator=g) * 0.2).to(DEV)
    rng = {k: v.to(DEV) for k, v in syn.draw_step_rng(syn.make_sys_param(**kw), B, seed=9).items()
           if k in ("jitter", "noise_c", "noise_sel", "noise_f")}
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MCNERF_FUSED_TAILS", mode)
        for p in m.nerf.parameters():
            p.grad = None
        rgb_c, rgb_f = m.nerf.render_rays_train(rd, ro, 25, 1.0, rng=rng)
        (rgb_c.sum() + (rgb_f ** 2).sum()).backward()
        res[mode] = (rgb_c.detach(), rgb_f.detach(), [p.grad.clone() for p in m.nerf.parameters()])
    assert torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])
    for x, y in zip(res["0"][2], res["1"][2]):
        torch.testing.assert_close(x, y, rtol=2e-5, atol=1e-8)      # fp32 path: atomics in its split-K sums

"""GPU parity tests: every C-ABI kernel against the CPU oracle (and the reference's golden vectors)
on identical inputs.  Tolerances are fp32 (the fp32 path only differs from the CPU by summation order
and libm ulps); the bf16 tensor-core path has its own file.  Run with `-m gpu` on a B200."""
import math

import pytest
import torch

from conftest import load_golden
from oracle import mcnerf_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda"


def ops():
    from mc_nerf_b200 import ops as _ops
    return _ops


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.cpu(), b.cpu(), rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def mods():
    return load_golden("modules.pt")


def test_library_loads_and_counts_launches():
    from mc_nerf_b200._lib import lib
    n0 = lib().launch_count()
    ops().SE3Fn.apply(torch.ones(3, 6, device=DEV))
    torch.cuda.synchronize()
    assert lib().launch_count() == n0 + 1


def test_camera_model_golden(mods):
    c = mods["cam"]
    w = {k: v.to(DEV).requires_grad_(True) for k, v in c["w"].items()}
    K, Kinv = ops().IntrinsicsFn.apply(w["weights_fx"], w["weights_fy"], w["weights_ux"], w["weights_uy"], c["H"], c["W"])
    close(K, c["K"])
    close(Kinv, c["Kinv"], rtol=1e-5, atol=1e-6)
    close(ops().SE3Fn.apply(w["weights_pose"]), c["pose"])
    close(ops().SE3Fn.apply(w["weights_pose_intr"]), c["calib"])
    b = mods["se3_big"]
    # |w| up to ~9 rad: the truncated alternating series cancels ~1e3x, so fp32 evaluation order shows
    # (terms reach 1e2 with a result of order 1: ~1e-5 relative to the largest term)
    close(ops().SE3Fn.apply(b["wu"].to(DEV)), b["Rt"], rtol=1e-4, atol=2e-4)


def test_camera_model_backward():
    g = torch.Generator().manual_seed(1)
    n = 7
    wu = (torch.randn(n, 6, generator=g) * 1.2)
    fx, fy = torch.rand(n, generator=g) + 0.5, -(torch.rand(n, generator=g) + 0.5)
    ux, uy = torch.rand(n, generator=g) + 0.5, torch.rand(n, generator=g) + 0.5
    gRt, gK, gKi = torch.randn(n, 3, 4, generator=g), torch.randn(n, 3, 3, generator=g), torch.randn(n, 3, 3, generator=g)
    # oracle
    lw = [t.clone().requires_grad_(True) for t in (wu, fx, fy, ux, uy)]
    Rt = orc.se3_to_SE3(lw[0])
    K = orc.intrinsics_from_weights(lw[1], lw[2], lw[3], lw[4], 12, 16)
    Ki = orc.inverse_intrinsics(K)
    ((Rt * gRt).sum() + (K * gK).sum() + (Ki * gKi).sum()).backward()
    # kernels
    dw = [t.to(DEV).requires_grad_(True) for t in (wu, fx, fy, ux, uy)]
    Rt2 = ops().SE3Fn.apply(dw[0])
    K2, Ki2 = ops().IntrinsicsFn.apply(dw[1], dw[2], dw[3], dw[4], 12, 16)
    ((Rt2 * gRt.to(DEV)).sum() + (K2 * gK.to(DEV)).sum() + (Ki2 * gKi.to(DEV)).sum()).backward()
    for a, b in zip(dw, lw):
        close(a.grad, b.grad, rtol=1e-4, atol=1e-5)


def test_raygen_golden_and_backward(mods):
    c, r = mods["cam"], mods["rays"]
    H, W = c["H"], c["W"]
    Kinv = c["Kinv"].to(DEV).requires_grad_(True)
    pose = c["pose"].to(DEV).requires_grad_(True)
    ro, rd = ops().RaygenFn.apply(Kinv, pose, r["img_id"], None, H * W, W)
    close(rd, r["rays_d"], rtol=1e-5, atol=1e-6)
    close(ro, r["rays_o"], rtol=1e-5, atol=1e-6)
    # backward vs oracle autograd, random pixel subset and mixed cameras
    g = torch.Generator().manual_seed(2)
    B = 200
    cam = torch.randint(0, 6, (B,), generator=g, dtype=torch.int32)
    pix = torch.randint(0, H * W, (B,), generator=g, dtype=torch.int32)
    go, gd = torch.randn(B, 3, generator=g), torch.randn(B, 3, generator=g)
    Ki_c, P_c = c["Kinv"].clone().requires_grad_(True), c["pose"].clone().requires_grad_(True)
    ros, rds = [], []
    for cam_i in range(6):
        d_all, o_all = orc.get_rays(P_c[cam_i], Ki_c[cam_i], H, W)
        ros.append(o_all)
        rds.append(d_all)
    ro_ref = torch.stack(ros)[cam.long(), pix.long()]
    rd_ref = torch.stack(rds)[cam.long(), pix.long()]
    ((ro_ref * go).sum() + (rd_ref * gd).sum()).backward()
    ro2, rd2 = ops().RaygenFn.apply(Kinv, pose, cam.to(DEV), pix.to(DEV), B, W)
    close(ro2, ro_ref.detach())
    close(rd2, rd_ref.detach())
    ((ro2 * go.to(DEV)).sum() + (rd2 * gd.to(DEV)).sum()).backward()
    close(Kinv.grad, Ki_c.grad, rtol=1e-4, atol=1e-4)
    close(pose.grad, P_c.grad, rtol=1e-4, atol=1e-4)
    # uniform-camera warp-reduced path
    Kinv.grad = None
    pose.grad = None
    ro3, rd3 = ops().RaygenFn.apply(Kinv, pose, 4, pix.to(DEV), B, W)
    ((ro3 * go.to(DEV)).sum() + (rd3 * gd.to(DEV)).sum()).backward()
    Ki_c.grad = None
    P_c.grad = None
    d_all, o_all = orc.get_rays(P_c[4], Ki_c[4], H, W)
    ((o_all[pix.long()] * go).sum() + (d_all[pix.long()] * gd).sum()).backward()
    close(Kinv.grad, Ki_c.grad, rtol=1e-4, atol=1e-4)
    close(pose.grad, P_c.grad, rtol=1e-4, atol=1e-4)


def test_encoding_golden_and_backward(mods):
    e = mods["enc"]
    x = e["x"].to(DEV).requires_grad_(True)
    enc = ops().EncodePointsFn.apply(x, e["L"], None)
    # arguments reach ~7*512 rad: sin/cos of a float32 argument, 2 ulp libm differences -> 1e-6 absolute
    close(enc, e["plain"], rtol=1e-5, atol=2e-6)
    for r, ref in e["barf"].items():
        bw = ops().barf_band_weights(r, e["barf_start"], e["barf_end"], e["L"])
        close(ops().EncodePointsFn.apply(x, e["L"], bw), ref, rtol=1e-5, atol=2e-6)
    g = torch.randn(enc.shape, generator=torch.Generator().manual_seed(3))
    bw = ops().barf_band_weights(0.5, e["barf_start"], e["barf_end"], e["L"])
    ops().EncodePointsFn.apply(x, e["L"], bw).backward(g.to(DEV))
    xc = e["x"].clone().requires_grad_(True)
    orc.sincos_encode(xc, e["L"], orc.barf_weights(0.5, e["barf_start"], e["barf_end"], e["L"])).backward(g)
    close(x.grad, xc.grad, rtol=1e-4, atol=1e-3)   # grads carry a 2^k factor (up to 512 * |g|)


def test_eval_sh_golden_and_backward(mods):
    s = mods["sh"]
    sh, d = s["sh"].to(DEV).requires_grad_(True), s["dirs"].to(DEV).requires_grad_(True)
    out = ops().EvalSHFn.apply(sh, d)
    close(out, s["out"])
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    out.backward(g.to(DEV))
    shc, dc = s["sh"].clone().requires_grad_(True), s["dirs"].clone().requires_grad_(True)
    orc.eval_sh_deg2(shc, dc).backward(g)
    close(sh.grad, shc.grad)
    close(d.grad, dc.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["small", "big"])
def test_mlp_f32_golden(mods, name):
    f = mods[f"mlp_{name}"]
    dep, wid, skips = f["cfg"]
    p = orc.init_mlp_params(dep, wid, skips, seed=f["seed"])
    names = ops().param_names(dep)
    plist = [p[k].to(DEV).requires_grad_(True) for k in names]
    x = f["x_enc"].to(DEV).requires_grad_(True)
    d = f["dirs"].to(DEV).requires_grad_(True)
    out = ops().MLPFn.apply(x, d, dep, wid, skips, *plist)
    close(out, f["out"], rtol=1e-4, atol=2e-5)
    out.backward(f["gout"].to(DEV))
    close(x.grad, f["g_x"], rtol=1e-3, atol=2e-5)
    close(d.grad, f["g_dirs"], rtol=1e-3, atol=2e-5)
    for k, t in zip(names, plist):
        n = f["g_params_norm"][k]
        assert abs(float(t.grad.norm()) - n) <= 1e-3 * max(1.0, n), k
        close(t.grad.reshape(-1)[:256], f["g_params_slice"][k], rtol=1e-3, atol=2e-5)
        if f["g_params"] is not None:
            close(t.grad, f["g_params"][k], rtol=1e-3, atol=2e-5)


def test_mlp_f32_ragged_rows():
    """row counts that are not tile multiples, a single row, and zero rows."""
    dep, wid, skips = 3, 32, (1,)
    p = orc.init_mlp_params(dep, wid, skips, seed=9)
    names = ops().param_names(dep)
    plist = [p[k].to(DEV) for k in names]
    g = torch.Generator().manual_seed(5)
    for M in (1, 129, 300):
        x = torch.randn(M, 63, generator=g)
        d = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
        out = ops().MLPFn.apply(x.to(DEV), d.to(DEV), dep, wid, skips, *plist)
        close(out, orc.mlp_forward(p, x, d, dep, skips), rtol=1e-4, atol=1e-5)
    out = ops().MLPFn.apply(torch.zeros(0, 63, device=DEV), torch.zeros(0, 3, device=DEV), dep, wid, skips, *plist)
    assert out.shape == (0, 4)


def test_compositing_golden(mods):
    c = mods["composite"]
    out4 = c["out4"].to(DEV).requires_grad_(True)
    rgb, dep, opa = ops().CompositeFn.apply(out4, c["noise"].to(DEV), c["rays_d"].to(DEV), c["z"].to(DEV), None,
                                            1.0, 8.0, True)
    close(rgb, c["rgb"])
    close(dep, c["depth"], rtol=1e-5, atol=1e-5)
    close(opa, c["opacity"])
    rgb.backward(c["g_rgb"].to(DEV))
    close(out4.grad, c["g_out4"], rtol=1e-4, atol=1e-6)
    s = mods["s2w"]
    w = ops().sigma2weights(s["sigmas"].to(DEV).contiguous(), s["noise"].to(DEV), deltas=orc.z_deltas(s["z"]).to(DEV).contiguous())
    close(w, s["w"])


@pytest.mark.parametrize("S,B", [(8, 5), (64, 33), (128, 17), (640, 3)])
def test_compositing_sizes(S, B):
    """one chunk, two lanes/ray, four lanes/ray and the reference default of 640 fine samples;
    z from (near, far, jitter) instead of an explicit tensor."""
    g = torch.Generator().manual_seed(S)
    out4 = torch.cat([torch.randn(B, S, 1, generator=g) * 3 - 1, torch.rand(B, S, 3, generator=g)], -1)
    rays_d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)
    jit = torch.rand(B, 1, generator=g) * (7.0 / S)
    noise = torch.randn(B, S, generator=g)
    z = torch.linspace(1.0, 8.0, S).expand(B, -1) + jit
    o_c = out4.clone().requires_grad_(True)
    rgb_r, dep_r, opa_r, w_r = orc.composite(o_c, rays_d, z, noise, True)
    grgb = torch.randn(B, 3, generator=g)
    rgb_r.backward(grgb)
    o_d = out4.to(DEV).requires_grad_(True)
    rgb, dep, opa = ops().CompositeFn.apply(o_d, noise.to(DEV), rays_d.to(DEV), None, jit.to(DEV), 1.0, 8.0, True)
    close(rgb, rgb_r.detach(), rtol=1e-4, atol=2e-6)
    close(dep, dep_r.detach(), rtol=1e-4, atol=1e-5)
    close(opa, opa_r.detach(), rtol=1e-4, atol=2e-6)
    rgb.backward(grgb.to(DEV))
    close(o_d.grad, o_c.grad, rtol=1e-3, atol=2e-6)


def test_selection_matches_nonzero():
    g = torch.Generator().manual_seed(6)
    for B, Sc, scale, thresh in ((37, 64, 2, 1e-3), (5, 8, 2, 1e-3), (9, 128, 5, 1e-3), (4, 16, 3, 10.0)):
        w = torch.rand(B, Sc, generator=g) * (torch.rand(B, Sc, generator=g) < 0.3)
        if thresh > 1:
            w = w * 0 + torch.rand(B, Sc, generator=g) * 1e-6       # max(w) < thresh: threshold becomes the max
        ref = orc.select_fine(w, thresh, scale)
        wd = w.to(DEV)
        wmax = wd.max().reshape(1).clone()
        idx, offs, n = ops().select_fine(wd, wmax, scale, thresh)
        n = int(n.item())
        assert n == ref.shape[0]
        flat_ref = (ref[:, 0] * (Sc * scale) + ref[:, 1]).int()
        assert torch.equal(idx[:n].cpu(), flat_ref)
        cnt = torch.bincount(ref[:, 0], minlength=B)
        assert torch.equal(offs.cpu()[1:] - offs.cpu()[:-1], cnt.int())
    # nothing selected is impossible (max always passes); everything selected:
    w = torch.ones(3, 8)
    idx, offs, n = ops().select_fine(w.to(DEV), torch.ones(1, device=DEV), 2, 1e-3)
    assert int(n.item()) == 48 and torch.equal(idx.cpu(), torch.arange(48, dtype=torch.int32))


def test_scatter_gather_fine():
    g = torch.Generator().manual_seed(7)
    n_dense = 50
    idx = torch.randperm(n_dense, generator=g)[:20].sort().values.int()
    src = torch.randn(20, 4, generator=g)
    s = src.to(DEV).requires_grad_(True)
    dense = ops().ScatterFineFn.apply(s, idx.to(DEV), n_dense, -20.0)
    ref = torch.cat([torch.full((n_dense, 1), -20.0), torch.ones(n_dense, 3)], 1)
    ref[idx.long()] = src
    assert torch.equal(dense.cpu(), ref)
    gd = torch.randn(n_dense, 4, generator=g)
    dense.backward(gd.to(DEV))
    assert torch.equal(s.grad.cpu(), gd[idx.long()])
    empty = ops().ScatterFineFn.apply(torch.zeros(0, 4, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV), 6, -20.0)
    assert torch.equal(empty.cpu(), ref[:6] * 0 + torch.tensor([-20.0, 1, 1, 1]))


def test_reprojection_golden_and_backward(mods):
    c, r = mods["cam"], mods["reproj"]
    K = c["K"].to(DEV).requires_grad_(True)
    Rt = c["calib"].to(DEV).requires_grad_(True)
    pix = ops().ReprojectFn.apply(r["wpts"].to(DEV), K, Rt)
    close(pix, r["out"], rtol=1e-5, atol=1e-4)
    g = torch.randn(pix.shape, generator=torch.Generator().manual_seed(8))
    pix.backward(g.to(DEV))
    Kc, Rc = c["K"].clone().requires_grad_(True), c["calib"].clone().requires_grad_(True)
    orc.reproject(r["wpts"], Kc, Rc).backward(g)
    close(K.grad, Kc.grad, rtol=1e-4, atol=1e-3)
    close(Rt.grad, Rc.grad, rtol=1e-4, atol=1e-2)


def test_device_side_sample_cap_keeps_a_uniform_subset():
    """ref: model/mc_nerf.py:630-632 - more than 128*B selected fine samples in training: keep a uniformly random
    subset of exactly 128*B.  Drawn on the device (render.select_and_cap) without the reference's host round trip."""
    from mc_nerf_b200 import render
    B, Sc, scale = 64, 64, 4
    K = B * 128
    cfg = render.RenderCfg(1.0, 8.0, Sc, scale, 10, True, -20.0, 1e-3, (8, 256, (4,)), (8, 256, (4,)))
    g = torch.Generator().manual_seed(3)
    out_c = torch.randn(B * Sc, 4, generator=g)
    out_c[:, 0] = -2.0              # thin uniform medium: transmittance stays high, most samples pass the threshold
    noise = torch.randn(B, Sc, generator=g).to(DEV)
    jitter = (torch.rand(B, generator=g) * 0.1).to(DEV)
    out_c = out_c.to(DEV)

    def weights(o):
        w_max = torch.zeros(1, device=DEV)
        w = ops.sigma2weights(o, noise, jitter=jitter, near=1.0, far=8.0, sigma_stride=4, n_rays=B, S=Sc, w_max=w_max)
        return w, w_max

    from mc_nerf_b200 import ops
    w_sel, w_max = weights(out_c)
    all_idx, n_all, n_all_dev, _ = render.select_and_cap(cfg, w_sel, w_max, B, train=False)
    n = int(n_all_dev.item())
    assert n > K and n_all == B * Sc * scale
    pool = all_idx[:n].cpu()
    first_half = 0.0
    draws = []
    for _ in range(6):
        idx, n_rows, n_dev, _ = render.select_and_cap(cfg, w_sel, w_max, B, train=True)
        kept = idx[:K].cpu()
        assert n_rows == K and int(n_dev.item()) == K and idx.shape[0] == K
        assert len(torch.unique(kept)) == K and bool(torch.isin(kept, pool).all())
        first_half += float((kept < pool[n // 2]).float().mean())      # pool is ascending: median split
        draws.append(kept)
    assert abs(first_half / 6 - 0.5) < 0.02
    assert not torch.equal(draws[0], draws[1])
    # the kept subset is emitted in ascending slot order, and a given seed reproduces it exactly
    assert bool((draws[0][1:] > draws[0][:-1]).all())
    seed = ops.draw_seed(DEV)
    a1, _, _, _ = render.select_and_cap(cfg, w_sel, w_max, B, train=True, seed=seed)
    a2, _, _, _ = render.select_and_cap(cfg, w_sel, w_max, B, train=True, seed=seed)
    assert torch.equal(a1, a2)
    # uniformity over many draws: every pool member is kept with probability K / n
    hits = torch.zeros(int(pool.max()) + 1)
    R = 40
    for _ in range(R):
        kept = render.select_and_cap(cfg, w_sel, w_max, B, train=True)[0].cpu().long()
        hits[kept] += 1
    p_keep = K / n
    freq = hits[pool.long()] / R
    assert abs(float(freq.mean()) - p_keep) < 1e-6                       # exactly K kept each time
    assert float(freq.std()) < 1.25 * (p_keep * (1 - p_keep) / R) ** 0.5   # binomial spread, no favoured slots
    # fewer than 128*B selected: everything survives (row order is free)
    sparse = out_c.clone()
    sparse[:, 0] = -30.0
    sparse[::7, 0] = 5.0
    w2, wm2 = weights(sparse)
    all2, _, n2_dev, _ = render.select_and_cap(cfg, w2, wm2, B, train=False)
    n2 = int(n2_dev.item())
    assert 0 < n2 <= K
    idx2, n_rows2, n_dev2, _ = render.select_and_cap(cfg, w2, wm2, B, train=True)
    assert n_rows2 == K and int(n_dev2.item()) == n2
    assert torch.equal(torch.sort(idx2[:n2]).values, all2[:n2])
    # sys_param `fine_sample_cap` >= Sf switches the cap off (bench.py --workload stress reports both): training then
    # selects exactly what the un-capped test path selects, per-ray offsets included
    cfg.fine_cap = Sc * scale
    idx3, n_rows3, n_dev3, offs3 = render.select_and_cap(cfg, w_sel, w_max, B, train=True)
    assert n_rows3 == B * Sc * scale and int(n_dev3.item()) == n and offs3 is not None
    assert torch.equal(idx3[:n], all_idx[:n])


@pytest.mark.parametrize("cap,n,K", [(1, 0, 1), (1, 1, 1), (7, 5, 3), (1000, 1000, 1), (1024, 1024, 1023), (1025, 1025, 1024),
                                     (5000, 4000, 4000), (5000, 4001, 4000), (70000, 65537, 1000), (300000, 299999, 150000),
                                     (300000, 20, 150000)])
def test_cap_select_edge_shapes(cap, n, K):
    """mcnerf_cap_select over awkward sizes: empty input, n <= K (everything kept, in order), n = K + 1, capacities that are
    not multiples of the block size, n much smaller than the capacity.  Exactly min(n, K) distinct entries, ascending,
    all from the valid prefix; entries beyond n are never read."""
    from mc_nerf_b200 import ops
    g = torch.Generator().manual_seed(cap + n + K)
    sel = (torch.randperm(cap, generator=g).to(torch.int32) * 3 + 1).to(DEV)
    sel[:n] = torch.sort(sel[:n]).values                 # the selection emits ascending indices
    sel[n:] = -12345                                      # poison: must not appear in the output
    n_dev = torch.tensor([n], dtype=torch.int32, device=DEV)
    seed = torch.tensor([cap * 7919 + n, -K], dtype=torch.int64, device=DEV)
    out, n_out = ops.cap_select(sel, n_dev, K, seed)
    m = min(n, K)
    assert int(n_out.item()) == m and out.shape[0] == K
    kept = out[:m].cpu()
    assert bool(torch.isin(kept, sel[:n].cpu()).all()) and len(torch.unique(kept)) == m
    assert m < 2 or bool((kept[1:] > kept[:-1]).all())
    if n <= K:
        assert torch.equal(kept, sel[:n].cpu())
    out2, _ = ops.cap_select(sel, n_dev, K, seed)
    assert torch.equal(out2[:m].cpu(), kept)              # same seed, same subset
    if n > K + 8:
        seed2 = seed + 1
        out3, _ = ops.cap_select(sel, n_dev, K, seed2)
        assert not torch.equal(out3[:m].cpu(), kept)


def test_p2p_allreduce_kernel_two_virtual_ranks_on_one_gpu():
    """mcnerf_allreduce_p2p (csrc/allreduce.cu): the two-shot NVLink all-reduce, exercised on ONE device by two
    "ranks" whose buffers both live on it and whose kernels run concurrently on two streams - the flag barriers,
    the slice split and the fused all-gather are the same code path as across GPUs (bench.py --gpus N checks the
    multi-GPU case: `ranks_identical`)."""
    import ctypes
    from mc_nerf_b200 import ops
    from mc_nerf_b200._lib import P2P, lib
    n_ctas, count, off = 8, 4 * 50001, 16
    g = torch.Generator().manual_seed(4)
    bufs = [torch.randn(off + count + 8, generator=g).to(DEV) for _ in range(2)]
    flags = [torch.zeros(64 * 2 * 8, dtype=torch.int32, device=DEV) for _ in range(2)]
    epochs = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ctxs = []
    for r in range(2):
        c = P2P()
        c.rank, c.n_ranks, c.n_ctas = r, 2, n_ctas
        for p in range(2):
            c.buf[p], c.flags[p] = bufs[p].data_ptr(), flags[p].data_ptr()
        c.epoch = epochs[r].data_ptr()
        ctxs.append(c)
    torch.cuda.synchronize()
    for it in range(3):                    # epochs advance: the flag pads are never reset
        before = [b.clone() for b in bufs]
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                lib().call("mcnerf_allreduce_p2p", ctypes.byref(ctxs[r]), off, count, 0.5, ops._stream())
        torch.cuda.synchronize()
        want = (before[0][off:off + count] + before[1][off:off + count]) * 0.5
        for r in range(2):
            assert torch.equal(bufs[r][off:off + count], want)                  # identical bits on both "ranks"
            assert torch.equal(bufs[r][:off], before[r][:off]) and torch.equal(bufs[r][off + count:], before[r][off + count:])
        assert int(epochs[0][0]) == it + 1 and int(epochs[1][n_ctas - 1]) == it + 1


def test_fused_camera_model_equals_the_separate_kernels():
    """mcnerf_camera_fwd / _bwd (one launch each way) against IntrinsicsFn + SE3Fn x2 + ReprojectFn, values and all six
    parameter gradients, with the extrinsics trainable (GLOBAL_OPTIM) and frozen (FINE_TUNE, ref: model/mc_nerf.py:87)."""
    from mc_nerf_b200 import ops
    n, P, H, W = 11, 5, 48, 64
    g = torch.Generator().manual_seed(6)
    base = dict(fx=torch.rand(n, generator=g) + 0.5, fy=-(torch.rand(n, generator=g) + 0.5), ux=torch.rand(n, generator=g) + 0.5,
                uy=torch.rand(n, generator=g) + 0.5, pose=torch.randn(n, 6, generator=g) * 0.7, calib=torch.randn(n, 6, generator=g))
    wpts = (torch.rand(1, n, P, 3, generator=g) - 0.5).to(DEV)
    g_kinv = torch.randn(n, 3, 3, generator=g).to(DEV)
    g_pose = torch.randn(n, 3, 4, generator=g).to(DEV)
    g_pix = torch.randn(1, n, P, 2, generator=g).to(DEV)
    for train_pose in (True, False):
        res = []
        for fused in (True, False):
            w = {k: v.clone().to(DEV).requires_grad_(k != "pose" or train_pose) for k, v in base.items()}
            if fused:
                K, Kinv, pose, calib, pix = ops.CameraTrainFn.apply(w["fx"], w["fy"], w["ux"], w["uy"], w["pose"], w["calib"],
                                                                    wpts, H, W)
            else:
                K, Kinv = ops.IntrinsicsFn.apply(w["fx"], w["fy"], w["ux"], w["uy"], H, W)
                pose, calib = ops.SE3Fn.apply(w["pose"]), ops.SE3Fn.apply(w["calib"])
                pix = ops.ReprojectFn.apply(wpts, K, calib)
            loss = (Kinv * g_kinv).sum() + (pix * g_pix).sum() + ((pose * g_pose).sum() if train_pose else 0.0)
            loss.backward()
            res.append(([K, Kinv, pose, calib, pix], {k: v.grad for k, v in w.items()}))
        for a, b in zip(res[0][0], res[1][0]):
            assert torch.equal(a.detach(), b.detach())
        for k in base:
            ga, gb = res[0][1][k], res[1][1][k]
            if k == "pose" and not train_pose:
                assert ga is None and gb is None
            else:
                close(ga, gb, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_eval_sh_every_degree_matches_reference(deg):
    """model.net_utils.eval_sh for the degrees the reference implements (0..4), values and gradients w.r.t. the
    coefficients and the directions, against the unmodified reference's autograd (tests/golden/sh_degrees.pt)."""
    from mc_nerf_b200.model.net_utils import eval_sh
    f = load_golden("sh_degrees.pt")[deg]
    sh = f["sh"].to(DEV).requires_grad_(True)
    d = f["dirs"].to(DEV).requires_grad_(True)
    out = eval_sh(deg, sh, d)
    close(out, f["out"], rtol=1e-5, atol=1e-6)
    out.backward(f["gout"].to(DEV))
    close(sh.grad, f["g_sh"], rtol=1e-5, atol=1e-6)
    close(d.grad, f["g_dirs"], rtol=1e-4, atol=2e-6)


def test_network_with_sh_degree_3_matches_reference():
    """CorseFine_NeRF with MLP_deg = 3 (48 SH coefficients; ref: model/net_block.py:63-65, 75-77) on the fp32 kernels:
    forward, input gradients and every parameter gradient against the unmodified reference."""
    from mc_nerf_b200 import synthetic as syn
    from mc_nerf_b200.model.net_block import CorseFine_NeRF
    f = load_golden("sh_degrees.pt")["mlp_deg3"]
    dep, wid, skips = f["cfg"]
    sp = syn.make_sys_param(n_cam=4, img_h=8, img_w=8, batch=8, samples=8, scale=2, coarse=(dep, wid, skips),
                            fine=(dep, wid, skips), deg=3, device=DEV)
    net = CorseFine_NeRF(sp, type="coarse").to(DEV)
    net.load_state_dict(orc.init_mlp_params(dep, wid, skips, deg=3, seed=f["seed"]))
    x = f["x_enc"].to(DEV).requires_grad_(True)
    d = f["dirs"].to(DEV).requires_grad_(True)
    out = net(x, d)
    close(out, f["out"], rtol=1e-4, atol=1e-5)
    out.backward(f["gout"].to(DEV))
    close(x.grad, f["g_x"], rtol=1e-3, atol=1e-5)
    close(d.grad, f["g_dirs"], rtol=1e-3, atol=1e-5)
    for k, v in net.named_parameters():
        close(v.grad, f["g_params"][k], rtol=1e-3, atol=1e-5)

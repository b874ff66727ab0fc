"""Generate golden fixtures by EXECUTING the unmodified reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Every random draw the reference makes (`torch.randn`, `torch.randperm`,
`Tensor.uniform_`) is replayed from explicit tensors, so the oracle and the CUDA
path can be fed the identical values.  Outputs:

  tests/golden/modules.pt  - per-function in/out pairs (a2-a5, a7, a10-a14 of SURVEY §8a)
  tests/golden/tiny.pt     - a complete tiny train step + test render, all tensors and all gradients
  tests/golden/cfg2.pt     - BASELINE configs[1], the BENCHED configuration (110 cams, 800x800, 4096 rays, 64+128, 8x256):
                             renders, loss, camera grads, MLP grad norms + slices + probe dot products
  tests/golden/cam_stage.pt - CAM_PARAM_EPOCH step (stage 1: reprojection only), every output and gradient
  tests/golden/cfg1.pt     - BASELINE config 1 (110 cams, 100x100, 1024 rays, 64+128, 8x256):
                             outputs, loss, camera grads, per-tensor MLP grad norms + slices.
                             Inputs are regenerated from seeds (checksums stored).
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from _ref_import import import_reference  # noqa: E402
from mc_nerf_b200 import synthetic as syn  # noqa: E402
from oracle import mcnerf_oracle as orc  # noqa: E402

MC_Model, NeRF_Model, MC_NeRF_Loss, net_block, net_utils = import_reference()


class Replay:
    """Patch torch RNG entry points so the reference consumes the given draws, in order."""

    def __init__(self, randn=(), randperm=(), uniform=()):
        self.q = dict(randn=list(randn), randperm=list(randperm), uniform=list(uniform))

    def __enter__(self):
        self._randn, self._randperm, self._uniform = torch.randn, torch.randperm, torch.Tensor.uniform_
        q = self.q

        def randn(*size, **kw):
            t = q["randn"].pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.clone()

        def randperm(n, **kw):
            t = q["randperm"].pop(0)
            assert t.shape[0] == n
            return t.clone()

        def uniform_(self_t, a=0.0, b=1.0):
            t = q["uniform"].pop(0)
            assert t.shape == self_t.shape and float(t.min()) >= a and float(t.max()) <= b
            return self_t.copy_(t)

        torch.randn, torch.randperm, torch.Tensor.uniform_ = randn, randperm, uniform_
        return self

    def __exit__(self, *a):
        torch.randn, torch.randperm, torch.Tensor.uniform_ = self._randn, self._randperm, self._uniform
        assert not any(self.q.values()), {k: len(v) for k, v in self.q.items()}


def checksum(t):
    t = t.double()
    return [float(t.sum()), float(t.abs().sum()), float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape)).sum() / t.numel())]


def build_reference(sp, cam_w, params_c, params_f):
    m = MC_Model(sp)
    with torch.no_grad():
        for k, v in cam_w.items():
            getattr(m, k).copy_(v)
    m.nerf.nerf_coarse.load_state_dict(params_c)
    m.nerf.nerf_fine.load_state_dict(params_f)
    return m


def run_train_step(sp, cam_w, params_c, params_f, batch, rng, stage, step_r):
    m = build_reference(sp, cam_w, params_c, params_f)
    loss_fn = MC_NeRF_Loss(sp)
    with Replay(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]], randperm=[rng["perm"]],
                uniform=[rng["jitter"]]):
        loss_dict, intr_show, pose_show, rays_valid = m(batch, 25, stage, step_r)
        loss = loss_fn(loss_dict, stage)
        loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    out = dict(loss=loss.detach().clone(), rgb_c=loss_dict["rgb"][0].detach().clone(),
               rgb_f=loss_dict["rgb"][1].detach().clone(), gt_sel=loss_dict["rgb"][2].detach().clone(),
               reproj=loss_dict["intr"][0].detach().clone(),
               K=intr_show[1].clone(), pose=pose_show[1].clone())
    return m, out, grads


def make_modules():
    torch.manual_seed(0)
    sp = syn.make_sys_param(n_cam=6, img_h=12, img_w=16, batch=32, samples=8, scale=2,
                            coarse=(4, 64, (2,)), fine=(4, 64, (2,)))
    m = MC_Model(sp)
    g = torch.Generator().manual_seed(5)
    fx = {}
    # a2-a4: camera model
    cam_w = dict(weights_pose=torch.randn(6, 6, generator=g) * 0.7, weights_pose_intr=torch.randn(6, 6, generator=g),
                 weights_fx=torch.rand(6, generator=g) + 0.5, weights_fy=-(torch.rand(6, generator=g) + 0.5),
                 weights_ux=torch.rand(6, generator=g) + 0.5, weights_uy=torch.rand(6, generator=g) + 0.5)
    with torch.no_grad():
        for k, v in cam_w.items():
            getattr(m, k).copy_(v)
    K, pose, calib = m.add_weights2param(True, True, True)
    Kinv = m.inverse_intrinsic(K)
    fx["cam"] = dict(w=cam_w, K=K.detach(), pose=pose.detach(), calib=calib.detach(), Kinv=Kinv.detach(), H=12, W=16)
    # large-angle twists exercise the truncated series
    big = torch.randn(8, 6, generator=g) * 3.0
    fx["se3_big"] = dict(wu=big, Rt=m.se3_to_SE3(big).detach())
    # a5: rays of camera 2
    rd, ro = m.get_rays(pose, torch.tensor([2]), Kinv)
    fx["rays"] = dict(img_id=2, rays_d=rd.detach(), rays_o=ro.detach())
    # a7: reprojection
    wpts = torch.rand(1, 6, 5, 3, generator=g) - 0.5
    fx["reproj"] = dict(wpts=wpts, out=m.get_reproject_pixels(wpts, K, calib).detach())
    # a10: encoding with and without the BARF window
    x = (torch.rand(40, 3, generator=g) - 0.5) * 14
    emb = m.nerf.emmbedding_xyz
    emb.barf_mode = False
    e0 = emb(x, 1.0)
    emb.barf_mode = True
    encs = {r: emb(x, r).detach() for r in (0.40, 0.5, 0.65, 0.75)}
    fx["enc"] = dict(x=x, plain=e0.detach(), barf=encs, barf_start=sp["barf_start"], barf_end=sp["barf_end"], L=10)
    # a12: eval_sh
    sh = torch.randn(40, 3, 9, generator=g)
    d = torch.nn.functional.normalize(torch.randn(40, 3, generator=g), dim=-1)
    fx["sh"] = dict(sh=sh, dirs=d, out=net_utils.eval_sh(2, sh, d))
    # a11: MLPs (default coarse 4x128 skip[2], and 8x256 skip[4])
    for name, (dep, wid, skips) in dict(small=(4, 128, (2,)), big=(8, 256, (4,))).items():
        p = orc.init_mlp_params(dep, wid, skips, seed=3)
        sp2 = dict(sp, coarse_MLP_depth=dep, coarse_MLP_width=wid, coarse_MLP_skip=list(skips))
        net = net_block.CorseFine_NeRF(sp2, type="coarse")
        net.load_state_dict(p)
        xe = e0.detach().clone().requires_grad_(True)
        dd = d.clone().requires_grad_(True)
        out = net(xe, dd)
        gout = torch.randn(out.shape, generator=g)
        out.backward(gout)
        fx[f"mlp_{name}"] = dict(cfg=(dep, wid, skips), seed=3, x_enc=e0.detach(), dirs=d, out=out.detach(), gout=gout,
                                 g_x=xe.grad.clone(), g_dirs=dd.grad.clone(),
                                 g_params=({k: v.grad.clone() for k, v in net.named_parameters()} if name == "small" else None),
                                 g_params_norm={k: float(v.grad.norm()) for k, v in net.named_parameters()},
                                 g_params_slice={k: v.grad.reshape(-1)[:256].clone() for k, v in net.named_parameters()})
    # a13/a14: compositing
    B, S = 9, 16
    out4 = torch.cat([torch.randn(B, S, 1, generator=g) * 4, torch.rand(B, S, 3, generator=g)], -1).requires_grad_(True)
    rays_d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1) * 1.1
    z = torch.linspace(1, 8, S).expand(B, -1) + torch.rand(B, 1, generator=g) * 0.1
    noise = torch.randn(B, S, generator=g)
    nm = m.nerf
    with Replay(randn=[noise]):
        w = nm.sigma2weights(orc.z_deltas(z), out4[..., 0].detach())
    fx["s2w"] = dict(sigmas=out4[..., 0].detach(), z=z, noise=noise, w=w)

    class _Ident(torch.nn.Module):
        def forward(self, x_enc, dirs):
            return self.out

    ident = _Ident()
    ident.out = out4.reshape(-1, 4)
    nm_S = nm.samples_c
    nm.samples_c = S
    with Replay(randn=[noise]):
        rgb, sig, _, dep, opa = nm.inference(ident, lambda x, r: x, 1.0, torch.zeros(B, S, 3), rays_d, z)
    nm.samples_c = nm_S
    grgb = torch.randn(B, 3, generator=g)
    rgb.backward(grgb)
    fx["composite"] = dict(out4=out4.detach(), rays_d=rays_d, z=z, noise=noise, rgb=rgb.detach(), depth=dep.detach(),
                           opacity=opa.detach(), g_rgb=grgb, g_out4=out4.grad.clone())
    return fx


def make_step(name, sp_kw, n_rays, stage, step_r, img_id, full, keep_grads=False):
    sp = syn.make_sys_param(**sp_kw)
    cam_w = syn.init_camera_weights(sp)
    pc = orc.init_mlp_params(sp["coarse_MLP_depth"], sp["coarse_MLP_width"], tuple(sp["coarse_MLP_skip"]), seed=42)
    pf = orc.init_mlp_params(sp["fine_MLP_depth"], sp["fine_MLP_width"], tuple(sp["fine_MLP_skip"]), seed=43)
    batch = syn.make_train_batch(sp, img_id=img_id)
    rng = syn.draw_step_rng(sp, n_rays, seed=123)
    t0 = time.time()
    m, out, grads = run_train_step(sp, cam_w, pc, pf, batch, rng, stage, step_r)
    dt = time.time() - t0
    # test-mode render of the same rays with the same nets
    rays_d, rays_o = m.get_rays(out["pose"], batch[1], m.inverse_intrinsic(out["K"]))
    rd, ro = rays_d[rng["rand_idx"]].detach(), rays_o[rng["rand_idx"]].detach()
    with torch.no_grad(), Replay(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]]):
        t_rgb, t_dep, t_opa = m.nerf.render_rays_test(rd, ro, m.nerf.nerf_coarse, m.nerf.nerf_fine)
    fx = dict(name=name, sp_kw=sp_kw, n_rays=n_rays, stage=stage, step_r=step_r, img_id=img_id, ref_seconds=dt,
              rays_d=rd if full else rd[:16], rays_o=ro if full else ro[:16],
              test=dict(rgb=t_rgb, depth=t_dep, opacity=t_opa), **out)
    cam_keys = ["weights_pose", "weights_pose_intr", "weights_fx", "weights_fy", "weights_ux", "weights_uy"]
    fx["g_cam"] = {k: grads[k] for k in cam_keys}
    if full:
        fx["g_mlp"] = {k: v for k, v in grads.items() if k.startswith("nerf.")}
        fx["inputs"] = dict(cam_w=cam_w, pc=pc, pf=pf, batch=batch, rng={k: v for k, v in rng.items() if k != "perm"})
    else:
        fx["g_mlp_norm"] = {k: float(v.norm()) for k, v in grads.items() if k.startswith("nerf.")}
        fx["g_mlp_slice"] = {k: v.reshape(-1)[:64].clone() for k, v in grads.items() if k.startswith("nerf.")}
        fx["checksums"] = dict(gt=checksum(batch[0]), noise_f=checksum(rng["noise_f"]), jitter=checksum(rng["jitter"]),
                               rand_idx=checksum(rng["rand_idx"]), w_c0=checksum(pc["xyz_encoding_1.0.weight"]),
                               w_f7=checksum(pf["xyz_encoding_8.0.weight"]), pose_w=checksum(cam_w["weights_pose"]))
    if keep_grads:
        fx["_grads"] = grads
    print(f"{name}: reference step {dt:.2f}s loss={float(out['loss']):.6f}")
    return fx


PROBES = 8


def probe_dots(name_index, g):
    """Dot products of a gradient tensor with PROBES seeded N(0,1) vectors: a direction-sensitive
    fingerprint (E[(u.e)^2] = |e|^2 for an error vector e) that stays tiny for the 631 836-parameter networks."""
    gen = torch.Generator().manual_seed(90000 + name_index)
    u = torch.randn(PROBES, g.numel(), generator=gen, dtype=torch.float64)
    return (u @ g.reshape(-1).double()).float()


def make_cfg2():
    """BASELINE configs[1] - the benched configuration: 110 cameras, 800x800, 4096 rays, 64+128, both nets 8x256.
    Inputs come from seeds (checksums stored); stored: renders, loss, all six camera gradients, and for every MLP
    tensor its norm, a 64-element slice and PROBES probe dot products."""
    kw = dict(n_cam=110, img_h=800, img_w=800, batch=4096, samples=64, scale=2, with_images=False)
    fx = make_step("cfg2", kw, 4096, "GLOBAL_OPTIM_EPOCH", 0.5, 3, False, keep_grads=True)
    grads = fx.pop("_grads")
    names = sorted(k for k in grads if k.startswith("nerf."))
    fx["g_mlp_probe"] = {k: probe_dots(i, grads[k]) for i, k in enumerate(names)}
    fx["g_mlp_small"] = {k: grads[k].clone() for k in names if grads[k].numel() <= 8192}     # biases, sigma.2, sh.2: in full
    fx["probe_check"] = checksum(torch.randn(4, 7, generator=torch.Generator().manual_seed(90000), dtype=torch.float64))
    return fx


def make_cfg2_bf16emu():
    """The benched configuration through the ORACLE with the tcgen05 path's rounding points emulated
    (tests/bf16_emu.py): what a correct bf16-in / fp32-accumulate implementation must produce, to which the kernels
    are compared tightly.  (Against the fp32 reference some gradients are ill-conditioned at random init - the coarse
    sigma head: the colours along a ray are nearly constant, so d rgb / d sigma is a small difference of large terms -
    and no bf16 forward can reproduce them to better than tens of percent; the emulation shows exactly that.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bf16_emu
    kw = dict(n_cam=110, img_h=800, img_w=800, batch=4096, samples=64, scale=2, with_images=False)
    sp = syn.make_sys_param(**kw)
    cfg = orc.cfg_from_sys_param(sp)
    cam = {k: v.clone().requires_grad_(True) for k, v in syn.init_camera_weights(sp).items()}
    pc = {k: v.clone().requires_grad_(True) for k, v in orc.init_mlp_params(*cfg["coarse"], seed=42).items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in orc.init_mlp_params(*cfg["fine"], seed=43).items()}
    batch = syn.make_train_batch(sp, img_id=3)
    rng = syn.draw_step_rng(sp, 4096, seed=123)
    plain = orc.mlp_forward
    orc.mlp_forward = lambda p, x, d, depth, skips, deg=2: bf16_emu.mlp_forward_bf16emu_trainable(p, x, d, depth, skips)
    try:
        out = orc.train_step(cam, pc, pf, cfg, batch, rng, step_r=0.5, stage="GLOBAL_OPTIM_EPOCH")
    finally:
        orc.mlp_forward = plain
    grads = {f"nerf.nerf_coarse.{k}": v.grad for k, v in pc.items()}
    grads.update({f"nerf.nerf_fine.{k}": v.grad for k, v in pf.items()})
    names = sorted(grads)
    return dict(name="cfg2_bf16emu", sp_kw=kw, n_rays=4096, stage="GLOBAL_OPTIM_EPOCH", step_r=0.5, img_id=3,
                loss=out["loss"], rgb_c=out["rgb_c"], rgb_f=out["rgb_f"],
                g_cam={k: v.grad.clone() for k, v in cam.items()},
                g_mlp_norm={k: float(grads[k].norm()) for k in names},
                g_mlp_probe={k: probe_dots(i, grads[k]) for i, k in enumerate(names)},
                g_mlp_small={k: grads[k].clone() for k in names if grads[k].numel() <= 8192})


def make_cam_stage():
    """CAM_PARAM_EPOCH (stage 1, ref model/mc_nerf.py:64-71, loss.py:18-26): reprojection of the calibration points
    through both pose sets, un-normalised loss, NeRF untouched.  No random draws."""
    kw = dict(n_cam=5, img_h=10, img_w=12, batch=24, samples=8, scale=2, coarse=(3, 32, (1,)), fine=(4, 64, (2,)))
    sp = syn.make_sys_param(**kw)
    cam_w = syn.init_camera_weights(sp)
    g = torch.Generator().manual_seed(77)
    cam_w["weights_pose_intr"] = cam_w["weights_pose"] + 0.05 * torch.randn(5, 6, generator=g)
    pc = orc.init_mlp_params(3, 32, (1,), seed=42)
    pf = orc.init_mlp_params(4, 64, (2,), seed=43)
    batch = syn.make_train_batch(sp, img_id=2)
    m = build_reference(sp, cam_w, pc, pf)
    loss_fn = MC_NeRF_Loss(sp)
    loss_dict, intr_show, pose_show, rays_valid = m(batch, 3, "CAM_PARAM_EPOCH", 0.1)
    loss = loss_fn(loss_dict, "CAM_PARAM_EPOCH")
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    assert m.opt_idx == 0 and "rgb" not in loss_dict
    return dict(name="cam_stage", sp_kw=kw, stage="CAM_PARAM_EPOCH", img_id=2,
                inputs=dict(cam_w=cam_w, pc=pc, pf=pf, batch=batch),
                loss=loss.detach().clone(), reproj_intr=loss_dict["intr"][0].detach().clone(),
                reproj_extr=loss_dict["extr"][0].detach().clone(), K=intr_show[1].clone(), pose=pose_show[1].clone(),
                rays_valid_d=rays_valid[0].clone(), rays_valid_o=rays_valid[1].clone(),
                g_cam={k: v for k, v in grads.items() if not k.startswith("nerf.")},
                g_mlp_none=all(v is None for k, v in grads.items() if k.startswith("nerf.")))


def make_sh_degrees():
    """eval_sh for every degree the reference implements (0..4, model/net_utils.py:103-191) with autograd gradients, and a
    CorseFine_NeRF with MLP_deg = 3 (model/net_block.py:63-65, 75-77: SH head of 3 * 16 coefficients), fwd + bwd."""
    g = torch.Generator().manual_seed(17)
    n = 33
    fx = {}
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    for deg in range(5):
        sh = torch.randn(n, 3, (deg + 1) ** 2, generator=g).requires_grad_(True)
        dd = d.clone().requires_grad_(True)
        out = net_utils.eval_sh(deg, sh, dd)
        gout = torch.randn(n, 3, generator=g)
        out.backward(gout)
        fx[deg] = dict(sh=sh.detach().clone(), dirs=d, out=out.detach(), gout=gout, g_sh=sh.grad.clone(),
                       g_dirs=(dd.grad.clone() if dd.grad is not None else torch.zeros_like(d)))
    sp = syn.make_sys_param(n_cam=4, img_h=8, img_w=8, batch=8, samples=8, scale=2, coarse=(3, 64, (1,)), fine=(3, 64, (1,)),
                            deg=3)
    p = orc.init_mlp_params(3, 64, (1,), deg=3, seed=5)
    net = net_block.CorseFine_NeRF(sp, type="coarse")
    net.load_state_dict(p)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 6
    xe = orc.sincos_encode(x, 10).requires_grad_(True)
    dd = d.clone().requires_grad_(True)
    out = net(xe, dd)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    fx["mlp_deg3"] = dict(cfg=(3, 64, (1,)), seed=5, x_enc=xe.detach().clone(), dirs=d, out=out.detach(), gout=gout,
                          g_x=xe.grad.clone(), g_dirs=dd.grad.clone(),
                          g_params={k: v.grad.clone() for k, v in net.named_parameters()})
    return fx


def make_radam():
    """Trajectory of the reference's RAdam (model/net_utils.py:10-101) over 14 steps: crosses the N_sma >= 5
    switch (step 6) and wraps the 10-slot step-size cache."""
    g = torch.Generator().manual_seed(8)
    p0 = [torch.randn(37, 5, generator=g), torch.randn(11, generator=g)]
    p = [t.clone().requires_grad_(True) for t in p0]
    lr, wd = 5e-3, 4e-4
    opt = net_utils.RAdam([dict(params=p[:1]), dict(params=p[1:], lr=lr * 0.5)], lr=lr, weight_decay=wd)
    grads, traj = [], {}
    for step in range(14):
        gs = [torch.randn(t.shape, generator=g) for t in p]
        for t, gg in zip(p, gs):
            t.grad = gg.clone()
        opt.step()
        grads.append(gs)
        traj[step] = [t.detach().clone() for t in p]
    return dict(p0=p0, lr=lr, wd=wd, grads=grads, traj=traj)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    if only:       # python tests/golden/make_golden.py cfg2 cam_stage : just these fixtures
        if "cfg2" in only:
            torch.save(make_cfg2(), os.path.join(HERE, "cfg2.pt"))
        if "cfg2_bf16emu" in only:
            torch.save(make_cfg2_bf16emu(), os.path.join(HERE, "cfg2_bf16emu.pt"))
        if "sh_degrees" in only:
            torch.save(make_sh_degrees(), os.path.join(HERE, "sh_degrees.pt"))
        if "cam_stage" in only:
            torch.save(make_cam_stage(), os.path.join(HERE, "cam_stage.pt"))
        sys.exit(0)
    torch.save(make_radam(), os.path.join(HERE, "radam.pt"))
    torch.save(make_modules(), os.path.join(HERE, "modules.pt"))
    tiny_kw = dict(n_cam=5, img_h=10, img_w=12, batch=24, samples=8, scale=2, coarse=(3, 32, (1,)), fine=(4, 64, (2,)))
    torch.save(make_step("tiny", tiny_kw, 24, "GLOBAL_OPTIM_EPOCH", 0.5, 3, True), os.path.join(HERE, "tiny.pt"))
    torch.save(make_step("tiny_ft", tiny_kw, 24, "FINE_TUNE_EPOCH", 0.9, 1, True), os.path.join(HERE, "tiny_ft.pt"))
    cfg1_kw = dict(n_cam=110, img_h=100, img_w=100, batch=1024, samples=64, scale=2)
    torch.save(make_step("cfg1", cfg1_kw, 1024, "GLOBAL_OPTIM_EPOCH", 0.5, 3, False), os.path.join(HERE, "cfg1.pt"))
    torch.save(make_cfg2(), os.path.join(HERE, "cfg2.pt"))
    torch.save(make_cam_stage(), os.path.join(HERE, "cam_stage.pt"))
    torch.save(make_cfg2_bf16emu(), os.path.join(HERE, "cfg2_bf16emu.pt"))
    torch.save(make_sh_degrees(), os.path.join(HERE, "sh_degrees.pt"))
    for f in ("modules.pt", "tiny.pt", "tiny_ft.pt", "cfg1.pt", "cfg2.pt", "cam_stage.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")

"""Parity at the BENCHED configuration - BASELINE configs[1]: 110 cameras, 800x800 images, 4096 rays per batch,
64 coarse + 128 fine samples, both networks 8x256, GLOBAL_OPTIM stage - against a golden produced by running the
UNMODIFIED reference on the same seeded inputs (tests/golden/make_golden.py::make_cfg2, ref model/mc_nerf.py:73-83,
model/loss.py:15-31, main.py:79-84).  Checked for the fp32 path AND the bf16 tcgen05 path bench.py times:
renders, loss, all six camera-parameter gradients, and every MLP gradient by norm, a 64-element slice and PROBES
random-probe dot products (direction-sensitive: the rms probe difference estimates |g - g_ref|).

Also: the bf16 `render_rays_test` outputs (rgb / depth / opacity) against the reference, and the CAM_PARAM_EPOCH
(stage 1) step against its own golden.

Every threshold below is <= 3x the value measured on a B200 (printed by the test; the measured numbers are in
DESIGN.md section 2).  A seeded input that does not regenerate identically FAILS the test - it never skips.
"""
import json
import os

import pytest
import torch

from conftest import load_golden, ROOT
from mc_nerf_b200 import synthetic as syn
from oracle import mcnerf_oracle as orc
from replay import Replay
from tests_checksum import checksum, probe_dots

pytestmark = pytest.mark.gpu
DEV = "cuda"

# Thresholds = at most 3x the values measured on a B200 in round 2 (gpurun_out/parity_cfg2_*.json, quoted per entry)
TOL = {
    # measured: loss 0, rgb max 3.6e-7 (0 rays over), cam 2.6e-5, MLP norm 6.1e-5 / probe 6.8e-4 (coarse sigma head
    # 1.8e-3), max-abs 2.4e-8
    "fp32": dict(loss=1e-6, rgb_tol=1e-5, rgb_bad=4, cam=8e-5, mlp_norm=2e-4, mlp_probe=2e-3, mlp_probe_ill=6e-3,
                 mlp_maxabs=8e-8),
    # measured: loss < 1e-7, rgb max 1.8e-5 (0 rays over), cam 6.6e-3 (weights_pose), MLP norm 2.8e-3 / probe 5.2e-2
    # (coarse sigma head 0.41, see ILL-CONDITIONED), max-abs 1.4e-6;
    # vs the emulated oracle: rgb 4.8e-6, cam 5.7e-4, probe 1.0e-2 (coarse sigma head 6.2e-2)
    "bf16": dict(loss=2e-5, rgb_tol=6e-5, rgb_bad=4, cam=2e-2, mlp_norm=9e-3, mlp_probe=1.6e-1, mlp_probe_ill=1.2,
                 mlp_maxabs=5e-6, emu_rgb=1.5e-5, emu_cam=2e-3, emu_probe=1.9e-1),
}


def _report(name, payload):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
            json.dump(payload, f, indent=1)
    except OSError:
        pass
    print(name, json.dumps(payload))


def _inputs(fx):
    sp0 = syn.make_sys_param(**fx["sp_kw"])
    cfg = orc.cfg_from_sys_param(sp0)
    cam_w = syn.init_camera_weights(sp0)
    pc = orc.init_mlp_params(*cfg["coarse"], seed=42)
    pf = orc.init_mlp_params(*cfg["fine"], seed=43)
    batch = syn.make_train_batch(sp0, img_id=fx["img_id"])
    rng = syn.draw_step_rng(sp0, fx["n_rays"], seed=123)
    cs = fx["checksums"]
    got = dict(gt=checksum(batch[0]), noise_f=checksum(rng["noise_f"]), jitter=checksum(rng["jitter"]),
               rand_idx=checksum(rng["rand_idx"]), w_c0=checksum(pc["xyz_encoding_1.0.weight"]),
               w_f7=checksum(pf["xyz_encoding_8.0.weight"]), pose_w=checksum(cam_w["weights_pose"]))
    for k in cs:
        assert all(abs(a - b) <= 1e-6 * max(1.0, abs(b)) for a, b in zip(got[k], cs[k])), \
            f"seeded input {k} does not regenerate on this torch build: the benched-config parity cannot be checked"
    return sp0, cam_w, pc, pf, batch, rng


def _model(fx, cam_w, pc, pf, precision):
    from mc_nerf_b200.model import MC_Model
    sp = syn.make_sys_param(device=DEV, **fx["sp_kw"])
    sp["mlp_precision"] = precision
    m = MC_Model(sp).to(DEV)
    with torch.no_grad():
        for k, v in cam_w.items():
            getattr(m, k).copy_(v)
    m.nerf.nerf_coarse.load_state_dict(pc)
    m.nerf.nerf_fine.load_state_dict(pf)
    return sp, m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_benched_config_train_step_matches_reference(precision):
    from mc_nerf_b200.model import MC_NeRF_Loss
    from mc_nerf_b200 import render
    fx = load_golden("cfg2.pt")
    sp0, cam_w, pc, pf, batch, rng = _inputs(fx)
    sp, m = _model(fx, cam_w, pc, pf, precision)
    if precision == "bf16":
        assert render.use_tc(m.nerf.render_cfg, m.nerf.render_cfg.fine)
    loss_fn = MC_NeRF_Loss(sp)
    with Replay(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]], randperm=[rng["perm"]],
                uniform=[rng["jitter"]]):
        loss_dict, _, _, _ = m(batch, 25, fx["stage"], fx["step_r"])
        loss = loss_fn(loss_dict, fx["stage"])
    loss.backward()
    tol = TOL[precision]
    rep = dict(precision=precision, loss_abs_err=abs(float(loss) - float(fx["loss"])))
    assert rep["loss_abs_err"] <= tol["loss"], rep
    # the fine-sample gate is discontinuous: a ray whose gate flipped under rounding may move
    for key, idx in (("rgb_c", 0), ("rgb_f", 1)):
        err = (loss_dict["rgb"][idx].detach().cpu() - fx[key]).abs()
        rep[key] = dict(max=float(err.max()), q999=float(err.flatten().quantile(0.999)),
                        rays_over=int((err > tol["rgb_tol"]).any(-1).sum()))
    named = dict(m.named_parameters())
    rep["cam"] = {k: float((named[k].grad.cpu() - g).norm() / g.norm().clamp_min(1e-12)) for k, g in fx["g_cam"].items()}
    names = sorted(fx["g_mlp_norm"])
    rep["mlp_norm"], rep["mlp_probe"], rep["mlp_slice"] = {}, {}, {}
    for i, k in enumerate(names):
        g = named[k].grad
        n = fx["g_mlp_norm"][k]
        rep["mlp_norm"][k] = abs(float(g.norm()) - n) / max(n, 1e-12)
        rep["mlp_probe"][k] = float((probe_dots(i, g) - fx["g_mlp_probe"][k]).pow(2).mean().sqrt()) / max(n, 1e-12)
        s_ref = fx["g_mlp_slice"][k]
        rep["mlp_slice"][k] = float((g.reshape(-1)[:64].cpu() - s_ref).norm() / s_ref.norm().clamp_min(1e-12))
    # north_star states the gradient tolerance as a max-abs error: checked in full on every small tensor
    # (biases, sigma.2, sh.2) and on the 64-element slices of the large ones
    rep["mlp_maxabs"] = {k: float((named[k].grad.cpu() - g).abs().max()) for k, g in fx["g_mlp_small"].items()}
    rep["mlp_maxabs_slice"] = {k: float((named[k].grad.reshape(-1)[:64].cpu() - fx["g_mlp_slice"][k]).abs().max())
                               for k in names}
    rep["mlp_absmax_ref"] = {k: float(g.abs().max()) for k, g in fx["g_mlp_small"].items()}
    ill = [k for k in names if k.startswith("nerf.nerf_coarse.sigma.")]         # see ILL-CONDITIONED below
    well = [k for k in names if k not in ill]
    rep["worst"] = dict(cam=max(rep["cam"].values()), mlp_norm=max(rep["mlp_norm"][k] for k in well),
                        mlp_probe=max(rep["mlp_probe"][k] for k in well),
                        mlp_probe_ill=max(rep["mlp_probe"][k] for k in ill),
                        mlp_maxabs=max(list(rep["mlp_maxabs"].values()) + list(rep["mlp_maxabs_slice"].values())))
    if precision == "bf16":
        # ... and against the oracle with the path's rounding points emulated (tests/bf16_emu.py): tight everywhere
        emu = load_golden("cfg2_bf16emu.pt")
        rep["emu"] = dict(loss=abs(float(loss.detach()) - float(emu["loss"])),
                          rgb_c=float((loss_dict["rgb"][0].detach().cpu() - emu["rgb_c"]).abs().max()),
                          rgb_f_q999=float((loss_dict["rgb"][1].detach().cpu() - emu["rgb_f"]).abs().flatten().quantile(0.999)),
                          cam={k: float((named[k].grad.cpu() - g).norm() / g.norm().clamp_min(1e-12))
                               for k, g in emu["g_cam"].items()},
                          mlp_probe={k: float((probe_dots(i, named[k].grad) - emu["g_mlp_probe"][k]).pow(2).mean().sqrt())
                                     / max(emu["g_mlp_norm"][k], 1e-12) for i, k in enumerate(names)})
        rep["worst"]["emu_probe"] = max(rep["emu"]["mlp_probe"].values())
        rep["worst"]["emu_cam"] = max(rep["emu"]["cam"].values())
    _report(f"cfg2_{precision}", rep)
    for key in ("rgb_c", "rgb_f"):
        assert rep[key]["rays_over"] <= tol["rgb_bad"], (key, rep[key])
    for k, v in rep["cam"].items():
        assert v <= tol["cam"], (k, v)
    for k in well:
        assert rep["mlp_norm"][k] <= tol["mlp_norm"], (k, rep["mlp_norm"][k])
        assert rep["mlp_probe"][k] <= tol["mlp_probe"], (k, rep["mlp_probe"][k])
    # ILL-CONDITIONED at random init: the coarse network's sigma head.  The colours along a coarse ray are nearly
    # constant, so d rgb / d sigma is a small difference of large terms; its gradient is ~1e-4 of the network's
    # gradient norm and ANY bf16 forward moves it by tens of percent (the emulated oracle differs from the fp32
    # reference by the same 0.12-0.37, tests/golden/make_golden.py::make_cfg2_bf16emu).  It is bounded in absolute
    # terms (north_star's form of the tolerance) and, for bf16, tightly against the emulation.
    for k in ill:
        assert rep["mlp_probe"][k] <= tol["mlp_probe_ill"], (k, rep["mlp_probe"][k])
    assert rep["worst"]["mlp_maxabs"] <= tol["mlp_maxabs"], rep["worst"]
    if precision == "bf16":
        assert rep["emu"]["loss"] <= 1e-6 and rep["emu"]["rgb_c"] <= tol["emu_rgb"], rep["emu"]
        assert rep["worst"]["emu_cam"] <= tol["emu_cam"] and rep["worst"]["emu_probe"] <= tol["emu_probe"], rep["worst"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_benched_config_test_render_matches_reference(precision):
    """render_rays_test (ref model/mc_nerf.py:648-680) of the same 4096 rays: rgb, depth and opacity."""
    fx = load_golden("cfg2.pt")
    sp0, cam_w, pc, pf, batch, rng = _inputs(fx)
    sp, m = _model(fx, cam_w, pc, pf, precision)
    # the golden's rays are those of the learnable camera model at the selected pixels
    with torch.no_grad():
        from mc_nerf_b200 import ops
        intr, pose, _ = m.add_weights2param(True, True, True)
        B = fx["n_rays"]
        cam = torch.full((B,), fx["img_id"], dtype=torch.int32, device=DEV)
        rays_o, rays_d = ops.RaygenFn.apply(m.inverse_intrinsic(intr), pose, cam,
                                            rng["rand_idx"].to(DEV).to(torch.int32), B, sp["data_img_w"])
        assert float((rays_d[:16].cpu() - fx["rays_d"]).abs().max()) < 1e-5
        with Replay(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]]):
            rgb, dep, opa = m.nerf.render_rays_test(rays_d.contiguous(), rays_o.contiguous(), m.nerf.nerf_coarse,
                                                    m.nerf.nerf_fine)
    t = fx["test"]
    rep = {}
    for name, a, b in (("rgb", rgb, t["rgb"]), ("depth", dep, t["depth"]), ("opacity", opa, t["opacity"])):
        err = (a.detach().cpu().reshape(b.shape) - b).abs()
        rep[name] = dict(max=float(err.max()), q999=float(err.flatten().quantile(0.999)), mean=float(err.mean()))
    _report(f"cfg2_test_render_{precision}", rep)
    # measured q999: fp32 rgb 2.9e-7 / depth 1.9e-6 / opacity 8.9e-7; bf16 rgb 1.35e-5 / depth 2.5e-5 / opacity 8.9e-7
    lim = dict(fp32=dict(rgb=1e-6, depth=6e-6, opacity=3e-6), bf16=dict(rgb=4e-5, depth=8e-5, opacity=3e-6))[precision]
    for k in ("rgb", "depth", "opacity"):
        assert rep[k]["q999"] <= lim[k], (k, rep[k])      # 99.9 % of the rays; flipped-gate rays excepted
        assert rep[k]["mean"] <= lim[k], (k, rep[k])
        assert rep[k]["max"] <= 1e-3 if k != "depth" else rep[k]["max"] <= 2e-2, (k, rep[k])


def test_tiny_test_render_bf16_matches_reference():
    """bf16 render_rays_test against the reference's tiny golden (narrow nets on the padded tensor-core path)."""
    fx = load_golden("tiny.pt")
    i = fx["inputs"]
    from mc_nerf_b200.model import MC_Model
    sp = syn.make_sys_param(device=DEV, **fx["sp_kw"])
    sp["mlp_precision"] = "bf16"
    m = MC_Model(sp).to(DEV)
    m.nerf.nerf_coarse.load_state_dict(i["pc"])
    m.nerf.nerf_fine.load_state_dict(i["pf"])
    rng = i["rng"]
    with torch.no_grad(), Replay(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]]):
        rgb, dep, opa = m.nerf.render_rays_test(fx["rays_d"].to(DEV), fx["rays_o"].to(DEV), m.nerf.nerf_coarse,
                                                m.nerf.nerf_fine)
    t = fx["test"]
    rep = dict(rgb=float((rgb.cpu() - t["rgb"]).abs().max()), depth=float((dep.cpu() - t["depth"]).abs().max()),
               opacity=float((opa.cpu() - t["opacity"]).abs().max()))
    _report("tiny_test_render_bf16", rep)
    # measured: rgb 4.7e-5, depth 9.3e-5, opacity 2.4e-7
    assert rep["rgb"] <= 1.5e-4 and rep["opacity"] <= 1e-6 and rep["depth"] <= 3e-4, rep


def test_camera_stage_step_matches_reference():
    """CAM_PARAM_EPOCH (stage 1, ref model/mc_nerf.py:64-71 + loss.py:18-26): both reprojections, the un-normalised
    loss, every camera gradient; the NeRF receives no gradient; opt_idx = 0."""
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss
    fx = load_golden("cam_stage.pt")
    i = fx["inputs"]
    sp = syn.make_sys_param(device=DEV, **fx["sp_kw"])
    sp["mlp_precision"] = "fp32"
    m = MC_Model(sp).to(DEV)
    with torch.no_grad():
        for k, v in i["cam_w"].items():
            getattr(m, k).copy_(v)
    m.nerf.nerf_coarse.load_state_dict(i["pc"])
    m.nerf.nerf_fine.load_state_dict(i["pf"])
    loss_dict, intr_show, pose_show, rays_valid = m(i["batch"], 3, "CAM_PARAM_EPOCH", 0.1)
    loss = MC_NeRF_Loss(sp)(loss_dict, "CAM_PARAM_EPOCH")
    loss.backward()
    assert m.opt_idx == 0 and "rgb" not in loss_dict and m.nerf.emmbedding_xyz.barf_mode is False

    def close(a, b, rtol, atol):
        torch.testing.assert_close(a.detach().cpu(), b, rtol=rtol, atol=atol)
    close(loss, fx["loss"], 1e-5, 1e-6)
    close(loss_dict["intr"][0], fx["reproj_intr"], 1e-4, 1e-3)
    close(loss_dict["extr"][0], fx["reproj_extr"], 1e-4, 1e-3)
    close(intr_show[1], fx["K"], 1e-5, 1e-5)
    close(pose_show[1], fx["pose"], 1e-5, 1e-6)
    close(rays_valid[0], fx["rays_valid_d"], 1e-5, 1e-6)
    close(rays_valid[1], fx["rays_valid_o"], 1e-5, 1e-6)
    named = dict(m.named_parameters())
    for k, g in fx["g_cam"].items():
        close(named[k].grad, g, 2e-3, 1e-7)
    assert all(p.grad is None for k, p in named.items() if k.startswith("nerf."))

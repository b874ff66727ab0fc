"""Shared harness of tests/test_main_integration*.py: run the reference's UNMODIFIED main.py (Model_Engine) on the
synthetic rig, with `data` replaced by baseline/synthetic_data.py and `model` either the reference's own package
(CPU) or the drop-in mc_nerf_b200.model (GPU)."""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import ref_loader, synthetic_data  # noqa: E402


def reference_available():
    if ref_loader.reference_root() is None and os.path.isdir("/root/reference"):
        from baseline import install_ref
        install_ref.install(verbose=False)
    return ref_loader.reference_root() is not None


def load_main(drop_in):
    model_pkg = None
    if drop_in:
        import importlib
        model_pkg = importlib.import_module("mc_nerf_b200.model")
    main = ref_loader.import_reference_main(model_package=model_pkg, data_module=synthetic_data)
    if drop_in:
        assert main.MC_Model.__module__.startswith("mc_nerf_b200."), main.MC_Model.__module__
    else:
        assert main.MC_Model.__module__ == "model.mc_nerf", main.MC_Model.__module__
    try:
        import matplotlib  # noqa: F401
        has_mpl = not isinstance(sys.modules["matplotlib"].__dict__.get("__getattr__"), type(load_main))
    except Exception:
        has_mpl = False
    if not has_mpl:
        # demo post-processing only (main.py:117-118): matplotlib's colour table is not installed here
        main.apply_depth_colormap = lambda depth, cmap="inferno": torch.clip(depth, 0, 1).expand(-1, 3)
    return main


class Recorder:
    """Wraps the engine's loss function to record (stage, loss, opt_idx) per step without touching main.py."""

    def __init__(self, engine):
        self.engine, self.log = engine, []
        self._inner = engine.loss_func

        def wrapped(loss_dict, epoch_type):
            loss = self._inner(loss_dict, epoch_type)
            self.log.append((epoch_type, float(loss.detach()), sorted(loss_dict)))
            return loss
        engine.loss_func = wrapped


def run_training(main, sp):
    """Model_Engine(sys_param).train_model(): 3 stages x steps_per_epoch steps, per-epoch checkpoint + validation."""
    main.sys_param = sp                      # main.py reads this module global inside train_model (main.py:60,72)
    torch.manual_seed(sp["seed"])
    engine = main.Model_Engine(sp)
    rec = Recorder(engine)
    before = {k: v.detach().clone() for k, v in engine.mc_nerf.state_dict().items()}
    engine.forward()
    after = engine.mc_nerf.state_dict()
    ckpts = sorted(glob.glob(os.path.join(sp["root_weight"], "train", "*.ckpt")), key=os.path.getmtime)
    return engine, rec, before, after, ckpts


def check_training(engine, rec, before, after, ckpts, steps):
    stages = [s for s, _, _ in rec.log]
    assert stages == ["CAM_PARAM_EPOCH"] * steps + ["GLOBAL_OPTIM_EPOCH"] * steps + ["FINE_TUNE_EPOCH"] * steps, stages
    assert all(l == l and abs(l) < 1e6 for _, l, _ in rec.log), rec.log           # finite
    assert rec.log[0][2] == ["extr", "intr"] and rec.log[steps][2] == ["intr", "rgb"]
    moved = {k: float((after[k].detach().cpu() - before[k].cpu()).abs().max()) for k in before}
    # every stage's optimiser moved its parameters: cameras (stage 1+), both networks (stage 2+)
    for k in ("weights_pose", "weights_pose_intr", "weights_fx", "weights_fy", "weights_ux", "weights_uy",
              "nerf.nerf_coarse.xyz_encoding_1.0.weight", "nerf.nerf_fine.sh.2.bias"):
        assert moved[k] > 0, (k, moved[k])
    assert len(ckpts) == 3
    ck = torch.load(ckpts[-1], map_location="cpu")
    assert set(ck) == {"model_nerf"} and set(ck["model_nerf"]) == set(before)
    assert engine.mc_nerf.opt_idx == 2


def run_demo(main, sp_demo):
    main.sys_param = sp_demo
    engine = main.Model_Engine(sp_demo)
    engine.forward()
    out = sorted(glob.glob(sp_demo["demo_render_pth"] + "_*"))
    assert out, "no render directory written"
    n = sp_demo["synthetic_rig"]["n_cam"]
    for sub in ("pred", "depth", "gt"):
        assert len(glob.glob(os.path.join(out[-1], sub, "*.png"))) == n, sub
    return out[-1]

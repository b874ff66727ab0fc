"""CPU: the harness of tests/test_main_integration_gpu.py drives the reference's UNMODIFIED main.py
(Model_Engine.train_model / generate_optimizer / test_model, ref main.py:54-95,97-173,176-208) with the reference's
OWN model package on the synthetic rig - so that what the GPU test proves about the drop-in package is a statement
about main.py and not about the harness.  Also: the drop-in package is importable through both documented routes."""
import os
import subprocess
import sys

import pytest

import main_harness as mh
from baseline import synthetic_data


@pytest.mark.skipif(not mh.reference_available(), reason="baseline/_ref not installed and /root/reference absent")
def test_reference_main_runs_on_the_synthetic_rig(tmp_path):
    main = mh.load_main(drop_in=False)
    sp = synthetic_data.engine_sys_param("cpu", n_cam=4, img=8, batch=16, samples=8, scale=2, steps_per_epoch=2,
                                         coarse=(2, 32, ()), fine=(3, 32, (1,)), root=str(tmp_path))
    engine, rec, before, after, ckpts = mh.run_training(main, sp)
    mh.check_training(engine, rec, before, after, ckpts, steps=2)
    sp_demo = synthetic_data.engine_sys_param("cpu", n_cam=4, img=8, batch=16, samples=8, scale=2, mode=1,
                                              coarse=(2, 32, ()), fine=(3, 32, (1,)), root=str(tmp_path))
    sp_demo["demo_ckpt"] = ckpts[-1]
    mh.run_demo(main, sp_demo)


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)


def test_drop_in_package_imports_under_the_name_model_via_symlink(tmp_path):
    """INTEGRATION.md route 1: `model` is a link to mc_nerf_b200/model inside a checkout of the reference."""
    os.symlink(os.path.join(ROOT, "mc_nerf_b200", "model"), tmp_path / "model")
    r = _run("import sys; sys.path.insert(0, '.'); import model; "
             "from model import MC_Model, MC_NeRF_Loss, RAdam, apply_depth_colormap; "
             "from model.external.pohsun_ssim import pytorch_ssim; print(MC_Model.__module__)", str(tmp_path))
    assert r.returncode == 0 and r.stdout.strip() == "model.mc_nerf", (r.stdout, r.stderr[-2000:])


def test_drop_in_package_imports_under_the_name_model_via_sys_modules(tmp_path):
    """INTEGRATION.md route 2: register the package under the name `model` before main.py is imported."""
    r = _run("import importlib, sys\n"
             "sys.modules['model'] = importlib.import_module('mc_nerf_b200.model')\n"
             "for sub in ('mc_nerf', 'net_block', 'net_utils', 'loss', 'external', 'external.pohsun_ssim'):\n"
             "    sys.modules[f'model.{sub}'] = importlib.import_module(f'mc_nerf_b200.model.{sub}')\n"
             "from model import MC_Model, MC_NeRF_Loss, RAdam, apply_depth_colormap\n"
             "from model.external.pohsun_ssim import pytorch_ssim\n"
             "print(MC_Model.__module__)", str(tmp_path))
    assert r.returncode == 0 and r.stdout.strip() == "mc_nerf_b200.model.mc_nerf", (r.stdout, r.stderr[-2000:])

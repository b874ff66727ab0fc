"""GPU: the reference's UNMODIFIED main.py drives the drop-in package.

`Model_Engine.train_model` (DDP wrap, `generate_optimizer`'s name-prefix split into three RAdam + ExponentialLR,
the step loop `zero_grad -> model(data, epoch, epoch_type, cur_ratio) -> loss -> backward -> optimizer.step`,
per-epoch `save_model / show_estimate_param / show_RT_est_results / valid_train`; ref main.py:54-95,176-208) runs
3 steps in each of the three stages, then `Model_Engine.test_model` (ref main.py:97-173) renders every test view from
the checkpoint the training wrote.  `model` is mc_nerf_b200.model, `data` is the synthetic rig
(baseline/synthetic_data.py); main.py itself is the byte-identical copy under baseline/_ref/ (git-ignored).
The 2-rank variant runs the same loop under DistributedDataParallel(find_unused_parameters=True) (ref main.py:61).
The harness is validated against the reference's own model package on CPU in tests/test_main_integration.py."""
import os
import sys

import pytest
import torch

import main_harness as mh
from baseline import synthetic_data

pytestmark = pytest.mark.gpu


def _need_ref():
    if not mh.reference_available():
        pytest.fail("baseline/_ref is missing: run `python baseline/install_ref.py` (or __graft_entry__.build()) in the "
                    "build container so that the reference's main.py travels to the GPU box")


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_reference_main_trains_and_renders_with_the_drop_in_package(tmp_path, precision):
    _need_ref()
    main = mh.load_main(drop_in=True)
    sp = synthetic_data.engine_sys_param("cuda", n_cam=6, img=16, batch=64, samples=16, scale=2, steps_per_epoch=3,
                                         root=str(tmp_path))
    sp["mlp_precision"] = precision
    engine, rec, before, after, ckpts = mh.run_training(main, sp)
    mh.check_training(engine, rec, before, after, ckpts, steps=3)
    from mc_nerf_b200.model import net_utils
    assert isinstance(engine.mc_nerf.nerf.nerf_fine, torch.nn.Module) and main.RAdam is net_utils.RAdam
    sp_demo = synthetic_data.engine_sys_param("cuda", n_cam=6, img=16, batch=64, samples=16, scale=2, mode=1,
                                              root=str(tmp_path))
    sp_demo["mlp_precision"] = precision
    sp_demo["demo_ckpt"] = ckpts[-1]
    mh.run_demo(main, sp_demo)


def _ddp_worker(rank, world, root, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)        # both ranks share cuda:0: gloo, not nccl
    try:
        main = mh.load_main(drop_in=True)
        sp = synthetic_data.engine_sys_param("cuda", n_cam=6, img=16, batch=64, samples=16, scale=2,
                                             steps_per_epoch=4, root=os.path.join(root, f"rank{rank}"))
        sp.update(distributed=True, gpu=0, rank=rank, world_size=world, mlp_precision="bf16")
        engine, rec, before, after, ckpts = mh.run_training(main, sp)
        steps = 4 // world
        stages = [s for s, _, _ in rec.log]
        assert stages == ["CAM_PARAM_EPOCH"] * steps + ["GLOBAL_OPTIM_EPOCH"] * steps + ["FINE_TUNE_EPOCH"] * steps
        # DDP averaged the gradients: every rank holds the same parameters after training
        flat = torch.cat([v.detach().reshape(-1).float().cpu() for v in after.values()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered[1:]), "ranks diverged under DDP"
        assert float((flat - torch.cat([v.reshape(-1).float().cpu() for v in before.values()])).abs().max()) > 0
        if rank == 0:
            assert len(ckpts) == 3
    finally:
        dist.destroy_process_group()


def test_reference_main_under_ddp_two_ranks(tmp_path):
    _need_ref()
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_ddp_worker, args=(2, str(tmp_path), port), nprocs=2, join=True)

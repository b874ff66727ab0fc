"""Stand-in for the reference's `data` package (ref: data/data_read.py) on the synthetic Ball-style rig.

The reference's loader reads a Blender dataset from disk and runs the AprilTag detector (cv2 + apriltag); neither the
data nor the detector exists here.  `Data_set` / `Data_loader` below have the constructor signatures, attributes and
item formats `main.py:27-46,56-57,99` uses, filled from mc_nerf_b200.synthetic instead of the disk:

  Data_set(sys_param).update_system_param        - sys_param + the keys data_read.py:267-283 appends
  Data_set[idx] (mode 0) -> (gt_rgbs[HW,3], img_id, intr_wpts[n,5,3], intr_pts[n,5,2], extr_wpts, extr_pts)
  Data_set[idx] (mode 1) -> (gt_rgbs[HW,3], img_idx)
  Data_loader(dataset, sys_param).dataloader["Shuffle_loader" | "Squence_loader"], .sampler

Test / benchmark infrastructure only (tests/test_main_integration*.py); the product never imports it.
"""
import torch
from torch.utils.data import DataLoader, DistributedSampler

from mc_nerf_b200 import synthetic as syn


class Data_set(torch.utils.data.Dataset):
    def __init__(self, sys_param):
        self.system_param = sys_param
        kw = sys_param["synthetic_rig"]             # dict(n_cam, img_h, img_w, steps_per_epoch, ...)
        self.steps = kw.get("steps_per_epoch", 3)
        base = syn.make_sys_param(n_cam=kw["n_cam"], img_h=kw["img_h"], img_w=kw["img_w"], batch=sys_param["batch"],
                                  samples=sys_param["samples"], scale=sys_param["scale"], seed=kw.get("seed", 4))
        for k in ("intr_mat", "intr_mat_inv", "data_numb", "gt_pose", "valid_pose", "test_pose", "valid_rgbs",
                  "data_img_h", "data_img_w", "train_json_file"):
            sys_param[k] = base[k]
        # data_read.py:338-351 (get_squence_info): the BARF window as fractions of the whole training run
        s1, s2, s3 = sys_param["stage1_epoch"], sys_param["stage2_epoch"], sys_param["stage3_epoch"]
        sys_param["epoch_squence"] = torch.tensor([s1, s2, s3], dtype=torch.long)
        sys_param["epoch_numb"] = total = s1 + s2 + s3
        start = float(s1) / float(total) + sys_param["barf_start"]
        end = float(s1 + s2) / float(total)
        sys_param["barf_start"], sys_param["barf_end"] = start, start + (end - start) * sys_param["barf_end"]
        self.update_system_param = sys_param
        n = kw["n_cam"]
        H, W = kw["img_h"], kw["img_w"]
        g = torch.Generator().manual_seed(kw.get("seed", 4) + 17)
        self.rgbs = torch.rand(n, H * W, 3, generator=g)
        self.intr_wpts = torch.rand(n, 5, 3, generator=g) - 0.5
        self.intr_pts = torch.rand(n, 5, 2, generator=g) * W
        self.extr_wpts = torch.rand(n, 5, 3, generator=g) - 0.5
        self.extr_pts = torch.rand(n, 5, 2, generator=g) * W
        self.n_cam = n

    def __len__(self):
        return self.steps if self.system_param["mode"] == 0 else self.n_cam

    def __getitem__(self, idx):
        cam = idx % self.n_cam
        if self.system_param["mode"] == 0:
            return (self.rgbs[cam], torch.tensor(cam, dtype=torch.long), self.intr_wpts, self.intr_pts,
                    self.extr_wpts, self.extr_pts)
        return self.rgbs[cam], torch.tensor(cam, dtype=torch.long)


class Data_loader:
    def __init__(self, dataset, sys_param):
        self.dataset, self.sys_param = dataset, sys_param
        if sys_param["distributed"]:
            self.sampler = DistributedSampler(dataset, shuffle=True)
            self.sampler_no_shuffle = DistributedSampler(dataset, shuffle=False)
        else:
            self.sampler = torch.utils.data.RandomSampler(dataset)
            self.sampler_no_shuffle = torch.utils.data.SequentialSampler(dataset)
        train = torch.utils.data.BatchSampler(self.sampler, 1, drop_last=True)
        val = torch.utils.data.BatchSampler(self.sampler_no_shuffle, 1, drop_last=False)
        pin = str(sys_param["device_type"]).startswith("cuda")
        self.dataloader = {"Shuffle_loader": DataLoader(dataset, batch_sampler=train, num_workers=0, pin_memory=pin),
                           "Squence_loader": DataLoader(dataset, batch_sampler=val, num_workers=0, pin_memory=pin)}


def engine_sys_param(device, n_cam=6, img=16, batch=64, samples=16, scale=2, steps_per_epoch=3, mode=0,
                     coarse=(8, 256, (4,)), fine=(8, 256, (4,)), root="."):
    """The flat dict config_read.py:21-74 builds from config.yaml + argparse, for the synthetic rig."""
    import os
    return dict(
        mode=mode, device_type=device, distributed=False, start_device=0, seed=42, batch=batch,
        stage1_epoch=1, stage2_epoch=1, stage3_epoch=1, stage1_lr=0.1, stage2_lr=5e-4, stage3_lr=2.5e-4,
        weight_d=4e-4, warmup_epoch=100, res_h=img, res_w=img, demo_ckpt="",
        root_weight=os.path.join(root, "weights"), root_out=os.path.join(root, "results"),
        demo_render_pth=os.path.join(root, "results", "img_rendered"), log_pth=os.path.join(root, "log"),
        tb_available=False, tb_pth="./tensorboard", tb_del=False, tag_size=1.0,
        data_name="synthetic", root_data="", barf_mask=False, barf_start=0.0, barf_end=1.0,
        near=1.0, far=8.0, samples=samples, scale=scale, grid_nerf=384, sigma_init=30.0, sigma_default=-20.0,
        sample_weight_thresh=1e-3, boader_min=-3.5, boader_max=3.5, white_back=True, emb_freqs_xyz=10,
        coarse_MLP_depth=coarse[0], coarse_MLP_width=coarse[1], coarse_MLP_skip=list(coarse[2]),
        fine_MLP_depth=fine[0], fine_MLP_width=fine[1], fine_MLP_skip=list(fine[2]), MLP_deg=2,
        synthetic_rig=dict(n_cam=n_cam, img_h=img, img_w=img, steps_per_epoch=steps_per_epoch),
    )

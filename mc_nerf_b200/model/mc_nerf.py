"""MC_Model and NeRF_Model with the reference's public surface (ref: model/mc_nerf.py), driven by
libmcnerf.so.  main.py --train / --demo and config.yaml use these classes unchanged:

  MC_Model(sys_param)(data, epoch, epoch_type, cur_ratio) -> (loss_dict, intr_show, pose_show, rays_valid)
  MC_Model(sys_param)(img_idx)                            -> (rgbs, depth, opacity) on the CPU   (demo mode)
  NeRF_Model.forward / render_rays_train / render_rays_test / inference / sigma2weights
  .nerf.save_model / .nerf.valid_train / .show_estimate_param / .show_RT_est_results / .opt_idx / .weights_pose

Differences that are deliberate: the train step generates only the `batch` selected rays
from (camera, pixel) instead of two full images of rays, and the per-step validation rays are produced
lazily when valid_train actually reads them.  Outputs are identical.
"""
import logging
import os
import time
from pathlib import Path

import torch
import torch.distributed as dist
import torch.nn as nn

from mc_nerf_b200 import ops, render
from .net_block import CorseFine_NeRF, SinCosEmbedding
from .net_utils import get_rank


class _LazyValidRays:
    """[rays_d, rays_o, rgbs] of the validation view, materialised on first access
    (ref: model/mc_nerf.py:97-99 builds all H*W rays every step; main.py reads them once per epoch)."""

    def __init__(self, model, img_id):
        self._model, self._img_id, self._val = model, img_id, None

    def _get(self):
        if self._val is None:
            m = self._model
            with torch.no_grad():
                rd, ro = m.get_rays(m.valid_pose, self._img_id, m.intr_val_inv.to(m.device))
            self._val = [rd.detach(), ro.detach(), m.valid_rgbs[self._img_id].detach()]
        return self._val

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return 3


class MC_Model(nn.Module):
    def __init__(self, sys_param):
        logging.info("Creating MC-NeRF Model...")
        super().__init__()
        self.sys_param = sys_param
        self.mode = sys_param["mode"]
        self.device = sys_param["device_type"]
        self.batch = sys_param["batch"]
        self.bound_min = sys_param["boader_min"]
        self.bound_max = sys_param["boader_max"]
        self.intr = sys_param["intr_mat"]
        self.intr_inv = sys_param["intr_mat_inv"]
        self.intr_train, self.intr_test, self.intr_val = self.intr
        self.intr_train_inv, self.intr_test_inv, self.intr_val_inv = self.intr_inv
        self.gt_pose = sys_param["gt_pose"].to(self.device)
        self.test_pose = sys_param["test_pose"].to(self.device)
        self.valid_pose = sys_param["valid_pose"].to(self.device)
        self.valid_rgbs = sys_param["valid_rgbs"].to(self.device)
        self.img_h = sys_param["data_img_h"]
        self.img_w = sys_param["data_img_w"]
        self.train_img_pth = sys_param["demo_render_pth"]
        self.data_name = sys_param["data_name"]
        self.data_numb = sys_param["data_numb"]
        self.train_numb, self.test_numb, self.val_numb = self.data_numb
        self.train_json_pth = sys_param["train_json_file"]
        self.register_parameters()
        self.nerf = NeRF_Model(sys_param).to(self.device)
        self.count_rays = 0
        self.opt_idx = 0
        self.last_epoch_type = 0
        self.wait_reset = 0
        self._intr_inv_adj = None
        self.init_show_figure(show_info=False)

    # ------------------------------------------------------------------ forward
    def forward(self, *args):
        if self.sys_param["mode"] == 0:
            return self._forward_train(*args)
        return self._forward_demo(*args)

    def _forward_train(self, *args):
        loss_dict = {}
        gt_rgbs, img_id, intr_wpts, intr_pts, extr_wpts, extr_pts, epoch, epoch_type, cur_ratio = \
            self.data2device(*args)
        img_id_host = args[0][1]            # camera id as the loader produced it (CPU tensor in main.py): no sync
        if epoch_type == "CAM_PARAM_EPOCH":          # ref: model/mc_nerf.py:64-71
            self.nerf.emmbedding_xyz.barf_mode = False
            self.intr_adj, self.pose_adj, self.calib_pose_adj = self.add_weights2param(True, True, True)
            loss_dict["intr"] = [self.get_reproject_pixels(intr_wpts, self.intr_adj, self.calib_pose_adj), intr_pts]
            loss_dict["extr"] = [self.get_reproject_pixels(extr_wpts, self.intr_adj, self.pose_adj), extr_pts]
            self.opt_idx = 0
        else:
            glob = epoch_type == "GLOBAL_OPTIM_EPOCH"    # ref: :73-83 (global) / :85-95 (fine tune)
            self.nerf.prefetch_weights()                 # bf16 weight images pack while the camera kernels run
            self.nerf.emmbedding_xyz.barf_mode = glob
            if intr_wpts.is_cuda and intr_wpts.dim() == 4 and intr_wpts.shape[0] == 1 and not intr_wpts.requires_grad:
                # the whole camera model in one launch (and one in backward): add_weights2param + get_reproject_pixels
                ws = [self.weights_fx, self.weights_fy, self.weights_ux, self.weights_uy]
                for w in ws:
                    w.requires_grad_(True)
                self.weights_pose.requires_grad_(glob)
                self.weights_pose_intr.requires_grad_(True)
                K, Kinv, self.pose_adj, self.calib_pose_adj, reproj = ops.CameraTrainFn.apply(
                    *ws, self.weights_pose, self.weights_pose_intr, intr_wpts, self.img_h, self.img_w)
                self.intr_adj, self._intr_inv_adj = K, (K, Kinv)
            else:
                self.intr_adj, self.pose_adj, self.calib_pose_adj = self.add_weights2param(True, glob, True)
                reproj = self.get_reproject_pixels(intr_wpts, self.intr_adj, self.calib_pose_adj)
            rays_d, rays_o, rand_idx = self.generate_train_rays(img_id_host)
            rgbs_c, rgbs_f = self.nerf(rays_d, rays_o, epoch, cur_ratio if glob else 1)
            if self._gt_event is not None:          # side-stream H2D of the image must have landed
                torch.cuda.current_stream().wait_event(self._gt_event)
                gt_rgbs.record_stream(torch.cuda.current_stream())
            gt_sel = gt_rgbs.reshape(-1, 3)[rand_idx]
            loss_dict["intr"] = [reproj, intr_pts]
            loss_dict["rgb"] = [rgbs_c, rgbs_f, gt_sel]
            self.opt_idx = 1 if glob else 2
        rays_valid = _LazyValidRays(self, img_id_host)
        if self.__dict__.get("_intr_train_dev") is None:
            self.__dict__["_intr_train_dev"] = self.intr_train.to(self.device).detach()
        intr_show = [self._intr_train_dev, self.intr_adj.detach()]
        pose_show = [self.gt_pose.detach(), self.pose_adj.detach()]
        self.last_epoch_type = epoch_type
        return loss_dict, intr_show, pose_show, rays_valid

    def _forward_demo(self, *args):
        """ref: model/mc_nerf.py:106-122 (chunk loop with three blocking `.cpu()` per chunk).  Here every chunk's three
        outputs are copied to PINNED host tensors on a side stream while the next chunk renders; one wait at the end.
        Fresh host tensors per call (main.py:119-120 keeps every view's result until all views are rendered)."""
        img_id = args[0] if len(args) == 1 else args
        rays_d, rays_o = self.get_rays(self.test_pose, img_id, self.intr_test_inv.to(self.device))
        n = rays_d.shape[0]
        if not rays_d.is_cuda:
            raise ops._lib.McnerfError("MC_Model demo mode needs a CUDA device (no CPU fallback)")
        host = [torch.empty((n, c), dtype=torch.float32, pin_memory=True) for c in (3, 1, 1)]
        if self.__dict__.get("_d2h_stream") is None:
            self.__dict__["_d2h_stream"] = torch.cuda.Stream(device=rays_d.device)
        side, main = self._d2h_stream, torch.cuda.current_stream(rays_d.device)
        for ii in range(0, n, self.batch):
            outs = self.nerf(rays_d[ii:ii + self.batch], rays_o[ii:ii + self.batch])
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for h, o in zip(host, outs):
                    o = o.detach()
                    h[ii:ii + o.shape[0]].copy_(o, non_blocking=True)
                    o.record_stream(side)
        side.synchronize()
        return host[0], host[1], host[2]

    # ------------------------------------------------------------------ rays
    def _cam_index(self, img_id, n_rays=None):
        """Camera selector for the ray kernels: a Python int when the id lives on the host, otherwise a per-ray
        int32 device tensor - never a device->host synchronisation (the reference indexes `pose[img_id]` on device)."""
        if isinstance(img_id, (tuple, list)):
            return self._cam_index(img_id[0], n_rays)
        if torch.is_tensor(img_id):
            if not img_id.is_cuda:
                return int(img_id.reshape(-1)[0])
            return img_id.reshape(-1)[:1].to(torch.int32).expand(n_rays).contiguous()
        return int(img_id)

    def get_rays(self, pose, img_id, intr_inv):
        """All H*W rays of camera img_id, row-major pixels with +0.5 centres.  ref: model/mc_nerf.py:124-145."""
        n = self.img_h * self.img_w
        cam = self._cam_index(img_id, n)
        rays_o, rays_d = ops.RaygenFn.apply(intr_inv.to(self.device), pose, cam, None, n, self.img_w)
        return rays_d, rays_o

    def generate_train_rays(self, img_id):
        """randperm(H*W)[:batch] exactly as generate_rand_rays draws it (ref: :327-345), but only the selected
        rays are generated, straight from (camera, pixel)."""
        n = self.img_h * self.img_w
        rand_idx, rand_idx32 = self._choose_pixels(n)
        cam = self._cam_index(img_id, rand_idx.shape[0])
        rays_o, rays_d = ops.RaygenFn.apply(self.inverse_intrinsic(self.intr_adj), self.pose_adj, cam,
                                            rand_idx32, rand_idx.shape[0], self.img_w)
        self.count_rays += 1
        return rays_d, rays_o, rand_idx

    def _choose_pixels(self, n):
        """-> (int64, int32) pixel indices = randperm(n)[:batch].  sys_param["pixel_sampler"]:
        "device" (default): libmcnerf's threshold-and-sort sampler (same distribution, 2 launches instead of the
        8-pass radix sort of all n keys; seeded from torch's CUDA generator, so torch.manual_seed governs it and it is
        CUDA-graph safe);  "randperm": torch.randperm itself, draw for draw as the reference (replay tests)."""
        if self.sys_param.get("pixel_sampler", "device") == "device" and torch.device(self.device).type == "cuda":
            ws = self.__dict__.get("_pixel_ws")
            if ws is None or ws[0] != (n, self.batch):
                ws = self.__dict__["_pixel_ws"] = ((n, self.batch), ops.sample_pixels_workspace(n, self.batch, self.device))
            if ws[1] is not None:
                seed = ops.draw_seed(self.device)
                # the same two key words serve the renderer's device-side draws of this step (other Philox streams)
                self.nerf.__dict__["_step_seed"] = seed
                return ops.sample_pixels(n, self.batch, seed, ws[1])
        rand_idx = torch.randperm(n, device=self.device)[:self.batch]
        return rand_idx, rand_idx.to(torch.int32)

    def generate_rand_rays(self, rays_d, rays_o, rand=True):
        """API-compatible subset selection on pre-generated rays.  ref: model/mc_nerf.py:327-345."""
        n = rays_d.shape[0]
        if rand:
            rand_idx = torch.randperm(n, device=self.device)[:self.batch]
        else:
            starts = list(range(0, n, self.batch))
            if self.count_rays == len(starts):
                self.count_rays = 0
            s = starts[self.count_rays]
            rand_idx = torch.arange(s, min(n, s + self.batch), device=self.device)
        self.count_rays += 1
        return rays_d[rand_idx], rays_o[rand_idx], rand_idx

    def get_reproject_pixels(self, tag_wpts, intr_adj, pose_adj):
        """Calibration-point reprojection px = K [R|t] X / z.  ref: model/mc_nerf.py:147-152."""
        if tag_wpts.is_cuda and tag_wpts.dim() == 4 and tag_wpts.shape[0] == 1 and not tag_wpts.requires_grad:
            return ops.ReprojectFn.apply(tag_wpts, intr_adj, pose_adj)
        cam = self.world2cam(self.world2hom(tag_wpts), pose_adj.unsqueeze(0))
        return self.cam2pix(cam, intr_adj.unsqueeze(0))

    # ------------------------------------------------------------------ learnable camera parameters
    def add_weights2param(self, intr=True, extr=True, calib_extr=False):
        return (self.add_weights2intr(self.img_h, self.img_w, adj=intr), self.add_weights2pose(adj=extr),
                self.add_weights2calib_pose(adj=calib_extr))

    def add_weights2intr(self, img_h, img_w, adj=True):
        ws = [self.weights_fx, self.weights_fy, self.weights_ux, self.weights_uy]
        for w in ws:
            w.requires_grad_(adj)
        K, Kinv = ops.IntrinsicsFn.apply(*ws, img_h, img_w)
        self._intr_inv_adj = (K, Kinv)
        return K

    def add_weights2pose(self, adj=True):
        return self.se3_to_SE3(self.weights_pose.requires_grad_(adj))

    def add_weights2calib_pose(self, adj=True):
        return self.se3_to_SE3(self.weights_pose_intr.requires_grad_(adj))

    def inverse_intrinsic(self, intr_mats):
        """ref: model/mc_nerf.py:204-210 (a Python loop of 110 torch.inverse calls); closed form in the
        intrinsics kernel when asked for the matrix add_weights2intr just produced."""
        if self._intr_inv_adj is not None and intr_mats is self._intr_inv_adj[0]:
            return self._intr_inv_adj[1]
        return torch.linalg.inv(intr_mats)

    def se3_to_SE3(self, wu):
        return ops.SE3Fn.apply(wu)

    # small homogeneous-coordinate helpers kept for API compatibility (ref: model/mc_nerf.py:213-267)
    @staticmethod
    def _hom(x):
        return torch.cat([x, torch.ones_like(x[..., :1])], dim=-1)

    def pix2hom(self, pixel_cord):
        return self._hom(pixel_cord)

    def cam2hom(self, cam_cord):
        return self._hom(cam_cord)

    def world2hom(self, world_cord):
        return self._hom(world_cord)

    def pix2cam(self, pix_cord, intr_inv_mat):
        return pix_cord @ intr_inv_mat.transpose(-2, -1)

    def cam2pix(self, cam_cord, intr_mat):
        pix = torch.cat([intr_mat, torch.zeros_like(intr_mat[..., :1])], dim=-1) @ cam_cord
        return (pix[..., :2, :] / pix[..., 2:, :]).transpose(-2, -1)

    def cam2world(self, cam_cord, pose):
        R_inv = pose[..., :3].transpose(-2, -1)
        pose_inv = torch.cat([R_inv, -R_inv @ pose[..., 3:]], -1)
        return cam_cord @ pose_inv.transpose(-2, -1)

    def world2cam(self, world_cord, pose):
        bottom = torch.zeros_like(pose[..., :1, :])
        bottom[..., 0, 3] = 1
        return torch.cat([pose, bottom], dim=-2) @ world_cord.transpose(-2, -1)

    def skew_symmetric(self, w):
        w0, w1, w2 = w.unbind(dim=-1)
        O = torch.zeros_like(w0)
        return torch.stack([torch.stack([O, -w2, w1], -1), torch.stack([w2, O, -w0], -1),
                            torch.stack([-w1, w0, O], -1)], -2)

    @staticmethod
    def _taylor(x, first_factor, nth):
        ans, denom = torch.zeros_like(x), 1.0
        for i in range(nth + 1):
            if i > 0 or first_factor > 0:
                denom *= (2 * i + first_factor) * (2 * i + first_factor + 1)
            ans = ans + (-1) ** i * x ** (2 * i) / denom
        return ans

    def taylor_A(self, x, nth=10):
        return self._taylor(x, 0, nth)     # sin(x)/x

    def taylor_B(self, x, nth=10):
        return self._taylor(x, 1, nth)     # (1-cos x)/x^2

    def taylor_C(self, x, nth=10):
        return self._taylor(x, 2, nth)     # (x-sin x)/x^3

    def compose_param2pose(self, param, pose):
        R_a, t_a, R_b, t_b = param[..., :3], param[..., 3:], pose[..., :3], pose[..., 3:]
        return torch.cat([R_b @ R_a, R_b @ t_a + t_b], -1)

    def register_parameters(self):
        """All-ones initial state, exactly the reference's (ref: model/mc_nerf.py:347-371)."""
        n = self.train_numb
        for name, shape in (("weights_pose", [n, 6]), ("weights_pose_intr", [n, 6]), ("weights_ux", [n]),
                            ("weights_uy", [n]), ("weights_fx", [n]), ("weights_fy", [n])):
            self.register_parameter(name, nn.Parameter(torch.ones(shape, device=self.device), requires_grad=True))

    def data2device(self, *args):
        dev = torch.device(self.device)

        def mv(t):
            return t if t.device == dev or (t.is_cuda and dev.type == "cuda" and dev.index is None) \
                else t.to(dev, non_blocking=True)
        gt_rgbs, img_id, intr_wpts, intr_pts, extr_wpts, extr_pts = args[0]
        # The ground-truth image (H*W*3 fp32: 7.7 MB at 800x800) is only needed for the loss at the END of the forward
        # pass: copy it on a side stream so the transfer overlaps ray generation and the coarse/fine MLPs.
        self._gt_event = None
        if dev.type == "cuda" and not gt_rgbs.is_cuda and gt_rgbs.is_pinned():
            if self.__dict__.get("_h2d_stream") is None:
                self.__dict__["_h2d_stream"] = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(self._h2d_stream):
                gt_dev = gt_rgbs.to(dev, non_blocking=True)
                self._gt_event = torch.cuda.Event()
                self._gt_event.record(self._h2d_stream)
        else:
            gt_dev = mv(gt_rgbs)
        return (gt_dev, mv(img_id), mv(intr_wpts), mv(intr_pts), mv(extr_wpts), mv(extr_pts),
                args[1], args[2], args[3])

    # ------------------------------------------------------------------ epoch-end reporting (not hot path)
    def show_estimate_param(self, intr_show, pose_show, epoch, epoch_type):
        """Mean |K - K_gt| and |RT - RT_gt| per epoch (ref: model/mc_nerf.py:388-407), logged as one line."""
        il = (intr_show[0] - intr_show[1]).abs()
        pl = (pose_show[0] - pose_show[1]).abs()
        row = dict(EPOCH=epoch, LOSS_FX=il[:, 0, 0].mean().item(), LOSS_FY=il[:, 1, 1].mean().item(),
                   LOSS_UX=il[:, 0, 2].mean().item(), LOSS_UY=il[:, 1, 2].mean().item(), LOSS_K=il.mean().item(),
                   LOSS_R=pl[..., :3].mean().item(), LOSS_T=pl[..., 3:].mean().item())
        logging.info("camera parameter error: " + " ".join(f"{k}={v:.5g}" for k, v in row.items()))
        return row

    def init_show_figure(self, show_info=True):
        """The reference opens a matplotlib 3-D figure (ref: :409-432); plotting is optional here."""
        self.fig = None

    def show_RT_est_results(self, epoch, epoch_type, mode="train"):
        """Camera-frustum plot of ground-truth vs estimated poses (ref: :448-497).  Drawn only when matplotlib
        is importable; a no-op otherwise (visualisation, outside the hot path)."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return None
        with torch.no_grad():
            est = self.se3_to_SE3(self.weights_pose.detach()).cpu()
            gt = self.gt_pose.cpu()
        fig = plt.figure(figsize=(6, 6))
        ax = fig.add_subplot(projection="3d")
        for P, c in ((gt, "g"), (est, "r")):
            centres = -(P[:, :, :3].transpose(1, 2) @ P[:, :, 3:]).squeeze(-1)
            ax.scatter(centres[:, 0], centres[:, 1], centres[:, 2], c=c, s=6)
        out = os.path.join(str(self.train_img_pth), str(self.data_name))
        os.makedirs(out, exist_ok=True)
        fig.savefig(os.path.join(out, f"poses_epoch_{epoch}.png"))
        plt.close(fig)


class NeRF_Model(nn.Module):
    def __init__(self, sys_param):
        logging.info("Creating NeRF Model...")
        super().__init__()
        self.sys_param = sys_param
        self.mode = sys_param["mode"]
        self.device = sys_param["device_type"]
        self.near, self.far = sys_param["near"], sys_param["far"]
        self.samples_c = sys_param["samples"]
        self.sample_scale = sys_param["scale"]
        self.samples_f = self.samples_c * self.sample_scale
        self.dim_sh = 3 * (sys_param["MLP_deg"] + 1) ** 2
        self.white_back = sys_param["white_back"]
        self.weights_pth = sys_param["root_weight"]
        self.train_img_pth = sys_param["demo_render_pth"]
        self.batch_test = sys_param["batch"]
        self.xyz_min, self.xyz_max = sys_param["boader_min"], sys_param["boader_max"]
        self.xyz_scope = self.xyz_max - self.xyz_min
        self.grid_nerf = sys_param["grid_nerf"]
        self.sigma_init = sys_param["sigma_init"]
        self.sigma_default = sys_param["sigma_default"]
        self.warmup_epoch = sys_param["warmup_epoch"]
        self.weight_thresh = sys_param["sample_weight_thresh"]
        self.render_h, self.render_w = sys_param["res_h"], sys_param["res_w"]
        self.z_vals_c = torch.linspace(self.near, self.far, self.samples_c, device=self.device)
        self.z_vals_f = torch.linspace(self.near, self.far, self.samples_f, device=self.device)
        self.global_step = 0
        self.emmbedding_xyz = SinCosEmbedding(sys_param)        # (sic) the reference's attribute name
        self.nerf_coarse = CorseFine_NeRF(sys_param, type="coarse")
        self.nerf_fine = CorseFine_NeRF(sys_param, type="fine")
        self.data_name = sys_param["data_name"]
        self.render_cfg = render.RenderCfg.from_sys_param(sys_param)
        if self.mode != 0:
            self.nerf_ckpt_name = sys_param["demo_ckpt"]
            ckpt = torch.load(Path(self.nerf_ckpt_name), map_location=self.device)
            self.nerf_coarse.load_state_dict(self.rewrite_nerf_ckpt(ckpt, coarse=True))
            self.nerf_fine.load_state_dict(self.rewrite_nerf_ckpt(ckpt))
            logging.info("Loading weights:{}".format(self.nerf_ckpt_name))

    def forward(self, *args):
        self.global_step += 1
        if self.mode == 0:
            rays_d, rays_o, cur_epoch, step_r = args
            return self.render_rays_train(rays_d, rays_o, cur_epoch, step_r, only_coarse=False)
        rays_d, rays_o = args
        return self.render_rays_test(rays_d, rays_o, self.nerf_coarse, self.nerf_fine)

    def use_device_band_weights(self, enable=True):
        """Keep the BARF window weights in a device buffer that the kernels read at run time instead of baking them
        into every launch - what a captured CUDA graph needs to follow `cur_ratio` (graph.GraphedTrainStep).  The
        caller then calls set_band_weights(step_r) before each training render."""
        if enable:
            self.__dict__["_band_w_dev"] = torch.ones(16, dtype=torch.float32, device=self.device)
        else:
            self.__dict__["_band_w_dev"] = None

    def set_band_weights(self, step_r):
        """fill the device buffer with the window weights of `step_r` (independent of the current barf_mode flag,
        which MC_Model.forward only sets when it runs)"""
        if self.__dict__.get("_band_w_dev") is None:
            return
        emb = self.emmbedding_xyz
        w = ops.barf_band_weights(float(step_r), emb.barf_start, emb.barf_end, emb.n_freqs)
        ops.store_floats(self._band_w_dev, w)          # values travel in the launch itself: stream-ordered, no staging

    def prefetch_weights(self):
        """start deriving the tensor-core weight images on a side stream (see render.prefetch_weights)"""
        if torch.device(self.device).type == "cuda":
            render.prefetch_weights(self.render_cfg, self.nerf_coarse.param_dict(), self.nerf_fine.param_dict(),
                                    need_bwd=torch.is_grad_enabled())

    # ------------------------------------------------------------------ fused render paths
    def render_rays_train(self, rays_d, rays_o, cur_epoch, step_r, only_coarse=False, rng=None, cap_perm=None):
        """ref: model/mc_nerf.py:598-646.  `rng` / `cap_perm` let tests inject the random draws."""
        if only_coarse:
            return self._coarse_only(rays_d, rays_o, step_r)
        band_w = self.emmbedding_xyz.band_weights(step_r)
        if band_w is not None and self.__dict__.get("_band_w_dev") is not None:
            # device-side weights (CUDA-graph replays follow a moving window).  A captured graph reads whatever
            # set_band_weights stored before the replay; an EAGER call refreshes the buffer from its own step_r, so
            # mixing graphed and eager steps on one model can never apply a stale window.
            if not torch.cuda.is_current_stream_capturing():
                self.set_band_weights(step_r)
            band_w = self._band_w_dev
        seed = self.__dict__.pop("_step_seed", None)
        if rng is None:
            rng = render.draw_rng(self.render_cfg, rays_d.shape[0], rays_d.device, True, seed=seed)
        rgb_c, rgb_f, _, _ = render.render(self.render_cfg, self.nerf_coarse.param_dict(), self.nerf_fine.param_dict(),
                                           rays_d, rays_o, True, band_w, rng, cap_perm)
        return rgb_c, rgb_f

    def render_rays_test(self, rays_d, rays_o, model_coarse, model_fine, rng=None):
        """ref: model/mc_nerf.py:648-680 (no jitter, step_r = 1, no 128-per-ray cap)."""
        band_w = self.emmbedding_xyz.band_weights(1)
        cfg = self.render_cfg
        if (model_coarse.cfg(), model_fine.cfg()) != (cfg.coarse, cfg.fine):
            cfg = render.RenderCfg(cfg.near, cfg.far, cfg.Sc, cfg.scale, cfg.n_freqs, cfg.white_back, cfg.sigma_default,
                                   cfg.thresh, model_coarse.cfg(), model_fine.cfg(), cfg.precision, cfg.device_rng)
            cfg.sh_dim = self.render_cfg.sh_dim
        _, rgb_f, depth_f, opa_f = render.render(cfg, model_coarse.param_dict(), model_fine.param_dict(),
                                                 rays_d, rays_o, False, band_w, rng, None)
        return rgb_f, depth_f, opa_f

    def _coarse_only(self, rays_d, rays_o, step_r):
        B = rays_d.shape[0]
        z = self.z_vals_c.clone().expand(B, -1) + torch.empty(B, 1, device=self.device).uniform_(
            0.0, (self.far - self.near) / self.samples_c)
        xyz = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(2)
        rgb, _, _, depth, _ = self.inference(self.nerf_coarse, self.emmbedding_xyz, step_r, xyz, rays_d, z)
        return rgb, None, depth

    # ------------------------------------------------------------------ module-level path (API parity)
    def inference(self, model, embedding_xyz, step_r, xyz, rays_d, z_vals, idx_render=None, coarse=True):
        """ref: model/mc_nerf.py:682-727.  Same arguments and return tuple (rgb, sigmas, xyz, depth, opacity);
        built from the individual kernels so that arbitrary callers (and autograd) keep working."""
        S = self.samples_c if coarse else self.samples_f
        B = rays_d.shape[0]
        view = rays_d.unsqueeze(1).expand(-1, S, -1)
        if idx_render is not None:
            view_s = view[idx_render[:, 0], idx_render[:, 1]]
            xyz = xyz[idx_render[:, 0], idx_render[:, 1]]
        else:
            xyz = xyz.reshape(-1, 3)
            view_s = view.reshape(-1, 3)
        out = model(embedding_xyz(xyz, step_r), view_s)
        if idx_render is not None:
            flat = (idx_render[:, 0] * S + idx_render[:, 1]).to(torch.int32)
            out = ops.ScatterFineFn.apply(out, flat, B * S, self.sigma_default)
        out = out.reshape(B, S, 4)
        noise = torch.randn((B, S), device=self.device)
        rgb, depth, opacity = ops.CompositeFn.apply(out, noise, rays_d, z_vals, None, self.near, self.far,
                                                    self.white_back)
        return rgb, out[..., 0], xyz, depth, opacity

    def sigma2weights(self, deltas, sigmas):
        """ref: model/mc_nerf.py:729-736 (draws its own N(0,1) density noise)."""
        noise = torch.randn(sigmas.shape, device=self.device)
        return ops.sigma2weights(sigmas.detach().contiguous(), noise, deltas=deltas.contiguous())

    # ------------------------------------------------------------------ checkpoints / validation (not hot path)
    def save_model(self, model, epoch):
        """{'model_nerf': MC_Model.state_dict()} -> ./weights/train/<data>-EPOCH-<e>-<time>.ckpt, rank 0.
        ref: model/mc_nerf.py:738-752 (format kept so checkpoints interchange with the reference)."""
        save_path = os.path.join(Path(self.weights_pth), Path("train"))
        stamp = time.strftime("%Y-%m-%d-%H-%M-%S.ckpt", time.localtime())
        self.model_name = "{}-EPOCH-{}-".format(self.data_name, epoch) + stamp
        self.file_path = os.path.join(Path(save_path), Path(self.model_name))
        os.makedirs(save_path, exist_ok=True)
        if not self.sys_param["distributed"] or dist.get_rank() == 0:
            torch.save({"model_nerf": model.state_dict()}, self.file_path)
        logging.info("\nSave model:{}".format(self.model_name))

    def rewrite_nerf_ckpt(self, nerf_ckpt_dict, coarse=False):
        """Strip everything up to and including 'nerf_coarse' / 'nerf_fine' from the keys.  ref: :815-837."""
        tag = "nerf_coarse" if coarse else "nerf_fine"
        out = {}
        for key, val in nerf_ckpt_dict["model_nerf"].items():
            parts = key.split(".")
            if tag in parts:
                out[".".join(parts[parts.index(tag) + 1:])] = val
        return out

    def valid_train(self, epoch, val_data, epoch_type):
        """Render one validation image from the checkpoint just written and log PSNR / SSIM (/ LPIPS when the
        package is installed).  ref: model/mc_nerf.py:754-813."""
        if epoch_type in ["CAM_PARAM_EPOCH"]:
            return 0
        if get_rank() == 0:
            rays_d, rays_o, gt_rgbs = val_data
            ckpt = torch.load(self.file_path, map_location=self.device)
            val_c = CorseFine_NeRF(self.sys_param, type="coarse").to(self.device)
            val_f = CorseFine_NeRF(self.sys_param, type="fine").to(self.device)
            val_c.load_state_dict(self.rewrite_nerf_ckpt(ckpt, coarse=True))
            val_f.load_state_dict(self.rewrite_nerf_ckpt(ckpt, coarse=False))
            rgbs, deps = [], []
            with torch.no_grad():
                for ii in range(0, rays_d.shape[0], self.batch_test):
                    r, d, _ = self.render_rays_test(rays_d[ii:ii + self.batch_test], rays_o[ii:ii + self.batch_test],
                                                    val_c, val_f)
                    rgbs.append(r)
                    deps.append(d)
            img = torch.cat(rgbs, 0).view(self.render_h, self.render_w, 3).cpu().permute(2, 0, 1)
            dep = torch.cat(deps, 0).view(self.render_h, self.render_w, 1).cpu().permute(2, 0, 1)
            gt = gt_rgbs.view(self.render_h, self.render_w, 3).cpu().permute(2, 0, 1)
            out_dir = os.path.join(Path(self.train_img_pth), Path(self.data_name))
            os.makedirs(out_dir, exist_ok=True)
            try:
                from torchvision import transforms
                transforms.ToPILImage()(img).convert("RGB").save(os.path.join(out_dir, f"epoch_{epoch}.png"))
                transforms.ToPILImage()(gt).convert("RGB").save(os.path.join(out_dir, f"epoch_{epoch}_gt.png"))
                transforms.ToPILImage()(dep).convert("L").save(os.path.join(out_dir, f"epoch_{epoch}_depth.png"))
            except Exception as e:  # image writing is best-effort
                logging.info("validation images not written: {}".format(e))
            logging.info("PSNR:{}".format(self.psnr_score(img, gt)))
            logging.info("SSIM:{}".format(self.ssim_score(img, gt)))
            try:
                logging.info("LPIPS:{}".format(self.lpips_score(img, gt)))
            except ImportError:
                logging.info("LPIPS: lpips package not installed")
        if self.sys_param["distributed"]:
            dist.barrier()

    def psnr_score(self, image_pred, image_gt, valid_mask=None, reduction="mean"):
        err = (image_pred - image_gt) ** 2
        if valid_mask is not None:
            err = err[valid_mask]
        return -10 * torch.log10(err.mean() if reduction == "mean" else err)

    def lpips_score(self, image_pred, image_gt):
        import lpips
        return lpips.LPIPS(net="alex")(image_pred * 2 - 1, image_gt * 2 - 1).item()

    def ssim_score(self, image_pred, image_gt):
        from .external.pohsun_ssim import pytorch_ssim
        return pytorch_ssim.ssim(image_pred.unsqueeze(0), image_gt.unsqueeze(0)).item()

"""MC_NeRF_Loss with the reference's interface (ref: model/loss.py).  Tiny tensors ([B,3] renders and
110x5 reprojected calibration points).  The rendering stages on the GPU use one fused kernel (loss value + its
gradients, libmcnerf `mcnerf_train_loss`); the camera-only stage and CPU tensors keep the plain torch expression."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MC_NeRF_Loss(nn.Module):
    def __init__(self, sys_param, tblogger=None):
        super().__init__()
        self.sys_param = sys_param
        self.tblogger = tblogger
        self.global_step = 0
        self.img_h = sys_param["data_img_h"]
        self.img_w = sys_param["data_img_w"]

    def forward(self, loss_dict, epoch_type):
        self.global_step += 1
        if ("rgb" in loss_dict and "extr" not in loss_dict and epoch_type != "CAM_PARAM_EPOCH"
                and loss_dict["rgb"][0].is_cuda and loss_dict["rgb"][0].dtype == torch.float32):
            # rendering stages on the GPU: the whole expression below, forward and backward, is one kernel
            from mc_nerf_b200 import ops
            rgb_c, rgb_f, gt = loss_dict["rgb"]
            px, px_gt = loss_dict["intr"] if "intr" in loss_dict else (None, None)
            return ops.TrainLossFn.apply(rgb_c, rgb_f, gt, px, px_gt, self.img_w, self.img_h, True)
        total = 0.0
        if "intr" in loss_dict:
            l_intr = self.get_reproject_loss(loss_dict["intr"])
            # stage 1 uses the raw value, later stages normalise it by its own magnitude (ref: model/loss.py:20-23)
            total = total + (l_intr if epoch_type == "CAM_PARAM_EPOCH" else l_intr / (l_intr.detach() + 1e-8))
        if "extr" in loss_dict:
            total = total + self.get_reproject_loss(loss_dict["extr"])
        if "rgb" in loss_dict:
            total = total + self.get_rgb_loss(loss_dict["rgb"])
        return total

    def get_rgb_loss(self, rgbs_list):
        rgb_c, rgb_f, gt = rgbs_list
        loss = F.mse_loss(rgb_c, gt)
        if rgb_f is not None:
            loss = loss + F.mse_loss(rgb_f, gt)
        return loss

    def get_reproject_loss(self, rpro_list):
        pd, gt = rpro_list
        lx = F.mse_loss(pd[..., 0] / self.img_w, gt[..., 0] / self.img_w)
        ly = F.mse_loss(pd[..., 1] / self.img_h, gt[..., 1] / self.img_h)
        return lx + ly

"""SinCosEmbedding and CorseFine_NeRF with the reference's constructor/forward signatures, parameter
names and checkpoint layout (ref: model/net_block.py), computed by libmcnerf.so kernels."""
import torch
import torch.nn as nn

from mc_nerf_b200 import ops


class SinCosEmbedding(nn.Module):
    """x [M,3] -> [M, 3+6L]: [x, per coordinate: sin(2^k c), cos(2^k c)], optional BARF window.
    ref: model/net_block.py:6-35."""

    def __init__(self, sys_params):
        super().__init__()
        self.sys_param = sys_params
        self.device = sys_params["device_type"]
        self.n_freqs = sys_params["emb_freqs_xyz"]
        self.barf_mode = sys_params["barf_mask"]
        self.barf_start = sys_params["barf_start"]
        self.barf_end = sys_params["barf_end"]
        self.in_channels = 3
        self.out_channels = self.in_channels * (2 * self.n_freqs + 1)
        self.freq_bands = 2 ** torch.linspace(0, self.n_freqs - 1, self.n_freqs, device=self.device)

    def band_weights(self, step_r):
        """The BARF weights the kernels take (None when the window is off)."""
        if not self.barf_mode:
            return None
        return ops.barf_band_weights(float(step_r), self.barf_start, self.barf_end, self.n_freqs)

    def forward(self, x, step_r):
        return ops.EncodePointsFn.apply(x, self.n_freqs, self.band_weights(step_r))


class CorseFine_NeRF(nn.Module):
    """D x W ReLU MLP with one input skip, sigma head (W->W->1) and SH head (W->W->3 (MLP_deg+1)^2) -> (sigma_raw, rgb).
    ref: model/net_block.py:37-78.  Submodule and parameter names match the reference's state_dict.  MLP_deg = 2 (the
    shipped value) runs on the tcgen05 path; other degrees on the fp32 kernels."""

    def __init__(self, sys_params, type="coarse"):
        super().__init__()
        self.in_channels_xyz = 3 * (2 * sys_params["emb_freqs_xyz"] + 1)
        self.deg = sys_params["MLP_deg"]
        self.depth = sys_params[f"{type}_MLP_depth"]
        self.width = sys_params[f"{type}_MLP_width"]
        self.skips = sys_params[f"{type}_MLP_skip"]
        if not 0 <= self.deg <= 4:
            raise ValueError("MLP_deg must be 0..4 (the degrees eval_sh implements, ref: model/net_utils.py:150)")
        for i in range(self.depth):
            if i == 0:
                k = self.in_channels_xyz
            elif i in self.skips:
                k = self.width + self.in_channels_xyz
            else:
                k = self.width
            setattr(self, f"xyz_encoding_{i+1}", nn.Sequential(nn.Linear(k, self.width), nn.ReLU(True)))
        self.sigma = nn.Sequential(nn.Linear(self.width, self.width), nn.ReLU(True), nn.Linear(self.width, 1))
        self.sh = nn.Sequential(nn.Linear(self.width, self.width), nn.ReLU(True),
                                nn.Linear(self.width, 3 * (self.deg + 1) ** 2))

    def param_dict(self):
        """reference state_dict name -> Parameter, in ops.param_names order (cached: Parameter objects are stable
        under .to(), load_state_dict and optimiser updates, which all work in place on .data)."""
        pd = self.__dict__.get("_param_dict")
        if pd is None:
            sd = dict(self.named_parameters())
            pd = {k: sd[k] for k in ops.param_names(self.depth)}
            self.__dict__["_param_dict"] = pd
        return pd

    def cfg(self):
        return (self.depth, self.width, tuple(self.skips))

    def forward(self, x, dirs):
        p = self.param_dict()
        return ops.MLPFn.apply(x, dirs, self.depth, self.width, tuple(self.skips), *p.values())

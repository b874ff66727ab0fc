"""Drop-in replacement for the reference's `model` package (ref: model/__init__.py:1-5).

Put the directory that CONTAINS this package first on sys.path under the name `model`
 and the reference's main.py / config.yaml drive it unchanged.
"""
from .mc_nerf import MC_Model, NeRF_Model
from .net_utils import RAdam
from .loss import MC_NeRF_Loss
from .net_utils import apply_depth_colormap
from .external.pohsun_ssim import pytorch_ssim

__all__ = ["MC_Model", "NeRF_Model", "RAdam", "MC_NeRF_Loss", "apply_depth_colormap", "pytorch_ssim"]

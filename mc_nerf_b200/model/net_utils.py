"""eval_sh, RAdam and the small helpers main.py imports, with the reference's signatures
(ref: model/net_utils.py).  eval_sh and RAdam.step run in libmcnerf.so kernels."""
import ctypes
import math

import torch
import torch.distributed as dist
from torch.optim.optimizer import Optimizer

from mc_nerf_b200 import ops
from mc_nerf_b200._lib import lib


def eval_sh(deg, sh, dirs):
    """Real spherical harmonics of degree 0..4 dotted with coefficients.
    sh [..., C, (deg+1)^2], dirs [..., 3] -> [..., C].  ref: model/net_utils.py:103-191 (the same hard-coded polynomials
    and constants, evaluated by one kernel; gradients flow into the coefficients and the directions)."""
    assert 0 <= deg <= 4
    assert (deg + 1) ** 2 == sh.shape[-1]
    if sh.shape[-2] != 3:
        raise NotImplementedError("eval_sh: libmcnerf evaluates 3 colour channels (the reference's only use)")
    lead = sh.shape[:-2]
    sh2 = sh.reshape(-1, 3, sh.shape[-1])
    d2 = dirs.expand(*lead, 3).reshape(-1, 3)
    return ops.EvalSHFn.apply(sh2, d2, deg).reshape(*lead, 3)


class RAdam(Optimizer):
    """Rectified Adam with the reference's exact schedule: 10-slot (step % 10) cache of (N_sma, step_size),
    N_sma >= 5 switch, SGD-like degenerate branch, `p -= wd*lr*p` decay (ref: model/net_utils.py:10-101).
    The per-tensor update is one fused kernel (mcnerf_radam_step) instead of ~10 ATen launches + fp32 copies."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        self.degenerated_to_sgd = degenerated_to_sgd
        self.grad_scale = 1.0     # every gradient is multiplied by this inside the update kernel (1 / world size when
                                  # the data-parallel all-reduce leaves the SUM in .grad: parallel.GradSync)
        if isinstance(params, (list, tuple)) and len(params) > 0 and isinstance(params[0], dict):
            for param in params:
                if "betas" in param and (param["betas"][0] != betas[0] or param["betas"][1] != betas[1]):
                    param["buffer"] = [[None, None, None] for _ in range(10)]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                        buffer=[[None, None, None] for _ in range(10)])
        super().__init__(params, defaults)

    @staticmethod
    def schedule(step, beta1, beta2, degenerated_to_sgd=True):
        """(N_sma, step_size) for a given step count (ref: model/net_utils.py:70-84)."""
        beta2_t = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma >= 5:
            ss = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) \
                / (1 - beta1 ** step)
        elif degenerated_to_sgd:
            ss = 1.0 / (1 - beta1 ** step)
        else:
            ss = -1
        return n_sma, ss

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stream = ops._stream()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step = {}                     # tensors that share a step count share every scalar of the update
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("RAdam does not support sparse gradients")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("RAdam (libmcnerf): contiguous fp32 parameters expected")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                state["step"] += 1
                by_step.setdefault(state["step"], []).append((p, state))
            for step, items in by_step.items():
                buffered = group["buffer"][int(step % 10)]
                if step == buffered[0]:
                    n_sma, step_size = buffered[1], buffered[2]
                else:
                    buffered[0] = step
                    n_sma, step_size = self.schedule(step, beta1, beta2, self.degenerated_to_sgd)
                    buffered[1], buffered[2] = n_sma, step_size
                mode = 1 if n_sma >= 5 else (2 if step_size > 0 else 0)
                n = len(items)
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p, _ in items]
                vp = ctypes.c_void_p * n
                lib().call("mcnerf_radam_multi", n,
                           vp(*[p.data_ptr() for p, _ in items]), vp(*[g.data_ptr() for g in grads]),
                           vp(*[st["exp_avg"].data_ptr() for _, st in items]),
                           vp(*[st["exp_avg_sq"].data_ptr() for _, st in items]),
                           (ctypes.c_int64 * n)(*[p.numel() for p, _ in items]),
                           float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                           float(group["weight_decay"]), float(step_size), mode, float(self.grad_scale), stream)
                # the kernel wrote the parameters through raw pointers: bump their version counters as an in-place
                # torch op would, so that derived caches (the packed bf16 weight images) and autograd notice
                torch.autograd.graph.increment_version([p for p, _ in items])
        return loss


def get_rank():
    if not is_dist_avail_and_initialized():
        return 0
    return dist.get_rank()


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def apply_colormap(image, cmap="viridis"):
    """image [...,1] in [0,1] -> RGB through a matplotlib listed colormap (ref: model/net_utils.py:205-217;
    demo post-processing, needs matplotlib)."""
    from matplotlib import cm
    table = torch.tensor(cm.get_cmap(cmap).colors).to(image.device)
    idx = (image * 255).long().clamp_(63, 255)
    return table[idx[..., 0]]


def apply_depth_colormap(depth, accumulation=None, near_plane=None, far_plane=None, cmap="turbo"):
    """ref: model/net_utils.py:219-231."""
    img = apply_colormap(torch.clip(depth, 0, 1), cmap=cmap)
    if accumulation is not None:
        img = img * accumulation + (1 - accumulation)
    return img

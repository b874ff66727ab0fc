"""The tcgen05 MLP path's rounding points emulated on top of the CPU oracle (test infrastructure).

bf16 weights / encoded inputs / per-layer activations with fp32 accumulation in the forward pass; in the backward
pass every dY tile is rounded to bf16 as the kernels store it, and weight gradients are formed from the bf16
activations and bf16 dY.  Used by tests/test_tc_mlp_gpu.py (kernel vs emulation, tight) and by
tests/golden/make_golden.py::make_cfg2_bf16emu (the benched configuration through the emulation)."""
import torch
import torch.nn.functional as F

from oracle import mcnerf_oracle as orc


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def mlp_forward_bf16emu(p, x_enc, dirs, depth, skips):
    """oracle.mlp_forward with the kernel's rounding points."""
    x = bf(x_enc)
    h = x
    for i in range(depth):
        if i in skips:
            h = torch.cat([x, h], -1)
        h = bf(F.relu(F.linear(h, bf(p[f"xyz_encoding_{i+1}.0.weight"]), p[f"xyz_encoding_{i+1}.0.bias"])))
    s = F.relu(F.linear(h, bf(p["sigma.0.weight"]), p["sigma.0.bias"]))           # kept fp32 in the epilogue
    sigma = F.linear(s, p["sigma.2.weight"], p["sigma.2.bias"])
    c = bf(F.relu(F.linear(h, bf(p["sh.0.weight"]), p["sh.0.bias"])))
    sh = F.linear(c, bf(p["sh.2.weight"]), p["sh.2.bias"])
    rgb = torch.sigmoid(orc.eval_sh_deg2(sh.reshape(-1, 3, 9), dirs))
    return torch.cat([sigma, rgb], -1)


class _RoundGrad(torch.autograd.Function):
    """identity whose backward rounds the gradient to bf16 (the kernels store every dY tile as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return bf(g)


class _BfSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient."""

    @staticmethod
    def forward(ctx, x):
        return bf(x)

    @staticmethod
    def backward(ctx, g):
        return g


def mlp_forward_bf16emu_trainable(p, x_enc, dirs, depth, skips):
    """same rounding points as the tcgen05 path in forward AND backward: bf16 weights/activations/dY tiles,
    fp32 accumulation; weight gradients are formed from the bf16 activations and bf16 dY."""
    rg, q = _RoundGrad.apply, _BfSTE.apply
    x = q(x_enc)
    h = x
    for i in range(depth):
        if i in skips:
            h = torch.cat([x, h], -1)
        h = q(F.relu(rg(F.linear(h, q(p[f"xyz_encoding_{i+1}.0.weight"]), p[f"xyz_encoding_{i+1}.0.bias"]))))
    s = F.relu(rg(F.linear(h, q(p["sigma.0.weight"]), p["sigma.0.bias"])))
    sigma = F.linear(s, p["sigma.2.weight"], p["sigma.2.bias"])
    c = q(F.relu(rg(F.linear(h, q(p["sh.0.weight"]), p["sh.0.bias"]))))
    sh = rg(F.linear(c, q(p["sh.2.weight"]), p["sh.2.bias"]))
    rgb = torch.sigmoid(orc.eval_sh_deg2(sh.reshape(-1, 3, 9), dirs))
    return torch.cat([sigma, rgb], -1)

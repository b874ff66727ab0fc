import torch


def checksum(t):
    """Order-sensitive checksum used to confirm seeded inputs regenerate identically."""
    t = t.double()
    idx = torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape)
    return [float(t.sum()), float(t.abs().sum()), float((t * idx).sum() / t.numel())]


PROBES = 8


def probe_dots(name_index, g):
    """Dot products of a gradient tensor with PROBES seeded N(0,1) vectors (same recipe as
    tests/golden/make_golden.py::probe_dots): for an error vector e, E[(u.e)^2] = |e|^2, so the rms difference of the
    probe products estimates |g - g_ref| - sensitive to direction, unlike a norm comparison."""
    gen = torch.Generator().manual_seed(90000 + name_index)
    u = torch.randn(PROBES, g.numel(), generator=gen, dtype=torch.float64)
    return (u @ g.detach().reshape(-1).double().cpu()).float()

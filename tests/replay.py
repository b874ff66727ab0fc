"""Replay explicit random draws through torch's RNG entry points (tests only)."""
import torch


class Replay:
    """While active, torch.randn / torch.randperm / Tensor.uniform_ return the queued tensors in order
    (moved to the requested device), so the product code consumes exactly the draws the reference did."""

    def __init__(self, randn=(), randperm=(), uniform=()):
        self.q = dict(randn=list(randn), randperm=list(randperm), uniform=list(uniform))

    def __enter__(self):
        self._saved = (torch.randn, torch.randperm, torch.Tensor.uniform_)
        q = self.q

        def randn(*size, **kw):
            t = q["randn"].pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.clone().to(kw.get("device", "cpu"))

        def randperm(n, **kw):
            t = q["randperm"].pop(0)
            assert t.shape[0] == n
            return t.clone().to(kw.get("device", "cpu"))

        def uniform_(self_t, a=0.0, b=1.0):
            t = q["uniform"].pop(0)
            assert t.shape == self_t.shape
            return self_t.copy_(t)

        torch.randn, torch.randperm, torch.Tensor.uniform_ = randn, randperm, uniform_
        return self

    def __exit__(self, *a):
        torch.randn, torch.randperm, torch.Tensor.uniform_ = self._saved
        if a[0] is None:
            assert not any(self.q.values()), {k: len(v) for k, v in self.q.items()}

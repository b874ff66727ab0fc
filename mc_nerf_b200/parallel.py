"""Data-parallel plumbing: rays shard across ranks (each rank renders its own camera's batch, the reference's DDP
semantic, ref: main.py:60-62, data/data_read.py:358-360); the only exchange is ONE all-reduce of a flat gradient
buffer (MLP + camera gradients, 5.06 MB at 8x256/8x256) per step.  torch.distributed only (NCCL on GPUs, gloo in
the CPU tests); no data-path collective."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """Averages the gradients of `params` across ranks with a single all-reduce.

    Parameters a rank did not touch this step (grad is None: e.g. `weights_pose` in the fine-tune stage, the MLPs
    in the camera stage) contribute zeros, which is what DistributedDataParallel(find_unused_parameters=True)
    does in the reference."""

    def __init__(self, params):
        self.params = [p for p in params]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self):
        n = world()
        if n == 1:
            return
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                self.flat[off:off + k].zero_()
            else:
                self.flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / n)
        off = 0
        for p in self.params:
            k = p.numel()
            g = self.flat[off:off + k].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += k


def shard_rays(n_rays, rank=None, n_ranks=None):
    """[begin, end) of this rank's contiguous slice of a ray batch (strong-scaling mode: equal slices)."""
    rank = dist.get_rank() if rank is None else rank
    n_ranks = world() if n_ranks is None else n_ranks
    per = n_rays // n_ranks
    return rank * per, (rank + 1) * per if rank < n_ranks - 1 else n_rays

"""Data-parallel plumbing: rays shard across ranks (each rank renders its own camera's batch, the reference's DDP
semantic, ref: main.py:60-62, data/data_read.py:358-360); the only exchange is ONE all-reduce of a flat gradient
buffer (MLP + camera gradients, 5.06 MB at 8x256/8x256) per step.  torch.distributed only (NCCL on GPUs, gloo in
the CPU tests); no data-path collective."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """Averages the gradients of `params` across ranks with as few all-reduces as the memory layout allows.

    The renderer's backward hands autograd the MLP gradients as views of ONE buffer per network (render.py), and
    autograd keeps those views as `.grad` - so each network's gradients are all-reduced IN PLACE through an alias
    of that buffer (no flatten / unflatten copies); the handful of small remaining tensors (camera parameters) go
    through one concatenated buffer.  Parameters a rank did not touch this step (grad is None: e.g. `weights_pose`
    in the fine-tune stage) contribute zeros, which is what DistributedDataParallel(find_unused_parameters=True)
    does in the reference (ref: main.py:61)."""

    def __init__(self, params):
        self.params = [p for p in params]

    @staticmethod
    def _runs(grads):
        """maximal runs of gradients that sit back to back in one storage -> [(first_index, count, numel)]"""
        runs, i = [], 0
        while i < len(grads):
            g = grads[i]
            j, end, numel = i + 1, g.data_ptr() + g.numel() * g.element_size(), g.numel()
            base = g.untyped_storage().data_ptr()
            while (j < len(grads) and grads[j].data_ptr() == end and grads[j].untyped_storage().data_ptr() == base
                   and grads[j].is_contiguous() and g.is_contiguous()):
                end += grads[j].numel() * grads[j].element_size()
                numel += grads[j].numel()
                j += 1
            runs.append((i, j - i, numel))
            i = j
        return runs

    def __call__(self):
        n = world()
        if n == 1:
            return
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        singles = []
        self.n_collectives = 0
        for first, count, numel in self._runs(grads):
            g0 = grads[first]
            if count > 1 and g0.is_contiguous():
                alias = torch.empty(0, dtype=g0.dtype, device=g0.device).set_(g0.untyped_storage(), g0.storage_offset(),
                                                                                (numel,))
                dist.all_reduce(alias, op=dist.ReduceOp.SUM)
                alias.mul_(1.0 / n)
                self.n_collectives += 1
            else:
                singles += grads[first:first + count]
        if singles:
            flat = torch.cat([g.reshape(-1) for g in singles])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / n)
            self.n_collectives += 1
            torch._foreach_copy_(singles, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in singles]), singles)])


def shard_rays(n_rays, rank=None, n_ranks=None):
    """[begin, end) of this rank's contiguous slice of a ray batch (strong-scaling mode: equal slices)."""
    rank = dist.get_rank() if rank is None else rank
    n_ranks = world() if n_ranks is None else n_ranks
    per = n_rays // n_ranks
    return rank * per, (rank + 1) * per if rank < n_ranks - 1 else n_rays

"""Data-parallel plumbing: rays shard across ranks (each rank renders its own camera's batch, the reference's DDP
semantic, ref: main.py:60-62, data/data_read.py:358-360); the only exchange is ONE all-reduce of a flat gradient
buffer (MLP + camera gradients, 5.06 MB at 8x256/8x256) per step.  torch.distributed only (NCCL on GPUs, gloo in
the CPU tests); no data-path collective."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """Averages the gradients of `params` across ranks with ONE all-reduce of a flat buffer laid out by parameter
    order and numel - never by a rank's local gradient memory layout, so every rank always issues the same
    collective (a gradient that is None on one rank only, or that lives in a private tensor there, cannot
    desynchronise the ranks).  Parameters a rank did not touch this step (grad is None: e.g. `weights_pose` in the
    fine-tune stage) contribute zeros, which is what DistributedDataParallel(find_unused_parameters=True) does in the
    reference (ref: main.py:61).  The copy in / copy out costs two small launches; the overlapped, copy-free path
    is GradSync below."""

    def __init__(self, params):
        self.params = [p for p in params]
        self.n_collectives = 0

    def __call__(self):
        n = world()
        if n == 1:
            return
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / n)
        self.n_collectives = 1
        torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])


def broadcast_parameters(model, src=0):
    """Every rank starts from rank `src`'s parameters - what DistributedDataParallel does at construction
    (ref: main.py:61 wraps the model after each rank seeded itself with 42 + rank, main.py:273-277)."""
    if world() == 1:
        return
    with torch.no_grad():
        for p in model.parameters():
            dist.broadcast(p.data, src=src)


def parameters_identical(model):
    """True iff all ranks hold bit-identical parameters (checked through the int32 bit patterns)."""
    if world() == 1:
        return True
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1).view(torch.int32) for p in model.parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))


class GradSync:
    """Gradient averaging overlapped with the backward pass (the reference gets this from DistributedDataParallel's
    bucketed all-reduce, ref: main.py:61,84).

    The renderer's backward finishes the FINE network first (render.RenderFn.backward); through render.GRAD_HOOK its
    flat gradient buffer is all-reduced on a communication stream right there, overlapping the coarse network's
    backward (~0.8 ms of kernels at the benched shape); the coarse buffer follows and overlaps the camera-model
    backward.  `finish()` - after backward - all-reduces what is left (the six camera tensors, 7 KB, as ONE coalesced
    NCCL launch; gradients of networks that did not go through the hook) and joins the communication stream.
    Buffers, order and sizes are fixed by the model, never by a rank's local gradient state, so all ranks always
    issue the same collectives.  The sum is NOT divided here: pass `grad_scale = 1 / world` to the optimiser
    (model.RAdam reads `opt.grad_scale`), which folds the division into its update kernel.

    Works eagerly and under CUDA-graph capture (append `finish` to GraphedTrainStep.after_backward: the fork and the
    join of the communication stream are then part of the captured graph).  Requires gradients to be None before
    backward (zero_grad(set_to_none=True), the default): the in-place reduction must not race an accumulation."""

    def __init__(self, model, overlap=True):
        self.model = model
        self.n = world()
        self.overlap = overlap
        self._hooked = set()
        self._comm = None
        self.n_collectives = 0
        named = list(model.named_parameters())
        self._net_params = {"coarse": [p for k, p in named if k.startswith("nerf.nerf_coarse.")],
                            "fine": [p for k, p in named if k.startswith("nerf.nerf_fine.")]}
        self._other = [p for k, p in named if not k.startswith("nerf.nerf_coarse.") and not k.startswith("nerf.nerf_fine.")]

    def install(self):
        from . import render
        render.GRAD_HOOK = self._on_ready if self.n > 1 else None
        return self

    def uninstall(self):
        from . import render
        render.GRAD_HOOK = None

    def _on_ready(self, name, flat):
        self.n_collectives += 1
        if not self.overlap or not flat.is_cuda:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        else:
            main = torch.cuda.current_stream(flat.device)
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=flat.device)
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.record_stream(self._comm)
        self._hooked.add(name)

    def finish(self):
        if self.n == 1:
            return
        rest = list(self._other)
        for name, ps in self._net_params.items():
            if name not in self._hooked:
                rest += ps
        self._hooked.clear()
        for p in rest:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in rest]
        if grads:
            dev = grads[0].device
            if dev.type == "cuda":
                with dist._coalescing_manager(device=dev, async_ops=False):     # one NCCL launch for all of them
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM)
            else:
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM)
            self.n_collectives += 1
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)

    def collectives_per_step(self, steps):
        return self.n_collectives / max(1, steps)


def shard_rays(n_rays, rank=None, n_ranks=None):
    """[begin, end) of this rank's contiguous slice of a ray batch (strong-scaling mode: equal slices)."""
    rank = dist.get_rank() if rank is None else rank
    n_ranks = world() if n_ranks is None else n_ranks
    per = n_rays // n_ranks
    return rank * per, (rank + 1) * per if rank < n_ranks - 1 else n_rays

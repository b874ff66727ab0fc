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
    backward.  `finish()` - after backward - reduces what is left (the six camera tensors, 7 KB; gradients of networks
    that did not go through the hook) and joins the communication stream.
    Buffers, order and sizes are fixed by the model, never by a rank's local gradient state, so all ranks always
    issue the same collectives.  The SUM is left in .grad: pass `grad_scale = 1 / world` to the optimiser (model.RAdam
    reads `opt.grad_scale`), which folds the division into its update kernel.

    transport = "p2p" (default on one NVSwitch box): the gradients live in SYMMETRIC memory mapped into every rank
      (torch.distributed._symmetric_memory does the handle exchange; render.GRAD_ALLOC places both networks' gradient
      buffers there, so nothing is copied) and each reduction is ONE launch of libmcnerf's two-shot NVLink all-reduce
      kernel (csrc/allreduce.cu): ~3 launches of NCCL cost 110-120 us of exposed time per step, this costs a fraction;
    transport = "nccl": torch.distributed all-reduces (any topology; also the gloo CPU tests).

    `install()` sets the renderer's process-global hooks (render.GRAD_HOOK / GRAD_ALLOC): one GradSync per process;
    `uninstall()` before rendering another model with gradients in the same process.
    Works eagerly and under CUDA-graph capture (append `finish` to GraphedTrainStep.after_backward: the fork and the
    join of the communication stream are then part of the captured graph).  Requires gradients to be None before
    backward (zero_grad(set_to_none=True), the default): the in-place reduction must not race an accumulation."""

    P2P_CTAS = 24

    def __init__(self, model, overlap=True, transport="nccl"):
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
        self.transport = transport if self.n > 1 else "nccl"
        self._p2p = None
        if self.transport == "p2p":
            self._setup_p2p()

    # ------------------------------------------------------------------ symmetric memory (p2p transport)
    def _setup_p2p(self):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from ._lib import P2P
        dev = next(self.model.parameters()).device
        pad4 = lambda n: (n + 3) // 4 * 4
        self._n_nets = sum(p.numel() for ps in self._net_params.values() for p in ps)
        self._cam_off = pad4(self._n_nets)
        self._rest_cap = pad4(sum(p.numel() for p in self._other)) + pad4(self._n_nets)      # room for un-hooked networks
        total = self._cam_off + self._rest_cap
        group = dist.group.WORLD
        self._sym = symm.empty(total, dtype=torch.float32, device=dev)
        self._sym.zero_()
        hdl = symm.rendezvous(self._sym, group)
        self._flags = symm.empty(64 * 2 * 8, dtype=torch.int32, device=dev)
        self._flags.zero_()
        hdl_f = symm.rendezvous(self._flags, group)
        self._epoch = torch.zeros(64, dtype=torch.int32, device=dev)
        ctx = P2P()
        ctx.rank, ctx.n_ranks, ctx.n_ctas = dist.get_rank(), self.n, self.P2P_CTAS
        for p in range(self.n):
            ctx.buf[p] = int(hdl.buffer_ptrs[p])
            ctx.flags[p] = int(hdl_f.buffer_ptrs[p])
        ctx.epoch = self._epoch.data_ptr()
        self._p2p, self._hdls = ctx, (hdl, hdl_f)
        torch.cuda.synchronize()
        dist.barrier()                       # every rank's buffers are zeroed and mapped before the first kernel

    def _alloc(self, numel, device):
        """render.GRAD_ALLOC: both networks' gradient buffer inside the symmetric region (zeroed)"""
        if self._p2p is None or numel != self._n_nets or device != self._sym.device:
            return None
        flat = self._sym[:numel]
        flat.zero_()
        return flat

    def _reduce(self, flat):
        """SUM-all-reduce a contiguous fp32 buffer on the current stream"""
        if self._p2p is not None and flat.is_cuda:
            base = self._sym.data_ptr()
            off = (flat.data_ptr() - base) // 4
            if 0 <= off and off + flat.numel() <= self._sym.numel() and off % 4 == 0 and flat.numel() % 4 == 0:
                import ctypes
                from . import ops
                from ._lib import lib
                lib().call("mcnerf_allreduce_p2p", ctypes.byref(self._p2p), int(off), int(flat.numel()), 1.0, ops._stream())
                return
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def install(self):
        from . import render
        render.GRAD_HOOK = self._on_ready if self.n > 1 else None
        render.GRAD_ALLOC = self._alloc if (self.n > 1 and self._p2p is not None) else None
        return self

    def uninstall(self):
        from . import render
        render.GRAD_HOOK = None
        render.GRAD_ALLOC = None

    def _on_ready(self, name, flat):
        self.n_collectives += 1
        if not self.overlap or not flat.is_cuda:
            self._reduce(flat)
        else:
            main = torch.cuda.current_stream(flat.device)
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=flat.device, priority=-1)      # its few CTAs go first when SMs free up
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                self._reduce(flat)
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
        if self._comm is not None:           # the p2p kernels of one rank must run in the same order on every rank
            torch.cuda.current_stream().wait_stream(self._comm)
        if grads:
            dev = grads[0].device
            n_rest = sum(g.numel() for g in grads)
            if self._p2p is not None and dev.type == "cuda" and n_rest <= self._rest_cap:
                # gather the small tensors into the symmetric region (one launch), reduce it (one launch), and let
                # .grad point at the reduced values (no copy back)
                region = self._sym[self._cam_off:self._cam_off + (n_rest + 3) // 4 * 4]
                torch.cat([g.reshape(-1) for g in grads], out=region[:n_rest])
                self._reduce(region)
                off = 0
                for p, g in zip(rest, grads):
                    p.grad = region[off:off + g.numel()].view_as(g)
                    off += g.numel()
            elif dev.type == "cuda":
                with dist._coalescing_manager(device=dev, async_ops=False):     # one NCCL launch for all of them
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM)
            else:
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM)
            self.n_collectives += 1

    def collectives_per_step(self, steps):
        return self.n_collectives / max(1, steps)


def shard_rays(n_rays, rank=None, n_ranks=None):
    """[begin, end) of this rank's contiguous slice of a ray batch (strong-scaling mode: equal slices)."""
    rank = dist.get_rank() if rank is None else rank
    n_ranks = world() if n_ranks is None else n_ranks
    per = n_rays // n_ranks
    return rank * per, (rank + 1) * per if rank < n_ranks - 1 else n_rays

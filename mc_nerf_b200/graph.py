"""CUDA-graph replay of the train step's device work: forward + loss + backward captured once, replayed per step.

The reference's loop (main.py:78-92) issues ~100 kernel launches per step from Python; on a B200 the kernels of one
4096-ray step take 3.9 ms while PyTorch needs ~3.1 ms of host time to issue them, so the host bounds multi-GPU
scaling and the end-to-end rate.  GraphedTrainStep keeps the reference's objects (MC_Model, MC_NeRF_Loss, the
optimiser) and replaces only HOW the launches are issued:

    step = GraphedTrainStep(model, loss_fn)
    for data in loader:                                   # data exactly as main.py hands it to the model
        loss = step(data, epoch, epoch_type, cur_ratio)   # copies inputs, replays the graph; .grad is populated
        optimizer.step()                                  # (gradient all-reduce first when distributed)

`step.prefetch(next_data)` right after a call uploads the next step's (pinned) inputs on a side stream while the graph
runs.

What is baked into a captured graph and therefore part of its cache key: epoch_type (the stage) and the input shapes.
cur_ratio is NOT baked: the BARF frequency weights it selects are kept in a device buffer (NeRF_Model.set_band_weights)
that the captured kernels read, so one graph serves the whole stage although cur_ratio changes every step.
Random draws are NOT baked: torch's CUDA generator is graph-aware, every replay consumes fresh Philox offsets in the
reference's draw order.  The camera id is a device
tensor, so one graph serves all cameras.  Gradients are produced by the graph into static buffers (the usual
whole-network-capture contract): do not call optimizer.zero_grad() between replays, and do not accumulate.
Drop every reference to losses / outputs of earlier EAGER steps of the same model before the first graphed step: a live
eager autograd graph keeps its AccumulateGrad nodes bound to the default stream, which breaks the capture.
The reference's 128-samples-per-ray cap (samples*scale > 128) is drawn on the device (render.select_and_cap), so those
configurations capture as well.
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, loss_fn, warmup=3):
        self.model, self.loss_fn, self.warmup = model, loss_fn, warmup
        self._graphs = {}
        self.after_backward = []   # callables run right after loss.backward() INSIDE the captured region (e.g. the join of
                                   # parallel.GradSync's communication stream + the camera-gradient all-reduce)
        self._staged = None       # (ids of the host tensors, device staging copies, upload-done event)
        self._consumed = None     # event: the last staging -> static copy has been issued on the main stream
        self._copy_stream = None

    def prefetch(self, data):
        """Start uploading the NEXT step's inputs (pinned host tensors) on a side stream while the current step runs;
        the next __call__ with the same tensors only pays a device-to-device copy into the graph's static inputs."""
        dev = torch.device(self.model.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        side = self._copy_stream
        if self._staged is None or any(s.shape != t.shape or s.dtype != t.dtype for s, t in zip(self._staged[1], data)):
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in data)
        else:
            bufs = self._staged[1]
        if self._consumed is not None:
            side.wait_event(self._consumed)            # the previous contents have been handed to the graph
        with torch.cuda.stream(side):
            for b, t in zip(bufs, data):
                b.copy_(t, non_blocking=True)
            done = side.record_event()
        self._staged = (tuple(id(t) for t in data), bufs, done)

    def _eager(self, static, key):
        epoch_type, epoch, ratio = key[0], self._call[0], self._call[1]
        loss_dict, _, _, _ = self.model(static, epoch, epoch_type, ratio)
        loss = self.loss_fn(loss_dict, epoch_type)
        loss.backward()
        for fn in self.after_backward:
            fn()
        return loss

    def _capture(self, data, key):
        m = self.model
        dev = torch.device(m.device)
        static = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev).copy_(t) for t in data)
        params = [p for p in m.parameters()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in params:
                    p.grad = None
                self._eager(static, key)
        torch.cuda.current_stream(dev).wait_stream(side)
        for p in params:
            p.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = self._eager(static, key)
        return g, static, loss

    def __call__(self, data, epoch, epoch_type, cur_ratio):
        # cur_ratio moves every step (main.py:80): the BARF weights it selects live in a device buffer the captured
        # kernels read, refreshed here; `epoch` is not used by the render (ref: model/mc_nerf.py:598-646)
        nerf = self.model.nerf
        if nerf.__dict__.get("_band_w_dev") is None:
            nerf.use_device_band_weights(True)
        nerf.set_band_weights(cur_ratio if epoch_type == "GLOBAL_OPTIM_EPOCH" else 1)
        self._call = (epoch, float(cur_ratio))
        key = (epoch_type, tuple(tuple(t.shape) for t in data))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(data, key)
        g, static, loss = ent
        staged = self._staged
        if staged is not None and staged[0] == tuple(id(t) for t in data):
            main = torch.cuda.current_stream()
            main.wait_event(staged[2])
            for s, b in zip(static, staged[1]):
                s.copy_(b, non_blocking=True)
            self._consumed = main.record_event()
            self._staged = (None, staged[1], None)      # keep the buffers, forget the contents
        else:
            for s, t in zip(static, data):
                s.copy_(t, non_blocking=True)
        g.replay()
        return loss

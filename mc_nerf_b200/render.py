"""Fused coarse+fine renderer: one autograd node for NeRF_Model.render_rays_train / render_rays_test.

Forward  : sample+encode (coarse grid) -> coarse MLP -> compositing -> selection weights (fresh noise)
           -> device-side threshold/compaction (no host sync) -> sample+encode (selected fine samples)
           -> fine MLP -> scatter into defaults -> compositing.
Backward : compositing bwd -> gather -> MLP bwd (dgrad + wgrad) -> encoding bwd -> per-ray (dL/do, dL/dd),
           for the fine and then the coarse branch.

ref: model/mc_nerf.py:598-736.  Random draws (jitter, three noise tensors) are explicit arguments so that the
caller decides where they come from (torch's generator in the reference's draw order, or fixtures in tests).
"""
import ctypes
import os

import torch

from . import ops
from ._lib import MlpGrads, lib

_p, _stream = ops._p, ops._stream


LAST = {}      # debug/bench: row counts of the most recent fine pass (device tensor: no sync is forced here)
GRAD_ALLOC = None  # callable(numel, device) -> zeroed flat fp32 buffer for BOTH networks' gradients, or None: lets the
                   # data-parallel layer place the gradients straight into NVLink-mapped symmetric memory (parallel.GradSync)
GRAD_HOOK = None   # callable(name, flat) or None: called inside RenderFn.backward, on the launching stream, as soon as
                   # ONE network's gradients ("fine", then "coarse") are final in their flat fp32 buffer - the gradient
                   # all-reduce of the fine network then overlaps the coarse network's backward (parallel.GradSync)


class RenderCfg:
    """Static renderer configuration (the sys_param keys NeRF_Model reads, ref: model/mc_nerf.py:547-571)."""

    def __init__(self, near, far, Sc, scale, n_freqs, white_back, sigma_default, thresh,
                 coarse, fine, precision="fp32", device_rng=False):
        self.near, self.far, self.Sc, self.scale = float(near), float(far), int(Sc), int(scale)
        self.Sf = self.Sc * self.scale
        self.n_freqs, self.white_back = int(n_freqs), bool(white_back)
        self.sigma_default, self.thresh = float(sigma_default), float(thresh)
        self.coarse, self.fine = coarse, fine            # (depth, width, skips)
        self.in_ch = 3 + 6 * self.n_freqs
        self.precision = precision
        # True: jitter and the three density-noise draws come from Philox inside the kernels (keyed from torch's CUDA
        # generator: same distributions as the reference's uniform_/randn draws, not the same values);
        # False: torch.rand / torch.randn in the reference's order (draw-for-draw replay, tests)
        self.device_rng = bool(device_rng)
        self.sh_dim = 27                           # 3 (MLP_deg + 1)^2; from_sys_param sets it
        # train-only cap on the selected fine samples, per ray (the reference hard-codes 128, model/mc_nerf.py:630;
        # sys_param key `fine_sample_cap` - a value >= Sf switches the cap off, bench.py --workload stress reports both)
        self.fine_cap = 128

    @staticmethod
    def from_sys_param(sp, precision=None):
        """`mlp_precision` (sys_param key, or env MCNERF_PRECISION): "bf16" = tcgen05 tensor-core path (default),
        "fp32" = CUDA-core exact-parity path.  Shapes the tensor-core kernels do not implement use fp32."""
        import os
        if precision is None:
            precision = sp.get("mlp_precision", os.environ.get("MCNERF_PRECISION", "bf16"))
        # `noise_sampler` ("device" | "torch") follows `pixel_sampler` unless given: "device" is the package default
        device_rng = sp.get("noise_sampler", "torch" if sp.get("pixel_sampler", "device") == "randperm" else "device") \
            == "device" and str(sp.get("device_type", "cuda")).startswith("cuda")
        cfg = RenderCfg(sp["near"], sp["far"], sp["samples"], sp["scale"], sp["emb_freqs_xyz"], sp["white_back"],
                        sp["sigma_default"], sp["sample_weight_thresh"],
                        (sp["coarse_MLP_depth"], sp["coarse_MLP_width"], tuple(sp["coarse_MLP_skip"])),
                        (sp["fine_MLP_depth"], sp["fine_MLP_width"], tuple(sp["fine_MLP_skip"])), precision, device_rng)
        cfg.sh_dim = 3 * (int(sp.get("MLP_deg", 2)) + 1) ** 2
        cfg.fine_cap = int(sp.get("fine_sample_cap", 128))
        return cfg


def _owner_cache(params):
    """Derived-data cache of ONE network (packed bf16 weight images, zero-padded shadow), stored on its first Parameter
    so that it lives and dies with the network.  (A global cache keyed by device address would hand a newly created
    network the images of a freed one that happened to occupy the same memory with the same version counters.)"""
    owner = next(iter(params.values()))
    cache = owner.__dict__.get("_mcnerf_cache")
    if cache is None:
        cache = owner.__dict__["_mcnerf_cache"] = {}
    return cache


def _tc_weights(ps, tensors, need_bwd, cache):
    tcw = cache.get("tcw")
    if tcw is None:
        tcw = cache["tcw"] = ops.TcWeights()
    return tcw.get(ps, tensors, need_bwd)


_SIDE = {}         # device index -> side stream for the weight packs


def prefetch_weights(cfg, params_c, params_f, need_bwd=True):
    """Start packing both networks' bf16 images on a side stream (no-op for networks on the fp32 path).  Call at the
    top of a train step; render() picks the images up where it would otherwise pack them."""
    for net, params in ((cfg.coarse, params_c), (cfg.fine, params_f)):
        if not use_tc(cfg, net) or net[1] != ops.TC_WIDTH:      # padded shadows are refreshed and packed in line
            continue
        tensors = {k: ops._f32(params[k]) for k in ops.param_names(net[0])}
        first = next(iter(tensors.values()))
        if not first.is_cuda:
            continue
        side = _SIDE.get(first.device.index)
        if side is None:
            side = _SIDE[first.device.index] = torch.cuda.Stream(device=first.device)
        ps = ops.make_mlp_params(tensors, net[0], net[1], net[2], in_ch=cfg.in_ch)
        cache = _owner_cache(params)
        tcw = cache.get("tcw")
        if tcw is None:
            tcw = cache["tcw"] = ops.TcWeights()
        tcw.prefetch(ps, tensors, need_bwd, side)


def use_tc(cfg, net, tensors=None):
    """bf16 tcgen05 path when asked for and the network shape is one the tensor-core kernels implement: width 256
    natively, narrower networks through a zero-padded 256-wide shadow (ops.PaddedNet)."""
    depth, width, skips = net
    if width != ops.TC_WIDTH and os.environ.get("MCNERF_TC_PAD", "1") == "0":      # A/B switch for measurements
        return False
    if getattr(cfg, "sh_dim", 27) != 27:           # the tensor-core SH epilogue is specialised for MLP_deg = 2
        return False
    return (cfg.precision == "bf16" and 8 <= width <= ops.TC_WIDTH and cfg.n_freqs == 10 and 2 <= depth <= 12
            and len([s for s in skips if 0 < s < depth]) <= 1)


def _tc_view(cfg, net, tensors, cache):
    """-> (net the kernels see, tensors the kernels see, PaddedNet or None, cache of the tensors the kernels see)"""
    depth, width, skips = net
    if not use_tc(cfg, net) or width == ops.TC_WIDTH:
        return net, tensors, None, cache
    pad = cache.get("pad")
    if pad is None or pad.narrow_shapes != {k: tuple(v.shape) for k, v in tensors.items()} \
            or pad.flat.device != next(iter(tensors.values())).device:
        pad = cache["pad"] = ops.PaddedNet(tensors, depth, width, cfg.in_ch)
    return (depth, ops.TC_WIDTH, skips), pad.refresh(tensors), pad, pad.cache


def _branch_fwd(cfg, net, tensors, rays_o, rays_d, jitter, S, band_w, sel_idx, n_rows, n_rows_dev, train=True,
                cache=None, ray_offsets=None):
    """encode + MLP for one branch.  Returns (out4 [n_rows,4], saved-for-backward tuple)."""
    depth, width, skips = net
    dev = rays_o.device
    B = rays_o.shape[0]
    smp = ops.make_sampling(cfg.near, cfg.far, S, cfg.n_freqs, band_w)
    if use_tc(cfg, net):
        # fused sampling + encoding + MLP + SH head, bf16 tcgen05 (mlp_tc_fwd.cu)
        ps = ops.make_mlp_params(tensors, depth, width, skips, in_ch=cfg.in_ch)
        tcw = _tc_weights(ps, tensors, train, cache if cache is not None else {})
        tin = ops.make_tc_input_rays(rays_o, rays_d, jitter, smp, sel_idx, n_rows, n_rows_dev, ray_offsets)
        out4 = torch.empty(n_rows, 4, device=dev)
        stash = ops.tc_stash(ps, n_rows, dev) if train else None
        ops.mlp_tc_fwd(ps, tcw, tin, out4, stash)
        return out4, ("tc", tcw, tin, stash)
    enc = torch.empty(n_rows, cfg.in_ch, device=dev)
    lib().call("mcnerf_encode_rays_fwd", _p(rays_o), _p(rays_d), _p(jitter), B, ctypes.byref(smp),
               _p(sel_idx, torch.int32), n_rows, _p(n_rows_dev, torch.int32), _p(enc), cfg.in_ch, _stream())
    ps = ops.make_mlp_params(tensors, depth, width, skips, in_ch=cfg.in_ch)
    ws = ops.mlp_f32_workspace(ps, n_rows, dev)
    d = ops.make_dirs(rays_d, sel_idx, S)
    out4 = torch.empty(n_rows, 4, device=dev)
    lib().call("mcnerf_mlp_f32_fwd", ctypes.byref(ps), _p(enc), cfg.in_ch, ctypes.byref(d), n_rows,
               _p(n_rows_dev, torch.int32), _p(out4), _p(ws), _stream())
    return out4, (enc, ws)


def _branch_bwd(cfg, net, tensors, grads, rays_o, rays_d, jitter, S, band_w, sel_idx, n_rows, n_rows_dev,
                saved, out4, g_out4, g_rays_o, g_rays_d):
    depth, width, skips = net
    B = rays_o.shape[0]
    ps = ops.make_mlp_params(tensors, depth, width, skips, in_ch=cfg.in_ch)
    gs = ops.fill_mlp_struct(MlpGrads(), grads, depth)
    if saved[0] == "tc":
        _, tcw, tin, stash = saved
        ws = ops.tc_bwd_workspace(ps, n_rows, rays_o.device)
        ops.mlp_tc_bwd(ps, tcw, tin, out4, g_out4, stash, ws, gs, g_rays_o=g_rays_o, g_rays_d=g_rays_d)
        return
    enc, ws = saved
    d = ops.make_dirs(rays_d, sel_idx, S)
    g_enc = torch.empty_like(enc)
    lib().call("mcnerf_mlp_f32_bwd", ctypes.byref(ps), _p(enc), cfg.in_ch, ctypes.byref(d), n_rows,
               _p(n_rows_dev, torch.int32), _p(g_out4), _p(ws), ctypes.byref(gs), _p(g_enc), _p(g_rays_d), _stream())
    smp = ops.make_sampling(cfg.near, cfg.far, S, cfg.n_freqs, band_w)
    lib().call("mcnerf_encode_rays_bwd", _p(rays_o), _p(rays_d), _p(jitter), B, ctypes.byref(smp),
               _p(sel_idx, torch.int32), n_rows, _p(n_rows_dev, torch.int32), _p(g_enc), cfg.in_ch,
               _p(g_rays_o), _p(g_rays_d), _stream())


def _flat_zero_grads(*nets):
    """zero gradients for every tensor of the given networks as views of ONE zero-filled buffer, in parameter order
    (1 fill instead of 48, and the gradient all-reduce sees a single contiguous run: parallel.FlatGradAllReduce)."""
    total = sum(v.numel() for tensors in nets for v in tensors.values())
    dev = next(iter(nets[0].values())).device
    flat = GRAD_ALLOC(total, dev) if GRAD_ALLOC is not None else None
    if flat is None:
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
    outs, off = [], 0
    for tensors in nets:
        out, begin = {}, off
        for k, v in tensors.items():
            out[k] = flat[off:off + v.numel()].view_as(v)
            off += v.numel()
        outs.append(out)
        outs[-1]["__flat__"] = flat[begin:off]
    return outs


def select_and_cap(cfg, w_sel, w_max, B, train, cap_perm=None, seed=None):
    """Device-side compaction of the selected fine samples from the selection weights; the reference's train-only
    128-per-ray cap (model/mc_nerf.py:630-632) is reproduced when it can trigger (Sf > cfg.fine_cap = 128).
    -> (sel_idx, n_rows capacity, n_rows_dev, sel_offsets or None when the cap reshuffled the rows)"""
    dev = w_sel.device
    sel_idx, offs, n_sel = ops.select_fine(w_sel, w_max, cfg.scale, cfg.thresh)
    n_rows, n_rows_dev = B * cfg.Sf, n_sel
    if train and cfg.Sf > cfg.fine_cap:
        offs = None                                # capped rows are a random subset: no per-ray structure left
        K = B * cfg.fine_cap
        if cap_perm is not None:                   # test hook: an explicit permutation of the n selected samples
            n = int(n_sel.item())                  # (host synchronisation, as in the reference)
            if n > K:
                sel_idx = sel_idx[:n][cap_perm[:K].to(dev)].contiguous()
                n_rows, n_rows_dev = K, None
            else:
                n_rows, n_rows_dev = n, None
        elif sel_idx.shape[0] > K:
            # The reference synchronises, draws torch.randperm(n) on the CPU (~40 ms for the shipped config's 3.4 M
            # selected samples) and keeps the first K.  Same uniform K-subset without leaving the device and without a
            # sort: seeded bijective keys + a two-level radix select (mcnerf_cap_select); n <= K keeps all n.
            if seed is None:
                seed = ops.draw_seed(dev)
            sel_idx, n_rows_dev = ops.cap_select(sel_idx, n_sel, K, seed)
            n_rows = K
    return sel_idx, n_rows, n_rows_dev, offs


class RenderFn(torch.autograd.Function):
    """(rays_d, rays_o, *coarse params, *fine params) -> rgb_c, rgb_f, depth_f, opacity_f  (all [B,*])."""

    @staticmethod
    def forward(ctx, cfg, caches, train, need_grad, band_w, rng, cap_perm, rays_d, rays_o, *params):
        rays_d, rays_o = ops._f32(rays_d), ops._f32(rays_o)
        B, dev = rays_d.shape[0], rays_d.device
        nc = len(ops.param_names(cfg.coarse[0]))
        tc = {k: ops._f32(v) for k, v in zip(ops.param_names(cfg.coarse[0]), params[:nc])}
        tf = {k: ops._f32(v) for k, v in zip(ops.param_names(cfg.fine[0]), params[nc:])}
        jitter = ops._f32(rng["jitter"]).reshape(-1) if (train and rng.get("jitter") is not None) else None
        seed = rng.get("seed")                     # device RNG mode: noise is generated inside the tail kernels
        noise_c, noise_sel, noise_f = (ops._f32(rng[k]) if rng.get(k) is not None else None
                                       for k in ("noise_c", "noise_sel", "noise_f"))
        net_c, run_c, pad_c, cache_c = _tc_view(cfg, cfg.coarse, tc, caches[0])    # what the kernels see (narrow nets:
        net_f, run_f, pad_f, cache_f = _tc_view(cfg, cfg.fine, tf, caches[1])      # the 256-wide shadow)
        # coarse
        # need_grad: decided by render() - grad mode is always off inside Function.forward, and ctx.needs_input_grad
        # stays True for parameters under torch.no_grad(), which would run the stash-writing training kernels in the
        # demo / validation renders
        out_c, saved_c = _branch_fwd(cfg, net_c, run_c, rays_o, rays_d, jitter, cfg.Sc, band_w, None, B * cfg.Sc, None,
                                     need_grad, cache_c)
        fused = cfg.Sc <= 256 and cfg.Sf <= 256 and os.environ.get("MCNERF_FUSED_TAILS", "1") != "0"
        if fused:     # colour + selection weights + their maximum in one pass over out_c (tails.cu)
            rgb_c, w_sel, w_max = ops.coarse_tail_fwd(out_c, noise_c, noise_sel, seed, jitter, B, cfg.near, cfg.far,
                                                      cfg.Sc, cfg.white_back)
        else:
            if seed is not None:
                noise_c = ops.philox_fill(seed, 1, B * cfg.Sc).view(B, cfg.Sc)
                noise_sel = ops.philox_fill(seed, 2, B * cfg.Sc).view(B, cfg.Sc)
            cc = ops.make_composite_cfg(cfg.near, cfg.far, cfg.Sc, cfg.white_back)
            rgb_c = torch.empty(B, 3, device=dev)
            lib().call("mcnerf_composite_fwd", _p(out_c), _p(noise_c), _p(rays_d), _p(jitter), None, B,
                       ctypes.byref(cc), _p(rgb_c), None, None, None, _stream())
            w_max = torch.zeros(1, device=dev)
            w_sel = ops.sigma2weights(out_c, noise_sel, jitter=jitter, near=cfg.near, far=cfg.far, sigma_stride=4,
                                      n_rays=B, S=cfg.Sc, w_max=w_max)
        # selection
        sel_idx, n_rows, n_rows_dev, offs = select_and_cap(cfg, w_sel, w_max, B, train, cap_perm, seed)
        LAST["n_rows"], LAST["n_rows_dev"] = n_rows, n_rows_dev
        # fine
        if n_rows > 0:
            out_sel, saved_f = _branch_fwd(cfg, net_f, run_f, rays_o, rays_d, jitter, cfg.Sf, band_w, sel_idx,
                                           n_rows, n_rows_dev, need_grad, cache_f, offs)
        else:
            out_sel, saved_f = torch.empty(0, 4, device=dev), None
        cf = ops.make_composite_cfg(cfg.near, cfg.far, cfg.Sf, cfg.white_back)
        rgb_f = torch.empty(B, 3, device=dev)
        depth_f = torch.empty(B, 1, device=dev)
        opa_f = torch.empty(B, 1, device=dev)
        compact = fused and offs is not None and n_rows > 0
        dense = None
        if compact:   # compositing straight from the compacted rows: no dense [B,Sf,4] tensor, no scatter
            lib().call("mcnerf_fine_tail_fwd", _p(out_sel), _p(w_sel), _p(w_max), cfg.thresh, cfg.scale,
                       _p(offs, torch.int32), _p(rays_d), _p(jitter), _p(noise_f), _p(seed, torch.int64), B,
                       ctypes.byref(cf), cfg.sigma_default, _p(rgb_f), _p(depth_f), _p(opa_f), _stream())
        else:
            if seed is not None and noise_f is None:
                noise_f = ops.philox_fill(seed, 3, B * cfg.Sf).view(B, cfg.Sf)
            dense = torch.empty(B * cfg.Sf, 4, device=dev)
            lib().call("mcnerf_scatter_fine", _p(out_sel) if n_rows else None, _p(sel_idx, torch.int32) if n_rows else None,
                       n_rows, _p(n_rows_dev, torch.int32), B * cfg.Sf, cfg.sigma_default, _p(dense), _stream())
            lib().call("mcnerf_composite_fwd", _p(dense), _p(noise_f), _p(rays_d), _p(jitter), None, B,
                       ctypes.byref(cf), _p(rgb_f), _p(depth_f), _p(opa_f), None, _stream())
        ctx.cfg, ctx.band_w, ctx.n_rows = cfg, band_w, n_rows
        ctx.tc, ctx.tf = run_c, run_f
        ctx.nets, ctx.pads = (net_c, net_f), (pad_c, pad_f)
        ctx.saved = (rays_d, rays_o, jitter, noise_c, noise_f, out_c, saved_c, sel_idx, n_rows_dev, dense, saved_f,
                     out_sel)
        ctx.tail = (fused, compact, seed, w_sel, w_max, offs)
        ctx.mark_non_differentiable(depth_f, opa_f)
        return rgb_c, rgb_f, depth_f, opa_f

    @staticmethod
    def backward(ctx, g_rgb_c, g_rgb_f, _gd, _go):
        cfg, band_w, n_rows = ctx.cfg, ctx.band_w, ctx.n_rows
        rays_d, rays_o, jitter, noise_c, noise_f, out_c, saved_c, sel_idx, n_rows_dev, dense, saved_f, out_sel = ctx.saved
        B, dev = rays_d.shape[0], rays_d.device
        tc, tf = ctx.tc, ctx.tf
        (net_c, net_f), (pad_c, pad_f) = ctx.nets, ctx.pads
        gc, gf = _flat_zero_grads(tc, tf)
        flat_c, flat_f = gc.pop("__flat__"), gf.pop("__flat__")
        g_od = torch.zeros(2, B, 3, device=dev)      # one fill for both ray-gradient accumulators
        g_o, g_d = g_od[0], g_od[1]
        fused, compact, seed, w_sel, w_max, offs = ctx.tail
        if g_rgb_f is not None and n_rows > 0:
            cf = ops.make_composite_cfg(cfg.near, cfg.far, cfg.Sf, cfg.white_back)
            g_sel = torch.empty(n_rows, 4, device=dev)
            if compact:    # gradients of the compacted rows directly: no dense gradient tensor, no gather
                lib().call("mcnerf_fine_tail_bwd", _p(out_sel), _p(w_sel), _p(w_max), cfg.thresh, cfg.scale,
                           _p(offs, torch.int32), _p(jitter), _p(noise_f), _p(seed, torch.int64), B, ctypes.byref(cf),
                           cfg.sigma_default, _p(ops._f32(g_rgb_f)), _p(g_sel), _stream())
            else:
                g_dense = torch.empty_like(dense)
                lib().call("mcnerf_composite_bwd", _p(dense), _p(noise_f), _p(jitter), None, B, ctypes.byref(cf),
                           _p(ops._f32(g_rgb_f)), _p(g_dense), _stream())
                lib().call("mcnerf_gather_fine", _p(g_dense), _p(sel_idx, torch.int32), n_rows,
                           _p(n_rows_dev, torch.int32), _p(g_sel), _stream())
            _branch_bwd(cfg, net_f, tf, gf, rays_o, rays_d, jitter, cfg.Sf, band_w, sel_idx, n_rows, n_rows_dev,
                        saved_f, out_sel, g_sel, g_o, g_d)
        if GRAD_HOOK is not None and pad_f is None:
            GRAD_HOOK("fine", flat_f)          # fine gradients are final: their all-reduce overlaps the coarse backward
        if g_rgb_c is not None:
            cc = ops.make_composite_cfg(cfg.near, cfg.far, cfg.Sc, cfg.white_back)
            g_out_c = torch.empty_like(out_c)
            if fused:
                lib().call("mcnerf_coarse_tail_bwd", _p(out_c), _p(noise_c), _p(seed, torch.int64), _p(jitter), B,
                           ctypes.byref(cc), _p(ops._f32(g_rgb_c)), _p(g_out_c), _stream())
            else:
                if seed is not None and noise_c is None:
                    noise_c = ops.philox_fill(seed, 1, B * cfg.Sc).view(B, cfg.Sc)
                lib().call("mcnerf_composite_bwd", _p(out_c), _p(noise_c), _p(jitter), None, B, ctypes.byref(cc),
                           _p(ops._f32(g_rgb_c)), _p(g_out_c), _stream())
            _branch_bwd(cfg, net_c, tc, gc, rays_o, rays_d, jitter, cfg.Sc, band_w, None, B * cfg.Sc, None,
                        saved_c, out_c, g_out_c, g_o, g_d)
        if GRAD_HOOK is not None and pad_c is None:
            GRAD_HOOK("coarse", flat_c)        # overlaps the camera-model backward that follows
        if pad_c is not None:
            gc = pad_c.unpad(gc)
        if pad_f is not None:
            gf = pad_f.unpad(gf)
        pg = [gc[k] for k in ops.param_names(cfg.coarse[0])] + [gf[k] for k in ops.param_names(cfg.fine[0])]
        return (None, None, None, None, None, None, None, g_d, g_o) + tuple(pg)


def draw_rng(cfg, B, device, train, seed=None):
    """The random inputs of one render.  cfg.device_rng: two key words from torch's generator (or `seed`) - jitter
    from Philox stream 4, the density noise generated inside the tail kernels.  Otherwise the reference's draws in the
    reference's order from torch's current generator:
    uniform_[B,1] (train only) -> randn[B,Sc] -> randn[B,Sc] -> randn[B,Sf]."""
    rng = {}
    if cfg.device_rng and torch.device(device).type == "cuda":
        rng["seed"] = seed if seed is not None else ops.draw_seed(device)
        if train:
            rng["jitter"] = ops.philox_fill(rng["seed"], 4, B, normal=False, lo=0.0, hi=(cfg.far - cfg.near) / cfg.Sc)
        return rng
    if train:
        rng["jitter"] = torch.empty(B, 1, device=device).uniform_(0.0, (cfg.far - cfg.near) / cfg.Sc)
    rng["noise_c"] = torch.randn((B, cfg.Sc), device=device)
    rng["noise_sel"] = torch.randn((B, cfg.Sc), device=device)
    rng["noise_f"] = torch.randn((B, cfg.Sf), device=device)
    return rng


def render(cfg, params_c, params_f, rays_d, rays_o, train, band_w=None, rng=None, cap_perm=None):
    """params_*: dicts of the reference's state_dict names -> tensors (leaf Parameters are fine)."""
    if rng is None:
        rng = draw_rng(cfg, rays_d.shape[0], rays_d.device, train)
    plist = [params_c[k] for k in ops.param_names(cfg.coarse[0])] + [params_f[k] for k in ops.param_names(cfg.fine[0])]
    need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in [rays_d, rays_o] + plist)
    caches = (_owner_cache(params_c), _owner_cache(params_f))
    return RenderFn.apply(cfg, caches, train, need_grad, band_w, rng, cap_perm, rays_d, rays_o, *plist)

"""PyTorch-facing wrappers over the C ABI: thin functions that pass raw device pointers + the current
stream to libmcnerf.so, and torch.autograd.Functions at the granularity of the reference's modules.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic of the hot path
happens in the CUDA library.  Nothing in this file computes on the CPU or falls back to torch ops.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import CompositeCfg, Dirs, MlpGrads, MlpParams, Sampling, TcInput, lib


def _stream():
    """current CUDA stream of the current device as a raw cudaStream_t (fast path: no Stream object is built)."""
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:      # private helpers moved: fall back to the public API
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.McnerfError("libmcnerf kernels need CUDA tensors (no CPU fallback)")
    if t.device.index != torch._C._cuda_getDevice():
        # launches go to the CURRENT device's current stream: a tensor of another GPU would be dereferenced there
        raise _lib.McnerfError(f"tensor on cuda:{t.device.index} but the current device is cuda:{torch._C._cuda_getDevice()}: "
                               "wrap the call in torch.cuda.device(...)")
    if t.dtype != dtype:
        raise _lib.McnerfError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.McnerfError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def _c(t):
    return t.contiguous() if t is not None else None


def _f32(t):
    if t.dtype == torch.float32 and t.is_contiguous():
        return t.detach()
    return t.detach().to(torch.float32).contiguous()


def make_sampling(near, far, S, n_freqs, band_w=None):
    """band_w: None (all ones), a sequence of floats (baked into the struct), or a float32 CUDA tensor of >= n_freqs
    weights that the kernels read at run time (CUDA-graph replays with a moving BARF window)."""
    s = Sampling()
    s.near_, s.far_, s.S, s.n_freqs = float(near), float(far), int(S), int(n_freqs)
    dynamic = torch.is_tensor(band_w)
    for k in range(_lib.MAX_FREQS):
        s.band_w[k] = float(band_w[k]) if (band_w is not None and not dynamic and k < n_freqs) else 1.0
    if dynamic:
        if band_w.numel() < n_freqs:
            raise _lib.McnerfError("band weight tensor shorter than n_freqs")
        s.band_w_dev = _p(band_w).value
        s._keep = band_w
    return s


def make_composite_cfg(near, far, S, white_back):
    c = CompositeCfg()
    c.near_, c.far_, c.S, c.white_back = float(near), float(far), int(S), int(bool(white_back))
    return c


def store_floats(dst, values):
    """dst[:len(values)] = values (<= 16 Python floats), stream-ordered, values passed by value in the launch"""
    n = len(values)
    lib().call("mcnerf_store_floats", _p(dst), (ctypes.c_float * n)(*[float(v) for v in values]), n, _stream())


def barf_band_weights(step_r, barf_start, barf_end, n_freqs):
    """Host-side evaluation of the 10 BARF window weights (ref: model/net_block.py:26-29).
    Returned as Python floats computed in fp32 the way torch does it."""
    alpha = (step_r - barf_start) / (barf_end - barf_start) * n_freqs
    k = torch.arange(n_freqs, dtype=torch.float32)
    w = (1 - (alpha - k).clamp_(min=0, max=1).mul_(torch.pi).cos_()) / 2
    return [float(x) for x in w]

# --------------------------------------------------------------------------- loss + pixel choice


class TrainLossFn(torch.autograd.Function):
    """Total loss of a rendering stage in ONE launch: normalised reprojection term + coarse/fine MSE, with the
    gradients produced by the same kernel.  ref: model/loss.py:15-58."""

    @staticmethod
    def forward(ctx, rgb_c, rgb_f, gt, px, px_gt, img_w, img_h, normalise):
        rc, g = _f32(rgb_c), _f32(gt)
        rf = _f32(rgb_f) if rgb_f is not None else None
        pxc = _f32(px) if px is not None else None
        pgt = _f32(px_gt).to(rc.device) if px is not None else None
        n, n_pts = rc.shape[0], (pxc.numel() // 2 if pxc is not None else 0)
        if rc.shape != g.shape or (rf is not None and rf.shape != g.shape) or rc.dim() != 2 or rc.shape[1] != 3:
            raise _lib.McnerfError(f"TrainLossFn: renders {tuple(rc.shape)} / ground truth {tuple(g.shape)} mismatch")
        out = torch.empty(3, device=rc.device)
        grads = torch.empty(6 * n + 2 * n_pts, device=rc.device)        # g_c | g_f | g_px in one buffer
        g_c, g_f, g_px = grads[:3 * n], grads[3 * n:6 * n], grads[6 * n:]
        lib().call("mcnerf_train_loss", _p(rc), _p(rf), _p(g), n, _p(pxc), _p(pgt), n_pts, int(img_w), int(img_h),
                   int(bool(normalise)), _p(out), _p(g_c), _p(g_f) if rf is not None else None,
                   _p(g_px) if pxc is not None else None, _stream())
        ctx.save_for_backward(grads)
        ctx.meta = (n, rgb_f is not None, tuple(px.shape) if px is not None else None)
        ctx.parts = out                       # (total, raw reprojection loss, rgb loss) for logging
        return out[0]

    @staticmethod
    def backward(ctx, go):
        (grads,) = ctx.saved_tensors
        n, has_f, px_shape = ctx.meta
        g = grads * go                        # one launch for all three gradients
        return (g[:3 * n].view(n, 3), g[3 * n:6 * n].view(n, 3) if has_f else None, None,
                g[6 * n:].view(px_shape) if px_shape is not None else None, None, None, None, None)


def sample_pixels_workspace(n, batch, device):
    """zeroed workspace for sample_pixels, or None when the batch is too large for the one-block sort."""
    sz = ctypes.c_size_t()
    if lib().cdll.mcnerf_sample_pixels_workspace(int(n), int(batch), ctypes.byref(sz)) != 0:
        return None
    return torch.zeros(sz.value, dtype=torch.uint8, device=device)


def sample_pixels(n, batch, seed, workspace):
    """first min(batch, n) entries of a random permutation of [0, n): (int64 [m], int32 [m]).
    ref: model/mc_nerf.py:327-345."""
    m = min(int(n), int(batch))
    idx64 = torch.empty(m, dtype=torch.int64, device=seed.device)
    idx32 = torch.empty(m, dtype=torch.int32, device=seed.device)
    lib().call("mcnerf_sample_pixels", int(n), int(batch), _p(seed, torch.int64), _p(workspace, torch.uint8),
               _p(idx64, torch.int64), _p(idx32, torch.int32), _stream())
    return idx64, idx32

# --------------------------------------------------------------------------- camera


class IntrinsicsFn(torch.autograd.Function):
    """(w_fx, w_fy, w_ux, w_uy) -> K [n,3,3], Kinv [n,3,3].  ref: model/mc_nerf.py:171-186, 204-210."""

    @staticmethod
    def forward(ctx, w_fx, w_fy, w_ux, w_uy, img_h, img_w):
        ws = [_f32(w) for w in (w_fx, w_fy, w_ux, w_uy)]
        n = ws[0].shape[0]
        K = torch.empty(n, 3, 3, device=ws[0].device)
        Kinv = torch.empty_like(K)
        lib().call("mcnerf_intrinsics_fwd", *[_p(w) for w in ws], n, img_h, img_w, _p(K), _p(Kinv), _stream())
        ctx.save_for_backward(*ws)
        ctx.hw = (img_h, img_w)
        return K, Kinv

    @staticmethod
    def backward(ctx, gK, gKinv):
        ws = ctx.saved_tensors
        n = ws[0].shape[0]
        gs = [torch.empty_like(ws[0]) for _ in range(4)]
        gK = _f32(gK) if gK is not None else None
        gKinv = _f32(gKinv) if gKinv is not None else None
        lib().call("mcnerf_intrinsics_bwd", *[_p(w) for w in ws], n, ctx.hw[0], ctx.hw[1], _p(gK), _p(gKinv),
                   *[_p(g) for g in gs], _stream())
        return gs[0], gs[1], gs[2], gs[3], None, None


class SE3Fn(torch.autograd.Function):
    """se(3) twist [n,6] -> world->camera [n,3,4].  ref: model/mc_nerf.py:269-316."""

    @staticmethod
    def forward(ctx, wu):
        wu = _f32(wu)
        shape = wu.shape[:-1]
        flat = wu.reshape(-1, 6)
        Rt = torch.empty(flat.shape[0], 3, 4, device=wu.device)
        lib().call("mcnerf_se3_fwd", _p(flat), flat.shape[0], _p(Rt), _stream())
        ctx.save_for_backward(flat)
        ctx.shape = shape
        return Rt.reshape(*shape, 3, 4)

    @staticmethod
    def backward(ctx, gRt):
        (flat,) = ctx.saved_tensors
        g = torch.empty_like(flat)
        lib().call("mcnerf_se3_bwd", _p(flat), _p(_f32(gRt).reshape(-1, 3, 4)), flat.shape[0], _p(g), _stream())
        return g.reshape(*ctx.shape, 6)


class ReprojectFn(torch.autograd.Function):
    """Calibration-point reprojection: wpts [1,n,P,3], K [n,3,3], Rt [n,3,4] -> pixels [1,n,P,2].
    ref: model/mc_nerf.py:147-152.  One launch each way instead of ~30 small ATen kernels."""

    @staticmethod
    def forward(ctx, wpts, K, Rt):
        wpts, K, Rt = _f32(wpts), _f32(K), _f32(Rt)
        n, P = K.shape[0], wpts.shape[-2]
        pix = torch.empty(*wpts.shape[:-1], 2, device=K.device)
        lib().call("mcnerf_reproject_fwd", _p(wpts), _p(K), _p(Rt), n, P, _p(pix), _stream())
        ctx.save_for_backward(wpts, K, Rt)
        return pix

    @staticmethod
    def backward(ctx, g):
        wpts, K, Rt = ctx.saved_tensors
        gK, gRt = torch.empty_like(K), torch.empty_like(Rt)
        lib().call("mcnerf_reproject_bwd", _p(wpts), _p(K), _p(Rt), _p(_f32(g)), K.shape[0], wpts.shape[-2],
                   _p(gK), _p(gRt), _stream())
        return None, gK, gRt


class CameraTrainFn(torch.autograd.Function):
    """The camera model of one train step in one launch each way: (w_fx, w_fy, w_ux, w_uy, w_pose, w_pose_intr, wpts)
    -> K, Kinv [n,3,3], pose, calib_pose [n,3,4], reprojected calibration pixels [1,n,P,2].
    ref: model/mc_nerf.py:75-76 / 87-88 (add_weights2param + get_reproject_pixels).  Gradients: Kinv and pose from the
    rays, the pixels from the loss; K and calib_pose are consumed by the reprojection inside."""

    @staticmethod
    def forward(ctx, w_fx, w_fy, w_ux, w_uy, w_pose, w_calib, wpts, img_h, img_w):
        ws = [_f32(w) for w in (w_fx, w_fy, w_ux, w_uy, w_pose, w_calib)]
        wp = _f32(wpts)
        n, P, dev = ws[0].shape[0], wp.shape[-2], ws[0].device
        K, Kinv = torch.empty(n, 3, 3, device=dev), torch.empty(n, 3, 3, device=dev)
        pose, calib = torch.empty(n, 3, 4, device=dev), torch.empty(n, 3, 4, device=dev)
        pix = torch.empty(*wp.shape[:-1], 2, device=dev)
        lib().call("mcnerf_camera_fwd", *[_p(w) for w in ws], _p(wp), n, P, int(img_h), int(img_w), _p(K), _p(Kinv),
                   _p(pose), _p(calib), _p(pix), _stream())
        ctx.save_for_backward(*ws, wp, K, calib)
        ctx.hw = (int(img_h), int(img_w))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(K, calib)      # consumed inside (reprojection); exposed for reporting only
        return K, Kinv, pose, calib, pix

    @staticmethod
    def backward(ctx, _gK, gKinv, gpose, _gcalib, gpix):
        *ws, wp, K, calib = ctx.saved_tensors
        n, P, dev = ws[0].shape[0], wp.shape[-2], ws[0].device
        need_pose = ctx.needs_input_grad[4] and gpose is not None
        out = torch.empty(4 * n + 12 * n + 21 * n, device=dev)          # g_fx|g_fy|g_ux|g_uy | g_pose | g_calib | scratch
        gf = [out[k * n:(k + 1) * n] for k in range(4)]
        g_wpose, g_wcalib = out[4 * n:10 * n].view(n, 6), out[10 * n:16 * n].view(n, 6)
        scratch = out[16 * n:]
        if gpix is None:
            gpix = torch.zeros(n * P * 2, device=dev)
        lib().call("mcnerf_camera_bwd", *[_p(w) for w in ws], _p(wp), _p(K), _p(calib), n, P, ctx.hw[0], ctx.hw[1],
                   _p(_f32(gKinv)) if gKinv is not None else None, _p(_f32(gpose)) if need_pose else None, _p(_f32(gpix)),
                   _p(scratch), *[_p(g) for g in gf], _p(g_wpose) if need_pose else None, _p(g_wcalib), _stream())
        return gf[0], gf[1], gf[2], gf[3], (g_wpose if need_pose else None), g_wcalib, None, None, None


class RaygenFn(torch.autograd.Function):
    """(Kinv [n,3,3], Rt [n,3,4], cam, pix) -> rays_o, rays_d [B,3].  ref: model/mc_nerf.py:124-145, 327-345.
    `cam` is an int or an int32 tensor [B]; `pix` is None (whole image, row-major) or an int32 tensor [B]."""

    @staticmethod
    def forward(ctx, Kinv, Rt, cam, pix, n_rays, img_w):
        Kinv, Rt = _f32(Kinv), _f32(Rt)
        cam_t = cam if torch.is_tensor(cam) else None
        cam_c = 0 if cam_t is not None else int(cam)
        ro = torch.empty(n_rays, 3, device=Kinv.device)
        rd = torch.empty_like(ro)
        lib().call("mcnerf_raygen_fwd", _p(Kinv), _p(Rt), _p(cam_t, torch.int32), cam_c, _p(pix, torch.int32),
                   n_rays, img_w, _p(ro), _p(rd), _stream())
        ctx.save_for_backward(Kinv, Rt, cam_t, pix)
        ctx.meta = (cam_c, n_rays, img_w)
        return ro, rd

    @staticmethod
    def backward(ctx, g_o, g_d):
        Kinv, Rt, cam_t, pix = ctx.saved_tensors
        cam_c, n_rays, img_w = ctx.meta
        acc = torch.zeros(Kinv.numel() + Rt.numel(), device=Kinv.device)      # one fill for both accumulators
        gK, gRt = acc[:Kinv.numel()].view_as(Kinv), acc[Kinv.numel():].view_as(Rt)
        g_o = _f32(g_o) if g_o is not None else torch.zeros(n_rays, 3, device=Kinv.device)
        g_d = _f32(g_d) if g_d is not None else torch.zeros(n_rays, 3, device=Kinv.device)
        lib().call("mcnerf_raygen_bwd", _p(Kinv), _p(Rt), _p(cam_t, torch.int32), cam_c, _p(pix, torch.int32),
                   n_rays, img_w, _p(g_o), _p(g_d), _p(gK), _p(gRt), _stream())
        return gK, gRt, None, None, None, None

# --------------------------------------------------------------------------- encoding


class EncodePointsFn(torch.autograd.Function):
    """SinCosEmbedding.forward on explicit points.  ref: model/net_block.py:20-35."""

    @staticmethod
    def forward(ctx, x, n_freqs, band_w):
        x = _f32(x)
        n = x.shape[0]
        enc = torch.empty(n, 3 + 6 * n_freqs, device=x.device)
        bw = (ctypes.c_float * n_freqs)(*band_w) if band_w is not None else None
        lib().call("mcnerf_encode_points_fwd", _p(x), n, n_freqs, ctypes.cast(bw, ctypes.c_void_p) if bw else None,
                   _p(enc), enc.shape[1], _stream())
        ctx.save_for_backward(x)
        ctx.meta = (n_freqs, band_w)
        return enc

    @staticmethod
    def backward(ctx, g_enc):
        (x,) = ctx.saved_tensors
        n_freqs, band_w = ctx.meta
        g_enc = _f32(g_enc)
        gx = torch.empty_like(x)
        bw = (ctypes.c_float * n_freqs)(*band_w) if band_w is not None else None
        lib().call("mcnerf_encode_points_bwd", _p(x), x.shape[0], n_freqs,
                   ctypes.cast(bw, ctypes.c_void_p) if bw else None, _p(g_enc), g_enc.shape[1], _p(gx), _stream())
        return gx, None, None

# --------------------------------------------------------------------------- MLP

TRUNK = "xyz_encoding_{}.0.{}"
HEAD_KEYS = [("W_sigma0", "sigma.0.weight"), ("b_sigma0", "sigma.0.bias"), ("W_sigma2", "sigma.2.weight"),
             ("b_sigma2", "sigma.2.bias"), ("W_sh0", "sh.0.weight"), ("b_sh0", "sh.0.bias"),
             ("W_sh2", "sh.2.weight"), ("b_sh2", "sh.2.bias")]


def param_names(depth):
    """state_dict key order of CorseFine_NeRF (ref: model/net_block.py:51-65)."""
    names = []
    for i in range(depth):
        names += [TRUNK.format(i + 1, "weight"), TRUNK.format(i + 1, "bias")]
    return names + [k for _, k in HEAD_KEYS]


def fill_mlp_struct(struct, tensors, depth):
    """tensors: dict name -> contiguous fp32 CUDA tensor."""
    for i in range(depth):
        struct.W[i] = tensors[TRUNK.format(i + 1, "weight")].data_ptr()
        struct.b[i] = tensors[TRUNK.format(i + 1, "bias")].data_ptr()
    for field, key in HEAD_KEYS:
        setattr(struct, field, tensors[key].data_ptr())
    return struct


def make_mlp_params(tensors, depth, width, skips, in_ch=63, sh_dim=None):
    p = MlpParams()
    if sh_dim is None:       # 3 (MLP_deg + 1)^2 output rows of the colour head (ref: model/net_block.py:63-65)
        sh_dim = int(tensors["sh.2.bias"].numel())
    p.depth, p.width, p.in_ch, p.sh_dim = depth, width, in_ch, sh_dim
    p.skip_mask = sum(1 << int(s) for s in skips if 0 < int(s) < depth)
    return fill_mlp_struct(p, tensors, depth)


def make_dirs(dirs, dir_idx=None, dir_S=0):
    d = Dirs()
    d.dirs = dirs.data_ptr()
    d.dir_idx = dir_idx.data_ptr() if dir_idx is not None else None
    d.dir_S = int(dir_S)
    return d


def mlp_f32_workspace(params_struct, n_rows, device):
    nbytes = lib().cdll.mcnerf_mlp_f32_workspace(ctypes.byref(params_struct), int(n_rows))
    return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)


class MLPFn(torch.autograd.Function):
    """CorseFine_NeRF.forward(x_enc [M,63], dirs [M,3]) -> [M,4].  ref: model/net_block.py:67-78.
    Differentiable in x_enc, dirs and every parameter (passed flat in param_names(depth) order)."""

    @staticmethod
    def forward(ctx, x_enc, dirs, depth, width, skips, *params):
        x_enc, dirs = _f32(x_enc), _f32(dirs)
        names = param_names(depth)
        tensors = {k: _f32(v) for k, v in zip(names, params)}
        M = x_enc.shape[0]
        ps = make_mlp_params(tensors, depth, width, skips, in_ch=x_enc.shape[1])
        out4 = torch.empty(M, 4, device=x_enc.device)
        ws = mlp_f32_workspace(ps, max(M, 1), x_enc.device)
        d = make_dirs(dirs)
        if M > 0:
            lib().call("mcnerf_mlp_f32_fwd", ctypes.byref(ps), _p(x_enc), x_enc.shape[1], ctypes.byref(d), M, None,
                       _p(out4), _p(ws), _stream())
        ctx.save_for_backward(x_enc, dirs, ws, *[tensors[k] for k in names])
        ctx.meta = (depth, width, tuple(skips))
        return out4

    @staticmethod
    def backward(ctx, g_out):
        x_enc, dirs, ws, *plist = ctx.saved_tensors
        depth, width, skips = ctx.meta
        names = param_names(depth)
        tensors = dict(zip(names, plist))
        ps = make_mlp_params(tensors, depth, width, skips, in_ch=x_enc.shape[1])
        grads = {k: torch.zeros_like(v) for k, v in tensors.items()}
        gs = fill_mlp_struct(MlpGrads(), grads, depth)
        M = x_enc.shape[0]
        g_x = torch.empty_like(x_enc)
        g_d = torch.zeros_like(dirs)
        d = make_dirs(dirs)
        if M > 0:
            lib().call("mcnerf_mlp_f32_bwd", ctypes.byref(ps), _p(x_enc), x_enc.shape[1], ctypes.byref(d), M, None,
                       _p(_f32(g_out)), _p(ws), ctypes.byref(gs), _p(g_x), _p(g_d), _stream())
        return (g_x, g_d, None, None, None) + tuple(grads[k] for k in names)


class EvalSHFn(torch.autograd.Function):
    """eval_sh(deg): sh [n,3,(deg+1)^2], dirs [n,3] -> [n,3], deg 0..4.  ref: model/net_utils.py:103-191."""

    @staticmethod
    def forward(ctx, sh, dirs, deg=2):
        sh, dirs = _f32(sh), _f32(dirs)
        n = dirs.shape[0]
        out = torch.empty(n, 3, device=sh.device)
        lib().call("mcnerf_eval_sh_deg_fwd", int(deg), _p(sh), _p(dirs), n, _p(out), _stream())
        ctx.save_for_backward(sh, dirs)
        ctx.deg = int(deg)
        return out

    @staticmethod
    def backward(ctx, g):
        sh, dirs = ctx.saved_tensors
        g_sh, g_d = torch.empty_like(sh), torch.empty_like(dirs)
        lib().call("mcnerf_eval_sh_deg_bwd", ctx.deg, _p(sh), _p(dirs), _p(_f32(g)), dirs.shape[0], _p(g_sh), _p(g_d),
                   _stream())
        return g_sh, g_d, None

# --------------------------------------------------------------------------- compositing / selection


class CompositeFn(torch.autograd.Function):
    """Tail of NeRF_Model.inference: out4 [B,S,4] -> rgb [B,3], depth [B,1], opacity [B,1].
    ref: model/mc_nerf.py:705-727.  Gradient flows from rgb into out4 only (the reference discards
    depth/opacity in training, model/mc_nerf.py:590-591)."""

    @staticmethod
    def forward(ctx, out4, noise, rays_d, z_vals, jitter, near, far, white_back):
        out4 = _f32(out4)
        B, S = out4.shape[0], out4.shape[1]
        cfg = make_composite_cfg(near, far, S, white_back)
        noise = _f32(noise) if noise is not None else None
        z_vals = _f32(z_vals) if z_vals is not None else None
        jitter = _f32(jitter).reshape(-1) if jitter is not None else None
        rgb = torch.empty(B, 3, device=out4.device)
        depth = torch.empty(B, 1, device=out4.device)
        opacity = torch.empty(B, 1, device=out4.device)
        lib().call("mcnerf_composite_fwd", _p(out4), _p(noise), _p(_f32(rays_d)), _p(jitter), _p(z_vals), B,
                   ctypes.byref(cfg), _p(rgb), _p(depth), _p(opacity), None, _stream())
        ctx.save_for_backward(out4, noise, z_vals, jitter)
        ctx.cfg = (near, far, S, white_back)
        ctx.mark_non_differentiable(depth, opacity)
        return rgb, depth, opacity

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_opacity):
        out4, noise, z_vals, jitter = ctx.saved_tensors
        near, far, S, white_back = ctx.cfg
        cfg = make_composite_cfg(near, far, S, white_back)
        g_out4 = torch.empty_like(out4)
        lib().call("mcnerf_composite_bwd", _p(out4), _p(noise), _p(jitter), _p(z_vals), out4.shape[0],
                   ctypes.byref(cfg), _p(_f32(g_rgb)), _p(g_out4), _stream())
        return g_out4, None, None, None, None, None, None, None


def sigma2weights(sigmas, noise, deltas=None, jitter=None, near=0.0, far=1.0, sigma_stride=1, n_rays=None, S=None,
                  w_max=None):
    """ref: model/mc_nerf.py:729-736.  Forward only (the reference uses its gradient only through the
    compositing above, which CompositeFn covers)."""
    if n_rays is None:
        n_rays, S = sigmas.shape[0], sigmas.shape[1]
    cfg = make_composite_cfg(near, far, S, True)
    w = torch.empty(n_rays, S, device=sigmas.device)
    lib().call("mcnerf_sigma2weights", _p(sigmas), sigma_stride, _p(noise), _p(jitter), None, _p(deltas), n_rays,
               ctypes.byref(cfg), _p(w), _p(w_max), _stream())
    return w


def philox_fill(seed, stream_id, n, normal=True, lo=0.0, hi=1.0):
    """n draws of Philox stream `stream_id` keyed by `seed` (int64 [2] on the device): N(0,1) or uniform(lo, hi)."""
    out = torch.empty(n, device=seed.device)
    lib().call("mcnerf_philox_fill", _p(seed, torch.int64), int(stream_id), int(n), int(bool(normal)), float(lo), float(hi),
               _p(out), _stream())
    return out


def draw_seed(device):
    """two int64 key words from torch's CUDA generator (torch.manual_seed governs every device-side draw; graph-safe)"""
    return torch.randint(-(1 << 62), 1 << 62, (2,), device=device, dtype=torch.int64)


def coarse_tail_fwd(out_c, noise_rgb, noise_sel, seed, jitter, B, near, far, S, white_back):
    """-> rgb [B,3], w_sel [B,S], w_max [1].  ref: model/mc_nerf.py:613-621, 719-727."""
    cfg = make_composite_cfg(near, far, S, white_back)
    rgb = torch.empty(B, 3, device=out_c.device)
    w_sel = torch.empty(B, S, device=out_c.device)
    w_max = torch.zeros(1, device=out_c.device)
    lib().call("mcnerf_coarse_tail_fwd", _p(out_c), _p(noise_rgb), _p(noise_sel), _p(seed, torch.int64), _p(jitter), B,
               ctypes.byref(cfg), _p(rgb), _p(w_sel), _p(w_max), _stream())
    return rgb, w_sel, w_max


def select_fine(weights, w_max, scale, thresh):
    """-> (sel_idx int32 [B*Sc*scale] capacity, sel_offsets int32 [B+1], n_sel int32 [1]); no host sync."""
    B, Sc = weights.shape
    dev = weights.device
    sel_idx = torch.empty(B * Sc * scale, dtype=torch.int32, device=dev)
    offs = torch.empty(B + 1, dtype=torch.int32, device=dev)
    n_sel = torch.empty(1, dtype=torch.int32, device=dev)
    lib().call("mcnerf_select_fine", _p(weights), _p(w_max), B, Sc, scale, float(thresh), _p(sel_idx, torch.int32),
               _p(offs, torch.int32), _p(n_sel, torch.int32), _stream())
    return sel_idx, offs, n_sel


def cap_select(sel_idx, n_sel, K, seed):
    """-> (out_idx int32 [K], n_out int32 [1]): a uniformly random K-subset of the first n_sel entries of sel_idx (all of
    them when n_sel <= K), ascending; no sort, no host sync.  ref: model/mc_nerf.py:630-632."""
    dev = sel_idx.device
    cap = int(sel_idx.shape[0])
    sz = ctypes.c_size_t()
    lib().call("mcnerf_cap_select_workspace", cap, ctypes.byref(sz))
    ws = torch.empty(sz.value, dtype=torch.uint8, device=dev)
    out = torch.empty(int(K), dtype=torch.int32, device=dev)
    n_out = torch.empty(1, dtype=torch.int32, device=dev)
    lib().call("mcnerf_cap_select", _p(sel_idx, torch.int32), _p(n_sel, torch.int32), cap, int(K), _p(seed, torch.int64),
               _p(out, torch.int32), _p(n_out, torch.int32), _p(ws, torch.uint8), _stream())
    return out, n_out


class ScatterFineFn(torch.autograd.Function):
    """out_dense[B*Sf,4] = defaults; out_dense[idx] = out_sel.  ref: model/mc_nerf.py:692-694, 700-701."""

    @staticmethod
    def forward(ctx, out_sel, idx, n_dense, sigma_default):
        out_sel = _f32(out_sel)
        dense = torch.empty(n_dense, 4, device=idx.device)
        n = out_sel.shape[0]
        lib().call("mcnerf_scatter_fine", _p(out_sel) if n else None, _p(idx, torch.int32) if n else None, n, None,
                   n_dense, float(sigma_default), _p(dense), _stream())
        ctx.save_for_backward(idx)
        ctx.n = n
        return dense

    @staticmethod
    def backward(ctx, g_dense):
        (idx,) = ctx.saved_tensors
        g_sel = torch.empty(ctx.n, 4, device=g_dense.device)
        lib().call("mcnerf_gather_fine", _p(_f32(g_dense)), _p(idx, torch.int32), ctx.n, None, _p(g_sel), _stream())
        return g_sel, None, None, None


# --------------------------------------------------------------------------- bf16 tcgen05 MLP path


TC_WIDTH = 256        # hidden width the tcgen05 kernels are built for


class PaddedNet:
    """256-wide zero-padded fp32 shadow of a network narrower than 256 (e.g. the reference's default 4x128 coarse net).
    Zero rows/columns change nothing in the arithmetic (a padded unit has pre-activation 0, output 0, gradient 0), so
    the tensor-core kernels run the shadow and the valid blocks of its gradients are the network's gradients.
    One launch copies every parameter into its shadow, one launch cuts every gradient back out."""

    def __init__(self, tensors, depth, width, in_ch):
        self.names = param_names(depth)
        dev = next(iter(tensors.values())).device
        wide_shapes = {}
        for name in self.names:
            shape = list(tensors[name].shape)
            head_out = name in ("sigma.2.weight", "sigma.2.bias", "sh.2.weight", "sh.2.bias")
            if not head_out:
                shape[0] = TC_WIDTH                                  # output features (weights and biases)
            if len(shape) == 2 and name != TRUNK.format(1, "weight"):   # input features: hidden, or encoding + hidden
                shape[1] = TC_WIDTH + (shape[1] - width)
            wide_shapes[name] = tuple(shape)
        total = sum(int(torch.Size(sh).numel()) for sh in wide_shapes.values())
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.wide, off = {}, 0
        for name in self.names:
            n = int(torch.Size(wide_shapes[name]).numel())
            self.wide[name] = self.flat[off:off + n].view(wide_shapes[name])
            off += n
        self.narrow_shapes = {name: tuple(tensors[name].shape) for name in self.names}
        self.key = None
        self.cache = {}               # derived data of the shadow itself (its packed bf16 images)

    def _copy(self, src, dst, shapes):
        n = len(self.names)
        vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
        rows = [shapes[k][0] if len(shapes[k]) == 2 else 1 for k in self.names]
        cols = [shapes[k][-1] for k in self.names]
        lib().call("mcnerf_copy_blocks", n, vp(*[src[k].data_ptr() for k in self.names]),
                   vp(*[dst[k].data_ptr() for k in self.names]), ip(*rows), ip(*cols),
                   ip(*[src[k].shape[-1] for k in self.names]), ip(*[dst[k].shape[-1] for k in self.names]), _stream())

    def refresh(self, tensors):
        """bring the shadow up to date with the parameters (after optimizer.step(); always under graph capture)"""
        key = tuple((t.data_ptr(), t._version) for t in tensors.values())
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing or key != self.key:
            self._copy(tensors, self.wide, self.narrow_shapes)
            torch.autograd.graph.increment_version(list(self.wide.values()))     # the packed images are now stale
            self.key = None if capturing else key
        return self.wide

    def unpad(self, wide_grads):
        """valid blocks of the 256-wide gradients as views of one contiguous buffer in parameter order"""
        total = sum(int(torch.Size(sh).numel()) for sh in self.narrow_shapes.values())
        flat = torch.empty(total, dtype=torch.float32, device=self.flat.device)
        out, off = {}, 0
        for name in self.names:
            n = int(torch.Size(self.narrow_shapes[name]).numel())
            out[name] = flat[off:off + n].view(self.narrow_shapes[name])
            off += n
        self._copy(wide_grads, out, self.narrow_shapes)
        return out


def tc_supported(ps):
    return bool(lib().cdll.mcnerf_mlp_tc_supported(ctypes.byref(ps)))


class TcWeights:
    """UMMA-ready bf16 weight images + fp32 bias block of one network (derived cache of the fp32 parameters;
    re-packed when any parameter's version counter changes, i.e. after optimizer.step())."""

    def __init__(self):
        self.key = None
        self.wf = self.wb = self.bias = None
        self.ready = None            # event of a pack issued ahead of time on a side stream (prefetch)

    def prefetch(self, ps, tensors, need_bwd, side):
        """Issue the pack on `side` (forked from the current stream) so that it overlaps the step's camera / pixel /
        RNG launches; the next get() on the main stream waits for it instead of packing."""
        main = torch.cuda.current_stream()
        side.wait_stream(main)        # everything issued so far (last backward reads wb, RAdam writes the weights)
        with torch.cuda.stream(side):
            self.get(ps, tensors, need_bwd)
            self.ready = (side.record_event(), tuple((t.data_ptr(), t._version) for t in tensors.values()))

    def get(self, ps, tensors, need_bwd=True):
        key = tuple((t.data_ptr(), t._version) for t in tensors.values())
        if self.ready is not None:
            event, packed_key = self.ready
            self.ready = None
            torch.cuda.current_stream().wait_event(event)
            if packed_key == key and (not need_bwd or self.wb is not None):
                return self
        # Under CUDA-graph capture the host-side version check would be frozen into the graph (a replay after
        # optimizer.step() would read stale images): always record the pack launch, and forget the key so that the
        # next eager call re-packs too.
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            key = None
        if capturing or key != self.key or (need_bwd and self.wb is None):
            dev = next(iter(tensors.values())).device
            sz = [ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()]
            lib().call("mcnerf_mlp_tc_pack_sizes", ctypes.byref(ps), *[ctypes.byref(x) for x in sz])
            if self.wf is None or self.wf.device != dev:
                self.wb = None
                self.wf = torch.empty(sz[0].value, dtype=torch.uint8, device=dev)
                self.bias = torch.empty(sz[2].value // 4, dtype=torch.float32, device=dev)
            if need_bwd and self.wb is None:
                self.wb = torch.empty(sz[1].value, dtype=torch.uint8, device=dev)
            lib().call("mcnerf_mlp_tc_pack", ctypes.byref(ps), _p(self.wf, torch.uint8),
                       _p(self.wb, torch.uint8) if self.wb is not None else None, _p(self.bias), _stream())
            self.key = key
        return self


def make_tc_input_rays(rays_o, rays_d, jitter, smp, sel_idx, n_rows, n_rows_dev, ray_offsets=None):
    t = TcInput()
    t.rays_o, t.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
    t.jitter = jitter.data_ptr() if jitter is not None else None
    t.n_rays = rays_o.shape[0]
    t.smp = smp
    t.sample_idx = sel_idx.data_ptr() if sel_idx is not None else None
    t.n_rows = int(n_rows)
    t.n_rows_dev = n_rows_dev.data_ptr() if n_rows_dev is not None else None
    t.x_enc, t.ld_enc, t.dirs_rows = None, 0, None
    # consecutive rows per ray (dense grid, or the compacted order of select_fine): ray gradients summed in a fixed order
    t.ray_offsets = ray_offsets.data_ptr() if ray_offsets is not None else None
    t.ordered_ray_grads = int(sel_idx is None or ray_offsets is not None)
    t._keep = (rays_o, rays_d, jitter, sel_idx, n_rows_dev, ray_offsets)      # the struct only holds raw pointers
    return t


def make_tc_input_enc(x_enc, dirs):
    t = TcInput()
    t.rays_o = t.rays_d = t.jitter = t.sample_idx = t.n_rows_dev = None
    t.n_rays = 0
    t.smp = make_sampling(0.0, 1.0, 2, 10)
    t.n_rows = x_enc.shape[0]
    t.x_enc, t.ld_enc, t.dirs_rows = x_enc.data_ptr(), x_enc.shape[1], dirs.data_ptr()
    t.ray_offsets, t.ordered_ray_grads = None, 0
    t._keep = (x_enc, dirs)
    return t


def tc_stash(ps, n_rows, device):
    n = lib().cdll.mcnerf_mlp_tc_stash_bytes(ctypes.byref(ps), int(n_rows))
    return torch.empty(n, dtype=torch.uint8, device=device)


def mlp_tc_fwd(ps, tcw, tcin, out4, stash=None):
    lib().call("mcnerf_mlp_tc_fwd", ctypes.byref(ps), _p(tcw.wf, torch.uint8), _p(tcw.bias), ctypes.byref(tcin),
               _p(out4), _p(stash, torch.uint8) if stash is not None else None, _stream())


def tc_bwd_workspace(ps, n_rows, device):
    n = lib().cdll.mcnerf_mlp_tc_bwd_workspace(ctypes.byref(ps), int(n_rows))
    return torch.empty(n, dtype=torch.uint8, device=device)


def mlp_tc_bwd(ps, tcw, tcin, out4, g_out4, stash, workspace, grads_struct, g_rays_o=None, g_rays_d=None,
               g_x_enc=None, g_dirs_rows=None):
    args = (ctypes.byref(ps), _p(tcw.wb, torch.uint8), _p(tcw.bias), ctypes.byref(tcin),
            _p(out4), _p(g_out4), _p(stash, torch.uint8), _p(workspace, torch.uint8), ctypes.byref(grads_struct),
            _p(g_rays_o), _p(g_rays_d), _p(g_x_enc), _p(g_dirs_rows), _stream())
    if lib().profiling() and os.environ.get("MCNERF_BWD_FUSED", "0") == "1":
        lib().call("mcnerf_mlp_tc_bwd", *args, label="mcnerf_mlp_tc_bwd[fused]")     # one launch (mlp_tc_bwd_fused.cu)
        return
    if lib().profiling():        # two-kernel mode: the chain and the weight-gradient kernels as two timed calls
        try:
            lib().cdll.mcnerf_mlp_tc_bwd_phases(1)
            lib().call("mcnerf_mlp_tc_bwd", *args, label="mcnerf_mlp_tc_bwd[chain]")
            lib().cdll.mcnerf_mlp_tc_bwd_phases(2)
            lib().call("mcnerf_mlp_tc_bwd", *args, label="mcnerf_mlp_tc_bwd[wgrad]")
        finally:
            lib().cdll.mcnerf_mlp_tc_bwd_phases(3)
        return
    lib().call("mcnerf_mlp_tc_bwd", *args)

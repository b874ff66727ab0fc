"""CPU ORACLE for the MC-NeRF train/render hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain PyTorch-CPU fp32 restatement of the reference's algorithm
(SkylerGao/MC_NeRF, pure Python/PyTorch).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product
(mc_nerf_b200/) never does and has no CPU fallback.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md §4).  This restatement is pinned against outputs of the UNMODIFIED
reference modules executed on CPU in the build container with every random draw
captured (tests/golden/make_golden.py -> tests/golden/*.pt; checked by
tests/test_oracle_golden.py).

Differences from the reference are limited to making randomness explicit: every
`torch.randn` / `uniform_` / `randperm` the reference draws internally is an
argument here (draw order: SURVEY.md §8c).  Gradients come from autograd over
this restatement, as in the reference.

Each function cites the reference lines it follows (paths relative to the
reference repository root).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- camera model


def intrinsics_from_weights(w_fx, w_fy, w_ux, w_uy, img_h, img_w):
    """K_i from the learnable scale weights.  model/mc_nerf.py:171-186.
    Note fy is initialised from the image WIDTH (line 173) and all four entries pass through abs()."""
    n = w_fx.shape[0]
    K = torch.zeros(n, 3, 3, dtype=w_fx.dtype, device=w_fx.device)
    K[:, 0, 0] = torch.abs(float(img_w) * w_fx)
    K[:, 1, 1] = torch.abs(float(img_w) * w_fy)
    K[:, 0, 2] = torch.abs(float(img_w) / 2 * w_ux)
    K[:, 1, 2] = torch.abs(float(img_h) / 2 * w_uy)
    K[:, 2, 2] = 1.0
    return K


def _series(theta, kind, nth=10):
    """11-term Taylor series of sin(t)/t, (1-cos t)/t^2, (t-sin t)/t^3.  model/mc_nerf.py:291-316."""
    ans = torch.zeros_like(theta)
    denom = 1.0
    for i in range(nth + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        else:
            denom *= (2 * i + 2) * (2 * i + 3)
        ans = ans + (-1) ** i * theta ** (2 * i) / denom
    return ans


def se3_to_SE3(wu):
    """[...,6] twist -> [...,3,4] world->camera [R|t].  model/mc_nerf.py:269-289."""
    w, u = wu[..., :3], wu[..., 3:]
    w0, w1, w2 = w.unbind(-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], -1),
                      torch.stack([w2, O, -w0], -1),
                      torch.stack([-w1, w0, O], -1)], -2)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3, dtype=wu.dtype, device=wu.device)
    A, B, C = _series(theta, "A"), _series(theta, "B"), _series(theta, "C")
    R = I + A * wx + B * wx @ wx
    V = I + B * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], -1)


def inverse_intrinsics(K):
    """Per-camera matrix inverse.  model/mc_nerf.py:204-210."""
    return torch.stack([k.inverse() for k in K], 0)


def get_rays(pose, Kinv, img_h, img_w):
    """All H*W rays of ONE camera.  pose [3,4] world->camera, Kinv [3,3].
    model/mc_nerf.py:124-145 with pix2cam :229-232 and cam2world :245-256.
    Returns rays_d, rays_o [H*W,3], row-major pixels, centres at +0.5."""
    ys = torch.arange(img_h, dtype=torch.float32, device=pose.device) + 0.5
    xs = torch.arange(img_w, dtype=torch.float32, device=pose.device) + 0.5
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([X, Y, torch.ones_like(X)], -1).reshape(-1, 3)
    cam = pix @ Kinv.transpose(-2, -1)
    R, t = pose[:, :3], pose[:, 3:]
    Rinv = R.transpose(-2, -1)
    tinv = -Rinv @ t
    pose_inv = torch.cat([Rinv, tinv], -1)                       # [3,4] camera->world
    cam_h = torch.cat([cam, torch.ones_like(cam[:, :1])], -1)
    org_h = torch.cat([torch.zeros_like(cam), torch.ones_like(cam[:, :1])], -1)
    world = cam_h @ pose_inv.transpose(-2, -1)
    rays_o = org_h @ pose_inv.transpose(-2, -1)
    rays_d = world - rays_o
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    return rays_d, rays_o


def reproject(wpts, K, pose):
    """Calibration-point reprojection.  wpts [1,N,P,3], K [N,3,3], pose [N,3,4] -> [1,N,P,2].
    model/mc_nerf.py:147-152, 236-241, 260-267."""
    hom = torch.cat([wpts, torch.ones_like(wpts[..., :1])], -1)                 # [1,N,P,4]
    cam = pose.unsqueeze(0) @ hom.transpose(-2, -1)                              # [1,N,3,P]
    pix = K.unsqueeze(0) @ cam
    pix = pix[..., :2, :] / pix[..., 2:, :]
    return pix.transpose(-2, -1)

# --------------------------------------------------------------------------- encoding + MLP


def barf_weights(step_r, barf_start, barf_end, n_freqs, device="cpu"):
    """Coarse-to-fine window over frequency bands.  model/net_block.py:26-29."""
    alpha = (step_r - barf_start) / (barf_end - barf_start) * n_freqs
    k = torch.arange(n_freqs, dtype=torch.float32, device=device)
    return (1 - torch.cos(math.pi * (alpha - k).clamp(0, 1))) / 2


def sincos_encode(x, n_freqs, barf_w=None):
    """[M,3] -> [M,3+6L]: [x, per coord: sin(2^k c) k<L, cos(2^k c) k<L].  model/net_block.py:20-35."""
    freqs = 2 ** torch.linspace(0, n_freqs - 1, n_freqs, device=x.device)
    spec = x[..., None] * freqs                                   # [M,3,L]
    enc = torch.stack([spec.sin(), spec.cos()], dim=-2)           # [M,3,2,L]
    if barf_w is not None:
        enc = enc * barf_w
    return torch.cat([x, enc.reshape(x.shape[0], -1)], -1)


SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)


def eval_sh_deg2(sh, dirs):
    """sh [M,3,9], dirs [M,3] -> [M,3].  model/net_utils.py:154-169 (deg <= 2 branch)."""
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    res = SH_C0 * sh[..., 0]
    res = res - SH_C1 * y * sh[..., 1] + SH_C1 * z * sh[..., 2] - SH_C1 * x * sh[..., 3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    res = (res + SH_C2[0] * xy * sh[..., 4] + SH_C2[1] * yz * sh[..., 5]
           + SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
           + SH_C2[3] * xz * sh[..., 7] + SH_C2[4] * (xx - yy) * sh[..., 8])
    return res


SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]
SH_C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
         -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


def eval_sh(deg, sh, dirs):
    """sh [M,3,(deg+1)^2], dirs [M,3] -> [M,3], degree 0..4.  model/net_utils.py:103-191 (every branch)."""
    res = SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        res = res - SH_C1 * y * sh[..., 1] + SH_C1 * z * sh[..., 2] - SH_C1 * x * sh[..., 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + SH_C2[0] * xy * sh[..., 4] + SH_C2[1] * yz * sh[..., 5]
               + SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
               + SH_C2[3] * xz * sh[..., 7] + SH_C2[4] * (xx - yy) * sh[..., 8])
    if deg > 2:
        res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + SH_C3[1] * xy * z * sh[..., 10]
               + SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11]
               + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
               + SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13]
               + SH_C3[5] * z * (xx - yy) * sh[..., 14] + SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    if deg > 3:
        res = (res + SH_C4[0] * xy * (xx - yy) * sh[..., 16] + SH_C4[1] * yz * (3 * xx - yy) * sh[..., 17]
               + SH_C4[2] * xy * (7 * zz - 1) * sh[..., 18] + SH_C4[3] * yz * (7 * zz - 3) * sh[..., 19]
               + SH_C4[4] * (zz * (35 * zz - 30) + 3) * sh[..., 20] + SH_C4[5] * xz * (7 * zz - 3) * sh[..., 21]
               + SH_C4[6] * (xx - yy) * (7 * zz - 1) * sh[..., 22] + SH_C4[7] * xz * (xx - 3 * yy) * sh[..., 23]
               + SH_C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)) * sh[..., 24])
    return res


def mlp_forward(params, x_enc, dirs, depth, skips, deg=2):
    """CorseFine_NeRF.forward.  model/net_block.py:67-78.
    `params` maps the reference's state_dict names ('xyz_encoding_1.0.weight', 'sigma.0.bias',
    'sh.2.weight', ...) to tensors.  Returns [M,4] = (sigma_raw, r, g, b)."""
    h = x_enc
    for i in range(depth):
        if i in skips:
            h = torch.cat([x_enc, h], -1)
        h = F.relu(F.linear(h, params[f"xyz_encoding_{i+1}.0.weight"], params[f"xyz_encoding_{i+1}.0.bias"]))
    s = F.relu(F.linear(h, params["sigma.0.weight"], params["sigma.0.bias"]))
    sigma = F.linear(s, params["sigma.2.weight"], params["sigma.2.bias"])
    c = F.relu(F.linear(h, params["sh.0.weight"], params["sh.0.bias"]))
    sh = F.linear(c, params["sh.2.weight"], params["sh.2.bias"])
    rgb = torch.sigmoid(eval_sh(deg, sh.reshape(-1, 3, (deg + 1) ** 2), dirs))
    return torch.cat([sigma, rgb], -1)


def init_mlp_params(depth, width, skips, n_freqs=10, deg=2, seed=42):
    """nn.Linear default init in the reference's construction order (model/net_block.py:51-65)."""
    import torch.nn as nn
    g = torch.get_rng_state()
    torch.manual_seed(seed)
    in_ch = 3 * (2 * n_freqs + 1)
    p = {}
    for i in range(depth):
        k = in_ch if i == 0 else (width + in_ch if i in skips else width)
        lin = nn.Linear(k, width)
        p[f"xyz_encoding_{i+1}.0.weight"], p[f"xyz_encoding_{i+1}.0.bias"] = lin.weight.detach(), lin.bias.detach()
    for name, (a, b) in (("sigma", (width, 1)), ("sh", (width, 3 * (deg + 1) ** 2))):
        l0, l2 = nn.Linear(width, a), nn.Linear(width, b)
        p[f"{name}.0.weight"], p[f"{name}.0.bias"] = l0.weight.detach(), l0.bias.detach()
        p[f"{name}.2.weight"], p[f"{name}.2.bias"] = l2.weight.detach(), l2.bias.detach()
    torch.set_rng_state(g)
    return p

# --------------------------------------------------------------------------- compositing


def sigma2weights(deltas, sigmas, noise):
    """Noisy alpha compositing weights.  model/mc_nerf.py:729-736 (noise = the randn it draws)."""
    s = sigmas + noise
    alphas = 1 - torch.exp(-deltas * F.softplus(s))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    return alphas * torch.cumprod(shifted, -1)[:, :-1]


def z_deltas(z_vals):
    """model/mc_nerf.py:708-710."""
    d = z_vals[:, 1:] - z_vals[:, :-1]
    return torch.cat([d, 1e10 * torch.ones_like(d[:, :1])], -1)


def composite(out4, rays_d, z_vals, noise, white_back=True):
    """Tail of NeRF_Model.inference.  model/mc_nerf.py:705-727.
    out4 [B,S,4] (sigma_raw, rgb).  Returns rgb [B,3], depth [B,1], opacity [B,1], weights [B,S]."""
    sigmas, rgbs = out4[..., 0], out4[..., 1:]
    ray_len = rays_d.norm(dim=-1, keepdim=True)
    deltas = z_deltas(z_vals)
    sd = F.softplus(sigmas) * (deltas * ray_len)
    alpha = 1 - torch.exp(-sd)
    T = torch.exp(-torch.cat([torch.zeros_like(sd[:, :1]), sd[:, :-1]], 1).cumsum(1))
    prob = (T * alpha)[..., None]
    opacity = prob.sum(1)
    depth = (z_vals.unsqueeze(-1) * prob).sum(1)
    w = sigma2weights(deltas, sigmas, noise)
    rgb = (w.unsqueeze(-1) * rgbs).sum(1)
    if white_back:
        rgb = rgb + 1 - w.sum(1).unsqueeze(-1)
    return rgb, depth, opacity, w


def select_fine(weights, thresh, scale, cap=None, cap_perm=None):
    """Threshold-gated fine selection.  model/mc_nerf.py:623-632 (train) / 663-667 (test).
    Returns idx [Msel,2] int64 (ray, fine index), ray-major ascending; `cap` (= 128*B in train)
    with the CPU randperm `cap_perm` reproduces lines 630-632."""
    thr = min(thresh, weights.max().item())
    idx = torch.nonzero(weights >= thr)
    idx = idx.unsqueeze(1).expand(-1, scale, -1).clone()
    idx[..., 1] = idx[..., 1] * scale + torch.arange(scale, device=weights.device).reshape(1, scale)
    idx = idx.reshape(-1, 2)
    if cap is not None and idx.shape[0] > cap:
        idx = idx[cap_perm[:cap]]
    return idx

# --------------------------------------------------------------------------- renderer


def _inference(params, cfg, net, xyz, rays_d, z_vals, noise, barf_w, idx=None):
    """NeRF_Model.inference.  model/mc_nerf.py:682-727."""
    depth, _, skips = cfg[net]
    B, S = z_vals.shape
    view = rays_d.unsqueeze(1).expand(-1, S, -1)
    if idx is not None:
        view_s, xyz_s = view[idx[:, 0], idx[:, 1]], xyz[idx[:, 0], idx[:, 1]]
        out = torch.cat([torch.full((B, S, 1), cfg["sigma_default"], device=xyz.device),
                         torch.full((B, S, 3), 1.0, device=xyz.device)], 2)
        res = mlp_forward(params, sincos_encode(xyz_s, cfg["n_freqs"], barf_w), view_s, depth, skips)
        out[idx[:, 0], idx[:, 1]] = res
    else:
        res = mlp_forward(params, sincos_encode(xyz.reshape(-1, 3), cfg["n_freqs"], barf_w),
                          view.reshape(-1, 3), depth, skips)
        out = res.reshape(B, S, 4)
    rgb, dep, opa, _ = composite(out, rays_d, z_vals, noise, cfg["white_back"])
    return rgb, out[..., 0], dep, opa, out


def render_rays(params_c, params_f, cfg, rays_d, rays_o, rng, step_r=1.0, barf=False, train=True,
                cap_perm=None, return_aux=False):
    """render_rays_train (model/mc_nerf.py:598-646) when train=True, else render_rays_test (:648-680).
    cfg: dict(near, far, Sc, scale, n_freqs, white_back, sigma_default, thresh, barf_start, barf_end,
              coarse=(depth,width,skips), fine=(...)).
    rng: dict(jitter [B,1] (train), noise_c, noise_sel [B,Sc], noise_f [B,Sf])."""
    B = rays_d.shape[0]
    Sc, Sf = cfg["Sc"], cfg["Sc"] * cfg["scale"]
    dev = rays_d.device
    z_c = torch.linspace(cfg["near"], cfg["far"], Sc, device=dev).expand(B, -1)
    z_f = torch.linspace(cfg["near"], cfg["far"], Sf, device=dev).expand(B, -1)
    if train:
        z_c = z_c + rng["jitter"]
        z_f = z_f + rng["jitter"]
    barf_w = barf_weights(step_r, cfg["barf_start"], cfg["barf_end"], cfg["n_freqs"], dev) if barf else None
    xyz_c = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_c.unsqueeze(2)
    rgb_c, sig_c, dep_c, opa_c, out_c = _inference(params_c, cfg, "coarse", xyz_c, rays_d, z_c, rng["noise_c"], barf_w)
    with torch.no_grad():
        w_sel = sigma2weights(z_deltas(z_c), sig_c.detach(), rng["noise_sel"])
    idx = select_fine(w_sel, cfg["thresh"], cfg["scale"], cap=(B * 128 if train else None), cap_perm=cap_perm)
    xyz_f = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_f.unsqueeze(2)
    rgb_f, sig_f, dep_f, opa_f, out_f = _inference(params_f, cfg, "fine", xyz_f, rays_d, z_f, rng["noise_f"], barf_w, idx=idx)
    if return_aux:
        return dict(rgb_c=rgb_c, rgb_f=rgb_f, depth_c=dep_c, opacity_c=opa_c, depth_f=dep_f, opacity_f=opa_f,
                    out_c=out_c, out_f=out_f, w_sel=w_sel, idx=idx)
    if train:
        return rgb_c, rgb_f
    return rgb_f, dep_f, opa_f

# --------------------------------------------------------------------------- loss + train step


def reproject_loss(pd, gt, img_h, img_w):
    """model/loss.py:45-58."""
    lx = F.mse_loss(pd[..., 0] / img_w, gt[..., 0] / img_w)
    ly = F.mse_loss(pd[..., 1] / img_h, gt[..., 1] / img_h)
    return lx + ly


def mc_nerf_loss(rgb_c, rgb_f, gt, reproj_intr, intr_pts, img_h, img_w, normalise_intr=True):
    """model/loss.py:15-43 for the GLOBAL_OPTIM / FINE_TUNE stages."""
    l_intr = reproject_loss(reproj_intr, intr_pts, img_h, img_w)
    loss = l_intr / (l_intr.detach() + 1e-8) if normalise_intr else l_intr
    return loss + F.mse_loss(rgb_c, gt) + F.mse_loss(rgb_f, gt)


def train_step(cam, params_c, params_f, cfg, batch, rng, step_r=0.5, stage="GLOBAL_OPTIM_EPOCH",
               cap_perm=None):
    """One MC_Model.forward + MC_NeRF_Loss + backward in the NeRF stages.
    model/mc_nerf.py:73-95, model/loss.py:15-31, main.py:79-84.
    cam: dict of the six learnable camera tensors (requires_grad set by the caller);
    batch: (gt_rgbs[1,HW,3], img_id[1], intr_wpts, intr_pts, extr_wpts, extr_pts);
    rng adds 'rand_idx' [B] (the randperm(HW)[:batch] of model/mc_nerf.py:329).
    Returns dict(loss, rgb_c, rgb_f); gradients are left in .grad of every leaf."""
    gt, img_id, intr_wpts, intr_pts, _, _ = batch
    H, W = cfg["img_h"], cfg["img_w"]
    barf = stage == "GLOBAL_OPTIM_EPOCH"
    K = intrinsics_from_weights(cam["weights_fx"], cam["weights_fy"], cam["weights_ux"], cam["weights_uy"], H, W)
    pose_w = cam["weights_pose"] if barf else cam["weights_pose"].detach()       # extr frozen in stage 3 (:87)
    pose = se3_to_SE3(pose_w)
    calib_pose = se3_to_SE3(cam["weights_pose_intr"])
    reproj = reproject(intr_wpts, K, calib_pose)
    i = int(img_id[0])
    rays_d, rays_o = get_rays(pose[i], inverse_intrinsics(K)[i], H, W)
    sel = rng["rand_idx"]
    rgb_c, rgb_f = render_rays(params_c, params_f, cfg, rays_d[sel], rays_o[sel], rng,
                               step_r=(step_r if barf else 1), barf=barf, train=True, cap_perm=cap_perm)
    gt_sel = gt.reshape(-1, 3)[sel]
    loss = mc_nerf_loss(rgb_c, rgb_f, gt_sel, reproj, intr_pts, H, W)
    loss.backward()
    return dict(loss=loss.detach(), rgb_c=rgb_c.detach(), rgb_f=rgb_f.detach())


def camera_stage_step(cam, cfg, batch):
    """One CAM_PARAM_EPOCH step (stage 1): model/mc_nerf.py:64-71 + model/loss.py:18-26.  Both calibration point
    sets are reprojected - the "intr" set through the calibration poses (weights_pose_intr), the "extr" set through
    the main poses - and the two reprojection losses are summed UN-normalised.  The NeRF is not evaluated.
    Returns dict(loss, reproj_intr, reproj_extr); gradients are left in .grad of the camera leaves."""
    _, _, intr_wpts, intr_pts, extr_wpts, extr_pts = batch
    H, W = cfg["img_h"], cfg["img_w"]
    K = intrinsics_from_weights(cam["weights_fx"], cam["weights_fy"], cam["weights_ux"], cam["weights_uy"], H, W)
    pose = se3_to_SE3(cam["weights_pose"])
    calib_pose = se3_to_SE3(cam["weights_pose_intr"])
    r_intr = reproject(intr_wpts, K, calib_pose)
    r_extr = reproject(extr_wpts, K, pose)
    loss = reproject_loss(r_intr, intr_pts, H, W) + reproject_loss(r_extr, extr_pts, H, W)
    loss.backward()
    return dict(loss=loss.detach(), reproj_intr=r_intr.detach(), reproj_extr=r_extr.detach(), K=K.detach(),
                pose=pose.detach())


def cfg_from_sys_param(sp):
    return dict(near=sp["near"], far=sp["far"], Sc=sp["samples"], scale=sp["scale"], n_freqs=sp["emb_freqs_xyz"],
                white_back=sp["white_back"], sigma_default=sp["sigma_default"], thresh=sp["sample_weight_thresh"],
                barf_start=sp["barf_start"], barf_end=sp["barf_end"], img_h=sp["data_img_h"], img_w=sp["data_img_w"],
                coarse=(sp["coarse_MLP_depth"], sp["coarse_MLP_width"], tuple(sp["coarse_MLP_skip"])),
                fine=(sp["fine_MLP_depth"], sp["fine_MLP_width"], tuple(sp["fine_MLP_skip"])))


# --------------------------------------------------------------------------- pixel choice (checker of the device sampler)


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on numpy uint64 lanes holding 32-bit words."""
    import numpy as np
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85, np.uint64(0xFFFFFFFF)
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def sample_pixels(n, batch, seed0, seed1):
    """randperm(n)[:batch] (model/mc_nerf.py:327-345) as the device sampler DEFINES its permutation: pixel i gets the
    composite key (64 - ceil(log2 n) random Philox bits | i) and the permutation is the argsort of ALL n composites.
    The kernel (mcnerf_sample_pixels) must return exactly the head of this full sort without performing it."""
    import numpy as np
    bits = 1
    while (1 << bits) < n:
        bits += 1
    s0, s1 = int(seed0) & 0xFFFFFFFFFFFFFFFF, int(seed1) & 0xFFFFFFFFFFFFFFFF
    i = np.arange(n, dtype=np.uint64)
    z = np.zeros(n, dtype=np.uint64)
    r0, r1, _, _ = _philox4x32_10(i, z, z + np.uint64(s1 & 0xFFFFFFFF), z + np.uint64(s1 >> 32),
                                  s0 & 0xFFFFFFFF, s0 >> 32)
    r = ((r0 << np.uint64(32)) | r1) >> np.uint64(bits)
    comp = (r << np.uint64(bits)) | i
    order = np.sort(comp)[:min(n, batch)]
    return (order & np.uint64((1 << bits) - 1)).astype(np.int64)


# --------------------------------------------------------------------------- device RNG streams (checker of philox.cuh)


def philox_stream(n, seed0, seed1, stream_id):
    """First two Philox4x32-10 output words of elements 0..n-1 of stream `stream_id`: counter
    (i_lo, stream | i_hi << 8, seed1_lo, seed1_hi), key (seed0_lo, seed0_hi) - mc_nerf_b200/csrc/philox.cuh."""
    import numpy as np
    s0, s1 = int(seed0) & 0xFFFFFFFFFFFFFFFF, int(seed1) & 0xFFFFFFFFFFFFFFFF
    i = np.arange(n, dtype=np.uint64)
    c0, c1 = i & np.uint64(0xFFFFFFFF), np.uint64(stream_id) | ((i >> np.uint64(32)) << np.uint64(8))
    c2 = np.full(n, s1 & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(n, s1 >> 32, dtype=np.uint64)
    r0, r1, _, _ = _philox4x32_10(c0, c1, c2, c3, s0 & 0xFFFFFFFF, s0 >> 32)
    return r0.astype(np.uint32), r1.astype(np.uint32)


def _u01(x):
    import numpy as np
    return (x.astype(np.float32) + np.float32(0.5)) * np.float32(2.3283064365386963e-10)


def philox_uniform(n, seed0, seed1, stream_id, lo=0.0, hi=1.0):
    import numpy as np
    r0, _ = philox_stream(n, seed0, seed1, stream_id)
    return np.float32(lo) + np.float32(hi - lo) * _u01(r0)


def philox_normal(n, seed0, seed1, stream_id):
    """N(0,1) by Box-Muller from the first two words (cos branch), as the render kernels draw their density noise in
    device-RNG mode (the reference draws torch.randn: same distribution, different values)."""
    import numpy as np
    r0, r1 = philox_stream(n, seed0, seed1, stream_id)
    u1 = np.minimum(_u01(r0), np.float32(0.99999994)).astype(np.float64)
    u2 = _u01(r1).astype(np.float64)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)

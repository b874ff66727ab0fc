"""Synthetic Ball-style multi-camera rig and inputs.

The reference ships no data; its Blender script (reference
synthetic_dataset_code/Ball.py:16-24,146-224) places 110 cameras on a sphere
of radius 3 looking at the origin (9 latitudes x 12 longitudes + 2 poles) with
an integer field of view drawn from [40, 80] degrees.  The loader turns each
camera-to-world matrix into a world->camera [R|t] with +z looking forward
(reference data/data_read.py:246-257) and the FOV into fx = fy = (W/2)/tan(fov/2)
(data/data_read.py:141-152).

Everything here is host-side numpy/torch on CPU tensors; the caller moves the
results to the device.  Used by tests, bench.py and smoke(); it never touches
the oracle or the CUDA library.
"""
import math

import numpy as np
import torch

N_CAM_BALL = 110


def ball_rig(n_cam=N_CAM_BALL, radius=3.0, seed=4, fov_min=40, fov_max=80):
    """Returns (w2c [n,3,4] float32, fov_deg [n] float32)."""
    rng = np.random.RandomState(seed)
    dirs = []
    for phi in np.linspace(-80.0, 80.0, 9):
        for theta in np.linspace(0.0, 360.0, 12, endpoint=False):
            dirs.append((phi, theta))
    dirs += [(-90.0, 0.0), (90.0, 0.0)]
    if n_cam <= len(dirs):
        # spread a smaller rig over the same sphere
        sel = np.linspace(0, len(dirs) - 1, n_cam).round().astype(int)
        dirs = [dirs[i] for i in sel]
    else:
        while len(dirs) < n_cam:
            dirs.append((rng.uniform(-80, 80), rng.uniform(0, 360)))
    poses = []
    for phi, theta in dirs:
        p, t = math.radians(phi), math.radians(theta)
        pos = radius * np.array([math.cos(p) * math.cos(t), math.cos(p) * math.sin(t), math.sin(p)])
        fwd = -pos / np.linalg.norm(pos)  # camera +z looks at the origin
        up = np.array([0.0, 0.0, 1.0])
        if abs(np.dot(fwd, up)) > 0.999:
            up = np.array([0.0, 1.0, 0.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        c2w_R = np.stack([right, down, fwd], axis=1)  # columns = camera axes in world
        w2c_R = c2w_R.T
        w2c_t = -w2c_R @ pos
        poses.append(np.concatenate([w2c_R, w2c_t[:, None]], axis=1))
    w2c = torch.tensor(np.stack(poses), dtype=torch.float32)
    fov = torch.tensor(rng.randint(fov_min, fov_max + 1, size=len(dirs)), dtype=torch.float32)
    return w2c, fov


def fov_to_intrinsics(fov_deg, img_h, img_w):
    """[n] degrees -> K [n,3,3] (reference data/data_read.py:141-152)."""
    f = torch.deg2rad(fov_deg.double())
    fx = (img_w / 2) / torch.tan(f / 2)
    fy = (img_h / 2) / torch.tan(f / 2)
    K = torch.zeros(len(fov_deg), 3, 3, dtype=torch.float64)
    K[:, 0, 0] = fx
    K[:, 1, 1] = fy
    K[:, 0, 2] = img_w / 2
    K[:, 1, 2] = img_h / 2
    K[:, 2, 2] = 1
    return K.float()


def _taylor(theta2, kind, nth=10):
    # same 11-term series the reference evaluates (model/mc_nerf.py:291-316)
    ans = torch.zeros_like(theta2)
    denom = 1.0
    for i in range(nth + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        else:
            denom *= (2 * i + 2) * (2 * i + 3)
        ans = ans + (-1) ** i * theta2 ** i / denom
    return ans


def se3_log(Rt):
    """Inverse of the reference's se3_to_SE3 (model/mc_nerf.py:269-281): [n,3,4] -> [n,6]."""
    Rt = Rt.double()
    R, t = Rt[:, :, :3], Rt[:, :, 3]
    cos = ((R.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2).clamp(-1 + 1e-12, 1 - 1e-12)
    theta = torch.acos(cos)
    axis = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], -1)
    s = (2 * torch.sin(theta)).clamp_min(1e-12)
    w = axis / s[:, None] * theta[:, None]
    O = torch.zeros_like(theta)
    wx = torch.stack([torch.stack([O, -w[:, 2], w[:, 1]], -1),
                      torch.stack([w[:, 2], O, -w[:, 0]], -1),
                      torch.stack([-w[:, 1], w[:, 0], O], -1)], -2)
    th2 = (theta ** 2)[:, None, None]
    V = torch.eye(3, dtype=torch.float64) + _taylor(th2, "B") * wx + _taylor(th2, "C") * wx @ wx
    u = torch.linalg.solve(V, t[:, :, None])[:, :, 0]
    return torch.cat([w, u], -1).float()


def make_sys_param(n_cam=N_CAM_BALL, img_h=100, img_w=100, batch=1024, samples=64, scale=2,
                   coarse=(8, 256, (4,)), fine=(8, 256, (4,)), emb_freqs=10, deg=2, device="cpu",
                   mode=0, near=1.0, far=8.0, seed=4, with_images=True, barf_start=20 / 52, barf_end=36 / 52,
                   weight_thresh=1e-3, pixel_sampler="randperm"):
    """The flat `sys_param` dict every reference constructor reads (SURVEY.md §5 'Config')."""
    w2c, fov = ball_rig(n_cam, seed=seed)
    K = fov_to_intrinsics(fov, img_h, img_w)
    Kinv = torch.linalg.inv(K)
    g = torch.Generator().manual_seed(seed + 1000)
    if with_images:
        valid_rgbs = torch.rand(n_cam, img_h * img_w, 3, generator=g)
    else:
        valid_rgbs = torch.zeros(n_cam, 1, 3)
    return dict(
        mode=mode, device_type=device, distributed=False, batch=batch, pixel_sampler=pixel_sampler,
        data_img_h=img_h, data_img_w=img_w, res_h=img_h, res_w=img_w,
        data_numb=[n_cam, n_cam, n_cam],
        intr_mat=[K, K.clone(), K.clone()], intr_mat_inv=[Kinv, Kinv.clone(), Kinv.clone()],
        gt_pose=w2c, test_pose=w2c.clone(), valid_pose=w2c.clone(), valid_rgbs=valid_rgbs,
        near=near, far=far, samples=samples, scale=scale, sample_weight_thresh=weight_thresh,
        sigma_default=-20.0, sigma_init=30.0, white_back=True, emb_freqs_xyz=emb_freqs,
        barf_mask=False, barf_start=barf_start, barf_end=barf_end, MLP_deg=deg,
        coarse_MLP_depth=coarse[0], coarse_MLP_width=coarse[1], coarse_MLP_skip=list(coarse[2]),
        fine_MLP_depth=fine[0], fine_MLP_width=fine[1], fine_MLP_skip=list(fine[2]),
        boader_min=-3.5, boader_max=3.5, grid_nerf=384, warmup_epoch=100,
        root_weight="./weights", demo_render_pth="./results", data_name="synthetic",
        train_json_file="", demo_ckpt="",
    )


def init_camera_weights(sys_param, seed=7, noise=0.01):
    """Benchmark initialisation of the learnables: the reference's
    all-ones init points every camera the same way, so start near the true rig."""
    w2c = sys_param["gt_pose"]
    K = sys_param["intr_mat"][0]
    W = sys_param["data_img_w"]
    g = torch.Generator().manual_seed(seed)
    n = w2c.shape[0]
    return dict(
        weights_pose=se3_log(w2c) + noise * torch.randn(n, 6, generator=g),
        weights_pose_intr=torch.ones(n, 6),
        weights_fx=K[:, 0, 0] / W,
        weights_fy=K[:, 1, 1] / W,   # sic: the reference scales fy by the WIDTH (model/mc_nerf.py:173)
        weights_ux=torch.ones(n),
        weights_uy=torch.ones(n),
    )


def make_train_batch(sys_param, img_id=3, seed=11):
    """One DataLoader item exactly as reference main.py:78-80 hands it to MC_Model.forward."""
    n = sys_param["data_numb"][0]
    H, W = sys_param["data_img_h"], sys_param["data_img_w"]
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(1, H * W, 3, generator=g)
    intr_wpts = torch.rand(1, n, 5, 3, generator=g) - 0.5
    intr_pts = torch.rand(1, n, 5, 2, generator=g) * W
    extr_wpts = torch.rand(1, n, 5, 3, generator=g) - 0.5
    extr_pts = torch.rand(1, n, 5, 2, generator=g) * W
    return (gt, torch.tensor([img_id], dtype=torch.long), intr_wpts, intr_pts, extr_wpts, extr_pts)


def draw_step_rng(sys_param, n_rays, seed=123, train=True):
    """The random draws one step consumes, in the reference's order:
    randperm(HW) -> uniform[B,1] -> randn[B,Sc] -> randn[B,Sc] -> randn[B,Sf]."""
    g = torch.Generator().manual_seed(seed)
    H, W = sys_param["data_img_h"], sys_param["data_img_w"]
    Sc = sys_param["samples"]
    Sf = Sc * sys_param["scale"]
    out = {}
    if train:
        out["perm"] = torch.randperm(H * W, generator=g)
        out["rand_idx"] = out["perm"][:n_rays]
        n_rays = out["rand_idx"].shape[0]
        out["jitter"] = torch.rand(n_rays, 1, generator=g) * ((sys_param["far"] - sys_param["near"]) / Sc)
    out["noise_c"] = torch.randn(n_rays, Sc, generator=g)
    out["noise_sel"] = torch.randn(n_rays, Sc, generator=g)
    out["noise_f"] = torch.randn(n_rays, Sf, generator=g)
    return out

/*
 * mcnerf.h - C ABI of libmcnerf.so: hand-written sm_100a CUDA kernels for the MC-NeRF
 * train/render hot path (reference: SkylerGao/MC_NeRF, model/mc_nerf.py, model/net_block.py,
 * model/net_utils.py - pure PyTorch, no FFI of its own).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; tensors are dense row-major
 *    fp32 unless stated; index tensors are int32.
 *  - every entry returns 0 on success, non-zero (a cudaError_t or MCNERF_E_*) on failure, never
 *    throws, never allocates: the caller owns all buffers including workspaces.
 *  - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns
 *    without synchronising.
 *  - gradient outputs named g_* are ACCUMULATED into (+=) when the doc says "accumulate",
 *    otherwise overwritten.
 *  - mcnerf_last_error() returns a thread-local description of the last failure.
 *
 * Each entry cites the reference lines it replaces ("ref:" = path in the reference repository).
 */
#ifndef MCNERF_H_
#define MCNERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCNERF_ABI_VERSION 1
#define MCNERF_MAX_DEPTH 16
#define MCNERF_MAX_FREQS 16

#define MCNERF_E_ARG 10001      /* bad argument (shape / null pointer / unsupported size) */
#define MCNERF_E_UNSUPPORTED 10002

int mcnerf_abi_version(void);
const char* mcnerf_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t mcnerf_launch_count(void);

/* ------------------------------------------------------------------ camera model (a2-a4, a7)
 * ref: model/mc_nerf.py:171-186 (add_weights2intr), :204-210 (inverse_intrinsic, closed form here),
 *      :269-316 (se3_to_SE3 with the 11-term Taylor series A,B,C). */
int mcnerf_intrinsics_fwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                          int n_cam, int img_h, int img_w,
                          float* K /*[n,3,3]*/, float* Kinv /*[n,3,3]*/, void* stream);
/* g_w* overwritten. gK / gKinv may be NULL. */
int mcnerf_intrinsics_bwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                          int n_cam, int img_h, int img_w, const float* gK, const float* gKinv,
                          float* g_fx, float* g_fy, float* g_ux, float* g_uy, void* stream);
int mcnerf_se3_fwd(const float* wu /*[n,6]*/, int n, float* Rt /*[n,3,4] world->camera*/, void* stream);
int mcnerf_se3_bwd(const float* wu, const float* gRt, int n, float* g_wu /*overwritten*/, void* stream);

/* Calibration-point reprojection (a7): pix[c,p] = (K_c [R|t]_c X_cp)_{xy} / z.  wpts [n_cam,n_pts,3] -> pix [n_cam,n_pts,2].
 * ref: model/mc_nerf.py:147-152 (get_reproject_pixels), :236-241 (cam2pix), :260-267 (world2cam).
 * bwd: gK [n_cam,3,3], gRt [n_cam,3,4] overwritten. */
int mcnerf_reproject_fwd(const float* wpts, const float* K, const float* Rt, int n_cam, int n_pts, float* pix,
                         void* stream);
int mcnerf_reproject_bwd(const float* wpts, const float* K, const float* Rt, const float* g_pix, int n_cam, int n_pts,
                         float* gK, float* gRt, void* stream);

/* The camera model of one train step fused (ref: model/mc_nerf.py:75-76, 87-88: add_weights2param +
 * get_reproject_pixels): K, Kinv [n,3,3], pose, calib_pose [n,3,4] (world->camera), pix [n,P,2] = reprojection of the
 * calibration points wpts [n,P,3] through (K, calib_pose).  One launch instead of four. */
int mcnerf_camera_fwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy, const float* w_pose,
                      const float* w_pose_calib, const float* wpts, int n_cam, int n_pts, int img_h, int img_w,
                      float* K, float* Kinv, float* pose, float* calib_pose, float* pix, void* stream);
/* Backward of the above from g_Kinv [n,3,3] (may be NULL), g_pose [n,3,4] (NULL with g_w_pose = NULL: extrinsics frozen,
 * ref: model/mc_nerf.py:87) and g_pix [n,P,2]; scratch: 21 * n_cam floats.  One launch instead of four. */
int mcnerf_camera_bwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy, const float* w_pose,
                      const float* w_pose_calib, const float* wpts, const float* K, const float* calib_pose, int n_cam,
                      int n_pts, int img_h, int img_w, const float* g_Kinv, const float* g_pose, const float* g_pix,
                      float* scratch, float* g_fx, float* g_fy, float* g_ux, float* g_uy, float* g_w_pose,
                      float* g_w_pose_calib, void* stream);

/* ------------------------------------------------------------------ ray generation (a5, a6)
 * ref: model/mc_nerf.py:124-145 (get_rays), :229-256 (pix2cam, cam2world), :327-345 (gather of the
 * randperm subset).  Ray r uses camera cam_id[r] (or `cam_const` when cam_id is NULL) and pixel
 * pix[r] = y*img_w + x (or r itself when pix is NULL: a full image in row-major order). */
int mcnerf_raygen_fwd(const float* Kinv, const float* Rt, const int32_t* cam_id, int cam_const,
                      const int32_t* pix, int n_rays, int img_w,
                      float* rays_o /*[n,3]*/, float* rays_d /*[n,3]*/, void* stream);
/* accumulate into gKinv [n_cam,3,3] and gRt [n_cam,3,4] (fused per-camera reduction, G5). */
int mcnerf_raygen_bwd(const float* Kinv, const float* Rt, const int32_t* cam_id, int cam_const,
                      const int32_t* pix, int n_rays, int img_w,
                      const float* g_rays_o, const float* g_rays_d,
                      float* gKinv, float* gRt, void* stream);

/* ------------------------------------------------------------------ sampling + encoding (a8-a10)
 * ref: model/mc_nerf.py:599-602, 633-635 (z = linspace(near,far,S) + per-ray jitter, x = o + d z),
 *      model/net_block.py:20-35 (sin/cos encoding, BARF window).
 * Row m of the output encodes one sample.  Which sample:
 *   sample_idx == NULL : m = ray*S + k            (dense; n_rows = n_rays*S)
 *   sample_idx != NULL : sample_idx[m] = ray*S + k (n_rows read from *n_rows_dev if non-NULL, else n_rows)
 * enc is [n_rows, ld_enc] with 3+6L valid columns (x,y,z, then per coordinate L sines, L cosines);
 * band_w[L] are the BARF weights (all 1 when the window is off). */
typedef struct {
  float near_, far_;
  int S;            /* samples per ray on this grid (coarse: Sc, fine: Sc*scale) */
  int n_freqs;      /* L */
  float band_w[MCNERF_MAX_FREQS];
  const float* band_w_dev;   /* optional: L floats in DEVICE memory that override band_w[] when non-NULL, so that a
                                captured CUDA graph follows a moving BARF window (the struct itself is baked in) */
} mcnerf_sampling;

int mcnerf_encode_rays_fwd(const float* rays_o, const float* rays_d, const float* jitter /*[n_rays] or NULL*/,
                           int n_rays, const mcnerf_sampling* smp,
                           const int32_t* sample_idx, int n_rows, const int32_t* n_rows_dev,
                           float* enc, int ld_enc, void* stream);
/* g_enc [n_rows, ld_enc] -> accumulate into g_rays_o, g_rays_d [n_rays,3]. */
int mcnerf_encode_rays_bwd(const float* rays_o, const float* rays_d, const float* jitter,
                           int n_rays, const mcnerf_sampling* smp,
                           const int32_t* sample_idx, int n_rows, const int32_t* n_rows_dev,
                           const float* g_enc, int ld_enc, float* g_rays_o, float* g_rays_d, void* stream);
/* SinCosEmbedding.forward on explicit points x [n,3] (ref: model/net_block.py:20-35). */
int mcnerf_encode_points_fwd(const float* x, int n, int n_freqs, const float* band_w_host,
                             float* enc, int ld_enc, void* stream);
int mcnerf_encode_points_bwd(const float* x, int n, int n_freqs, const float* band_w_host,
                             const float* g_enc, int ld_enc, float* g_x /*[n,3] overwritten*/, void* stream);

/* ------------------------------------------------------------------ NeRF MLP (a11, a12)
 * ref: model/net_block.py:37-78 (CorseFine_NeRF), model/net_utils.py:103-191 (eval_sh, deg 2).
 * Parameters are the reference's own tensors: row-major [out,in] fp32 weights and [out] biases. */
typedef struct {
  int depth, width, in_ch;     /* in_ch = 3+6L = 63 */
  uint32_t skip_mask;          /* bit i set: layer i (0-based) takes cat([x_enc, h]) */
  int sh_dim;                  /* 3*(deg+1)^2 = 27; only deg 2 is implemented */
  const float* W[MCNERF_MAX_DEPTH];
  const float* b[MCNERF_MAX_DEPTH];
  const float *W_sigma0, *b_sigma0, *W_sigma2, *b_sigma2;
  const float *W_sh0, *b_sh0, *W_sh2, *b_sh2;
} mcnerf_mlp_params;

typedef struct {               /* same shapes as the parameters; accumulated into */
  float* W[MCNERF_MAX_DEPTH];
  float* b[MCNERF_MAX_DEPTH];
  float *W_sigma0, *b_sigma0, *W_sigma2, *b_sigma2;
  float *W_sh0, *b_sh0, *W_sh2, *b_sh2;
} mcnerf_mlp_grads;

/* How row m finds its view direction: dirs[m] (per_row), dirs[ray_of_row[m]/S...]:
 *   dir_idx == NULL && dir_S == 0 : dirs is [n_rows,3]
 *   dir_idx == NULL && dir_S  > 0 : dirs is [n_rays,3], ray = m / dir_S
 *   dir_idx != NULL               : dirs is [n_rays,3], ray = dir_idx[m] / dir_S  (dir_idx = sample_idx) */
typedef struct {
  const float* dirs;
  const int32_t* dir_idx;
  int dir_S;
} mcnerf_dirs;

/* workspace size in BYTES for n_rows rows (activations kept for the backward pass) */
size_t mcnerf_mlp_f32_workspace(const mcnerf_mlp_params* p, int n_rows);
/* fp32 CUDA-core path (exact-parity mode; any depth/width/skips).  out4 [n_rows,4] = (sigma_raw, r, g, b). */
int mcnerf_mlp_f32_fwd(const mcnerf_mlp_params* p, const float* x_enc, int ld_enc, const mcnerf_dirs* d,
                       int n_rows, const int32_t* n_rows_dev, float* out4, void* workspace, void* stream);
/* g_x_enc [n_rows, ld_enc] overwritten (may be NULL); g_dirs accumulated: [n_rows,3] or [n_rays,3] matching `d`
 * (may be NULL); parameter grads accumulated. */
int mcnerf_mlp_f32_bwd(const mcnerf_mlp_params* p, const float* x_enc, int ld_enc, const mcnerf_dirs* d,
                       int n_rows, const int32_t* n_rows_dev, const float* g_out4, void* workspace,
                       const mcnerf_mlp_grads* g, float* g_x_enc, float* g_dirs, void* stream);
/* eval_sh standalone, degree 0..4: sh [n,3,(deg+1)^2], dirs [n,3] -> out [n,3]  (ref: model/net_utils.py:103-191) */
int mcnerf_eval_sh_deg_fwd(int deg, const float* sh, const float* dirs, int n, float* out, void* stream);
int mcnerf_eval_sh_deg_bwd(int deg, const float* sh, const float* dirs, const float* g_out, int n, float* g_sh,
                           float* g_dirs, void* stream);
/* the degree-2 form (sh [n,3,9]) */
int mcnerf_eval_sh_fwd(const float* sh, const float* dirs, int n, float* out, void* stream);
int mcnerf_eval_sh_bwd(const float* sh, const float* dirs, const float* g_out, int n,
                       float* g_sh, float* g_dirs, void* stream);

/* ------------------------------------------------------------------ compositing (a13, a14)
 * ref: model/mc_nerf.py:705-736.  One warp per ray, warp-scan over samples.
 * z values: z_vals [B,S] if non-NULL else near + k*(far-near)/(S-1) (+ jitter[ray]); delta_last = 1e10.
 *  (i)  noise-free:  alpha = 1-exp(-softplus(sigma)*delta*|d|), T = exp(-excl.cumsum) -> depth, opacity
 *  (ii) noisy:       w = alpha' * excl.cumprod(1-alpha'+1e-10), alpha' = 1-exp(-delta*softplus(sigma+noise))
 *                    rgb = sum w c (+ 1 - sum w when white_back)
 * out4 is [B,S,4]; noise [B,S] or NULL (= 0). */
typedef struct {
  float near_, far_;
  int S;
  int white_back;
} mcnerf_composite_cfg;

int mcnerf_composite_fwd(const float* out4, const float* noise, const float* rays_d, const float* jitter,
                         const float* z_vals, int n_rays, const mcnerf_composite_cfg* cfg,
                         float* rgb /*[B,3]*/, float* depth /*[B] or NULL*/, float* opacity /*[B] or NULL*/,
                         float* weights /*[B,S] or NULL*/, void* stream);
/* g_out4 [B,S,4] overwritten.  Only d/d(rgb) is propagated (the reference discards depth/opacity in training,
 * ref: model/mc_nerf.py:590-591,646). */
int mcnerf_composite_bwd(const float* out4, const float* noise, const float* jitter, const float* z_vals,
                         int n_rays, const mcnerf_composite_cfg* cfg, const float* g_rgb,
                         float* g_out4, void* stream);
/* sigma2weights alone (ref: model/mc_nerf.py:729-736): sigma read with stride `sigma_stride` floats
 * (4 to read it out of out4, 1 for a dense [B,S] tensor); also folds max(w) into *w_max (atomic; caller zeroes). */
int mcnerf_sigma2weights(const float* sigma, int sigma_stride, const float* noise, const float* jitter,
                         const float* z_vals, const float* deltas /*[B,S] or NULL (explicit deltas)*/,
                         int n_rays, const mcnerf_composite_cfg* cfg,
                         float* weights, float* w_max, void* stream);

/* ------------------------------------------------------------------ fused tails + device RNG (a13-a15, perf mode)
 * ref: model/mc_nerf.py:601 (jitter uniform_), :619/:662/:719 (three torch.randn density-noise draws), :613-621,
 * :688-701, :705-736.  Noise arguments: an explicit [B,S] tensor (parity mode: the reference's draws replayed), or NULL
 * with `seed` = two device-side int64 words -> N(0,1) from Philox4x32-10 generated (and, in the backward pass,
 * re-generated) inside the kernel; both NULL = no noise.  Streams: 1 coarse colour, 2 selection, 3 fine colour, 4 jitter. */
/* out[i] = N(0,1) (normal != 0) or uniform in (lo, hi) of Philox stream `stream_id`, element i */
int mcnerf_philox_fill(const int64_t* seed, int stream_id, int64_t n, int normal, float lo, float hi, float* out,
                       void* stream);
/* coarse tail: ONE pass over out4 [B,S,4]: rgb [B,3] = noisy compositing (+ white background), w_sel [B,S] = selection
 * weights from an independent noise draw, *w_max = max(w_sel) folded in atomically (caller zeroes).  S <= 256. */
int mcnerf_coarse_tail_fwd(const float* out4, const float* noise_rgb, const float* noise_sel, const int64_t* seed,
                           const float* jitter, int n_rays, const mcnerf_composite_cfg* cfg, float* rgb, float* w_sel,
                           float* w_max, void* stream);
/* g_out4 [B,S,4] overwritten (same noise as the forward: noise_rgb, or stream 1 of `seed`) */
int mcnerf_coarse_tail_bwd(const float* out4, const float* noise_rgb, const int64_t* seed, const float* jitter,
                           int n_rays, const mcnerf_composite_cfg* cfg, const float* g_rgb, float* g_out4, void* stream);
/* fine tail: compositing of the fine grid (cfg->S = Sf samples) straight from the COMPACTED rows out_sel of the
 * selected samples, in the order mcnerf_select_fine emits them (ray r: sel_offsets[r] + rank * scale + j); unselected
 * samples are (sigma_default, 1, 1, 1).  w_sel [B,Sf/scale], *w_max, thresh as given to mcnerf_select_fine. */
int mcnerf_fine_tail_fwd(const float* out_sel, const float* w_sel, const float* w_max, float thresh, int scale,
                         const int32_t* sel_offsets, const float* rays_d, const float* jitter, const float* noise,
                         const int64_t* seed, int n_rays, const mcnerf_composite_cfg* cfg, float sigma_default,
                         float* rgb, float* depth /*[B] or NULL*/, float* opacity /*[B] or NULL*/, void* stream);
/* g_sel: gradients w.r.t. the compacted rows (rows of unselected samples do not exist; rows beyond the count untouched) */
int mcnerf_fine_tail_bwd(const float* out_sel, const float* w_sel, const float* w_max, float thresh, int scale,
                         const int32_t* sel_offsets, const float* jitter, const float* noise, const int64_t* seed,
                         int n_rays, const mcnerf_composite_cfg* cfg, float sigma_default, const float* g_rgb,
                         float* g_sel, void* stream);

/* ------------------------------------------------------------------ fine-sample selection (a15)
 * ref: model/mc_nerf.py:623-629, 663-667.  keep coarse sample (r,i) iff w[r,i] >= min(thresh, *w_max);
 * emits flat fine indices r*Sf + i*scale + j (ray-major, ascending - the order torch.nonzero gives),
 * *n_sel = count, sel_offsets[r] = first position of ray r (B+1 entries).  No host synchronisation. */
int mcnerf_select_fine(const float* weights, const float* w_max, int n_rays, int Sc, int scale, float thresh,
                       int32_t* sel_idx /*[B*Sc*scale] capacity*/, int32_t* sel_offsets /*[B+1]*/,
                       int32_t* n_sel, void* stream);

/* Train-only cap on the selected fine samples (ref: model/mc_nerf.py:630-632: `torch.randperm(n)[:K]` on the CPU after a
 * host synchronisation): out_idx[0 .. min(n,K)) = a uniformly random K-subset of sel_idx[0 .. n) (all of them when
 * n <= K), in ascending slot order; *n_out_dev = min(n, K).  n = min(*n_sel_dev, capacity).  Keys are a seeded bijection
 * of the slot index, the K-th smallest is found by a two-level radix select: no sort, no host synchronisation,
 * deterministic for a given `seed` (two words, as for mcnerf_philox_fill).  workspace: mcnerf_cap_select_workspace bytes. */
int mcnerf_cap_select_workspace(int capacity, size_t* bytes);
int mcnerf_cap_select(const int32_t* sel_idx, const int32_t* n_sel_dev, int capacity, int K, const int64_t* seed,
                      int32_t* out_idx, int32_t* n_out_dev, void* workspace, void* stream);
/* out_dense [B*Sf,4] = (sigma_default,1,1,1) everywhere, then out_dense[sel_idx[m]] = out_sel[m]
 * (ref: model/mc_nerf.py:692-694, 700-701). */
int mcnerf_scatter_fine(const float* out_sel, const int32_t* sel_idx, int n_sel, const int32_t* n_sel_dev,
                        int n_dense_rows, float sigma_default, float* out_dense, void* stream);
/* g_sel[m] = g_dense[sel_idx[m]] */
int mcnerf_gather_fine(const float* g_dense, const int32_t* sel_idx, int n_sel, const int32_t* n_sel_dev,
                       float* g_sel, void* stream);

/* ------------------------------------------------------------------ loss seed + optimiser (f-1, f-2)
 * ref: model/loss.py:33-43.  loss += mean((rgb_c-gt)^2) + mean((rgb_f-gt)^2) (atomic into *loss);
 * g_c = 2(rgb_c-gt)/(3B)*grad_scale, g_f likewise.  gt_idx (may be NULL) gathers gt rows. */
int mcnerf_rgb_loss(const float* rgb_c, const float* rgb_f, const float* gt, const int32_t* gt_idx, int n_rays,
                    float grad_scale, float* loss, float* g_c, float* g_f, void* stream);
/* ref: model/loss.py:15-58 (MC_NeRF_Loss.forward of the stages that render).  ONE single-block launch:
 *   l_px  = mean(((px_x-gt_x)/W)^2) + mean(((px_y-gt_y)/H)^2) over the n_pts reprojected calibration points
 *           (px == NULL: no reprojection term), divided by (l_px.detach() + 1e-8) when normalise != 0;
 *   l_rgb = mean((rgb_c-gt)^2) + mean((rgb_f-gt)^2) over n_rays*3 elements (rgb_f may be NULL);
 *   out3  = {l_px[/norm] + l_rgb, l_px, l_rgb};  g_c, g_f [n_rays,3], g_px [n_pts,2] = d out3[0] / d input. */
int mcnerf_train_loss(const float* rgb_c, const float* rgb_f, const float* gt, int n_rays, const float* px,
                      const float* px_gt, int n_pts, int img_w, int img_h, int normalise, float* out3,
                      float* g_c, float* g_f, float* g_px, void* stream);

/* ref: model/mc_nerf.py:327-345 (generate_rand_rays: torch.randperm(H*W)[:batch]).  The first min(batch, n)
 * entries of a uniform random permutation of [0, n) without sorting all n keys: Philox4x32-10 keyed by the two
 * device-side int64 `seed` words, threshold filter, one-block sort of the few thousand survivors.
 * workspace: mcnerf_sample_pixels_workspace() bytes, ZEROED once by the caller before the first call (the kernels
 * leave it zeroed).  out_idx int64 [min(batch,n)]; out_idx32 (optional) the same values as int32. */
int mcnerf_sample_pixels_workspace(int n, int batch, size_t* bytes);
int mcnerf_sample_pixels(int n, int batch, const int64_t* seed, void* workspace, int64_t* out_idx,
                         int32_t* out_idx32, void* stream);

/* dst[0..n) = host_values[0..n), n <= 16, passed by value in a one-block launch: stream-ordered refresh of a few
 * device-side scalars (mcnerf_sampling.band_w_dev) with no staging buffer for the host to race with. */
int mcnerf_store_floats(float* dst, const float* host_values, int n, void* stream);

/* dst_j[r*dst_ld_j + c] = src_j[r*src_ld_j + c] for r < rows_j, c < cols_j, j < n_jobs, in one launch per 64 jobs.
 * Used to zero-pad the parameters of a network narrower than 256 into 256-wide shadows for the tensor-core path
 * (zero rows/columns leave the arithmetic of ref: model/net_block.py:67-78 unchanged) and to cut the valid blocks
 * back out of the 256-wide gradients.  All seven arrays are HOST arrays of length n_jobs. */
int mcnerf_copy_blocks(int n_jobs, const float* const* src, float* const* dst, const int* rows, const int* cols,
                       const int* src_ld, const int* dst_ld, void* stream);

/* ref: model/net_utils.py:10-101 (RAdam.step) on one flat buffer.  The host evaluates the scalar schedule
 * (N_sma, step_size - incl. the 10-slot cache semantics) and passes it in:
 *   mode 1 (N_sma >= 5): p -= wd*lr*p ; p -= step_size*lr * m/(sqrt(v)+eps)
 *   mode 2 (degenerate): p -= wd*lr*p ; p -= step_size*lr * m
 *   mode 0: moments only. */
int mcnerf_radam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay,
                      float step_size, int mode, float grad_scale, void* stream);

/* The same update applied to n_tensors tensors sharing all scalars, in one launch per 64 tensors.
 * p, g, exp_avg, exp_avg_sq, numel are HOST arrays (of device pointers / element counts). */
int mcnerf_radam_multi(int n_tensors, float* const* p, const float* const* g, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float step_size, int mode, float grad_scale, void* stream);

/* ------------------------------------------------------------------ NeRF MLP, bf16 tcgen05 path (a8-a12 fused)
 * Throughput path for width-256 networks (depth 2..12, at most one input skip, L = 10, deg 2): sampling +
 * encoding + MLP + SH head fused in one persistent kernel; bf16 operands, fp32 accumulation in TMEM.
 * ref: model/mc_nerf.py:599-602,633-635 ; model/net_block.py:20-35,67-78 ; model/net_utils.py:154-169.
 * Weights are first packed (once per optimiser step) into UMMA-ready bf16 images. */
typedef struct {
  /* rays mode (x_enc == NULL): row m is sample sample_idx[m] (or m) = ray*S + k of the grid `smp` */
  const float *rays_o, *rays_d, *jitter;
  int n_rays;
  mcnerf_sampling smp;
  const int32_t* sample_idx;
  int n_rows;                    /* capacity / row count */
  const int32_t* n_rows_dev;     /* optional device-side row count (<= n_rows) */
  /* explicit mode: encodings x_enc [n_rows, ld_enc] (63 valid columns) and per-row view directions [n_rows,3] */
  const float* x_enc;
  int ld_enc;
  const float* dirs_rows;
  /* rays mode, backward only: when the rows of a ray are consecutive - no sample_idx (ray r owns rows [r*S, (r+1)*S)) or
   * ray_offsets [n_rays+1] given (ray r owns rows [ray_offsets[r], ray_offsets[r+1]), the order mcnerf_select_fine
   * emits) - set ordered_ray_grads = 1: per-ray gradients are then summed in a fixed order (bit-reproducible run to
   * run) instead of with fp32 atomics. */
  const int32_t* ray_offsets;
  int ordered_ray_grads;
} mcnerf_tc_input;

int mcnerf_mlp_tc_supported(const mcnerf_mlp_params* p);   /* 1 if this network shape can use the tcgen05 path */
int mcnerf_mlp_tc_pack_sizes(const mcnerf_mlp_params* p, size_t* wf_bytes, size_t* wb_bytes, size_t* bias_bytes);
/* wf: forward image, wb: transposed image for the backward pass (may be NULL), bias: fp32 bias block */
int mcnerf_mlp_tc_pack(const mcnerf_mlp_params* p, void* wf, void* wb, float* bias, void* stream);
/* bytes of the activation stash the training forward writes for n_rows rows */
size_t mcnerf_mlp_tc_stash_bytes(const mcnerf_mlp_params* p, int n_rows);
/* out4 [n_rows,4] = (sigma_raw, r, g, b).  stash == NULL: inference (activations never leave the SM). */
int mcnerf_mlp_tc_fwd(const mcnerf_mlp_params* p, const void* wf, const float* bias, const mcnerf_tc_input* in,
                      float* out4, void* stash, void* stream);

/* bytes of scratch the backward pass needs (dY stash + weight-gradient partials) */
size_t mcnerf_mlp_tc_bwd_workspace(const mcnerf_mlp_params* p, int n_rows);
/* Backward of mcnerf_mlp_tc_fwd (training forward with `stash`).  wb: transposed weight image from
 * mcnerf_mlp_tc_pack.  Parameter gradients are accumulated into `g`.  rays mode: accumulates dL/d(rays_o),
 * dL/d(rays_d) [n_rays,3]; explicit mode: overwrites g_x_enc [n_rows, ld_enc] and g_dirs_rows [n_rows,3]. */
int mcnerf_mlp_tc_bwd(const mcnerf_mlp_params* p, const void* wb, const float* bias, const mcnerf_tc_input* in,
                      const float* out4, const float* g_out4, const void* stash, void* workspace,
                      const mcnerf_mlp_grads* g, float* g_rays_o, float* g_rays_d, float* g_x_enc,
                      float* g_dirs_rows, void* stream);
/* Measurement aid (bench.py's per-kernel CUDA-event timing): select which phases the next mcnerf_mlp_tc_bwd calls
 * launch - bit 0: data-gradient chain (+ ray gradients), bit 1: weight/bias gradients.  Default 3 (both); the two
 * phases of one backward may be issued as two calls (1 then 2) with identical arguments.  Process-global. */
int mcnerf_mlp_tc_bwd_phases(int mask);


/* ------------------------------------------------------------------ gradient all-reduce over NVLink peer memory (e)
 * ref: main.py:61,84 (DistributedDataParallel's NCCL all-reduce of the MLP + camera gradients).  Two-shot all-reduce in
 * ONE kernel over symmetric buffers mapped into every rank of one NVSwitch box: buf[p] / flags[p] are rank p's buffer and
 * flag pad as seen from THIS process (peer pointers; the caller exchanges the handles, e.g. with
 * torch.distributed._symmetric_memory), flags zeroed once: n_ctas * 2 * 8 uint32 per rank; epoch: n_ctas zeroed uint32
 * in local memory.  buf[*][offset .. offset + count) = scale * sum over ranks, bit-identical on every rank.
 * Every rank must issue the same sequence of calls; the kernel needs its n_ctas CTAs resident on every rank. */
typedef struct {
  int rank, n_ranks, n_ctas;
  void* buf[8];
  void* flags[8];
  void* epoch;
} mcnerf_p2p;
int mcnerf_allreduce_p2p(const mcnerf_p2p* ctx, int64_t offset, int64_t count, float scale, void* stream);

/* ------------------------------------------------------------------ tensor-core self test
 * One 128xN tcgen05 tile: D = A B^T (mn_major = 0: A [128,K], B [N,K] bf16, K-major operands) or
 * D = A^T B (mn_major = 1: A [K,128], B [K,N], MN-major operands).  Used by tests/ to pin the UMMA
 * shared-memory / instruction descriptor conventions the fused MLP kernels rely on. */
int mcnerf_tc_selftest(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major, void* stream);
/* MMA issue-rate probe (one CTA): the K/16 MMAs repeated `reps` times on resident operands; cycles_out (device,
 * 2 x int64) = clock ticks to issue them / until the commit barrier fired. */
int mcnerf_tc_mma_rate(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major, int reps,
                       long long* cycles_out, void* stream);

/* CTA-pair tile: D[256,N] = A[256,K] B[N,K]^T through tcgen05.mma.cta_group::2 (cluster of 2 CTAs, each holding
 * 128 rows of A and N/2 rows of B); reps/cycles_out as in mcnerf_tc_mma_rate (cycles_out may be NULL).
 * bias (N floats, may be NULL): D += bias[n] through one extra MMA against a broadcast "ones" operand - the way the
 * fused forward kernel adds its biases. */
int mcnerf_tc_selftest2(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int reps,
                        long long* cycles_out, const float* bias, void* stream);

/* The same CTA-pair product with the A operand read from TENSOR MEMORY (tcgen05.mma "[a_tmem]" form): pins the TMEM
 * layout of a bf16 A operand (row = lane, one 32-bit column = the pair k even | k odd << 16).  n_split (low 2 bits) = 2
 * computes the product as two N/2-wide passes into the two column halves of the accumulator; bg > 0 adds background
 * tcgen05.ld/st traffic on unused columns (port-contention probe); cycles_out as in mcnerf_tc_mma_rate.
 * Probe-only flag bits of n_split (results then serve timing, not the product): bits 4-5 = 1 background loads only,
 * 2 stores only; bit 8 = the fused kernel's placement (A at columns [0,128), every pass accumulates at [128,256));
 * bit 9 = a tcgen05.commit after every pass. */
int mcnerf_tc_selftest_ts(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int n_split, int reps, int bg,
                          long long* cycles_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCNERF_H_ */

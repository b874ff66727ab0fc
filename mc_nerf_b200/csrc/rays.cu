// Ray generation from (camera, pixel) and its backward with the fused per-camera gradient reduction;
// stratified sampling + sin/cos positional encoding (fp32 stand-alone form) and its backward.
// ref: model/mc_nerf.py:124-145, 229-256 (get_rays), :599-602/633-635 (sampling),
//      model/net_block.py:20-35 (SinCosEmbedding).
#include "common.cuh"

namespace {

struct RayCam { float q[9]; float R[9]; float t[3]; };

__device__ __forceinline__ RayCam load_cam(const float* __restrict__ Kinv, const float* __restrict__ Rt, int c) {
  RayCam k;
#pragma unroll
  for (int i = 0; i < 9; ++i) k.q[i] = Kinv[9 * c + i];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int j = 0; j < 3; ++j) k.R[3 * r + j] = Rt[12 * c + 4 * r + j];
    k.t[r] = Rt[12 * c + 4 * r + 3];
  }
  return k;
}

__global__ void raygen_fwd_k(const float* __restrict__ Kinv, const float* __restrict__ Rt,
                             const int32_t* __restrict__ cam_id, int cam_const, const int32_t* __restrict__ pix,
                             int n, int img_w, float* __restrict__ ro, float* __restrict__ rd) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int c = cam_id ? cam_id[r] : cam_const;
  int p = pix ? pix[r] : r;
  RayCam k = load_cam(Kinv, Rt, c);
  float X = (float)(p % img_w) + 0.5f, Y = (float)(p / img_w) + 0.5f;
  // camera-space point p = Kinv (X, Y, 1)
  float pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc[i] = X * k.q[3 * i] + Y * k.q[3 * i + 1] + k.q[3 * i + 2];
  // o = -R^T t ; world = R^T p + o ; d = (world - o)/|.|  (same operation order as the reference)
  float o[3], v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o[j] = -(k.R[j] * k.t[0] + k.R[3 + j] * k.t[1] + k.R[6 + j] * k.t[2]);
    float wld = pc[0] * k.R[j] + pc[1] * k.R[3 + j] + pc[2] * k.R[6 + j] + o[j];
    v[j] = wld - o[j];
  }
  float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    ro[3 * r + j] = o[j];
    rd[3 * r + j] = v[j] * inv;
  }
}

// Backward: per ray 9 (Kinv) + 12 (Rt) partials; reduced inside the warp when the whole warp looks at one
// camera (the training case: one image per step), then one atomicAdd per warp and component.
__global__ void raygen_bwd_k(const float* __restrict__ Kinv, const float* __restrict__ Rt,
                             const int32_t* __restrict__ cam_id, int cam_const, const int32_t* __restrict__ pix,
                             int n, int img_w, const float* __restrict__ go, const float* __restrict__ gd,
                             float* __restrict__ gKinv, float* __restrict__ gRt) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = r < n;
  int rr = live ? r : n - 1;
  int c = cam_id ? cam_id[rr] : cam_const;
  int p = pix ? pix[rr] : rr;
  RayCam k = load_cam(Kinv, Rt, c);
  float X = (float)(p % img_w) + 0.5f, Y = (float)(p / img_w) + 0.5f;
  float pc[3], v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc[i] = X * k.q[3 * i] + Y * k.q[3 * i + 1] + k.q[3 * i + 2];
#pragma unroll
  for (int j = 0; j < 3; ++j) v[j] = pc[0] * k.R[j] + pc[1] * k.R[3 + j] + pc[2] * k.R[6 + j];
  float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  float d[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
  float g_d[3], g_o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    g_d[j] = live ? gd[3 * rr + j] : 0.f;
    g_o[j] = live ? go[3 * rr + j] : 0.f;
  }
  float dot = d[0] * g_d[0] + d[1] * g_d[1] + d[2] * g_d[2];
  float gv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gv[j] = (g_d[j] - d[j] * dot) * inv;
  float part[21];
  // v_j = sum_i R_ij p_i  ->  gR_ij = p_i gv_j ; gp_i = sum_j R_ij gv_j
  // o_j = -sum_i R_ij t_i ->  gR_ij -= t_i go_j ; gt_i = -sum_j R_ij go_j
  float gp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gp[i] = k.R[3 * i] * gv[0] + k.R[3 * i + 1] * gv[1] + k.R[3 * i + 2] * gv[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) part[9 + 4 * i + j] = pc[i] * gv[j] - k.t[i] * g_o[j];
    part[9 + 4 * i + 3] = -(k.R[3 * i] * g_o[0] + k.R[3 * i + 1] * g_o[1] + k.R[3 * i + 2] * g_o[2]);
  }
  // p_i = X q_i0 + Y q_i1 + q_i2
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    part[3 * i] = gp[i] * X;
    part[3 * i + 1] = gp[i] * Y;
    part[3 * i + 2] = gp[i];
  }
  int c0 = __shfl_sync(0xffffffffu, c, 0);
  bool uniform = __all_sync(0xffffffffu, c == c0);
  if (uniform) {
#pragma unroll
    for (int q = 0; q < 21; ++q) {
      float s = warp_sum(part[q]);
      if ((threadIdx.x & 31) == 0) {
        if (q < 9) atomicAdd(gKinv + 9 * c0 + q, s);
        else atomicAdd(gRt + 12 * c0 + (q - 9), s);
      }
    }
  } else if (live) {
#pragma unroll
    for (int q = 0; q < 21; ++q) {
      if (q < 9) atomicAdd(gKinv + 9 * c + q, part[q]);
      else atomicAdd(gRt + 12 * c + (q - 9), part[q]);
    }
  }
}

// ---------------------------------------------------------------- encoding
// One thread per (row, coordinate): writes x_c, L sines and L cosines.
// Accurate sincosf (arguments reach |x|*2^9 ~ 4000 rad); never the fast intrinsics.
__device__ __forceinline__ void encode_coord(float xc, int L, const float* bw, float* __restrict__ row, int c) {
  row[c] = xc;
  float* s = row + 3 + c * 2 * L;
  float f = 1.f;
  for (int k = 0; k < L; ++k) {
    float sn, cs;
    sincosf(xc * f, &sn, &cs);
    s[k] = sn * bw[k];
    s[L + k] = cs * bw[k];
    f *= 2.f;
  }
}

__device__ __forceinline__ float encode_coord_bwd(float xc, int L, const float* bw, const float* __restrict__ grow, int c) {
  float g = grow[c];
  const float* gs = grow + 3 + c * 2 * L;
  float f = 1.f;
  for (int k = 0; k < L; ++k) {
    float sn, cs;
    sincosf(xc * f, &sn, &cs);
    g += bw[k] * f * (gs[k] * cs - gs[L + k] * sn);
    f *= 2.f;
  }
  return g;
}

__global__ void encode_points_fwd_k(const float* __restrict__ x, int n, mcnerf_sampling smp, float* __restrict__ enc,
                                    int ld) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int m = t / 3, c = t - 3 * m;
  encode_coord(x[t], smp.n_freqs, (smp.band_w_dev ? smp.band_w_dev : smp.band_w), enc + (size_t)m * ld, c);
}

__global__ void encode_points_bwd_k(const float* __restrict__ x, int n, mcnerf_sampling smp,
                                    const float* __restrict__ genc, int ld, float* __restrict__ gx) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int m = t / 3, c = t - 3 * m;
  gx[t] = encode_coord_bwd(x[t], smp.n_freqs, (smp.band_w_dev ? smp.band_w_dev : smp.band_w), genc + (size_t)m * ld, c);
}

__device__ __forceinline__ float sample_z(const mcnerf_sampling& smp, int k, float jit) {
  return linspace_f(smp.near_, smp.far_, smp.S, k) + jit;
}

__global__ void encode_rays_fwd_k(const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ jitter, mcnerf_sampling smp,
                                  const int32_t* __restrict__ sidx, int n_rows, const int32_t* __restrict__ n_rows_dev,
                                  float* __restrict__ enc, int ld) {
  int rows = n_rows_dev ? min(*n_rows_dev, n_rows) : n_rows;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * rows) return;
  int m = t / 3, c = t - 3 * m;
  int flat = sidx ? sidx[m] : m;
  int ray = flat / smp.S, k = flat - ray * smp.S;
  float z = sample_z(smp, k, jitter ? jitter[ray] : 0.f);
  float xc = ro[3 * ray + c] + rd[3 * ray + c] * z;
  encode_coord(xc, smp.n_freqs, (smp.band_w_dev ? smp.band_w_dev : smp.band_w), enc + (size_t)m * ld, c);
}

// One thread per row: gradient wrt the sample position, then a segmented warp reduction over rows that
// share a ray (rows are ray-major) and one atomicAdd per (segment, component).
__global__ void encode_rays_bwd_k(const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ jitter, mcnerf_sampling smp,
                                  const int32_t* __restrict__ sidx, int n_rows, const int32_t* __restrict__ n_rows_dev,
                                  const float* __restrict__ genc, int ld, float* __restrict__ gro,
                                  float* __restrict__ grd) {
  int rows = n_rows_dev ? min(*n_rows_dev, n_rows) : n_rows;
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = m < rows;
  int ray = -1 - (int)(threadIdx.x & 31);   // distinct dead ids: never merged
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
    int flat = sidx ? sidx[m] : m;
    ray = flat / smp.S;
    int k = flat - ray * smp.S;
    float z = sample_z(smp, k, jitter ? jitter[ray] : 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float xc = ro[3 * ray + c] + rd[3 * ray + c] * z;
      float g = encode_coord_bwd(xc, smp.n_freqs, (smp.band_w_dev ? smp.band_w_dev : smp.band_w), genc + (size_t)m * ld, c);
      v[c] = g;
      v[3 + c] = g * z;
    }
  }
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int r2 = __shfl_down_sync(0xffffffffu, ray, o);
    bool take = (lane + o < 32) && (r2 == ray);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      float x2 = __shfl_down_sync(0xffffffffu, v[q], o);
      if (take) v[q] += x2;
    }
  }
  int rprev = __shfl_up_sync(0xffffffffu, ray, 1);
  bool head = (lane == 0) || (rprev != ray);
  if (live && head) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicAdd(gro + 3 * ray + c, v[c]);
      atomicAdd(grd + 3 * ray + c, v[3 + c]);
    }
  }
}

}  // namespace

extern "C" int mcnerf_raygen_fwd(const float* Kinv, const float* Rt, const int32_t* cam_id, int cam_const,
                                 const int32_t* pix, int n_rays, int img_w, float* rays_o, float* rays_d, void* stream) {
  MC_ARG(Kinv && Rt && rays_o && rays_d && n_rays > 0 && img_w > 0);
  raygen_fwd_k<<<cdiv(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(Kinv, Rt, cam_id, cam_const, pix, n_rays, img_w,
                                                                     rays_o, rays_d);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_raygen_bwd(const float* Kinv, const float* Rt, const int32_t* cam_id, int cam_const,
                                 const int32_t* pix, int n_rays, int img_w, const float* g_rays_o,
                                 const float* g_rays_d, float* gKinv, float* gRt, void* stream) {
  MC_ARG(Kinv && Rt && g_rays_o && g_rays_d && gKinv && gRt && n_rays > 0 && img_w > 0);
  raygen_bwd_k<<<cdiv(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(Kinv, Rt, cam_id, cam_const, pix, n_rays, img_w,
                                                                     g_rays_o, g_rays_d, gKinv, gRt);
  MC_LAUNCHED();
  return 0;
}

static int check_sampling(const mcnerf_sampling* s) {
  MC_ARG(s && s->S >= 2 && s->n_freqs >= 1 && s->n_freqs <= MCNERF_MAX_FREQS);
  return 0;
}

extern "C" int mcnerf_encode_rays_fwd(const float* rays_o, const float* rays_d, const float* jitter, int n_rays,
                                      const mcnerf_sampling* smp, const int32_t* sample_idx, int n_rows,
                                      const int32_t* n_rows_dev, float* enc, int ld_enc, void* stream) {
  if (int e = check_sampling(smp)) return e;
  MC_ARG(rays_o && rays_d && enc && n_rays > 0 && n_rows >= 0 && ld_enc >= 3 + 6 * smp->n_freqs);
  if (n_rows == 0) return 0;
  encode_rays_fwd_k<<<cdiv(3 * (int64_t)n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, jitter, *smp, sample_idx, n_rows, n_rows_dev, enc, ld_enc);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_encode_rays_bwd(const float* rays_o, const float* rays_d, const float* jitter, int n_rays,
                                      const mcnerf_sampling* smp, const int32_t* sample_idx, int n_rows,
                                      const int32_t* n_rows_dev, const float* g_enc, int ld_enc, float* g_rays_o,
                                      float* g_rays_d, void* stream) {
  if (int e = check_sampling(smp)) return e;
  MC_ARG(rays_o && rays_d && g_enc && g_rays_o && g_rays_d && n_rays > 0 && n_rows >= 0 &&
         ld_enc >= 3 + 6 * smp->n_freqs);
  if (n_rows == 0) return 0;
  encode_rays_bwd_k<<<cdiv(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, jitter, *smp, sample_idx,
                                                                          n_rows, n_rows_dev, g_enc, ld_enc, g_rays_o,
                                                                          g_rays_d);
  MC_LAUNCHED();
  return 0;
}

static int points_sampling(int n_freqs, const float* band_w_host, mcnerf_sampling* s) {
  MC_ARG(n_freqs >= 1 && n_freqs <= MCNERF_MAX_FREQS);
  s->near_ = 0.f; s->far_ = 1.f; s->S = 2; s->n_freqs = n_freqs; s->band_w_dev = nullptr;
  for (int k = 0; k < MCNERF_MAX_FREQS; ++k) s->band_w[k] = (band_w_host && k < n_freqs) ? band_w_host[k] : 1.f;
  return 0;
}

extern "C" int mcnerf_encode_points_fwd(const float* x, int n, int n_freqs, const float* band_w_host, float* enc,
                                        int ld_enc, void* stream) {
  mcnerf_sampling s;
  if (int e = points_sampling(n_freqs, band_w_host, &s)) return e;
  if (n == 0) return 0;
  MC_ARG(x && enc && n > 0 && ld_enc >= 3 + 6 * n_freqs);
  encode_points_fwd_k<<<cdiv(3 * (int64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, s, enc, ld_enc);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_encode_points_bwd(const float* x, int n, int n_freqs, const float* band_w_host,
                                        const float* g_enc, int ld_enc, float* g_x, void* stream) {
  mcnerf_sampling s;
  if (int e = points_sampling(n_freqs, band_w_host, &s)) return e;
  if (n == 0) return 0;
  MC_ARG(x && g_enc && g_x && n > 0 && ld_enc >= 3 + 6 * n_freqs);
  encode_points_bwd_k<<<cdiv(3 * (int64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, s, g_enc, ld_enc, g_x);
  MC_LAUNCHED();
  return 0;
}

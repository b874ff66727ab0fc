// Loss seed (MSE on the two renders) and the fused multi-tensor RAdam update on flat buffers.
// ref: model/loss.py:33-43 ; model/net_utils.py:10-101.
#include "common.cuh"

namespace {

__global__ void rgb_loss_k(const float* __restrict__ rc, const float* __restrict__ rf, const float* __restrict__ gt,
                           const int32_t* __restrict__ gt_idx, int n_rays, float grad_scale, float* __restrict__ loss,
                           float* __restrict__ gc, float* __restrict__ gf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 3 * n_rays;
  float part = 0.f;
  if (t < n) {
    int r = t / 3, c = t - 3 * r;
    float g = gt_idx ? gt[3 * (size_t)gt_idx[r] + c] : gt[t];
    float dc = rc[t] - g, df = rf[t] - g;
    float k = 2.f / (float)n * grad_scale;
    gc[t] = k * dc;
    gf[t] = k * df;
    part = (dc * dc + df * df) / (float)n;
  }
  part = warp_sum(part);
  __shared__ float sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = part;
  __syncthreads();
  if (wid == 0) {
    float v = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0 && loss) atomicAdd(loss, v);
  }
}

__global__ void radam_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                        float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                        float step_size, int mode, float gscale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * gscale;
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    float mi = m[i] * b1 + (1.f - b1) * gi;
    v[i] = vi;
    m[i] = mi;
    if (mode == 0) continue;
    float pi = p[i];
    if (wd != 0.f) pi += pi * (-wd * lr);
    if (mode == 1) pi += (-step_size * lr) * (mi / (sqrtf(vi) + eps));
    else pi += (-step_size * lr) * mi;
    p[i] = pi;
  }
}

constexpr int RADAM_MAX_TENSORS = 64;
struct RadamMulti {
  int n;
  float* p[RADAM_MAX_TENSORS];
  const float* g[RADAM_MAX_TENSORS];
  float* m[RADAM_MAX_TENSORS];
  float* v[RADAM_MAX_TENSORS];
  int64_t numel[RADAM_MAX_TENSORS];
  int blk_prefix[RADAM_MAX_TENSORS + 1];     // prefix of 256-element blocks per tensor
};

// One launch for a whole parameter group: block b serves 256 consecutive elements of one tensor.
__global__ void __launch_bounds__(256) radam_multi_k(const __grid_constant__ RadamMulti a, float lr, float b1, float b2,
                                                     float eps, float wd, float step_size, int mode, float gscale) {
  int t = 0;
  while ((int)blockIdx.x >= a.blk_prefix[t + 1]) ++t;
  const int64_t i = (int64_t)((int)blockIdx.x - a.blk_prefix[t]) * 256 + threadIdx.x;
  if (i >= a.numel[t]) return;
  float* p = a.p[t];
  float* m = a.m[t];
  float* v = a.v[t];
  const float gi = a.g[t][i] * gscale;
  const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
  const float mi = m[i] * b1 + (1.f - b1) * gi;
  v[i] = vi;
  m[i] = mi;
  if (mode == 0) return;
  float pi = p[i];
  if (wd != 0.f) pi += pi * (-wd * lr);
  if (mode == 1) pi += (-step_size * lr) * (mi / (sqrtf(vi) + eps));
  else pi += (-step_size * lr) * mi;
  p[i] = pi;
}

}  // namespace

extern "C" int mcnerf_rgb_loss(const float* rgb_c, const float* rgb_f, const float* gt, const int32_t* gt_idx,
                               int n_rays, float grad_scale, float* loss, float* g_c, float* g_f, void* stream) {
  MC_ARG(rgb_c && rgb_f && gt && g_c && g_f && n_rays > 0);
  rgb_loss_k<<<cdiv(3 * (int64_t)n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rgb_c, rgb_f, gt, gt_idx, n_rays,
                                                                                 grad_scale, loss, g_c, g_f);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_radam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, float step_size, int mode,
                                 float grad_scale, void* stream) {
  MC_ARG(p && g && exp_avg && exp_avg_sq && n > 0 && mode >= 0 && mode <= 2);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  radam_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                     step_size, mode, grad_scale);
  MC_LAUNCHED();
  return 0;
}

// Same update for up to 64 tensors that share (lr, betas, eps, weight decay, step) in ONE launch; longer lists
// are processed in slices.  p/g/m/v are HOST arrays of device pointers.
extern "C" int mcnerf_radam_multi(int n_tensors, float* const* p, const float* const* g, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, float step_size, int mode, float grad_scale,
                                  void* stream) {
  MC_ARG(n_tensors >= 0 && mode >= 0 && mode <= 2);
  if (n_tensors == 0) return 0;
  MC_ARG(p && g && exp_avg && exp_avg_sq && numel);
  for (int base = 0; base < n_tensors; base += RADAM_MAX_TENSORS) {
    RadamMulti a;
    a.n = n_tensors - base < RADAM_MAX_TENSORS ? n_tensors - base : RADAM_MAX_TENSORS;
    int blocks = 0;
    for (int t = 0; t < a.n; ++t) {
      MC_ARG(p[base + t] && g[base + t] && exp_avg[base + t] && exp_avg_sq[base + t] && numel[base + t] > 0);
      a.p[t] = p[base + t]; a.g[t] = g[base + t]; a.m[t] = exp_avg[base + t]; a.v[t] = exp_avg_sq[base + t];
      a.numel[t] = numel[base + t];
      a.blk_prefix[t] = blocks;
      blocks += (int)((numel[base + t] + 255) / 256);
    }
    for (int t = a.n; t <= RADAM_MAX_TENSORS; ++t) a.blk_prefix[t] = blocks;
    radam_multi_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, lr, beta1, beta2, eps, weight_decay, step_size, mode,
                                                             grad_scale);
    MC_LAUNCHED();
  }
  return 0;
}

// Alpha compositing along rays: one warp per ray, warp-scan over the samples, coalesced float4 access.
// ref: model/mc_nerf.py:705-736 (tail of NeRF_Model.inference, sigma2weights).
// HBM-bound: algorithmic bytes per sample = 16 (out4) + 4 (noise) read, (+4 weights written when asked);
// backward = 20 read + 16 written.
#include "common.cuh"

namespace {

constexpr int WARPS = 4;

__device__ __forceinline__ float scan_add_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float scan_mul_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

struct ZSrc {
  const float* z_vals;   // [B,S] or null
  const float* deltas;   // [B,S] or null (explicit deltas, sigma2weights API)
  float near_, far_, jit;
  int S;
  size_t row;
  __device__ __forceinline__ float z(int k) const {
    return z_vals ? z_vals[row + k] : linspace_f(near_, far_, S, k) + jit;
  }
  __device__ __forceinline__ float delta(int k) const {
    if (deltas) return deltas[row + k];
    return (k == S - 1) ? 1e10f : z(k + 1) - z(k);
  }
};

__device__ __forceinline__ ZSrc make_z(const float* z_vals, const float* deltas, const float* jitter, int ray,
                                       const mcnerf_composite_cfg& cfg) {
  ZSrc s;
  s.z_vals = z_vals; s.deltas = deltas; s.near_ = cfg.near_; s.far_ = cfg.far_; s.S = cfg.S;
  s.jit = jitter ? jitter[ray] : 0.f;
  s.row = (size_t)ray * cfg.S;
  return s;
}

__global__ void __launch_bounds__(WARPS * 32)
composite_fwd_k(const float4* __restrict__ out4, const float* __restrict__ noise, const float* __restrict__ rays_d,
                const float* __restrict__ jitter, const float* __restrict__ z_vals, int n_rays,
                mcnerf_composite_cfg cfg, float* __restrict__ rgb, float* __restrict__ depth,
                float* __restrict__ opacity, float* __restrict__ weights) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  ZSrc zs = make_z(z_vals, nullptr, jitter, ray, cfg);
  float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  float len = sqrtf(dx * dx + dy * dy + dz * dz);
  float tau_carry = 0.f;    // sum of tau over previous chunks
  float T_carry = 1.f;      // product of (1-alpha'+1e-10) over previous chunks
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_op = 0.f, acc_dp = 0.f;
  for (int base = 0; base < S; base += 32) {
    int k = base + lane;
    bool live = k < S;
    float4 o = live ? out4[zs.row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    float zk = live ? zs.z(k) : 0.f;
    float dl = live ? zs.delta(k) : 0.f;
    // (i) noise-free transmittance
    float tau = live ? softplus_f(o.x) * (dl * len) : 0.f;
    float tau_inc = scan_add_incl(tau, lane);
    // exclusive prefix by shifting, NOT tau_inc - tau: the last sample's tau is ~1e10 and would cancel
    float tau_exc = __shfl_up_sync(0xffffffffu, tau_inc, 1);
    if (lane == 0) tau_exc = 0.f;
    float T = expf(-(tau_carry + tau_exc));
    float pa = live ? T * (1.f - expf(-tau)) : 0.f;
    acc_op += pa;
    acc_dp += pa * zk;
    tau_carry += __shfl_sync(0xffffffffu, tau_inc, 31);
    // (ii) noisy weights
    float nz = (live && noise) ? noise[zs.row + k] : 0.f;
    float al = live ? 1.f - expf(-dl * softplus_f(o.x + nz)) : 0.f;
    float f = live ? (1.f - al + 1e-10f) : 1.f;
    float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    float w = al * (T_carry * f_exc);
    T_carry *= __shfl_sync(0xffffffffu, f_inc, 31);
    if (live && weights) weights[zs.row + k] = w;
    acc_w += w;
    acc_r += w * o.y;
    acc_g += w * o.z;
    acc_b += w * o.w;
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b); acc_w = warp_sum(acc_w);
  acc_op = warp_sum(acc_op); acc_dp = warp_sum(acc_dp);
  if (lane == 0) {
    float bg = cfg.white_back ? 1.f - acc_w : 0.f;
    rgb[3 * ray] = acc_r + bg;
    rgb[3 * ray + 1] = acc_g + bg;
    rgb[3 * ray + 2] = acc_b + bg;
    if (depth) depth[ray] = acc_dp;
    if (opacity) opacity[ray] = acc_op;
  }
}

// Backward of the noisy path.  With q_k = g.c_k - sum(g) (white background) and f_k = 1-alpha_k+1e-10:
//   dL/dalpha_k = T_k (q_k - R_k),  R_{k-1} = q_k alpha_k + f_k R_k,  R_{S-1} = 0   (division-free)
//   dL/dsigma_k = dL/dalpha_k (1-alpha_k) delta_k sigmoid(sigma_k + n_k),  dL/dc_k = w_k g.
// Pass 1 walks the chunks forward to get each chunk's transmittance carry, pass 2 walks them backward
// running the reverse affine scan.  S <= 32*MAXC.
constexpr int MAXC = 32;

__global__ void __launch_bounds__(WARPS * 32)
composite_bwd_k(const float4* __restrict__ out4, const float* __restrict__ noise, const float* __restrict__ jitter,
                const float* __restrict__ z_vals, int n_rays, mcnerf_composite_cfg cfg,
                const float* __restrict__ g_rgb, float4* __restrict__ g_out4) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  ZSrc zs = make_z(z_vals, nullptr, jitter, ray, cfg);
  float gr = g_rgb[3 * ray], gg = g_rgb[3 * ray + 1], gb = g_rgb[3 * ray + 2];
  float gsum = cfg.white_back ? (gr + gg + gb) : 0.f;
  int nchunk = (S + 31) / 32;
  // pass 1: product of f over each chunk (lane c keeps chunk c's incoming carry)
  float carry_in = 1.f;     // lane c: T at the start of chunk c
  {
    float run = 1.f;
    for (int c = 0; c < nchunk; ++c) {
      int k = c * 32 + lane;
      bool live = k < S;
      float f = 1.f;
      if (live) {
        float sg = out4[zs.row + k].x + (noise ? noise[zs.row + k] : 0.f);
        float al = 1.f - expf(-zs.delta(k) * softplus_f(sg));
        f = 1.f - al + 1e-10f;
      }
      float f_inc = scan_mul_incl(f, lane);
      if (lane == c) carry_in = run;
      run *= __shfl_sync(0xffffffffu, f_inc, 31);
    }
  }
  float R_carry = 0.f;
  for (int c = nchunk - 1; c >= 0; --c) {
    int k = c * 32 + lane;
    bool live = k < S;
    float4 o = live ? out4[zs.row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    float dl = live ? zs.delta(k) : 0.f;
    float sg = o.x + ((live && noise) ? noise[zs.row + k] : 0.f);
    float om = live ? expf(-dl * softplus_f(sg)) : 1.f;   // 1 - alpha
    float al = 1.f - om;
    float f = live ? (1.f - al + 1e-10f) : 1.f;           // rounded exactly as the forward pass does
    float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    float T = __shfl_sync(0xffffffffu, carry_in, c) * f_exc;
    float q = live ? (gr * o.y + gg * o.z + gb * o.w - gsum) : 0.f;
    // reverse inclusive affine scan of R -> a + f R
    float a = q * al, ff = f;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      float a2 = __shfl_down_sync(0xffffffffu, a, s);
      float f2 = __shfl_down_sync(0xffffffffu, ff, s);
      if (lane + s < 32) {
        a = a + ff * a2;
        ff = ff * f2;
      }
    }
    float E = a + ff * R_carry;                  // phi_k o ... o phi_last (0)
    float R = __shfl_down_sync(0xffffffffu, E, 1);
    if (lane == 31) R = R_carry;
    R_carry = __shfl_sync(0xffffffffu, E, 0);
    if (live) {
      float w = al * T;
      float dalpha = T * (q - R);
      float dsig = dalpha * om * dl * sigmoid_f(sg);
      g_out4[zs.row + k] = make_float4(dsig, w * gr, w * gg, w * gb);
    }
  }
}

__global__ void __launch_bounds__(WARPS * 32)
sigma2weights_k(const float* __restrict__ sigma, int stride, const float* __restrict__ noise,
                const float* __restrict__ jitter, const float* __restrict__ z_vals, const float* __restrict__ deltas,
                int n_rays, mcnerf_composite_cfg cfg, float* __restrict__ weights, float* __restrict__ w_max) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  float wm = 0.f;
  if (ray < n_rays) {
    const int S = cfg.S;
    ZSrc zs = make_z(z_vals, deltas, jitter, ray, cfg);
    float T_carry = 1.f;
    for (int base = 0; base < S; base += 32) {
      int k = base + lane;
      bool live = k < S;
      float al = 0.f;
      if (live) {
        float sg = sigma[(zs.row + k) * stride] + (noise ? noise[zs.row + k] : 0.f);
        al = 1.f - expf(-zs.delta(k) * softplus_f(sg));
      }
      float f = live ? (1.f - al + 1e-10f) : 1.f;
      float f_inc = scan_mul_incl(f, lane);
      float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
      if (lane == 0) f_exc = 1.f;
      float w = al * (T_carry * f_exc);
      T_carry *= __shfl_sync(0xffffffffu, f_inc, 31);
      if (live) {
        weights[zs.row + k] = w;
        wm = fmaxf(wm, w);
      }
    }
  }
  if (w_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    // weights are >= 0, so the int ordering of the bit pattern is the float ordering
    if (lane == 0 && ray < n_rays) atomicMax((int*)w_max, __float_as_int(wm));
  }
}

// ---- register-cached variants for S <= 32*NC (NC <= 8): every load of the ray is issued up front (maximum
// memory-level parallelism per warp) and each byte is read exactly once - the generic backward reads out4 twice
// (pass 1 only needs sigma, a 4-of-16-byte strided access).  Same arithmetic, same order, bit-identical results.
template <int NC>
__global__ void __launch_bounds__(WARPS * 32)
composite_fwd_c_k(const float4* __restrict__ out4, const float* __restrict__ noise, const float* __restrict__ rays_d,
                  const float* __restrict__ jitter, const float* __restrict__ z_vals, int n_rays,
                  mcnerf_composite_cfg cfg, float* __restrict__ rgb, float* __restrict__ depth,
                  float* __restrict__ opacity, float* __restrict__ weights) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  ZSrc zs = make_z(z_vals, nullptr, jitter, ray, cfg);
  float4 oc[NC];
  float nc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int k = c * 32 + lane;
    bool live = k < S;
    oc[c] = live ? out4[zs.row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    nc[c] = (live && noise) ? noise[zs.row + k] : 0.f;
  }
  float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  float len = sqrtf(dx * dx + dy * dy + dz * dz);
  float tau_carry = 0.f, T_carry = 1.f;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_op = 0.f, acc_dp = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c * 32 >= S) break;
    int k = c * 32 + lane;
    bool live = k < S;
    float4 o = oc[c];
    float zk = live ? zs.z(k) : 0.f;
    float dl = live ? zs.delta(k) : 0.f;
    float tau = live ? softplus_f(o.x) * (dl * len) : 0.f;
    float tau_inc = scan_add_incl(tau, lane);
    float tau_exc = __shfl_up_sync(0xffffffffu, tau_inc, 1);
    if (lane == 0) tau_exc = 0.f;
    float T = expf(-(tau_carry + tau_exc));
    float pa = live ? T * (1.f - expf(-tau)) : 0.f;
    acc_op += pa;
    acc_dp += pa * zk;
    tau_carry += __shfl_sync(0xffffffffu, tau_inc, 31);
    float al = live ? 1.f - expf(-dl * softplus_f(o.x + nc[c])) : 0.f;
    float f = live ? (1.f - al + 1e-10f) : 1.f;
    float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    float w = al * (T_carry * f_exc);
    T_carry *= __shfl_sync(0xffffffffu, f_inc, 31);
    if (live && weights) weights[zs.row + k] = w;
    acc_w += w;
    acc_r += w * o.y;
    acc_g += w * o.z;
    acc_b += w * o.w;
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b); acc_w = warp_sum(acc_w);
  acc_op = warp_sum(acc_op); acc_dp = warp_sum(acc_dp);
  if (lane == 0) {
    float bg = cfg.white_back ? 1.f - acc_w : 0.f;
    rgb[3 * ray] = acc_r + bg;
    rgb[3 * ray + 1] = acc_g + bg;
    rgb[3 * ray + 2] = acc_b + bg;
    if (depth) depth[ray] = acc_dp;
    if (opacity) opacity[ray] = acc_op;
  }
}

template <int NC>
__global__ void __launch_bounds__(WARPS * 32)
composite_bwd_c_k(const float4* __restrict__ out4, const float* __restrict__ noise, const float* __restrict__ jitter,
                  const float* __restrict__ z_vals, int n_rays, mcnerf_composite_cfg cfg,
                  const float* __restrict__ g_rgb, float4* __restrict__ g_out4) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  ZSrc zs = make_z(z_vals, nullptr, jitter, ray, cfg);
  float4 oc[NC];
  float sgc[NC];        // sigma + noise
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int k = c * 32 + lane;
    bool live = k < S;
    oc[c] = live ? out4[zs.row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    sgc[c] = oc[c].x + ((live && noise) ? noise[zs.row + k] : 0.f);
  }
  float gr = g_rgb[3 * ray], gg = g_rgb[3 * ray + 1], gb = g_rgb[3 * ray + 2];
  float gsum = cfg.white_back ? (gr + gg + gb) : 0.f;
  // pass 1: per chunk 1-alpha, f and the transmittance entering the chunk
  float omc[NC], Tin[NC];
  {
    float run = 1.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int k = c * 32 + lane;
      bool live = k < S;
      omc[c] = live ? expf(-zs.delta(k) * softplus_f(sgc[c])) : 1.f;
      float f = live ? (1.f - (1.f - omc[c]) + 1e-10f) : 1.f;
      float f_inc = scan_mul_incl(f, lane);
      Tin[c] = run;
      run *= __shfl_sync(0xffffffffu, f_inc, 31);
    }
  }
  float R_carry = 0.f;
#pragma unroll
  for (int c = NC - 1; c >= 0; --c) {
    if (c * 32 >= S) continue;
    int k = c * 32 + lane;
    bool live = k < S;
    float4 o = oc[c];
    float dl = live ? zs.delta(k) : 0.f;
    float om = omc[c];
    float al = 1.f - om;
    float f = live ? (1.f - al + 1e-10f) : 1.f;
    float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    float T = Tin[c] * f_exc;
    float q = live ? (gr * o.y + gg * o.z + gb * o.w - gsum) : 0.f;
    float a = q * al, ff = f;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      float a2 = __shfl_down_sync(0xffffffffu, a, s);
      float f2 = __shfl_down_sync(0xffffffffu, ff, s);
      if (lane + s < 32) {
        a = a + ff * a2;
        ff = ff * f2;
      }
    }
    float E = a + ff * R_carry;
    float R = __shfl_down_sync(0xffffffffu, E, 1);
    if (lane == 31) R = R_carry;
    R_carry = __shfl_sync(0xffffffffu, E, 0);
    if (live) {
      float w = al * T;
      float dalpha = T * (q - R);
      float dsig = dalpha * om * dl * sigmoid_f(sgc[c]);
      g_out4[zs.row + k] = make_float4(dsig, w * gr, w * gg, w * gb);
    }
  }
}

}  // namespace

static int check_cfg(const mcnerf_composite_cfg* cfg) {
  MC_ARG(cfg && cfg->S >= 2 && cfg->S <= 32 * MAXC);
  return 0;
}

extern "C" int mcnerf_composite_fwd(const float* out4, const float* noise, const float* rays_d, const float* jitter,
                                    const float* z_vals, int n_rays, const mcnerf_composite_cfg* cfg, float* rgb,
                                    float* depth, float* opacity, float* weights, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  MC_ARG(out4 && rays_d && rgb && n_rays > 0 && ((uintptr_t)out4 & 15) == 0);
  const int nc = (cfg->S + 31) / 32;
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define FWD_C(NC) composite_fwd_c_k<NC><<<grid, block, 0, st>>>((const float4*)out4, noise, rays_d, jitter, z_vals, n_rays, *cfg, rgb, depth, opacity, weights)
  // beyond 8 chunks the cached registers cost more occupancy than the single pass saves (measured at S = 320)
  if (nc <= 2) FWD_C(2); else if (nc <= 4) FWD_C(4); else if (nc <= 6) FWD_C(6); else if (nc <= 8) FWD_C(8);
  else composite_fwd_k<<<grid, block, 0, st>>>((const float4*)out4, noise, rays_d, jitter, z_vals, n_rays, *cfg, rgb, depth,
                                                opacity, weights);
#undef FWD_C
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_composite_bwd(const float* out4, const float* noise, const float* jitter, const float* z_vals,
                                    int n_rays, const mcnerf_composite_cfg* cfg, const float* g_rgb, float* g_out4,
                                    void* stream) {
  if (int e = check_cfg(cfg)) return e;
  MC_ARG(out4 && g_rgb && g_out4 && n_rays > 0 && ((uintptr_t)out4 & 15) == 0 && ((uintptr_t)g_out4 & 15) == 0);
  const int nc = (cfg->S + 31) / 32;
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define BWD_C(NC) composite_bwd_c_k<NC><<<grid, block, 0, st>>>((const float4*)out4, noise, jitter, z_vals, n_rays, *cfg, g_rgb, (float4*)g_out4)
  if (nc <= 2) BWD_C(2); else if (nc <= 4) BWD_C(4); else if (nc <= 6) BWD_C(6); else if (nc <= 8) BWD_C(8);
  else composite_bwd_k<<<grid, block, 0, st>>>((const float4*)out4, noise, jitter, z_vals, n_rays, *cfg, g_rgb, (float4*)g_out4);
#undef BWD_C
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_sigma2weights(const float* sigma, int sigma_stride, const float* noise, const float* jitter,
                                    const float* z_vals, const float* deltas, int n_rays,
                                    const mcnerf_composite_cfg* cfg, float* weights, float* w_max, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  MC_ARG(sigma && weights && n_rays > 0 && sigma_stride >= 1);
  sigma2weights_k<<<cdiv(n_rays, WARPS), WARPS * 32, 0, (cudaStream_t)stream>>>(
      sigma, sigma_stride, noise, jitter, z_vals, deltas, n_rays, *cfg, weights, w_max);
  MC_LAUNCHED();
  return 0;
}

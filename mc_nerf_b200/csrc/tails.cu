// Fused "tails" of the coarse and the fine branch of a render (what follows each MLP evaluation), one warp per ray:
//
//  coarse tail : ONE pass over the coarse network's outputs produces the noisy-composited colour (ref:
//                model/mc_nerf.py:719-727) AND the fine-sample selection weights with their own noise draw + their
//                global maximum (ref: :613-621 / :658-662, sigma2weights :729-736) - the unfused path read out_c three
//                times (composite_fwd, sigma2weights, select).
//  fine tail   : compositing of the fine grid straight from the COMPACTED MLP output (the rows of the selected samples
//                only): whether fine sample j of a ray was selected, and where its row is, follows from the selection
//                weights and the ray's offset (select.cu), so the default-filled dense [B,Sf,4] tensor, its scatter and,
//                in the backward pass, the dense gradient and its gather never exist (ref: :688-701, 705-727).
//
// Density noise: explicit tensors (parity mode: the reference's torch.randn draws are replayed) or, when a seed is
// given instead, N(0,1) from Philox4x32-10 generated in the kernel and RE-generated in the backward pass (philox.cuh) -
// no [B,S] fp32 tensors through HBM and no ATen RNG launches.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int WARPS = 4;

__device__ __forceinline__ float scan_mul_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float scan_add_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

struct ZGrid {       // z_k = linspace(near, far, S)[k] + jitter ; delta_k = z_{k+1} - z_k, last 1e10 (ref: :599-602, 708-710)
  float near_, far_, jit;
  int S;
  __device__ __forceinline__ float z(int k) const { return linspace_f(near_, far_, S, k) + jit; }
  __device__ __forceinline__ float delta(int k) const { return (k == S - 1) ? 1e10f : z(k + 1) - z(k); }
};

struct Noise {       // element (ray, k) of a [B,S] noise tensor, or of Philox stream `stream`, or zero
  const float* t;
  PhiloxKey key;
  bool rng;
  int stream;
  __device__ __forceinline__ float at(size_t idx) const {
    return t ? t[idx] : (rng ? philox_normal(key, stream, idx) : 0.f);
  }
};
__device__ __forceinline__ Noise make_noise(const float* t, const int64_t* seed, int stream) {
  Noise n;
  n.t = t; n.rng = (t == nullptr && seed != nullptr); n.stream = stream;
  if (n.rng) n.key = philox_key(seed);
  return n;
}

// ------------------------------------------------------------------------------------------------ coarse tail
template <int NC>
__global__ void __launch_bounds__(WARPS * 32)
coarse_tail_fwd_k(const float4* __restrict__ out4, const float* __restrict__ noise_rgb, const float* __restrict__ noise_sel,
                  const int64_t* __restrict__ seed, const float* __restrict__ jitter, int n_rays, mcnerf_composite_cfg cfg,
                  float* __restrict__ rgb, float* __restrict__ w_sel, float* __restrict__ w_max) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  const size_t row = (size_t)ray * S;
  ZGrid zg{cfg.near_, cfg.far_, jitter ? jitter[ray] : 0.f, S};
  const Noise n1 = make_noise(noise_rgb, seed, MC_STREAM_NOISE_C), n2 = make_noise(noise_sel, seed, MC_STREAM_NOISE_SEL);
  float4 oc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k = c * 32 + lane;
    oc[c] = k < S ? out4[row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float T1 = 1.f, T2 = 1.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, wm = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c * 32 >= S) break;
    const int k = c * 32 + lane;
    const bool live = k < S;
    const float4 o = oc[c];
    const float dl = live ? zg.delta(k) : 0.f;
    // colour weights (noise draw 1) and selection weights (independent noise draw 2): same sigma, same deltas
    const float a1 = live ? 1.f - expf(-dl * softplus_f(o.x + n1.at(row + k))) : 0.f;
    const float a2 = live ? 1.f - expf(-dl * softplus_f(o.x + n2.at(row + k))) : 0.f;
    const float f1 = live ? (1.f - a1 + 1e-10f) : 1.f, f2 = live ? (1.f - a2 + 1e-10f) : 1.f;
    const float f1_inc = scan_mul_incl(f1, lane), f2_inc = scan_mul_incl(f2, lane);
    float f1_exc = __shfl_up_sync(0xffffffffu, f1_inc, 1), f2_exc = __shfl_up_sync(0xffffffffu, f2_inc, 1);
    if (lane == 0) { f1_exc = 1.f; f2_exc = 1.f; }
    const float w1 = a1 * (T1 * f1_exc), w2 = a2 * (T2 * f2_exc);
    T1 *= __shfl_sync(0xffffffffu, f1_inc, 31);
    T2 *= __shfl_sync(0xffffffffu, f2_inc, 31);
    if (live) {
      w_sel[row + k] = w2;
      wm = fmaxf(wm, w2);
    }
    acc_w += w1;
    acc_r += w1 * o.y;
    acc_g += w1 * o.z;
    acc_b += w1 * o.w;
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b); acc_w = warp_sum(acc_w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
  if (lane == 0) {
    const float bg = cfg.white_back ? 1.f - acc_w : 0.f;
    rgb[3 * ray] = acc_r + bg;
    rgb[3 * ray + 1] = acc_g + bg;
    rgb[3 * ray + 2] = acc_b + bg;
    atomicMax((int*)w_max, __float_as_int(wm));      // weights are >= 0: int order of the bit pattern = float order
  }
}

// ------------------------------------------------------------------------------------------------ fine tail
// Which fine samples of this ray were selected, and the rank of each among the ray's selected ones: bit i of keep[cc]
// <=> coarse sample 32 cc + i passed the threshold (ballots are warp-uniform, so every lane holds all masks).
template <int MC>
struct KeepMask {
  unsigned m[MC];
  int pre[MC];
  __device__ __forceinline__ void build(const float* __restrict__ w_row, int Sc, float thr, int lane) {
    int run = 0;
#pragma unroll
    for (int cc = 0; cc < MC; ++cc) {
      const int i = cc * 32 + lane;
      m[cc] = __ballot_sync(0xffffffffu, (i < Sc) && (w_row[i] >= thr));
      pre[cc] = run;
      run += __popc(m[cc]);
    }
  }
  // coarse sample i -> (selected?, rank among the selected coarse samples of the ray)
  __device__ __forceinline__ bool rank(int i, int& r) const {
    const int cc = i >> 5, bit = i & 31;
    unsigned mm = 0;
    int pp = 0;
#pragma unroll
    for (int c = 0; c < MC; ++c)
      if (c == cc) { mm = m[c]; pp = pre[c]; }
    r = pp + __popc(mm & ((1u << bit) - 1u));
    return (mm >> bit) & 1u;
  }
};

template <int NC>
__global__ void __launch_bounds__(WARPS * 32)
fine_tail_fwd_k(const float4* __restrict__ out_sel, const float* __restrict__ w_sel, const float* __restrict__ w_max,
                float thresh, int scale, const int32_t* __restrict__ sel_offsets, const float* __restrict__ rays_d,
                const float* __restrict__ jitter, const float* __restrict__ noise, const int64_t* __restrict__ seed,
                int n_rays, mcnerf_composite_cfg cfg, float sigma_default, float* __restrict__ rgb,
                float* __restrict__ depth, float* __restrict__ opacity) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S, Sc = S / scale;
  const size_t row = (size_t)ray * S;
  ZGrid zg{cfg.near_, cfg.far_, jitter ? jitter[ray] : 0.f, S};
  const Noise nz = make_noise(noise, seed, MC_STREAM_NOISE_F);
  KeepMask<NC> km;
  km.build(w_sel + (size_t)ray * Sc, Sc, fminf(thresh, *w_max), lane);
  const int base = sel_offsets[ray];
  float4 oc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k = c * 32 + lane;
    oc[c] = make_float4(sigma_default, 1.f, 1.f, 1.f);       // unselected samples: white far plane (ref: :692-694)
    if (k < S) {
      const int i = k / scale;
      int r;
      if (km.rank(i, r)) oc[c] = out_sel[base + r * scale + (k - i * scale)];
    }
  }
  const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  const float len = sqrtf(dx * dx + dy * dy + dz * dz);
  float tau_carry = 0.f, T_carry = 1.f;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_op = 0.f, acc_dp = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c * 32 >= S) break;
    const int k = c * 32 + lane;
    const bool live = k < S;
    const float4 o = oc[c];
    const float zk = live ? zg.z(k) : 0.f;
    const float dl = live ? zg.delta(k) : 0.f;
    const float tau = live ? softplus_f(o.x) * (dl * len) : 0.f;            // (i) noise-free: depth / opacity
    const float tau_inc = scan_add_incl(tau, lane);
    float tau_exc = __shfl_up_sync(0xffffffffu, tau_inc, 1);                // by shifting, never inc - tau (tau ~ 1e10)
    if (lane == 0) tau_exc = 0.f;
    const float T = expf(-(tau_carry + tau_exc));
    const float pa = live ? T * (1.f - expf(-tau)) : 0.f;
    acc_op += pa;
    acc_dp += pa * zk;
    tau_carry += __shfl_sync(0xffffffffu, tau_inc, 31);
    const float al = live ? 1.f - expf(-dl * softplus_f(o.x + nz.at(row + k))) : 0.f;     // (ii) noisy: colour
    const float f = live ? (1.f - al + 1e-10f) : 1.f;
    const float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    const float w = al * (T_carry * f_exc);
    T_carry *= __shfl_sync(0xffffffffu, f_inc, 31);
    acc_w += w;
    acc_r += w * o.y;
    acc_g += w * o.z;
    acc_b += w * o.w;
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b); acc_w = warp_sum(acc_w);
  acc_op = warp_sum(acc_op); acc_dp = warp_sum(acc_dp);
  if (lane == 0) {
    const float bg = cfg.white_back ? 1.f - acc_w : 0.f;
    rgb[3 * ray] = acc_r + bg;
    rgb[3 * ray + 1] = acc_g + bg;
    rgb[3 * ray + 2] = acc_b + bg;
    if (depth) depth[ray] = acc_dp;
    if (opacity) opacity[ray] = acc_op;
  }
}

// Backward of the noisy compositing (same recurrences as composite.cu::composite_bwd_c_k), sample values fetched from
// `src` (dense [B,S,4], COMPACT = false) or from the compacted rows (COMPACT = true); gradients written in the same
// layout.  Unselected samples receive no gradient.
template <int NC, bool COMPACT>
__global__ void __launch_bounds__(WARPS * 32)
tail_bwd_k(const float4* __restrict__ src, const float* __restrict__ w_sel, const float* __restrict__ w_max, float thresh,
           int scale, const int32_t* __restrict__ sel_offsets, const float* __restrict__ jitter,
           const float* __restrict__ noise, const int64_t* __restrict__ seed, int stream_id, int n_rays,
           mcnerf_composite_cfg cfg, float sigma_default, const float* __restrict__ g_rgb, float4* __restrict__ g_out) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = cfg.S;
  const size_t row = (size_t)ray * S;
  ZGrid zg{cfg.near_, cfg.far_, jitter ? jitter[ray] : 0.f, S};
  const Noise nz = make_noise(noise, seed, stream_id);
  float4 oc[NC];
  float sgc[NC];
  int pos[NC];          // row of the sample in the compacted list, -1: not selected (COMPACT only)
  if (COMPACT) {
    const int Sc = S / scale;
    KeepMask<NC> km;
    km.build(w_sel + (size_t)ray * Sc, Sc, fminf(thresh, *w_max), lane);
    const int base = sel_offsets[ray];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k = c * 32 + lane;
      oc[c] = make_float4(sigma_default, 1.f, 1.f, 1.f);
      pos[c] = -1;
      if (k < S) {
        const int i = k / scale;
        int r;
        if (km.rank(i, r)) {
          pos[c] = base + r * scale + (k - i * scale);
          oc[c] = src[pos[c]];
        }
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k = c * 32 + lane;
      oc[c] = k < S ? src[row + k] : make_float4(0.f, 0.f, 0.f, 0.f);
      pos[c] = k < S ? (int)0 : -1;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k = c * 32 + lane;
    sgc[c] = oc[c].x + (k < S ? nz.at(row + k) : 0.f);
  }
  const float gr = g_rgb[3 * ray], gg = g_rgb[3 * ray + 1], gb = g_rgb[3 * ray + 2];
  const float gsum = cfg.white_back ? (gr + gg + gb) : 0.f;
  float omc[NC], Tin[NC];
  {
    float run = 1.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k = c * 32 + lane;
      const bool live = k < S;
      omc[c] = live ? expf(-zg.delta(k) * softplus_f(sgc[c])) : 1.f;
      const float f = live ? (1.f - (1.f - omc[c]) + 1e-10f) : 1.f;
      const float f_inc = scan_mul_incl(f, lane);
      Tin[c] = run;
      run *= __shfl_sync(0xffffffffu, f_inc, 31);
    }
  }
  float R_carry = 0.f;
#pragma unroll
  for (int c = NC - 1; c >= 0; --c) {
    if (c * 32 >= S) continue;
    const int k = c * 32 + lane;
    const bool live = k < S;
    const float4 o = oc[c];
    const float dl = live ? zg.delta(k) : 0.f;
    const float om = omc[c];
    const float al = 1.f - om;
    const float f = live ? (1.f - al + 1e-10f) : 1.f;
    const float f_inc = scan_mul_incl(f, lane);
    float f_exc = __shfl_up_sync(0xffffffffu, f_inc, 1);
    if (lane == 0) f_exc = 1.f;
    const float T = Tin[c] * f_exc;
    const float q = live ? (gr * o.y + gg * o.z + gb * o.w - gsum) : 0.f;
    float a = q * al, ff = f;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const float a2 = __shfl_down_sync(0xffffffffu, a, s);
      const float f2 = __shfl_down_sync(0xffffffffu, ff, s);
      if (lane + s < 32) {
        a = a + ff * a2;
        ff = ff * f2;
      }
    }
    const float E = a + ff * R_carry;
    float R = __shfl_down_sync(0xffffffffu, E, 1);
    if (lane == 31) R = R_carry;
    R_carry = __shfl_sync(0xffffffffu, E, 0);
    if (live && pos[c] >= 0) {
      const float w = al * T;
      const float dalpha = T * (q - R);
      const float dsig = dalpha * om * dl * sigmoid_f(sgc[c]);
      g_out[COMPACT ? (size_t)pos[c] : row + k] = make_float4(dsig, w * gr, w * gg, w * gb);
    }
  }
}

__global__ void philox_fill_k(const int64_t* __restrict__ seed, int stream_id, int64_t n, int normal, float lo, float hi,
                              float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PhiloxKey k = philox_key(seed);
  out[i] = normal ? philox_normal(k, stream_id, (uint64_t)i) : lo + (hi - lo) * philox_uniform(k, stream_id, (uint64_t)i);
}

}  // namespace

#define TAIL_NC(S) (((S) + 31) / 32)

extern "C" int mcnerf_philox_fill(const int64_t* seed, int stream_id, int64_t n, int normal, float lo, float hi,
                                  float* out, void* stream) {
  MC_ARG(seed && out && n >= 0 && stream_id >= 0 && stream_id < 256);
  if (n == 0) return 0;
  philox_fill_k<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, n, normal, lo, hi, out);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_coarse_tail_fwd(const float* out4, const float* noise_rgb, const float* noise_sel,
                                      const int64_t* seed, const float* jitter, int n_rays,
                                      const mcnerf_composite_cfg* cfg, float* rgb, float* w_sel, float* w_max,
                                      void* stream) {
  MC_ARG(cfg && cfg->S >= 2 && cfg->S <= 256 && out4 && rgb && w_sel && w_max && n_rays > 0 && ((uintptr_t)out4 & 15) == 0);
  const int nc = TAIL_NC(cfg->S);
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define CT(NC) coarse_tail_fwd_k<NC><<<grid, block, 0, st>>>((const float4*)out4, noise_rgb, noise_sel, seed, jitter, n_rays, *cfg, rgb, w_sel, w_max)
  if (nc <= 2) CT(2); else if (nc <= 4) CT(4); else CT(8);
#undef CT
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_coarse_tail_bwd(const float* out4, const float* noise_rgb, const int64_t* seed, const float* jitter,
                                      int n_rays, const mcnerf_composite_cfg* cfg, const float* g_rgb, float* g_out4,
                                      void* stream) {
  MC_ARG(cfg && cfg->S >= 2 && cfg->S <= 256 && out4 && g_rgb && g_out4 && n_rays > 0 && ((uintptr_t)out4 & 15) == 0 &&
         ((uintptr_t)g_out4 & 15) == 0);
  const int nc = TAIL_NC(cfg->S);
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define CB(NC) tail_bwd_k<NC, false><<<grid, block, 0, st>>>((const float4*)out4, nullptr, nullptr, 0.f, 1, nullptr, jitter, noise_rgb, seed, MC_STREAM_NOISE_C, n_rays, *cfg, 0.f, g_rgb, (float4*)g_out4)
  if (nc <= 2) CB(2); else if (nc <= 4) CB(4); else CB(8);
#undef CB
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_fine_tail_fwd(const float* out_sel, const float* w_sel, const float* w_max, float thresh, int scale,
                                    const int32_t* sel_offsets, const float* rays_d, const float* jitter,
                                    const float* noise, const int64_t* seed, int n_rays, const mcnerf_composite_cfg* cfg,
                                    float sigma_default, float* rgb, float* depth, float* opacity, void* stream) {
  MC_ARG(cfg && cfg->S >= 2 && cfg->S <= 256 && scale >= 1 && cfg->S % scale == 0 && out_sel && w_sel && w_max &&
         sel_offsets && rays_d && rgb && n_rays > 0 && ((uintptr_t)out_sel & 15) == 0);
  const int nc = TAIL_NC(cfg->S);
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define FT(NC) fine_tail_fwd_k<NC><<<grid, block, 0, st>>>((const float4*)out_sel, w_sel, w_max, thresh, scale, sel_offsets, rays_d, jitter, noise, seed, n_rays, *cfg, sigma_default, rgb, depth, opacity)
  if (nc <= 2) FT(2); else if (nc <= 4) FT(4); else FT(8);
#undef FT
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_fine_tail_bwd(const float* out_sel, const float* w_sel, const float* w_max, float thresh, int scale,
                                    const int32_t* sel_offsets, const float* jitter, const float* noise,
                                    const int64_t* seed, int n_rays, const mcnerf_composite_cfg* cfg, float sigma_default,
                                    const float* g_rgb, float* g_sel, void* stream) {
  MC_ARG(cfg && cfg->S >= 2 && cfg->S <= 256 && scale >= 1 && cfg->S % scale == 0 && out_sel && w_sel && w_max &&
         sel_offsets && g_rgb && g_sel && n_rays > 0 && ((uintptr_t)out_sel & 15) == 0 && ((uintptr_t)g_sel & 15) == 0);
  const int nc = TAIL_NC(cfg->S);
  const dim3 grid(cdiv(n_rays, WARPS)), block(WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define FB(NC) tail_bwd_k<NC, true><<<grid, block, 0, st>>>((const float4*)out_sel, w_sel, w_max, thresh, scale, sel_offsets, jitter, noise, seed, MC_STREAM_NOISE_F, n_rays, *cfg, sigma_default, g_rgb, (float4*)g_sel)
  if (nc <= 2) FB(2); else if (nc <= 4) FB(4); else FB(8);
#undef FB
  MC_LAUNCHED();
  return 0;
}

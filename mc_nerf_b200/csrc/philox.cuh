// Philox4x32-10 (Salmon et al., SC'11) and the draws the render kernels derive from it in "device RNG" mode.
// Keyed by two device-side int64 words drawn from torch's CUDA generator (so torch.manual_seed governs it and a captured
// CUDA graph gets fresh draws per replay); element i of stream s is counter (i_lo, s | i_hi << 8, seed1_lo, seed1_hi),
// key (seed0_lo, seed0_hi).  Stream 0 is the pixel sampler (misc.cu); 1: density noise of the coarse render,
// 2: of the fine-sample selection, 3: of the fine render (ref draws: model/mc_nerf.py:719, 619/662, 719); 4: the
// per-ray jitter (ref: model/mc_nerf.py:601).
#pragma once
#include <stdint.h>

enum { MC_STREAM_PIXELS = 0, MC_STREAM_NOISE_C = 1, MC_STREAM_NOISE_SEL = 2, MC_STREAM_NOISE_F = 3, MC_STREAM_JITTER = 4 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

struct PhiloxKey {
  uint32_t k0, k1, c2, c3;
};
__device__ __forceinline__ PhiloxKey philox_key(const int64_t* seed) {
  const uint64_t s0 = (uint64_t)seed[0], s1 = (uint64_t)seed[1];
  PhiloxKey k;
  k.k0 = (uint32_t)s0; k.k1 = (uint32_t)(s0 >> 32); k.c2 = (uint32_t)s1; k.c3 = (uint32_t)(s1 >> 32);
  return k;
}
__device__ __forceinline__ uint4 philox_at(const PhiloxKey& k, int stream, uint64_t idx) {
  return philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)stream | ((uint32_t)(idx >> 32) << 8), k.c2, k.c3),
                       make_uint2(k.k0, k.k1));
}
// uniform in (0, 1): 32 random bits centred in their bin
__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; }
// N(0,1) by Box-Muller from the first two words
__device__ __forceinline__ float philox_normal(const PhiloxKey& k, int stream, uint64_t idx) {
  const uint4 r = philox_at(k, stream, idx);
  const float u1 = fminf(u01(r.x), 0.99999994f), u2 = u01(r.y);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}
__device__ __forceinline__ float philox_uniform(const PhiloxKey& k, int stream, uint64_t idx) {
  return u01(philox_at(k, stream, idx).x);
}

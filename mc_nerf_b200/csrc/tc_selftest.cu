// One-tile tcgen05 GEMM used by the test-suite to pin the UMMA descriptor conventions on real hardware:
//   D[128,N] = A[128,K] * B[N,K]^T     (bf16 in, fp32 out; the forward-layer orientation, both K-major)
//   D[128,N] = A^T * B   with A stored [K,128] and B stored [K,N] (both MN-major; the weight-gradient orientation)
// Operands are staged by ordinary loads into the SWIZZLE_NONE core-matrix layout the fused MLP kernels use.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

// canonical tile image: element (row r of the M/N index, reduction index k) lives at
//   (k/8)*rows*16 + r*16 + (k%8)*2   bytes            (k-group planes of [rows x 16 B])
__device__ __forceinline__ uint32_t canon_off(int r, int k, int rows) { return (k >> 3) * rows * 16 + r * 16 + (k & 7) * 2; }

__global__ void __launch_bounds__(128) tc_selftest_k(const __nv_bfloat16* __restrict__ A,
                                                     const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N,
                                                     int K, int variant, int reps, long long* cycles) {
  const int mn_major = variant & 1;
  const bool swap_ls = (variant & 2) != 0;   // diagnostic: exchange the lead/stride byte offsets
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;                   // 128 x K
  uint8_t* sB = smem + 128 * K * 2;     // N x K
  const int tid = threadIdx.x, warp = tid >> 5;
  if (!mn_major) {
    // K-major: core matrix = 8 rows x 8 consecutive k
    for (int i = tid; i < 128 * K; i += 128) {
      int r = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sA + canon_off(r, k, 128)) = A[r * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
      int r = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sB + canon_off(r, k, N)) = B[r * K + k];
    }
  } else {
    // MN-major: global A is [K,128] (m contiguous), B is [K,N].  Core matrix = 8 k x 8 consecutive m:
    // image: m-group planes of [K x 16 B]: (m/8)*K*16 + k*16 + (m%8)*2
    for (int i = tid; i < 128 * K; i += 128) {
      int k = i / 128, m = i % 128;
      *reinterpret_cast<__nv_bfloat16*>(sA + (m >> 3) * K * 16 + k * 16 + (m & 7) * 2) = A[k * 128 + m];
    }
    for (int i = tid; i < N * K; i += 128) {
      int k = i / N, n = i % N;
      *reinterpret_cast<__nv_bfloat16*>(sB + (n >> 3) * K * 16 + k * 16 + (n & 7) * 2) = B[k * N + n];
    }
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::umma_idesc_bf16(128, N, mn_major, mn_major);
    const long long t0 = clock64();
    if (variant & 4) {
      // tight issue loop (K-major): descriptor words precomputed, one add per operand and MMA
      const uint32_t a_hi = tc::umma_desc_hi(128), b_hi = a_hi;
      const uint32_t a_lo0 = tc::umma_desc_lo(tc::smem_u32(sA), 128 * 16), b_lo0 = tc::umma_desc_lo(tc::smem_u32(sB), N * 16);
      const uint32_t a_inc = (2 * 128 * 16) >> 4, b_inc = (2 * N * 16) >> 4;
      const int nk = K / 16;
      for (int rep = 0; rep < reps; ++rep) {
        uint32_t a_lo = a_lo0, b_lo = b_lo0;
        tc::umma_bf16_w(tmem, a_lo, a_hi, b_lo, b_hi, idesc, rep > 0);
#pragma unroll 4
        for (int k = 1; k < nk; ++k) {
          a_lo += a_inc; b_lo += b_inc;
          tc::umma_bf16_w(tmem, a_lo, a_hi, b_lo, b_hi, idesc, true);
        }
      }
    } else
    for (int rep = 0; rep < reps; ++rep)
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint64_t da, db;
      if (!mn_major) {
        // K direction: next k-group plane (rows*16 B); M/N direction: next 8 rows (128 B)
        da = tc::umma_desc(tc::smem_u32(sA) + (k0 >> 3) * 128 * 16, 128 * 16, 128);
        db = tc::umma_desc(tc::smem_u32(sB) + (k0 >> 3) * N * 16, N * 16, 128);
      } else {
        // K direction: next 8 k (128 B); MN direction: next m-group plane (K*16 B)
        da = tc::umma_desc(tc::smem_u32(sA) + k0 * 16, 128, K * 16);
        db = tc::umma_desc(tc::smem_u32(sB) + k0 * 16, 128, K * 16);
      }
      if (swap_ls) {   // diagnostic only: lead/stride exchanged (must give a wrong product)
        uint32_t la = !mn_major ? 128 * 16 : 128, lb = !mn_major ? N * 16 : 128;
        uint32_t ta = !mn_major ? 128 : K * 16, tb = !mn_major ? 128 : K * 16;
        uint32_t sa = !mn_major ? tc::smem_u32(sA) + (k0 >> 3) * 128 * 16 : tc::smem_u32(sA) + k0 * 16;
        uint32_t sb = !mn_major ? tc::smem_u32(sB) + (k0 >> 3) * N * 16 : tc::smem_u32(sB) + k0 * 16;
        da = tc::umma_desc(sa, ta, la);
        db = tc::umma_desc(sb, tb, lb);
      }
      tc::umma_bf16(tmem, da, db, idesc, (k0 | rep) > 0);
    }
    tc::umma_commit(&bar);
    const long long t1 = clock64();
    tc::mbar_wait(&bar, 0);
    if (cycles) { cycles[0] = t1 - t0; cycles[1] = clock64() - t0; }
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  const int row = tid;   // TMEM lane == row of D
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// CTA-pair variant: D[256,N] = A[256,K] * B[N,K]^T with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA.
// CTA r stages A rows [128r,+128) and B rows [N/2*r, +N/2) (K-major canonical images) and reads back D rows [128r,+128).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
tc_selftest2_k(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N, int K,
               int reps, long long* cycles, const float* __restrict__ bias) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t rank = tc::cluster_ctarank();
  const int NH = N / 2;
  uint8_t* sA = smem;                   // 128 x K
  uint8_t* sB = smem + 128 * K * 2;     // N/2 x K
  uint8_t* sOnes = sB + NH * K * 2;     // 2 core matrices: 8 rows x (1, 1, 0 ... 0) and zeros
  uint8_t* sBias = sOnes + 256;         // N/2 x 16: (hi(b), lo(b), 0 ... 0)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + canon_off(r, k, 128)) = A[(rank * 128 + r) * K + k];
  }
  if (bias) {
    // the bias rides in the accumulator: one extra K=16 MMA whose A operand is a broadcast "ones" block (every
    // 8-row group of the 128 rows reads the SAME 128-byte core matrix: stride-byte offset 0) and whose B operand
    // holds the bias split into two bf16 (hi + lo reproduces the fp32 value to ~2^-17)
    for (int i = tid; i < 128; i += 128) {
      int r = i >> 3 & 7, j = i & 7;   // (unused r) fill 2 core matrices = 128 bf16
      (void)r;
      reinterpret_cast<__nv_bfloat16*>(sOnes)[i] = __float2bfloat16((i < 64 && j < 2) ? 1.f : 0.f);
    }
    for (int i = tid; i < NH * 16; i += 128) {
      int r = i / 16, k = i % 16;
      float b = bias[rank * NH + r];
      __nv_bfloat16 hi = __float2bfloat16(b);
      float v = k == 0 ? __bfloat162float(hi) : k == 1 ? b - __bfloat162float(hi) : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(sBias + canon_off(r, k, NH)) = __float2bfloat16(v);
    }
  }
  for (int i = tid; i < NH * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + canon_off(r, k, NH)) = B[(rank * NH + r) * K + k];
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc2(&tmem_base, 256);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();              // both CTAs' operands are staged, barriers initialised
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = tc::umma_idesc_bf16(256, N);
    const uint32_t hi = tc::umma_desc_hi(128);
    const uint32_t a_lo0 = tc::umma_desc_lo(tc::smem_u32(sA), 128 * 16), b_lo0 = tc::umma_desc_lo(tc::smem_u32(sB), NH * 16);
    const uint32_t a_inc = (2 * 128 * 16) >> 4, b_inc = (2 * NH * 16) >> 4;
    const int nk = K / 16;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      tc::umma2_bf16_w(tmem, a_lo, hi, b_lo, hi, idesc, rep > 0);
#pragma unroll 4
      for (int k = 1; k < nk; ++k) {
        a_lo += a_inc; b_lo += b_inc;
        tc::umma2_bf16_w(tmem, a_lo, hi, b_lo, hi, idesc, true);
      }
    }
    if (bias)
      tc::umma2_bf16_w(tmem, tc::umma_desc_lo(tc::smem_u32(sOnes), 128), tc::umma_desc_hi(0),
                       tc::umma_desc_lo(tc::smem_u32(sBias), NH * 16), hi, idesc, true);
    tc::umma2_commit_multicast_addr(tc::smem_u32(&bar), (uint16_t)3);
    const long long t1 = clock64();
    tc::mbar_wait(&bar, 0);
    if (cycles) { cycles[0] = t1 - t0; cycles[1] = clock64() - t0; }
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();              // neither CTA frees tensor memory while the other may still be using the pair
  if (warp == 0) tc::tmem_dealloc2(tmem, 256);
}

// CTA-pair product with the A operand in TENSOR MEMORY: D[256,N] = A[256,K] * B[N,K]^T.  Each CTA writes its 128 rows
// of A (bf16 pairs, row = lane) into columns [256, 256 + K/2) of its tensor memory with tcgen05.st and stages its N/2
// rows of B in shared memory; the leader issues K/16 .ts MMAs per pass.  n_split = 2 runs the product as two N/2-wide
// passes into column halves of the accumulator (the schedule of a kernel that drains one half while the other is
// being computed).  bg > 0: warps 4-7 keep tcgen05.ld/st traffic on columns [384, 512) going meanwhile (port probe).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256)
tc_selftest_ts_k(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N, int K,
                 int n_split_flags, int reps, int bg, long long* cycles) {
  const int n_split = n_split_flags & 3, bg_mode = (n_split_flags >> 4) & 3;     // bg_mode 1: loads only, 2: stores only
  const int kl = (n_split_flags >> 8) & 1;   // 1: the fused kernel's layout - A at columns [0,128), every pass accumulates at [128,256) (timing only)
  const uint32_t a_col = kl ? 0 : 256;
  const int commit_each = (n_split_flags >> 9) & 1;   // 1: a tcgen05.commit (to a barrier nobody waits on) after every pass
  __shared__ uint64_t bar2;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t rank = tc::cluster_ctarank();
  const int NH = N / 2;                 // rows of B in this CTA
  const int NP = NH / n_split;          // ... per pass
  uint8_t* sB = smem;                   // n_split images of [NP x K]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // pass h of the pair covers D columns [h*N/n_split, +N/n_split): this CTA supplies B rows h*(N/n_split) + rank*NP + r
  for (int i = tid; i < NH * K; i += blockDim.x) {
    int rr = i / K, k = i % K, h = rr / NP, r = rr % NP;
    *reinterpret_cast<__nv_bfloat16*>(sB + (size_t)h * NP * K * 2 + canon_off(r, k, NP)) =
        B[(size_t)(h * (N / n_split) + rank * NP + r) * K + k];
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&bar2, 1u << 19);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc2(&tmem_base, 512);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp < 4) {
    const int row = warp * 32 + lane;
    const __nv_bfloat16* arow = A + (size_t)(rank * 128 + row) * K;
    for (int c = 0; c < K / 2; c += 8) {
      uint32_t w[8];
      for (int j = 0; j < 8; ++j) {
        uint32_t lo = __bfloat16_as_ushort(arow[2 * (c + j)]), hi = __bfloat16_as_ushort(arow[2 * (c + j) + 1]);
        w[j] = lo | (hi << 16);
      }
      tc::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col + c, w);
    }
    tc::tmem_st_wait();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tcgen05_fence_after();
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = tc::umma_idesc_bf16(256, N / n_split);
    const uint32_t hi = tc::umma_desc_hi(128);
    const uint32_t b_inc = (2 * NP * 16) >> 4;
    const int nk = K / 16;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int h = 0; h < n_split; ++h) {
        uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(sB + (size_t)h * NP * K * 2), NP * 16);
        uint32_t a_t = tmem + a_col;
        const uint32_t d_t = tmem + (kl ? 128 : h * (N / n_split));
        tc::umma2_bf16_ts(d_t, a_t, b_lo, hi, idesc, rep > 0);
#pragma unroll 4
        for (int k = 1; k < nk; ++k) {
          a_t += 8; b_lo += b_inc;
          tc::umma2_bf16_ts(d_t, a_t, b_lo, hi, idesc, true);
        }
        if (commit_each) tc::umma2_commit_multicast_addr(tc::smem_u32(&bar2), (uint16_t)3);
      }
    tc::umma2_commit_multicast_addr(tc::smem_u32(&bar), (uint16_t)3);
    const long long t1 = clock64();
    tc::mbar_wait(&bar, 0);
    if (cycles) { cycles[0] = t1 - t0; cycles[1] = clock64() - t0; }
  }
  if (warp >= 4 && bg > 0) {
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    uint32_t v[32], w[8];
    for (int j = 0; j < 8; ++j) w[j] = j;
    for (int i = 0; i < bg; ++i) {
      if (bg_mode != 2) {
        tc::tmem_ld32(t + (i & 3) * 32, v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 8; ++j) w[j] += v[j] + v[j + 8] + v[j + 16] + v[j + 24];
      }
      if (bg_mode != 1) {
        tc::tmem_st8(t + (i & 3) * 32, w);
        tc::tmem_st8(t + (i & 3) * 32 + 8, w);
        if (bg_mode == 2) tc::tmem_st_wait();
      }
    }
    tc::tmem_st_wait();
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tc::tmem_ld_wait();
      for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  if (warp == 0) tc::tmem_dealloc2(tmem, 512);
}

}  // namespace

extern "C" int mcnerf_tc_selftest_ts(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int n_split, int reps,
                                     int bg, long long* cycles_out, void* stream) {
  MC_ARG(A_bf16 && B_bf16 && D && N >= 64 && N <= 256 && N % 64 == 0 && K >= 16 && K <= 256 && K % 16 == 0 && reps >= 1);
  MC_ARG((n_split & 3) == 1 || (n_split & 3) == 2);
  size_t smem = (size_t)(N / 2) * K * 2;
  MC_CUDA(cudaFuncSetAttribute(tc_selftest_ts_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_ts_k<<<2, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A_bf16, (const __nv_bfloat16*)B_bf16, D, N, K,
                                                            n_split, reps, bg, cycles_out);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_tc_selftest2(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int reps,
                                   long long* cycles_out, const float* bias, void* stream) {
  MC_ARG(A_bf16 && B_bf16 && D && N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0 && reps >= 1);
  size_t smem = (size_t)(128 + N / 2) * K * 2 + 256 + (size_t)(N / 2) * 32;
  MC_ARG(smem <= 200 * 1024);
  MC_CUDA(cudaFuncSetAttribute(tc_selftest2_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest2_k<<<2, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A_bf16, (const __nv_bfloat16*)B_bf16, D, N, K,
                                                          reps, cycles_out, bias);
  MC_LAUNCHED();
  return 0;
}

static int selftest_launch(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major, int reps,
                           long long* cycles, void* stream);

// MMA issue-rate probe: the same K/16 MMAs repeated `reps` times on resident operands (result is reps x the product).
// cycles_out[0] = clock64 ticks to ISSUE them, [1] = until the commit barrier fired (device pointer, 2 x int64).
extern "C" int mcnerf_tc_mma_rate(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major,
                                  int reps, long long* cycles_out, void* stream) {
  return selftest_launch(A_bf16, B_bf16, D, N, K, mn_major, reps, cycles_out, stream);
}

extern "C" int mcnerf_tc_selftest(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major,
                                  void* stream) {
  return selftest_launch(A_bf16, B_bf16, D, N, K, mn_major, 1, nullptr, stream);
}

static int selftest_launch(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int mn_major, int reps,
                           long long* cycles, void* stream) {
  MC_ARG(A_bf16 && B_bf16 && D && N >= 16 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0);
  size_t smem = (size_t)(128 + N) * K * 2;
  MC_ARG(smem <= 200 * 1024);
  MC_CUDA(cudaFuncSetAttribute(tc_selftest_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_k<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A_bf16, (const __nv_bfloat16*)B_bf16, D, N,
                                                         K, mn_major, reps, cycles);
  MC_LAUNCHED();
  return 0;
}

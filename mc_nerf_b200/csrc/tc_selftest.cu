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

}  // namespace

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

// Gradient all-reduce over NVLink peer memory (one 8 x B200 NVSwitch box), hand-written: ONE kernel per call, two-shot.
//   every rank's gradient buffer is symmetric memory mapped into all ranks (torch.distributed._symmetric_memory does the
//   CUDA VMM / handle exchange; this file only sees raw peer pointers);
//   barrier 1 (system-scope release/acquire flags in the peers' flag pads): every rank's local gradients are complete;
//   rank r reduces slice r: 128-bit P2P loads of the slice from all N ranks, summed in rank order (every rank computes
//   bit-identical sums), and stores the result into ALL N buffers (fused all-gather) - slices are disjoint, so nobody reads
//   what somebody else writes;
//   barrier 2: all stores have landed; the buffers now hold the same sums on every rank.
// Why not NCCL here: the message is 2.5-5 MB once or twice per 0.8-3 ms step - latency-bound.  NCCL's three launches cost
// ~110-120 us of exposed time per step at N = 2..8 (bench.py `allreduce_exposed_us`); this kernel moves 2 x count x 4 bytes
// per rank over NVLink (770 GB/s measured per direction) plus two flag round trips.
// ref: the reference relies on DistributedDataParallel's bucketed NCCL all-reduce (main.py:61,84).
#include "common.cuh"

namespace {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_THREADS = 512;

struct ArArgs {
  float* buf[AR_MAX_RANKS];
  uint32_t* flags[AR_MAX_RANKS];      // per rank: [n_ctas][2 phases][AR_MAX_RANKS] counters
  uint32_t* epoch;                    // local: [n_ctas] number of completed calls
  int rank, n;
  long long off, count;               // floats; off and count multiples of 4
  float scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all ranks' CTA `blockIdx.x` meet: thread p tells rank p "I am here" and waits for rank p's word in the local pad
__device__ __forceinline__ void cross_gpu_barrier(const ArArgs& a, int phase, uint32_t e) {
  __syncthreads();
  if ((int)threadIdx.x < a.n) {
    const int p = threadIdx.x;
    const size_t slot = ((size_t)blockIdx.x * 2 + phase) * AR_MAX_RANKS;
    __threadfence_system();
    st_release_sys(a.flags[p] + slot + a.rank, e);
    const uint32_t* mine = a.flags[a.rank] + slot + p;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {
        printf("mcnerf: all-reduce barrier timeout: rank %d waits for rank %d (cta %d phase %d epoch %u)\n", a.rank, p,
               blockIdx.x, phase, e);
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_p2p_k(const __grid_constant__ ArArgs a) {
  const uint32_t e = a.epoch[blockIdx.x] + 1;
  cross_gpu_barrier(a, 0, e);
  // slice of this rank, in float4 units, split over the CTAs
  const long long n4 = a.count / 4;
  const long long s0 = n4 * a.rank / a.n, s1 = n4 * (a.rank + 1) / a.n;
  const float4* src[AR_MAX_RANKS];
  float4* dst[AR_MAX_RANKS];
#pragma unroll
  for (int p = 0; p < AR_MAX_RANKS; ++p) {
    src[p] = reinterpret_cast<const float4*>(a.buf[p < a.n ? p : 0] + a.off);
    dst[p] = reinterpret_cast<float4*>(a.buf[p < a.n ? p : 0] + a.off);
  }
  for (long long i = s0 + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < s1; i += (long long)gridDim.x * AR_THREADS) {
    float4 v[AR_MAX_RANKS];
#pragma unroll
    for (int p = 0; p < AR_MAX_RANKS; ++p)
      if (p < a.n) v[p] = src[p][i];                       // all loads in flight before the first add
    float4 s = v[0];
#pragma unroll
    for (int p = 1; p < AR_MAX_RANKS; ++p)
      if (p < a.n) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }      // rank order: same bits everywhere
    s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
    for (int p = 0; p < AR_MAX_RANKS; ++p)
      if (p < a.n) dst[p][i] = s;
  }
  cross_gpu_barrier(a, 1, e);
  if (threadIdx.x == 0) a.epoch[blockIdx.x] = e;
}

}  // namespace

extern "C" int mcnerf_allreduce_p2p(const mcnerf_p2p* ctx, int64_t offset, int64_t count, float scale, void* stream) {
  MC_ARG(ctx && ctx->n_ranks >= 1 && ctx->n_ranks <= AR_MAX_RANKS && ctx->rank >= 0 && ctx->rank < ctx->n_ranks &&
         ctx->n_ctas >= 1 && ctx->n_ctas <= 64 && ctx->epoch && offset >= 0 && count >= 0 && offset % 4 == 0 && count % 4 == 0);
  if (count == 0 || ctx->n_ranks == 1) return 0;
  ArArgs a;
  for (int p = 0; p < AR_MAX_RANKS; ++p) {
    a.buf[p] = (float*)ctx->buf[p < ctx->n_ranks ? p : 0];
    a.flags[p] = (uint32_t*)ctx->flags[p < ctx->n_ranks ? p : 0];
    MC_ARG(a.buf[p] && a.flags[p] && ((uintptr_t)a.buf[p] & 15) == 0);
  }
  a.epoch = (uint32_t*)ctx->epoch;
  a.rank = ctx->rank; a.n = ctx->n_ranks; a.off = offset; a.count = count; a.scale = scale;
  allreduce_p2p_k<<<ctx->n_ctas, AR_THREADS, 0, (cudaStream_t)stream>>>(a);
  MC_LAUNCHED();
  return 0;
}

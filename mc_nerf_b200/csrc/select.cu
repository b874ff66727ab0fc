// Threshold-gated fine-sample selection without host synchronisation, and the scatter/gather that
// moves MLP results between the compacted and the dense [B,Sf,4] layouts.
// ref: model/mc_nerf.py:623-629 / 663-667 (nonzero + x scale expansion), :692-694, 700-701 (defaults + index_put).
#include "common.cuh"

namespace {

// per-ray count of kept coarse samples; one warp per ray
__global__ void select_count_k(const float* __restrict__ w, const float* __restrict__ w_max, int n_rays, int Sc,
                               float thresh, int scale, int32_t* __restrict__ counts) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  float thr = fminf(thresh, *w_max);
  int cnt = 0;
  for (int k = lane; k < Sc; k += 32) cnt += (w[(size_t)ray * Sc + k] >= thr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) counts[ray] = cnt * scale;
}

// single-block exclusive scan over rays (n_rays <= a few 100k: one block of 1024 threads strides it)
__global__ void select_scan_k(const int32_t* __restrict__ counts, int n_rays, int32_t* __restrict__ offsets,
                              int32_t* __restrict__ total) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_rays; base += blockDim.x) {
    int i = base + tid;
    int v = i < n_rays ? counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int t = warp_tot[lane];
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;   // exclusive prefix of warp totals
    }
    __syncthreads();
    int c = carry;
    if (i < n_rays) offsets[i] = c + warp_tot[wid] + inc - v;
    __syncthreads();
    if (tid == blockDim.x - 1) carry = c + warp_tot[wid] + inc;
    __syncthreads();
  }
  if (tid == 0) {
    offsets[n_rays] = carry;
    *total = carry;
  }
}

__global__ void select_emit_k(const float* __restrict__ w, const float* __restrict__ w_max, int n_rays, int Sc,
                              float thresh, int scale, const int32_t* __restrict__ offsets,
                              int32_t* __restrict__ sel_idx) {
  int lane = threadIdx.x & 31;
  int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  float thr = fminf(thresh, *w_max);
  int pos = offsets[ray];
  int Sf = Sc * scale;
  for (int base = 0; base < Sc; base += 32) {
    int k = base + lane;
    bool keep = (k < Sc) && (w[(size_t)ray * Sc + k] >= thr);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      int p = pos + __popc(m & ((1u << lane) - 1u)) * scale;
      for (int j = 0; j < scale; ++j) sel_idx[p + j] = ray * Sf + k * scale + j;
    }
    pos += __popc(m) * scale;
  }
}

__global__ void fill_default_k(float4* __restrict__ out, int n, float sigma_default) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(sigma_default, 1.f, 1.f, 1.f);
}

__global__ void scatter_k(const float4* __restrict__ src, const int32_t* __restrict__ idx, int n,
                          const int32_t* __restrict__ n_dev, float4* __restrict__ dst) {
  int rows = n_dev ? min(*n_dev, n) : n;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) dst[idx[i]] = src[i];
}

__global__ void gather_k(const float4* __restrict__ src, const int32_t* __restrict__ idx, int n,
                         const int32_t* __restrict__ n_dev, float4* __restrict__ dst) {
  int rows = n_dev ? min(*n_dev, n) : n;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) dst[i] = src[idx[i]];
}

// ------------------------------------------------------------------ uniform K-subset of the selected samples
// ref: model/mc_nerf.py:630-632 (train-only cap: `rand_idx = torch.randperm(n)[:K]` on the CPU, after a host sync).
// Same distribution without leaving the device and without sorting: slot i < n gets the 32-bit key
//   key(i) = mix(i * 0x9E3779B1 + seed0) ^ seed1,   mix = the murmur3 finaliser,
// which is a BIJECTION of i, so all keys are distinct; the K-th smallest key T is found by a two-level radix select
// (65536-bin histograms of the high, then the low 16 bits) and {i : key(i) <= T} - exactly K slots - is compacted in
// ascending i.  n <= K keeps everything.  Deterministic for a given seed.
__device__ __forceinline__ uint32_t cap_key(uint32_t i, uint32_t s0, uint32_t s1) {
  uint32_t h = i * 0x9E3779B1u + s0;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h ^ s1;
}
struct CapState { uint32_t b1, below1, all, T, n_out; };
constexpr int CAP_BINS = 65536, CAP_BLOCK = 1024;

__global__ void cap_hist_k(const int32_t* __restrict__ n_dev, int capacity, const int64_t* __restrict__ seed, int level,
                           const CapState* __restrict__ st, uint32_t* __restrict__ hist) {
  const int n = min(*n_dev, capacity);
  const uint32_t s0 = (uint32_t)seed[0], s1 = (uint32_t)seed[1];
  if (level == 1 && st->all) return;
  const uint32_t b1 = level == 1 ? st->b1 : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t k = cap_key((uint32_t)i, s0, s1);
    if (level == 0) atomicAdd(&hist[k >> 16], 1u);
    else if ((k >> 16) == b1) atomicAdd(&hist[k & 0xFFFFu], 1u);
  }
}

// one block: smallest bin b whose inclusive prefix reaches the wanted rank; level 0 -> (b1, below1, all), level 1 -> T, n_out
__global__ void __launch_bounds__(CAP_BLOCK) cap_pick_k(const uint32_t* __restrict__ hist, int K, int level, CapState* st,
                                                        int32_t* __restrict__ n_out_dev) {
  __shared__ uint32_t part[CAP_BLOCK];
  __shared__ uint32_t warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int PER = CAP_BINS / CAP_BLOCK;       // 64 consecutive bins per thread
  uint32_t sum = 0;
  for (int j = 0; j < PER; ++j) sum += hist[tid * PER + j];
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;
  }
  __syncthreads();
  const uint32_t excl = warp_tot[wid] + inc - sum;      // keys in bins before this thread's
  part[tid] = excl;
  __syncthreads();
  const uint32_t total = part[CAP_BLOCK - 1] + (tid == CAP_BLOCK - 1 ? sum : 0);
  __shared__ uint32_t total_s;
  if (tid == CAP_BLOCK - 1) total_s = total;
  __syncthreads();
  uint32_t want;                                        // 1-based rank inside this histogram
  if (level == 0) {
    if (total_s <= (uint32_t)K) {
      if (tid == 0) { st->all = 1; st->b1 = 0; st->below1 = 0; st->T = 0xFFFFFFFFu; st->n_out = total_s; *n_out_dev = (int32_t)total_s; }
      return;
    }
    want = (uint32_t)K;
  } else {
    if (st->all) return;
    want = (uint32_t)K - st->below1;
  }
  if (excl < want && want <= excl + sum) {              // the wanted rank falls into this thread's 64 bins
    uint32_t c = excl;
    for (int j = 0; j < PER; ++j) {
      const uint32_t h = hist[tid * PER + j];
      if (c + h >= want) {
        if (level == 0) { st->all = 0; st->b1 = (uint32_t)(tid * PER + j); st->below1 = c; }
        else { st->T = (st->b1 << 16) | (uint32_t)(tid * PER + j); st->n_out = (uint32_t)K; *n_out_dev = K; }
        break;
      }
      c += h;
    }
  }
}

// per block of CAP_BLOCK slots: how many are kept
__global__ void __launch_bounds__(CAP_BLOCK) cap_count_k(const int32_t* __restrict__ n_dev, int capacity,
                                                         const int64_t* __restrict__ seed, const CapState* __restrict__ st,
                                                         int32_t* __restrict__ block_counts) {
  const int n = min(*n_dev, capacity);
  const int i = blockIdx.x * CAP_BLOCK + threadIdx.x;
  const bool keep = i < n && cap_key((uint32_t)i, (uint32_t)seed[0], (uint32_t)seed[1]) <= st->T;
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(CAP_BLOCK) cap_write_k(const int32_t* __restrict__ sel_idx, const int32_t* __restrict__ n_dev,
                                                         int capacity, const int64_t* __restrict__ seed,
                                                         const CapState* __restrict__ st, const int32_t* __restrict__ block_offs,
                                                         int K, int32_t* __restrict__ out_idx) {
  __shared__ int warp_tot[32];
  const int n = min(*n_dev, capacity);
  const int i = blockIdx.x * CAP_BLOCK + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool keep = i < n && cap_key((uint32_t)i, (uint32_t)seed[0], (uint32_t)seed[1]) <= st->T;
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_tot[wid] = __popc(m);
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;
  }
  __syncthreads();
  if (keep) {
    const int pos = block_offs[blockIdx.x] + warp_tot[wid] + __popc(m & ((1u << lane) - 1));
    if (pos < K) out_idx[pos] = sel_idx[i];
  }
}

}  // namespace

extern "C" int mcnerf_select_fine(const float* weights, const float* w_max, int n_rays, int Sc, int scale,
                                  float thresh, int32_t* sel_idx, int32_t* sel_offsets, int32_t* n_sel, void* stream) {
  MC_ARG(weights && w_max && sel_idx && sel_offsets && n_sel && n_rays > 0 && Sc > 0 && scale > 0);
  cudaStream_t st = (cudaStream_t)stream;
  // counts are staged in sel_offsets[0..B) and scanned in place via a temporary shift: use sel_idx tail as scratch
  int32_t* counts = sel_idx + (size_t)n_rays * Sc * scale - n_rays;   // last B entries of the index buffer
  select_count_k<<<cdiv(n_rays, 8), 256, 0, st>>>(weights, w_max, n_rays, Sc, thresh, scale, counts);
  MC_LAUNCHED();
  select_scan_k<<<1, 1024, 0, st>>>(counts, n_rays, sel_offsets, n_sel);
  MC_LAUNCHED();
  select_emit_k<<<cdiv(n_rays, 8), 256, 0, st>>>(weights, w_max, n_rays, Sc, thresh, scale, sel_offsets, sel_idx);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_scatter_fine(const float* out_sel, const int32_t* sel_idx, int n_sel, const int32_t* n_sel_dev,
                                   int n_dense_rows, float sigma_default, float* out_dense, void* stream) {
  MC_ARG(out_dense && n_dense_rows > 0 && n_sel >= 0 && ((uintptr_t)out_dense & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  fill_default_k<<<cdiv(n_dense_rows, 256), 256, 0, st>>>((float4*)out_dense, n_dense_rows, sigma_default);
  MC_LAUNCHED();
  if (n_sel > 0) {
    MC_ARG(out_sel && sel_idx && ((uintptr_t)out_sel & 15) == 0);
    scatter_k<<<cdiv(n_sel, 256), 256, 0, st>>>((const float4*)out_sel, sel_idx, n_sel, n_sel_dev, (float4*)out_dense);
    MC_LAUNCHED();
  }
  return 0;
}

extern "C" int mcnerf_gather_fine(const float* g_dense, const int32_t* sel_idx, int n_sel, const int32_t* n_sel_dev,
                                  float* g_sel, void* stream) {
  MC_ARG(n_sel >= 0);
  if (n_sel == 0) return 0;
  MC_ARG(g_dense && sel_idx && g_sel && ((uintptr_t)g_dense & 15) == 0 && ((uintptr_t)g_sel & 15) == 0);
  gather_k<<<cdiv(n_sel, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)g_dense, sel_idx, n_sel, n_sel_dev,
                                                               (float4*)g_sel);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_cap_select_workspace(int capacity, size_t* bytes) {
  MC_ARG(capacity >= 0 && bytes);
  const size_t nb = (size_t)(capacity + CAP_BLOCK - 1) / CAP_BLOCK;
  *bytes = 2 * (size_t)CAP_BINS * 4 + 64 + (2 * nb + 2) * 4;
  return 0;
}

extern "C" int mcnerf_cap_select(const int32_t* sel_idx, const int32_t* n_sel_dev, int capacity, int K, const int64_t* seed,
                                 int32_t* out_idx, int32_t* n_out_dev, void* workspace, void* stream) {
  MC_ARG(sel_idx && n_sel_dev && seed && out_idx && n_out_dev && workspace && capacity > 0 && K > 0);
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* hist1 = (uint32_t*)workspace;
  uint32_t* hist2 = hist1 + CAP_BINS;
  CapState* state = (CapState*)(hist2 + CAP_BINS);
  int32_t* counts = (int32_t*)((uint8_t*)state + 64);
  const int nb = (capacity + CAP_BLOCK - 1) / CAP_BLOCK;
  int32_t* offs = counts + nb + 1;
  MC_CUDA(cudaMemsetAsync(workspace, 0, 2 * (size_t)CAP_BINS * 4 + 64, st));
  const int hb = nb < 1184 ? nb : 1184;                  // 8 blocks per SM, grid-stride
  cap_hist_k<<<hb, CAP_BLOCK, 0, st>>>(n_sel_dev, capacity, seed, 0, state, hist1);
  MC_LAUNCHED();
  cap_pick_k<<<1, CAP_BLOCK, 0, st>>>(hist1, K, 0, state, n_out_dev);
  MC_LAUNCHED();
  cap_hist_k<<<hb, CAP_BLOCK, 0, st>>>(n_sel_dev, capacity, seed, 1, state, hist2);
  MC_LAUNCHED();
  cap_pick_k<<<1, CAP_BLOCK, 0, st>>>(hist2, K, 1, state, n_out_dev);
  MC_LAUNCHED();
  cap_count_k<<<nb, CAP_BLOCK, 0, st>>>(n_sel_dev, capacity, seed, state, counts);
  MC_LAUNCHED();
  int32_t* total = offs + nb;                            // offs has nb + 1 entries
  select_scan_k<<<1, 1024, 0, st>>>(counts, nb, offs, total);
  MC_LAUNCHED();
  cap_write_k<<<nb, CAP_BLOCK, 0, st>>>(sel_idx, n_sel_dev, capacity, seed, state, offs, K, out_idx);
  MC_LAUNCHED();
  return 0;
}

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

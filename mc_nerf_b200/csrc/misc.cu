// Loss seed (MSE on the two renders) and the fused multi-tensor RAdam update on flat buffers.
// ref: model/loss.py:33-43 ; model/net_utils.py:10-101.
#include "common.cuh"
#include "philox.cuh"

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


// ------------------------------------------------------------------------------------------------------------
// Whole train-stage loss in one single-block launch (ref: model/loss.py:15-58): reprojection term of the 110x5
// calibration points (optionally normalised by its own detached magnitude), MSE of the coarse and fine renders, and
// the gradients of the total with respect to the renders and the reprojected pixels.
//   out[0] = total, out[1] = raw reprojection loss, out[2] = rgb loss
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();                    // sm may still be read from the previous reduction
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float r = lane < NT / 32 ? sm[lane] : 0.f;
  r = warp_sum(r);                    // every warp reduces the same NT/32 partials: same value in every thread
  return r;
}

constexpr int LOSS_THREADS = 1024;

__global__ void __launch_bounds__(LOSS_THREADS)
train_loss_k(const float* __restrict__ rc, const float* __restrict__ rf, const float* __restrict__ gt, int n_rays,
             const float* __restrict__ px, const float* __restrict__ px_gt, int n_pts, float inv_w, float inv_h,
             int normalise, float* __restrict__ out, float* __restrict__ g_c, float* __restrict__ g_f,
             float* __restrict__ g_px) {
  __shared__ float sm[32];
  const int tid = threadIdx.x;
  float l_px = 0.f, px_scale = 1.f;
  if (px) {
    float sx = 0.f, sy = 0.f;
    for (int i = tid; i < n_pts; i += LOSS_THREADS) {
      float dx = px[2 * i] * inv_w - px_gt[2 * i] * inv_w, dy = px[2 * i + 1] * inv_h - px_gt[2 * i + 1] * inv_h;
      sx += dx * dx;
      sy += dy * dy;
    }
    sx = block_sum<LOSS_THREADS>(sx, sm);
    sy = block_sum<LOSS_THREADS>(sy, sm);
    l_px = sx / (float)n_pts + sy / (float)n_pts;
    if (normalise) px_scale = 1.f / (l_px + 1e-8f);
    const float kx = 2.f * inv_w / (float)n_pts * px_scale, ky = 2.f * inv_h / (float)n_pts * px_scale;
    for (int i = tid; i < n_pts; i += LOSS_THREADS) {
      g_px[2 * i] = kx * (px[2 * i] * inv_w - px_gt[2 * i] * inv_w);
      g_px[2 * i + 1] = ky * (px[2 * i + 1] * inv_h - px_gt[2 * i + 1] * inv_h);
    }
  }
  const int n = 3 * n_rays;
  const float k = 2.f / (float)n;
  float sc = 0.f, sf = 0.f;
  for (int t = tid; t < n; t += LOSS_THREADS) {
    const float g = gt[t];
    const float dc = rc[t] - g;
    g_c[t] = k * dc;
    sc += dc * dc;
    if (rf) {
      const float df = rf[t] - g;
      g_f[t] = k * df;
      sf += df * df;
    }
  }
  sc = block_sum<LOSS_THREADS>(sc, sm);
  sf = block_sum<LOSS_THREADS>(sf, sm);
  if (tid == 0) {
    const float l_rgb = sc / (float)n + sf / (float)n;
    out[0] = l_px * px_scale + l_rgb;
    out[1] = l_px;
    out[2] = l_rgb;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Pixel choice of a train step: the first `batch` entries of a uniform random permutation of [0, n)
// (ref: model/mc_nerf.py:327-345, torch.randperm(H*W)[:batch]) WITHOUT sorting all n keys.  Every pixel i gets a
// 64-bit composite (random high bits from Philox4x32-10 keyed by a device-side seed | i in the low bits); the
// `batch` smallest composites in ascending order are exactly the head of the permutation that sorting all n random
// keys would give.  Pass 1 keeps the pixels whose random bits fall under a threshold chosen so that
// batch + 8 sqrt(batch) + 16 are expected (P[fewer than batch] < 1e-12); pass 2 ranks those few thousand candidates
// (one block, bucket counting sort) and emits the head.
__device__ __forceinline__ uint64_t pixel_key(int i, const int64_t* seed, int idx_bits) {
  const uint64_t s0 = (uint64_t)seed[0], s1 = (uint64_t)seed[1];
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, (uint32_t)s1, (uint32_t)(s1 >> 32)),
                                make_uint2((uint32_t)s0, (uint32_t)(s0 >> 32)));
  return ((((uint64_t)r.x << 32) | r.y) >> idx_bits);          // 64 - idx_bits random bits
}

__global__ void pixel_keys_k(int n, const int64_t* __restrict__ seed, int idx_bits, uint64_t tau, int cap,
                             uint64_t* __restrict__ cand, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = pixel_key(i, seed, idx_bits);
  if (r < tau) {
    const int slot = atomicAdd(count, 1);
    if (slot < cap) cand[slot] = (r << idx_bits) | (uint64_t)i;
  }
}

// One block.  The candidates' random bits are uniform below tau, so a 1024-bucket counting sort on their top bits
// (histogram -> scan -> scatter into bucket order) leaves ~c/1024 keys per bucket; the rank of a key is its bucket's
// start plus the number of smaller keys inside the bucket.  A few passes over c*8 bytes, against 91 full passes of
// a bitonic network.
__global__ void __launch_bounds__(1024)
pixel_pick_k(int n, int n_out, const int64_t* __restrict__ seed, int idx_bits, uint64_t tau, int shift, int cap,
             const uint64_t* __restrict__ cand, uint64_t* __restrict__ cand2, int* __restrict__ count,
             int64_t* __restrict__ out64, int32_t* __restrict__ out32) {
  __shared__ int start[1025];
  __shared__ int cursor[1024];
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c = min(*count, cap);
  start[tid] = 0;
  cursor[tid] = 0;
  __syncthreads();
  for (int i = tid; i < c; i += 1024) atomicAdd(&start[min((int)((cand[i] >> idx_bits) >> shift), 1023)], 1);
  __syncthreads();
  const int v = start[tid];
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    const int w = wsum[lane];
    int z = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += y;
    }
    wsum[lane] = z - w;
  }
  __syncthreads();
  const int excl = x - v + wsum[wid];
  start[tid] = excl;
  if (tid == 1023) start[1024] = excl + v;
  __syncthreads();
  for (int i = tid; i < c; i += 1024) {
    const uint64_t key = cand[i];
    const int b = min((int)((key >> idx_bits) >> shift), 1023);
    cand2[start[b] + atomicAdd(&cursor[b], 1)] = key;
  }
  __syncthreads();
  const uint64_t mask = (1ull << idx_bits) - 1;
  for (int p = tid; p < c; p += 1024) {
    const uint64_t e = cand2[p];
    const int b = min((int)((e >> idx_bits) >> shift), 1023);
    const int lo = start[b], hi = start[b + 1];
    int rank = lo;
    for (int q = lo; q < hi; ++q) rank += cand2[q] < e;
    if (rank < n_out) {
      const int64_t idx = (int64_t)(e & mask);
      out64[rank] = idx;
      if (out32) out32[rank] = (int32_t)idx;
    }
  }
  // Fewer candidates than requested (probability < 1e-12 per call): top up with the lowest-numbered pixels that were
  // NOT candidates, which keeps the indices distinct.
  const int have = min(c, n_out);
  if (have < n_out && tid == 0) {
    int w = have;
    for (int i = 0; i < n && w < n_out; ++i)
      if (pixel_key(i, seed, idx_bits) >= tau) {
        out64[w] = i;
        if (out32) out32[w] = i;
        ++w;
      }
  }
  __syncthreads();
  if (tid == 0) *count = 0;       // ready for the next call (and the next graph replay)
}


// ------------------------------------------------------------------------------------------------------------
// dst[r*dst_ld + c] = src[r*src_ld + c] for the top-left rows x cols block of up to 64 matrices in one launch:
// zero-pads a narrow network's parameters into 256-wide shadows (tensor-core path for widths < 256) and cuts the
// valid blocks back out of the 256-wide gradients.
constexpr int COPY_MAX_JOBS = 64;
struct CopyBlocks {
  int n;
  const float* src[COPY_MAX_JOBS];
  float* dst[COPY_MAX_JOBS];
  int rows[COPY_MAX_JOBS], cols[COPY_MAX_JOBS], src_ld[COPY_MAX_JOBS], dst_ld[COPY_MAX_JOBS];
  int blk_prefix[COPY_MAX_JOBS + 1];
};

__global__ void copy_blocks_k(const __grid_constant__ CopyBlocks a) {
  int job = 0;
  while ((int)blockIdx.x >= a.blk_prefix[job + 1]) ++job;
  const int i = ((int)blockIdx.x - a.blk_prefix[job]) * blockDim.x + threadIdx.x;
  const int cols = a.cols[job];
  if (i >= a.rows[job] * cols) return;
  const int r = i / cols, c = i - r * cols;
  a.dst[job][(size_t)r * a.dst_ld[job] + c] = a.src[job][(size_t)r * a.src_ld[job] + c];
}


// dst[0..n) = values passed BY VALUE in the launch (n <= 16): a stream-ordered way to refresh a few device-side
// scalars (the BARF band weights read by captured graphs) without a pinned staging buffer that the host could
// overwrite before the copy engine reads it.
struct FloatPack { float v[16]; };
__global__ void store_floats_k(float* __restrict__ dst, FloatPack p, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = p.v[threadIdx.x];
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

extern "C" int mcnerf_train_loss(const float* rgb_c, const float* rgb_f, const float* gt, int n_rays, const float* px,
                                 const float* px_gt, int n_pts, int img_w, int img_h, int normalise, float* out3,
                                 float* g_c, float* g_f, float* g_px, void* stream) {
  MC_ARG(rgb_c && gt && g_c && out3 && n_rays > 0 && (!rgb_f || g_f));
  MC_ARG(!px || (px_gt && g_px && n_pts > 0 && img_w > 0 && img_h > 0));
  train_loss_k<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(rgb_c, rgb_f, gt, n_rays, px, px_gt, n_pts,
                                                             1.f / (float)img_w, 1.f / (float)img_h, normalise, out3,
                                                             g_c, g_f, g_px);
  MC_LAUNCHED();
  return 0;
}

static int pixel_plan(int n, int batch, int* idx_bits, uint64_t* tau, int* cap, int* shift) {
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  const int n_out = batch < n ? batch : n;
  const double root = sqrt((double)n_out);
  double want = n_out + 8.0 * root + 16.0;
  int c = 1024;
  while (c < n_out + 16.0 * root + 64.0) c <<= 1;
  const int rand_bits = 64 - bits;
  if (want >= n) {              // nearly everything is asked for: every pixel is a candidate
    *tau = ~0ull;
    c = 1024;
    while (c < n) c <<= 1;
  } else {
    *tau = (uint64_t)((want / (double)n) * ldexp(1.0, rand_bits));
  }
  uint64_t range_max = *tau - 1;                       // random bits r satisfy r <= range_max
  if (rand_bits < 64 && range_max > (1ull << rand_bits) - 1) range_max = (1ull << rand_bits) - 1;
  int bl = 0;
  while (bl < 64 && (range_max >> bl) != 0) ++bl;
  *shift = bl > 10 ? bl - 10 : 0;                       // r >> shift < 1024: the bucket of the counting sort
  *idx_bits = bits;
  *cap = c;
  return n_out;
}

extern "C" int mcnerf_sample_pixels_workspace(int n, int batch, size_t* bytes) {
  MC_ARG(n > 0 && batch > 0 && bytes);
  int bits, cap, shift;
  uint64_t tau;
  pixel_plan(n, batch, &bits, &tau, &cap, &shift);
  if (cap > 16384) {
    mcnerf_set_error("mcnerf_sample_pixels: batch %d of %d pixels needs %d sort slots (max 16384)", batch, n, cap);
    return MCNERF_E_ARG;
  }
  *bytes = 16 + (size_t)cap * 16;          /* counter | candidates | candidates in bucket order */
  return 0;
}

extern "C" int mcnerf_sample_pixels(int n, int batch, const int64_t* seed, void* workspace, int64_t* out_idx,
                                    int32_t* out_idx32, void* stream) {
  MC_ARG(n > 0 && batch > 0 && seed && workspace && out_idx);
  int bits, cap, shift;
  uint64_t tau;
  const int n_out = pixel_plan(n, batch, &bits, &tau, &cap, &shift);
  MC_ARG(cap <= 16384);
  int* count = (int*)workspace;                                   // zero on first use, reset by pixel_pick_k
  uint64_t* cand = (uint64_t*)((char*)workspace + 16);
  pixel_keys_k<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, seed, bits, tau, cap, cand, count);
  MC_LAUNCHED();
  pixel_pick_k<<<1, 1024, 0, (cudaStream_t)stream>>>(n, n_out, seed, bits, tau, shift, cap, cand, cand + cap, count,
                                                     out_idx, out_idx32);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_copy_blocks(int n_jobs, const float* const* src, float* const* dst, const int* rows,
                                  const int* cols, const int* src_ld, const int* dst_ld, void* stream) {
  MC_ARG(n_jobs >= 0);
  if (n_jobs == 0) return 0;
  MC_ARG(src && dst && rows && cols && src_ld && dst_ld);
  for (int base = 0; base < n_jobs; base += COPY_MAX_JOBS) {
    CopyBlocks a;
    a.n = n_jobs - base < COPY_MAX_JOBS ? n_jobs - base : COPY_MAX_JOBS;
    int blocks = 0;
    for (int t = 0; t < a.n; ++t) {
      const int j = base + t;
      MC_ARG(src[j] && dst[j] && rows[j] > 0 && cols[j] > 0 && src_ld[j] >= cols[j] && dst_ld[j] >= cols[j]);
      a.src[t] = src[j]; a.dst[t] = dst[j]; a.rows[t] = rows[j]; a.cols[t] = cols[j];
      a.src_ld[t] = src_ld[j]; a.dst_ld[t] = dst_ld[j];
      a.blk_prefix[t] = blocks;
      blocks += cdiv((int64_t)rows[j] * cols[j], 256);
    }
    for (int t = a.n; t <= COPY_MAX_JOBS; ++t) a.blk_prefix[t] = blocks;
    copy_blocks_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    MC_LAUNCHED();
  }
  return 0;
}

extern "C" int mcnerf_store_floats(float* dst, const float* host_values, int n, void* stream) {
  MC_ARG(dst && host_values && n >= 1 && n <= 16);
  FloatPack p;
  for (int i = 0; i < 16; ++i) p.v[i] = i < n ? host_values[i] : 0.f;
  store_floats_k<<<1, 32, 0, (cudaStream_t)stream>>>(dst, p, n);
  MC_LAUNCHED();
  return 0;
}

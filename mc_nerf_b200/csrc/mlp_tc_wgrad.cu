// bf16 tcgen05 NeRF MLP, weight gradients: dW_l = dY_l^T X_l with the reduction running over ALL sample rows.
// Both operands are the UMMA tile images the forward / backward-chain kernels left in HBM (activation stash and
// dY stash); read as MN-major operands (reduction index = row) they need no transposition.
// Grid = one CTA per SM, partitioned over the (layer, tile-range) jobs in proportion to their HBM traffic:
//   warp 0   : producer - every lane issues 1 KB bulk copies (one k-group plane of 64 rows each) into a 3-stage ring
//   warp 1   : MMA issuer - M = 256 output features as two 128-lane halves, N = input features (<= 256), K = 64 rows
//              per stage; fp32 accumulators fill TMEM (2 x 256 columns)
//   warps 2-5: epilogue - TMEM -> per-CTA partial in a scratch buffer; a second small kernel sums the partials into
//              the fp32 parameter gradients (deterministic, no fp32 atomics) and column-sums dY for the biases.
// HBM-bound by construction (64 KB of operands per 2 x 4 MMAs): see DESIGN.md for the roofline.
#include <math.h>
#include <stdlib.h>
#include "mlp_tc.cuh"

namespace mlptc {


constexpr int WG_ROWS = 64;                          // rows per stage
constexpr int WG_PLANE = WG_ROWS * 16;               // 1 KB: one k-group plane of a 64-row half tile
constexpr int WG_STAGE = 2 * 32 * WG_PLANE;          // 64 KB: A region (32 planes) + B region (<= 32 planes)
constexpr int WG_NSTAGE = 3;
constexpr int SMEM_WG = WG_NSTAGE * WG_STAGE + 256;
constexpr int WG_THREADS = 192;

struct WArgs {
  WPlan plan;
  int n_slots;
  const uint8_t *stash, *stash_enc, *dy, *dy_head;
  int n_rows;
  const int32_t* n_rows_dev;
  float* scratch;                    // [gridDim.x][2][128][256] fp32 partials
  int cta_begin[MAX_STEPS + 1];
  float* bias_part;                  // [gridDim.x][256] per-CTA column sums (bias gradients), reduced by wgrad_reduce_k
};

struct __align__(16) WBars {
  uint64_t full[WG_NSTAGE], empty[WG_NSTAGE], acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1) mlp_tc_wgrad_k(const __grid_constant__ WArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  WBars* bars = reinterpret_cast<WBars*>(smem + WG_NSTAGE * WG_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int n_tiles = (rows + TM - 1) / TM;
  int job = 0;
  while (job + 1 < a.plan.n_jobs && (int)blockIdx.x >= a.cta_begin[job + 1]) ++job;
  const WJob& wj = a.plan.j[job];
  const int part = blockIdx.x - a.cta_begin[job], nparts = a.cta_begin[job + 1] - a.cta_begin[job];
  const int t_begin = (int)((long long)n_tiles * part / nparts), t_end = (int)((long long)n_tiles * (part + 1) / nparts);
  const int n_stages_total = (t_end - t_begin) * 2;
  const int N = wj.N;

  if (tid == 0) {
    for (int i = 0; i < WG_NSTAGE; ++i) { tc::mbar_init(&bars->full[i], 1); tc::mbar_init(&bars->empty[i], 5); }
    tc::mbar_init(&bars->acc_full, 1);
    tc::mbar_init_fence();
  }
  if (warp == 1) tc::tmem_alloc(&bars->tmem_base, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // operand sources: each 64-row half of a tile image is one contiguous [plane][64 rows][16 B] block.
  // A: all 32 planes (256 features) = 32 KB; B: N/8 planes starting at plane b_plane0 of a b_planes-plane image.
  const size_t tile_stride = (size_t)a.n_slots * ACT_BYTES;
  const uint8_t *a_base, *b_base;
  size_t a_stride, b_stride;
  int b_planes, b_plane0 = 0;
  if (!wj.transposed) {
    a_base = a.dy + (size_t)wj.dy_slot * ACT_BYTES; a_stride = tile_stride;
    if (wj.x_slot < 0) { b_base = a.stash_enc; b_stride = ENC_BYTES; b_planes = 8; }
    else { b_base = a.stash + (size_t)wj.x_slot * ACT_BYTES; b_stride = tile_stride; b_planes = 32; }
  } else {
    a_base = a.stash + (size_t)wj.x_slot * ACT_BYTES; a_stride = tile_stride;
    b_base = a.dy_head; b_stride = HEAD_BYTES; b_planes = 4; b_plane0 = wj.head_col0 / 8;
  }
  const int nb_planes = N / 8;
  const uint32_t a_bytes = 32 * WG_PLANE, b_bytes = (uint32_t)nb_planes * WG_PLANE;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      for (int i = 0; i < n_stages_total; ++i) {
        const int tile = t_begin + (i >> 1), half = i & 1;
        tc::mbar_wait(&bars->empty[stage], par ^ 1);
        tc::mbar_arrive_expect_tx(&bars->full[stage], a_bytes + b_bytes);
        uint8_t* sdst = smem + stage * WG_STAGE;
        tc::bulk_g2s(sdst, a_base + (size_t)tile * a_stride + (size_t)half * a_bytes, a_bytes, &bars->full[stage]);
        tc::bulk_g2s(sdst + 32 * WG_PLANE,
                     b_base + (size_t)tile * b_stride + (size_t)half * b_planes * WG_PLANE + (size_t)b_plane0 * WG_PLANE,
                     b_bytes, &bars->full[stage]);
        if (++stage == WG_NSTAGE) { stage = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_stages_total > 0) {
      const uint32_t idesc = tc::umma_idesc_bf16(128, N, 1, 1);
      // both operands MN-major: lead (K-direction) offset 128 B, stride (MN-direction) offset = one plane
      const uint32_t hi = tc::umma_desc_hi(WG_PLANE);
      const uint32_t s_lo0 = tc::umma_desc_lo(tc::smem_u32(smem), 128);
      const uint32_t full0 = tc::smem_u32(&bars->full[0]), empty0 = tc::smem_u32(&bars->empty[0]);
      int stage = 0;
      uint32_t par = 0;
      for (int i = 0; i < n_stages_total; ++i) {
        tc::mbar_wait_addr(full0 + stage * 8, par);
        tc::tcgen05_fence_after();
        const uint32_t a_lo = s_lo0 + stage * (WG_STAGE >> 4), b_lo = a_lo + ((32 * WG_PLANE) >> 4);
#pragma unroll
        for (int k16 = 0; k16 < WG_ROWS / 16; ++k16) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tc::umma_bf16_w(tmem + h * 256, a_lo + ((h * 16 * WG_PLANE + k16 * 256) >> 4), hi, b_lo + ((k16 * 256) >> 4), hi,
                            idesc, (i | k16) != 0);
        }
        tc::umma_commit_addr(empty0 + stage * 8);
        if (++stage == WG_NSTAGE) { stage = 0; par ^= 1; }
      }
      tc::umma_commit(&bars->acc_full);
    }
  } else {
    // ---- warps 2-5.  While the MMAs run: bias gradients = column sums of the dY operand sitting in the ring
    // (warp w owns 8 of the 32 planes; lane l owns rows l and l+32).  Afterwards: TMEM -> scratch partials.
    const int w4 = warp - 2;
    const int bm = wj.bias_mode;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    {
      int stage = 0;
      uint32_t par = 0;
      for (int i = 0; i < n_stages_total; ++i) {
        tc::mbar_wait(&bars->full[stage], par);
        const uint8_t* sA = smem + stage * WG_STAGE;
        if (bm == 1) {
#pragma unroll
          for (int pl = 0; pl < 8; ++pl)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const uint4 v = *reinterpret_cast<const uint4*>(sA + (w4 * 8 + pl) * WG_PLANE + (lane + 32 * rr) * 16);
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[pl][2 * e] += __uint_as_float(w[e] << 16);
                acc[pl][2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
              }
            }
        } else if (bm >= 2 && w4 < nb_planes) {
          const uint8_t* sB = sA + 32 * WG_PLANE;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const uint4 v = *reinterpret_cast<const uint4*>(sB + w4 * WG_PLANE + (lane + 32 * rr) * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[0][2 * e] += __uint_as_float(w[e] << 16);
              acc[0][2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->empty[stage]);
        if (++stage == WG_NSTAGE) { stage = 0; par ^= 1; }
      }
    }
    // per-CTA partial column sums (no atomics: the reduction kernel adds them in a fixed order)
    float* bpart = a.bias_part + (size_t)blockIdx.x * 256;
    if (bm == 1) {
#pragma unroll
      for (int pl = 0; pl < 8; ++pl)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = warp_sum(acc[pl][j]);
          if (lane == 0) bpart[(w4 * 8 + pl) * 8 + j] = s;
        }
    } else if (bm >= 2 && w4 < nb_planes) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = warp_sum(acc[0][j]);
        if (lane == 0) bpart[(b_plane0 + w4) * 8 + j] = s;          // indexed by head-tile column
      }
    }
    // ---- epilogue proper: lane quarter (warp % 4), both halves
    float* out = a.scratch + (size_t)blockIdx.x * 2 * 128 * 256;
    const int m = (warp & 3) * 32 + lane;
    if (n_stages_total > 0) {
      tc::mbar_wait(&bars->acc_full, 0);
      tc::tcgen05_fence_after();
    }
    for (int h = 0; h < 2; ++h) {
      float* row = out + ((size_t)h * 128 + m) * 256;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        if (n_stages_total > 0) {
          tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + h * 256 + c0, v);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        // N is a multiple of 16: the last block of a 16-wide job only owns 16 valid columns
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + 4 * i < N)
            *reinterpret_cast<float4*>(row + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                       __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// Sum the per-CTA partials of every job into the parameter gradients (+=), mapping the padded / transposed
// accumulator coordinates back to the reference's [out,in] layout.
struct RArgs {
  WPlan plan;
  const float* scratch;
  int cta_begin[MAX_STEPS + 1];
  float* dst[MAX_STEPS];
  int ld[MAX_STEPS];
  const float* bias_part;            // [CTA][256] partial column sums
  float* db[MAX_STEPS];              // bias gradient of each job (null: none)
};

__global__ void __launch_bounds__(256) wgrad_reduce_k(const __grid_constant__ RArgs a) {
  const int job = blockIdx.y;
  const WJob& wj = a.plan.j[job];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // (m, n) of the 256 x N accumulator
  if (blockIdx.x == 0 && a.db[job]) {
    // bias gradient of this job: column sums of the dY operand (mode 1), of the head tile's SH columns (2) or its
    // g_sigma column (3), summed over the job's CTAs in a fixed order
    const int t = threadIdx.x;
    const bool mine = wj.bias_mode == 1 || (wj.bias_mode == 2 && t < 27) || (wj.bias_mode == 3 && t == 31);
    if (mine) {
      float s = 0.f;
      for (int p = a.cta_begin[job]; p < a.cta_begin[job + 1]; ++p) s += a.bias_part[(size_t)p * 256 + t];
      a.db[job][wj.bias_mode == 3 ? 0 : t] += s;
    }
  }
  const int N = wj.N;
  if (idx >= 256 * N) return;
  const int m = idx / N, n = idx - m * N;
  float s = 0.f;
  for (int p = a.cta_begin[job]; p < a.cta_begin[job + 1]; ++p)
    s += a.scratch[((size_t)p * 2 + (m >> 7)) * 128 * 256 + (size_t)(m & 127) * 256 + n];
  if (!wj.transposed) {
    if (n < wj.n_valid) a.dst[job][(size_t)m * a.ld[job] + wj.col_off + n] += s;     // m = out feature, n = in feature
  } else if (wj.N == 16) {
    if (n == 15) a.dst[job][m] += s;                                                 // head column 31 = g_sigma
  } else {
    if (n < 27) a.dst[job][(size_t)n * a.ld[job] + m] += s;                          // m = in feature, n = out feature
  }
}

}  // namespace mlptc

using namespace mlptc;

size_t mlp_tc_wgrad_scratch_bytes(int n_ctas) { return (size_t)n_ctas * 2 * 128 * 256 * sizeof(float); }

int mlp_tc_wgrad_launch(const mcnerf_mlp_params* p, const PackLayout& L, const uint8_t* stash, const uint8_t* stash_enc,
                        const uint8_t* dy, const uint8_t* dy_head, float* scratch, float* bias_part, int n_rows,
                        const int32_t* n_rows_dev, const mcnerf_mlp_grads* g, cudaStream_t st) {
  const int D = p->depth;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  WArgs a;
  a.plan = L.wg;
  a.n_slots = D + 2;
  a.stash = stash; a.stash_enc = stash_enc; a.dy = dy; a.dy_head = dy_head;
  a.n_rows = n_rows; a.n_rows_dev = n_rows_dev;
  a.scratch = scratch;
  a.bias_part = bias_part;
  // CTAs per job.  Every job streams n_tiles x 2 stages of (32 + N/8) KB and the kernel ends with the slowest job.
  // With the HBM pipe saturated a CTA's bandwidth share follows the bytes it keeps in flight (3 stages), so a job's
  // time per stage lies between "proportional to its stage size" and "the same for every job": weight = stage size
  // ^ alpha.  Measured (backward, 786 k rows): alpha 1: 2.505 ms, alpha 0.75 / 0.5 / 0.25 / 0: 2.430-2.435 ms = the HBM
  // floor (chain + 8.6 GB at the copy peak); 0.5 is the default (env MCNERF_WG_ALPHA for measurements).  The counts
  // are integers: greedy makespan minimisation (start with one CTA per job, give the next CTA to the job with the
  // largest weight per CTA); proportional rounding had left the sigma.2 job 15 % over the mean.
  const int nj = L.wg.n_jobs;
  MC_ARG(nj <= sms);
  int used = nj, cnt[MAX_STEPS];
  double w[MAX_STEPS];
  static const double alpha = getenv("MCNERF_WG_ALPHA") ? atof(getenv("MCNERF_WG_ALPHA")) : 0.5;
  for (int j = 0; j < nj; ++j) { cnt[j] = 1; w[j] = pow(32 + L.wg.j[j].N / 8, alpha); }
  for (; used < sms; ++used) {
    int best = 0;
    for (int j = 1; j < nj; ++j) if (w[j] / cnt[j] > w[best] / cnt[best]) best = j;
    ++cnt[best];
  }
  a.cta_begin[0] = 0;
  for (int j = 0; j < nj; ++j) a.cta_begin[j + 1] = a.cta_begin[j] + cnt[j];
  float* dbs[MAX_STEPS];
  for (int j = 0; j < nj; ++j) {
    const int which = L.wg.j[j].which;
    float* db = nullptr;
    if (L.wg.j[j].bias_mode == 1) db = which < D ? g->b[which] : (which == D ? g->b_sigma0 : g->b_sh0);
    else if (L.wg.j[j].bias_mode == 2) db = g->b_sh2;
    else if (L.wg.j[j].bias_mode == 3) db = g->b_sigma2;
    dbs[j] = db;
  }
  // per device / context attribute: set on every call (cheap), not once per process
  MC_CUDA(cudaFuncSetAttribute(mlp_tc_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WG));
  mlp_tc_wgrad_k<<<used, WG_THREADS, SMEM_WG, st>>>(a);
  MC_LAUNCHED();

  RArgs r;
  r.plan = L.wg;
  r.scratch = scratch;
  for (int j = 0; j <= nj; ++j) r.cta_begin[j] = a.cta_begin[j];
  auto wptr = [&](int which, int* ld) -> float* {
    if (which < D) {
      *ld = which == 0 ? 63 : ((p->skip_mask >> which & 1u) ? 63 + WID : WID);
      return g->W[which];
    }
    *ld = WID;
    if (which == D) return g->W_sigma0;
    if (which == D + 1) return g->W_sh0;
    if (which == D + 2) return g->W_sh2;
    return g->W_sigma2;
  };
  for (int j = 0; j < nj; ++j) r.dst[j] = wptr(L.wg.j[j].which, &r.ld[j]);
  r.bias_part = bias_part;
  for (int j = 0; j < nj; ++j) r.db[j] = dbs[j];
  wgrad_reduce_k<<<dim3(256, nj), 256, 0, st>>>(r);
  MC_LAUNCHED();

  return 0;
}

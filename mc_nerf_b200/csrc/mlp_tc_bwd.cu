// bf16 tcgen05 NeRF MLP, backward chain (data gradients): ONE persistent kernel that walks a 128-row tile from
// d(out4) back to the ray, entirely on-chip:
//   SH/sigmoid head backward (CUDA cores) -> dgrad GEMMs through sh.2, sh.0 + sigma.0, trunk layers D-1..1,
//   layer 0 + the skip layer's encoding part (tcgen05, transposed weight images streamed by bulk copies)
//   -> sin/cos encoding backward -> per-ray (dL/do, dL/dd) with a segmented warp reduction + atomics.
// Runs on CTA pairs (tcgen05 cta_group::2) exactly like the forward kernel (mlp_tc_fwd.cu).
// ReLU masks come from the forward activation stash; every layer's dY tile is written (bf16, UMMA tile image)
// to the dY stash that the weight-gradient kernel (mlp_tc_wgrad.cu) consumes.
// ref: autograd of model/net_block.py:67-78, model/net_utils.py:154-169, model/net_block.py:20-35.
#include <stdlib.h>
#include "mlp_tc_chain.cuh"

namespace mlptc {

// Ordered per-ray sums of the chain kernel's segment partials: ray r owns the consecutive rows [b, e); its segments
// start at b and at every multiple of 32 inside (b, e) (a warp of the chain kernel covers 32 consecutive rows).
// One thread per ray, ascending row order: bit-reproducible, unlike fp32 atomics.
__global__ void ray_grad_reduce_k(const float* __restrict__ seg_part, const int32_t* __restrict__ ray_offsets, int S,
                                  int n_rays, int n_rows, const int32_t* __restrict__ n_rows_dev,
                                  float* __restrict__ g_o, float* __restrict__ g_d) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const int rows = n_rows_dev ? min(*n_rows_dev, n_rows) : n_rows;
  int b = ray_offsets ? ray_offsets[r] : r * S, e = ray_offsets ? ray_offsets[r + 1] : (r + 1) * S;
  e = min(e, rows);
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int h = b; h < e; h = (h / 32 + 1) * 32) {
    const float* sp = seg_part + (size_t)h * 9;
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] += sp[c];
  }
  if (b < e) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g_o[3 * r + c] += acc[c];
      g_d[3 * r + c] += acc[3 + c] + acc[6 + c];
    }
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1) mlp_tc_bwd_k(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  chain_role<false>(a, smem, (int)gridDim.x, nullptr, 0);
}

}  // namespace mlptc

using namespace mlptc;

size_t mlp_tc_wgrad_scratch_bytes(int n_ctas);
int mlp_tc_wgrad_launch(const mcnerf_mlp_params* p, const PackLayout& L, const uint8_t* stash, const uint8_t* stash_enc,
                        const uint8_t* dy, const uint8_t* dy_head, float* scratch, float* bias_part, int n_rows,
                        const int32_t* n_rows_dev, const mcnerf_mlp_grads* g, cudaStream_t st);
size_t mlp_tc_fused_extra_bytes(int n_rows, int n_slots);
int mlp_tc_bwd_fused_launch(const mcnerf_mlp_params* p, const PackLayout& L, const BwdArgs& chain, const uint8_t* stash,
                            const uint8_t* stash_enc, float* scratch, void* extra, const mcnerf_mlp_grads* g,
                            int sms, cudaStream_t st);
constexpr int WG_MAX_CTAS = 160;

extern "C" size_t mcnerf_mlp_tc_bwd_workspace(const mcnerf_mlp_params* p, int n_rows) {
  if (!mcnerf_mlp_tc_supported(p) || n_rows <= 0) return 0;
  return stash_tiles(n_rows) * ((size_t)(p->depth + 2) * ACT_BYTES + HEAD_BYTES) + mlp_tc_wgrad_scratch_bytes(WG_MAX_CTAS) +
         mlp_tc_fused_extra_bytes(n_rows, p->depth + 2) + stash_tiles(n_rows) * (size_t)TM * 9 * sizeof(float);
}

// Measurement aid: which phases mcnerf_mlp_tc_bwd launches (bit 0: data-gradient chain, bit 1: weight gradients).
static int g_bwd_phases = 3;
extern "C" int mcnerf_mlp_tc_bwd_phases(int mask) {
  MC_ARG(mask >= 1 && mask <= 3);
  g_bwd_phases = mask;
  return 0;
}

extern "C" int mcnerf_mlp_tc_bwd(const mcnerf_mlp_params* p, const void* wb, const float* bias,
                                 const mcnerf_tc_input* in, const float* out4, const float* g_out4, const void* stash,
                                 void* workspace, const mcnerf_mlp_grads* g, float* g_rays_o, float* g_rays_d,
                                 float* g_x_enc, float* g_dirs_rows, void* stream) {
  PackLayout L;
  if (int e = build_layout(p, &L)) return e;
  MC_ARG(in && in->n_rows >= 0);
  if (in->n_rows == 0) return 0;
  MC_ARG(wb && bias && out4 && g_out4 && stash && workspace && g);
  const bool explicit_mode = in->x_enc != nullptr;
  MC_ARG(explicit_mode ? (in->dirs_rows && g_x_enc && g_dirs_rows && in->ld_enc >= 63)
                       : (in->rays_o && in->rays_d && g_rays_o && g_rays_d && in->smp.n_freqs == 10));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stash_tiles(in->n_rows);
  const int n_slots = p->depth + 2;
  BwdArgs a;
  a.plan = L.bwd;
  a.wb = (const uint8_t*)wb;
  a.bias = bias;
  a.sig2_off = L.sig2_off;
  a.n_slots = n_slots;
  const uint8_t* stash_act = (const uint8_t*)stash;
  const uint8_t* stash_enc = stash_act + tiles * (size_t)n_slots * ACT_BYTES;
  a.stash_sh = (const float*)(stash_enc + tiles * ENC_BYTES);
  a.stash_bits = (const uint8_t*)(a.stash_sh + tiles * (size_t)TM * SH_LD);
  a.dy = (uint8_t*)workspace;
  a.dy_head = a.dy + tiles * (size_t)n_slots * ACT_BYTES;
  a.g_out4 = g_out4; a.out4 = out4;
  a.rays_o = in->rays_o; a.rays_d = in->rays_d; a.jitter = in->jitter; a.smp = in->smp;
  a.sel_idx = in->sample_idx; a.n_rows = in->n_rows; a.n_rows_dev = in->n_rows_dev;
  a.x_enc = in->x_enc; a.ld_enc = in->ld_enc; a.dirs_rows = in->dirs_rows;
  a.g_rays_o = g_rays_o; a.g_rays_d = g_rays_d; a.g_x_enc = g_x_enc; a.g_dirs_rows = g_dirs_rows;
  // ordered ray-gradient sums: the last region of the workspace holds one 9-float partial per row
  const bool ordered = !explicit_mode && in->ordered_ray_grads && (in->sample_idx == nullptr || in->ray_offsets != nullptr);
  a.seg_part = ordered ? (float*)((uint8_t*)workspace + tiles * ((size_t)n_slots * ACT_BYTES + HEAD_BYTES) +
                                  mlp_tc_wgrad_scratch_bytes(WG_MAX_CTAS) + mlp_tc_fused_extra_bytes(in->n_rows, n_slots))
                       : nullptr;
  // per device / context attribute: set on every call (cheap), not once per process
  MC_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
  const int n_tiles = (in->n_rows + TM - 1) / TM;
  const int n_pairs = (n_tiles + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Default: two back-to-back kernels (data-gradient chain, then weight gradients).  MCNERF_BWD_FUSED=1 selects the
  // experimental single launch in which chain CTA pairs hand every dY tile to weight-gradient CTA pairs through L2
  // (mlp_tc_bwd_fused.cu): same results up to summation order, measured 3-40 % SLOWER on B200.
  auto reduce_rays = [&]() -> int {
    if (!a.seg_part) return 0;
    ray_grad_reduce_k<<<cdiv(in->n_rays, 128), 128, 0, st>>>(a.seg_part, in->ray_offsets, in->smp.S, in->n_rays, in->n_rows,
                                                             in->n_rows_dev, g_rays_o, g_rays_d);
    MC_LAUNCHED();
    return 0;
  };
  const bool fused = getenv("MCNERF_BWD_FUSED") && atoi(getenv("MCNERF_BWD_FUSED")) == 1;
  if (fused && g_bwd_phases == 3) {
    MC_ARG(sms <= WG_MAX_CTAS);
    float* scratch = (float*)(a.dy_head + tiles * HEAD_BYTES);
    void* extra = (uint8_t*)scratch + mlp_tc_wgrad_scratch_bytes(WG_MAX_CTAS);
    if (int e = mlp_tc_bwd_fused_launch(p, L, a, stash_act, stash_enc, scratch, extra, g, sms, st)) return e;
    return reduce_rays();
  }
  if (g_bwd_phases & 1) {
    int grid = n_pairs < sms ? n_pairs : sms;
    grid = (grid + 1) & ~1;                                  // whole clusters of 2
    if (grid > sms) grid = sms & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(BWD_THREADS);
    cfg.dynamicSmemBytes = SMEM_BWD;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_bwd_k, a));
    MC_LAUNCHED();
    if (int e = reduce_rays()) return e;
  }
  if (!(g_bwd_phases & 2)) return 0;
  // weight gradients (tcgen05, reduction over all rows) and bias gradients (column sums of the dY stash)
  MC_ARG(sms <= WG_MAX_CTAS);
  float* scratch = (float*)(a.dy_head + tiles * HEAD_BYTES);
  float* bias_part = (float*)((uint8_t*)scratch + mlp_tc_wgrad_scratch_bytes(WG_MAX_CTAS));     // [CTA][256], in the extra region
  return mlp_tc_wgrad_launch(p, L, stash_act, stash_enc, a.dy, a.dy_head, scratch, bias_part, in->n_rows, in->n_rows_dev, g, st);
}

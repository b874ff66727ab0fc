// bf16 tcgen05 NeRF MLP, FUSED backward: the data-gradient chain (mlp_tc_chain.cuh) and the weight gradients in ONE
// persistent launch, spatially pipelined over the SMs.
//
//   chain CTA pairs (the first n_chain CTAs): walk their 128-row tiles from d(out4) down to the rays exactly as the
//       stand-alone chain kernel does, write every layer's dY tile (bf16 UMMA tile image) to the dY buffer and then
//       release-increment that tile's readiness counter;
//   weight-gradient CTA pairs (the rest): each pair OWNS the fp32 accumulators of a small group of layers in tensor
//       memory for the whole launch (a 256x256 dW is 256 TMEM columns in each CTA of a cta_group::2 pair, so a pair
//       holds two layers), waits for a tile's counter, fetches the dY tile - still resident in L2, it was written
//       microseconds ago by another SM - and the matching activation tile of the forward stash (HBM) with bulk
//       async copies, and accumulates dW += dY^T X over its share of the tiles (MN-major operands: no transposes).
//
// Why: run back to back, the chain kernel (tensor-bound) writes 5.2 KB of dY per MLP evaluation to HBM and the
// weight-gradient kernel (HBM-bound, 97 % of the copy peak) reads it back together with 5.5 KB of activations.  Here
// the dY hand-off goes through L2, the HBM-bound and the tensor-bound halves of the backward pass overlap, and the
// weight-gradient side reads each operand once per group (sigma.0 / sh.0 share their input tile).
//
// Progress: chain CTAs never wait for weight-gradient CTAs, so there is no cycle; the launch needs every CTA resident
// (grid <= what cudaOccupancyMaxActiveClusters reports, checked by the host).  All waits are bounded and trap.
// ref: autograd of model/net_block.py:67-78.
#include <math.h>
#include <stdlib.h>
#include "mlp_tc_chain.cuh"

namespace mlptc {

constexpr int WG_HALF = 64;                  // rows per stage (half a tile)
constexpr int WG_PL = WG_HALF * 16;          // 1 KB: one k-group plane of a half tile
constexpr int WGF_MAX_BLK = 6, WGF_MAX_OP = 4, WGF_MAX_BIAS = 3, WGF_MAX_GROUPS = 8, WGF_MAX_STAGES = 6;
constexpr int WGF_BIAS_WARPS = 16;

struct WBlkD {
  const uint8_t* base;       // plane 0, half 0, tile 0 of this operand's image
  size_t tile_stride;        // bytes between tiles
  uint32_t half_stride;      // bytes between the two 64-row halves of a tile image (= planes of the image x 1 KB)
  uint32_t plane_off[2];     // byte offset of the first plane THIS CTA (cluster rank 0 / 1) loads, inside a half
  uint32_t bytes;            // bytes per stage per CTA
  uint32_t smem_off;         // offset inside the stage buffer
  int flag_slot;             // readiness counter to wait for (-1: forward stash, always there)
  uint32_t flag_need;
  int discard;               // 1: these lines are dead after this read (hint L2 not to write them back)
};
struct WOpD {
  uint32_t a_off, b_off;     // stage-relative offsets of the A (M = 256 split over the pair) and B (N/2 per CTA) blocks
  int N;                     // MMA N (16 .. 256)
  int tmem_col;              // accumulator columns [tmem_col, tmem_col + N) in each CTA
};
struct WBiasD {
  uint32_t smem_off;         // operand block whose column sums over all rows are a bias gradient
  int n_planes;              // planes of that block in this CTA (8 features each)
};
struct WGroupD {
  int n_blk, n_op, n_bias, n_stage;
  uint32_t stage_bytes;      // per CTA
  int used_cols;             // accumulator columns in use
  int pair_begin, pair_end;  // weight-gradient pairs [begin, end) of this group (tile t belongs to pair begin + t % n)
  WBlkD blk[WGF_MAX_BLK];
  WOpD op[WGF_MAX_OP];
  WBiasD bias[WGF_MAX_BIAS];
};

struct FusedArgs {
  BwdArgs chain;
  int n_chain_ctas;
  uint32_t* flags;           // [tiles][n_slots + 1] readiness counters, zeroed before the launch
  int fstr;
  int n_groups;
  int do_discard;
  int pf_dist;               // stages of L2 prefetch distance for the forward-stash operands (0: off)
  float* scratch;            // [wgrad CTA][128 lanes][512 columns] fp32 partial accumulators
  float* bias_scratch;       // [wgrad CTA][WGF_MAX_BIAS][128] fp32 partial column sums
  long long* stats;          // debug (MCNERF_FUSED_STATS=1): per CTA [t0 ns, t1 ns, poll, empty wait, full wait, stages, group, -]
  WGroupD grp[WGF_MAX_GROUPS];
};

struct __align__(16) WFBars {
  uint64_t full[WGF_MAX_STAGES], empty[WGF_MAX_STAGES], acc_full;
  uint32_t tmem_base;
  int ready_tiles;           // tiles of this pair's sequence whose operands are complete (published by the scout warp)
};
constexpr int W_SCOUT = 18;
constexpr int WGF_RING_BYTES = SMEM_BWD - 512;       // the ring starts at offset 0; barriers live in the last 512 bytes
static_assert(sizeof(WFBars) <= 512, "barrier block");

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_cta_smem(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(tc::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_smem(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(tc::smem_u32(p)), "r"(v) : "memory");
}

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void wgrad_role(const FusedArgs& fa, uint8_t* smem) {
  WFBars* bars = reinterpret_cast<WFBars*>(smem + WGF_RING_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int wcta = (int)blockIdx.x - fa.n_chain_ctas, wpair = wcta >> 1;
  int g = 0;
  while (g + 1 < fa.n_groups && wpair >= fa.grp[g].pair_end) ++g;
  const WGroupD& G = fa.grp[g];
  const int part = wpair - G.pair_begin, nparts = G.pair_end - G.pair_begin;
  const int rows = fa.chain.n_rows_dev ? min(*fa.chain.n_rows_dev, fa.chain.n_rows) : fa.chain.n_rows;
  const int n_tiles = (rows + TM - 1) / TM;
  const int n_my = part < n_tiles ? (n_tiles - part + nparts - 1) / nparts : 0;      // tiles part, part + nparts, ...
  const int n_stages_total = 2 * n_my;
  const int NS = G.n_stage;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc::mbar_init(&bars->full[i], crank == 0 ? 2 : 1);           // own expect_tx arrive (+ the peer's relay)
      tc::mbar_init(&bars->empty[i], 1 + WGF_BIAS_WARPS);          // MMA commit + the column-sum warps
    }
    tc::mbar_init(&bars->acc_full, 1);
    bars->ready_tiles = 0;
    tc::mbar_init_fence();
  }
  if (warp == BW_MMA) tc::tmem_alloc2(&bars->tmem_base, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (fa.stats && tid == 0) { fa.stats[blockIdx.x * 8] = globaltimer_ns(); fa.stats[blockIdx.x * 8 + 5] = n_stages_total; fa.stats[blockIdx.x * 8 + 6] = g; }
  if (warp == BW_PROD) {
    // ---------------------------------------------------------------- producer: this CTA's planes of every block
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      long long t_poll = 0, t_empty = 0;
      int known = 0;
      const int pf_dist = fa.pf_dist;
      for (int i = 0; i < n_stages_total; ++i) {
        const int tile = part + (i >> 1) * nparts, half = i & 1;
        const long long tp0 = fa.stats ? clock64() : 0;
        if (half == 0 && (i >> 1) >= known) {
          // wait for the scout warp: it polls the readiness counters of 32 tiles ahead in parallel, so that the
          // global round trips never sit between two stages of this loop
          const long long t0 = clock64();
          while ((known = ld_acquire_cta_smem(&bars->ready_tiles)) <= (i >> 1)) {
            if (clock64() - t0 > 20000000000LL) {
              printf("mcnerf: fused backward: tile %d never became ready (block %d)\n", tile, blockIdx.x);
              __trap();
            }
          }
          // the tile images were written through the generic proxy by other SMs; the bulk copies below read them
          // through the async proxy
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        const long long tp1 = fa.stats ? clock64() : 0;
        tc::mbar_wait(&bars->empty[stage], par ^ 1);
        if (fa.stats) { t_poll += tp1 - tp0; t_empty += clock64() - tp1; }
        if (pf_dist > 0 && i + pf_dist < n_stages_total) {
          // forward-stash operands (always there) are pulled into L2 a few tiles ahead, so that the copies into shared
          // memory below see L2 latency instead of HBM latency: the ring holds only ~2 stages in flight per SM
          const int pt = part + ((i + pf_dist) >> 1) * nparts, ph = (i + pf_dist) & 1;
          for (int b = 0; b < G.n_blk; ++b) {
            const WBlkD& B = G.blk[b];
            if (B.flag_slot >= 0) continue;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(B.base + (size_t)pt * B.tile_stride +
                         (size_t)ph * B.half_stride + B.plane_off[crank]), "r"(B.bytes) : "memory");
          }
        }
        tc::mbar_arrive_expect_tx(&bars->full[stage], G.stage_bytes);
        uint8_t* sdst = smem + (size_t)stage * G.stage_bytes;
        for (int b = 0; b < G.n_blk; ++b) {
          const WBlkD& B = G.blk[b];
          tc::bulk_g2s(sdst + B.smem_off,
                       B.base + (size_t)tile * B.tile_stride + (size_t)half * B.half_stride + B.plane_off[crank], B.bytes,
                       &bars->full[stage]);
        }
        if (++stage == NS) { stage = 0; par ^= 1; }
      }
      if (fa.stats) { fa.stats[blockIdx.x * 8 + 2] = t_poll; fa.stats[blockIdx.x * 8 + 3] = t_empty; }
    }
  } else if (warp == BW_MMA) {
    if (lane == 0 && crank == 0) {
      // -------------------------------------------------------------- MMA issuer (leader): M = 256 over the pair
      if (n_stages_total > 0) {
        const uint32_t hi = tc::umma_desc_hi(WG_PL);               // MN-direction stride: one plane
        const uint32_t s_lo0 = tc::umma_desc_lo(tc::smem_u32(smem), 128);      // K-direction (rows) stride: 128 B
        const uint32_t full0 = tc::smem_u32(&bars->full[0]), empty0 = tc::smem_u32(&bars->empty[0]);
        uint32_t a_rel[WGF_MAX_OP], b_rel[WGF_MAX_OP], idesc[WGF_MAX_OP], dcol[WGF_MAX_OP];
#pragma unroll
        for (int o = 0; o < WGF_MAX_OP; ++o) {
          const WOpD& op = G.op[o < G.n_op ? o : 0];
          a_rel[o] = op.a_off >> 4; b_rel[o] = op.b_off >> 4;
          idesc[o] = tc::umma_idesc_bf16(2 * TM, op.N, 1, 1);
          dcol[o] = tmem + op.tmem_col;
        }
        const uint32_t stage_inc = G.stage_bytes >> 4;
        int stage = 0;
        uint32_t par = 0;
        long long t_full = 0;
        for (int i = 0; i < n_stages_total; ++i) {
          const long long tf0 = fa.stats ? clock64() : 0;
          tc::mbar_wait_addr(full0 + stage * 8, par);
          if (fa.stats) t_full += clock64() - tf0;
          tc::tcgen05_fence_after();
          const uint32_t s_lo = s_lo0 + stage * stage_inc;
#pragma unroll
          for (int o = 0; o < WGF_MAX_OP; ++o) {
            if (o < G.n_op) {
#pragma unroll
              for (int k16 = 0; k16 < WG_HALF / 16; ++k16)
                tc::umma2_bf16_w(dcol[o], s_lo + a_rel[o] + k16 * 16, hi, s_lo + b_rel[o] + k16 * 16, hi, idesc[o],
                                 (i | k16) != 0);
            }
          }
          tc::umma2_commit_multicast_addr(empty0 + stage * 8, (uint16_t)3);
          if (++stage == NS) { stage = 0; par ^= 1; }
        }
        tc::umma2_commit_multicast_addr(tc::smem_u32(&bars->acc_full), (uint16_t)3);
        if (fa.stats) fa.stats[blockIdx.x * 8 + 4] = t_full;
      }
    } else if (lane == 0) {
      // peer: tell the leader when this CTA's blocks of a stage have landed
      const uint32_t full0 = tc::smem_u32(&bars->full[0]);
      const uint32_t leader_full0 = tc::mapa(full0, 0);
      int stage = 0;
      uint32_t par = 0;
      for (int i = 0; i < n_stages_total; ++i) {
        tc::mbar_wait_addr(full0 + stage * 8, par);
        tc::mbar_arrive_remote(leader_full0 + stage * 8);
        if (++stage == NS) { stage = 0; par ^= 1; }
      }
    }
  } else if (warp == W_SCOUT) {
    // ---------------------------------------------------------------- scout: lane l watches the readiness counters of
    // tile (base + l) of this pair's sequence; the leading run of complete tiles is published to the producer
    for (int base = 0; base < n_my; base += 32) {
      const int k = base + lane, wn = min(32, n_my - base);
      const int tile = part + k * nparts;
      bool done = k >= n_my;
      int published = 0;
      const long long t0 = clock64();
      while (published < wn) {
        if (!done) {
          bool ok = true;
          for (int b = 0; b < G.n_blk; ++b)
            if (G.blk[b].flag_slot >= 0)
              ok &= ld_relaxed_u32(fa.flags + (size_t)tile * fa.fstr + G.blk[b].flag_slot) >= G.blk[b].flag_need;
          if (ok) {
            __threadfence();        // acquire: the writers released the counters after their data
            done = true;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, done);
        const int n = min(m == 0xffffffffu ? 32 : __ffs(~m) - 1, wn);
        if (n > published) {
          published = n;
          if (lane == 0) st_release_cta_smem(&bars->ready_tiles, base + n);
        }
        if (clock64() - t0 > 20000000000LL) {
          if (!done) printf("mcnerf: fused backward: tile %d never became ready (block %d)\n", tile, blockIdx.x);
          __trap();
        }
      }
    }
  } else if (warp < WGF_BIAS_WARPS) {
    // ---------------------------------------------------------------- column sums (bias gradients) while the MMAs
    // run: warp w owns plane w of every bias block (8 features), lane l rows l and l + 32 of the 64-row stage
    float acc[WGF_MAX_BIAS][8];
#pragma unroll
    for (int t = 0; t < WGF_MAX_BIAS; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    {
      int stage = 0;
      uint32_t par = 0;
      for (int i = 0; i < n_stages_total; ++i) {
        tc::mbar_wait(&bars->full[stage], par);
        const uint8_t* sA = smem + (size_t)stage * G.stage_bytes;
        if (fa.do_discard) {
          // optional: a dY tile is dead once its copy has landed here (nobody else reads it) - tell L2 that the
          // lines need not be written back to HBM.  The 16 warps share the 128-byte lines of every such block.
          const int tile = part + (i >> 1) * nparts, half = i & 1;
          for (int b = 0; b < G.n_blk; ++b) {
            const WBlkD& B = G.blk[b];
            if (!B.discard) continue;
            const uint8_t* src = B.base + (size_t)tile * B.tile_stride + (size_t)half * B.half_stride + B.plane_off[crank];
            for (uint32_t off = (uint32_t)(warp * 32 + lane) * 128u; off < B.bytes; off += WGF_BIAS_WARPS * 32 * 128u)
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(src + off) : "memory");
          }
        }
#pragma unroll
        for (int t = 0; t < WGF_MAX_BIAS; ++t) {
          if (t < G.n_bias && warp < G.bias[t].n_planes) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const uint4 v = *reinterpret_cast<const uint4*>(sA + G.bias[t].smem_off + warp * WG_PL + (lane + 32 * rr) * 16);
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[t][2 * e] += __uint_as_float(w[e] << 16);
                acc[t][2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->empty[stage]);
        if (++stage == NS) { stage = 0; par ^= 1; }
      }
    }
    float* bs = fa.bias_scratch + (size_t)wcta * WGF_MAX_BIAS * 128;
#pragma unroll
    for (int t = 0; t < WGF_MAX_BIAS; ++t) {
      if (t < G.n_bias && warp < G.bias[t].n_planes) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = warp_sum(acc[t][j]);          // fixed butterfly: deterministic
          if (lane == 0) bs[t * 128 + warp * 8 + j] = s;
        }
      }
    }
    // ---- accumulators -> per-CTA partials: lane quarter = warp % 4, column quarter = warp / 4
    float* out = fa.scratch + (size_t)wcta * 128 * 512;
    const int lq = warp & 3, cq = warp >> 2, m = lq * 32 + lane;
    if (n_stages_total > 0) {
      tc::mbar_wait(&bars->acc_full, 0);
      tc::tcgen05_fence_after();
    }
    for (int c0 = cq * 128; c0 < cq * 128 + 128 && c0 < G.used_cols; c0 += 32) {
      uint32_t v[32];
      if (n_stages_total > 0) {
        tc::tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      float* row = out + (size_t)m * 512 + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(row + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (fa.stats && tid == 0) fa.stats[blockIdx.x * 8 + 1] = globaltimer_ns();
  tc::cluster_sync();
  if (warp == BW_MMA) tc::tmem_dealloc2(tmem, 512);
}

__global__ void __launch_bounds__(BWD_THREADS, 1) mlp_tc_bwd_fused_k(const __grid_constant__ FusedArgs fa) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((int)blockIdx.x < fa.n_chain_ctas) {
    if (fa.stats && threadIdx.x == 0) { fa.stats[blockIdx.x * 8] = globaltimer_ns(); fa.stats[blockIdx.x * 8 + 6] = -1; }
    chain_role<true>(fa.chain, smem, fa.n_chain_ctas, fa.flags, fa.fstr);
    if (fa.stats && threadIdx.x == 0) fa.stats[blockIdx.x * 8 + 1] = globaltimer_ns();
  } else {
    wgrad_role(fa, smem);
  }
}

// ------------------------------------------------------------------------------------------------- reduction
// Sum the per-CTA partials of every accumulated product into the parameter gradients (+=), mapping accumulator
// coordinates back to the reference's [out,in] layout; bias gradients likewise from the column-sum partials.
// Fixed summation order: deterministic.
struct FRedOp {
  int pair_begin, pair_end;  // weight-gradient pairs holding partials of this product
  int tmem_col, N;           // accumulator columns; M is always 256 (rows 0-127 in CTA rank 0, 128-255 in rank 1)
  int kind;                  // 0: dW[m = out][n = in] ; 1: transposed dW[n = out][m = in] (sh.2) ; 2: sigma.2 (column 15)
                             // 3: bias from an A block (128 features per CTA) ; 4: sh.2 bias (head columns, 16 per CTA)
                             // 5: sigma.2 bias (head column 31 = rank 1, index 7)
  int bias_task;
  float* dst;
  int ld, col_off, n_valid;
};
struct FRedArgs {
  int n_ops;
  const float* scratch;
  const float* bias_scratch;
  FRedOp op[48];
};

__global__ void __launch_bounds__(256) fused_reduce_k(const __grid_constant__ FRedArgs a) {
  const FRedOp& op = a.op[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (op.kind >= 3) {
    const int n = op.kind == 3 ? 256 : (op.kind == 4 ? 32 : 1);
    if (idx >= n) return;
    int rank, i;
    if (op.kind == 3) { rank = idx >> 7; i = idx & 127; }
    else if (op.kind == 4) { rank = idx >> 4; i = idx & 15; }
    else { rank = 1; i = 7; }
    float s = 0.f;
    for (int p = op.pair_begin; p < op.pair_end; ++p)
      s += a.bias_scratch[((size_t)(2 * p + rank) * WGF_MAX_BIAS + op.bias_task) * 128 + i];
    if (op.kind == 4 && idx >= 27) return;
    op.dst[idx] += s;
    return;
  }
  const int N = op.N;
  if (idx >= 256 * N) return;
  const int m = idx / N, n = idx - m * N;
  float s = 0.f;
  for (int p = op.pair_begin; p < op.pair_end; ++p)
    s += a.scratch[((size_t)(2 * p + (m >> 7)) * 128 + (m & 127)) * 512 + op.tmem_col + n];
  if (op.kind == 0) {
    if (n < op.n_valid) op.dst[(size_t)m * op.ld + op.col_off + n] += s;      // m = out feature, n = in feature
  } else if (op.kind == 1) {
    if (n < 27) op.dst[(size_t)n * op.ld + m] += s;                           // m = in feature, n = out feature
  } else {
    if (n == 15) op.dst[m] += s;                                              // head column 31 = g_sigma
  }
}

}  // namespace mlptc

using namespace mlptc;

size_t mlp_tc_fused_extra_bytes(int n_rows, int n_slots) {
  return stash_tiles(n_rows) * (size_t)(n_slots + 1) * sizeof(uint32_t) + 256 +
         (size_t)160 * WGF_MAX_BIAS * 128 * sizeof(float);
}

// Host-side plan: groups of products that share a pair's tensor memory and ring, and the split of the SMs.
namespace {

struct HostOp { int a_blk, b_blk, N, kind, which, col_off, n_valid, bias_task; };
struct HostGroup {
  int n_blk = 0, n_op = 0, n_bias = 0, cols = 0;
  uint32_t bytes = 0;
  double mma_cycles = 0;     // per stage
  WBlkD blk[WGF_MAX_BLK];
  HostOp op[WGF_MAX_OP];
  int bias_blk[WGF_MAX_BIAS], bias_kind[WGF_MAX_BIAS], bias_which[WGF_MAX_BIAS];
};

}  // namespace

int mlp_tc_bwd_fused_launch(const mcnerf_mlp_params* p, const PackLayout& L, const BwdArgs& chain, const uint8_t* stash,
                            const uint8_t* stash_enc, float* scratch, void* extra, const mcnerf_mlp_grads* g,
                            int sms, cudaStream_t st) {
  const int D = p->depth, n_slots = D + 2;
  const size_t tiles = stash_tiles(chain.n_rows);
  const size_t tile_stride = (size_t)n_slots * ACT_BYTES;
  int skip = -1;
  for (int l = 1; l < D; ++l) if (p->skip_mask >> l & 1u) skip = l;
  // readiness arrivals per dY slot
  uint32_t need[MAX_STEPS + 2] = {0};
  for (int j = 0; j < L.bwd.n_jobs; ++j) {
    if (L.bwd.j[j].kind == BK_MASK_STORE) need[L.bwd.j[j].dy_slot] = READY_STASH;
    if (L.bwd.j[j].kind == BK_SIGMA_INJECT) need[L.bwd.j[j].dy_slot] = READY_SIGMA;
  }
  // operand block constructors: `rank_planes` = planes each CTA loads, CTA rank r starts at plane p0 + r * rank_planes
  auto blk_dy = [&](int slot) {          // A operand: this CTA's 128 of the 256 dY features
    WBlkD b{};
    b.base = chain.dy + (size_t)slot * ACT_BYTES; b.tile_stride = tile_stride; b.half_stride = 32 * WG_PL;
    b.plane_off[0] = 0; b.plane_off[1] = 16 * WG_PL; b.bytes = 16 * WG_PL;
    b.flag_slot = slot; b.flag_need = need[slot]; b.discard = slot != skip;      // the chain re-reads the skip layer's dY
    return b;
  };
  auto blk_act = [&](int slot, bool as_a, int N) {      // forward stash tile: A (128 features per CTA) or B (N/2 per CTA)
    WBlkD b{};
    b.base = stash + (size_t)slot * ACT_BYTES; b.tile_stride = tile_stride; b.half_stride = 32 * WG_PL;
    const int per = as_a ? 16 : N / 16;
    b.plane_off[0] = 0; b.plane_off[1] = per * WG_PL; b.bytes = per * WG_PL; b.flag_slot = -1;
    return b;
  };
  auto blk_enc = [&]() {                                // encoding tile as B, N = 64: 4 planes per CTA
    WBlkD b{};
    b.base = stash_enc; b.tile_stride = ENC_BYTES; b.half_stride = 8 * WG_PL;
    b.plane_off[0] = 0; b.plane_off[1] = 4 * WG_PL; b.bytes = 4 * WG_PL; b.flag_slot = -1;
    return b;
  };
  auto blk_head = [&](int col0, int N) {                // head-gradient tile as B: columns col0 .. col0 + N, N/2 per CTA
    WBlkD b{};
    b.base = chain.dy_head; b.tile_stride = HEAD_BYTES; b.half_stride = 4 * WG_PL;
    b.plane_off[0] = (col0 / 8) * WG_PL; b.plane_off[1] = (col0 / 8) * WG_PL + (N / 16) * WG_PL;
    b.bytes = (N / 16) * WG_PL; b.flag_slot = n_slots; b.flag_need = READY_HEAD; b.discard = 0;     // used by two products
    return b;
  };
  HostGroup groups[WGF_MAX_GROUPS];
  int ng = 0;
  auto add_blk = [&](HostGroup& G, WBlkD b) {
    b.smem_off = G.bytes;
    G.bytes += b.bytes;
    G.blk[G.n_blk] = b;
    return G.n_blk++;
  };
  auto add_op = [&](HostGroup& G, int a, int b, int N, int kind, int which, int col_off, int n_valid) {
    HostOp& o = G.op[G.n_op++];
    o.a_blk = a; o.b_blk = b; o.N = N; o.kind = kind; o.which = which; o.col_off = col_off; o.n_valid = n_valid;
    o.bias_task = -1;
    G.cols += N;
    G.mma_cycles += (WG_HALF / 16) * (N / 2.0);     // a pair MMA of M = 256, N, K = 16 takes N/2 cycles
    return &o;
  };
  auto add_bias = [&](HostGroup& G, int blk, int kind, int which) {
    G.bias_blk[G.n_bias] = blk; G.bias_kind[G.n_bias] = kind; G.bias_which[G.n_bias] = which;
    return G.n_bias++;
  };
  // 1. sigma.0 / sh.0: same input tile
  {
    HostGroup& G = groups[ng++];
    const int a0 = add_blk(G, blk_dy(D)), a1 = add_blk(G, blk_dy(D + 1)), b = add_blk(G, blk_act(D - 1, false, WID));
    add_op(G, a0, b, WID, 0, D, 0, WID);
    add_op(G, a1, b, WID, 0, D + 1, 0, WID);
    add_bias(G, a0, 3, D);
    add_bias(G, a1, 3, D + 1);
  }
  // 2. trunk layers 1 .. D-1 (hidden-input part), two per pair
  int bigs[MAX_STEPS], nb = 0;
  for (int l = 1; l < D; ++l) bigs[nb++] = l;
  int i = 0;
  for (; i + 1 < nb; i += 2) {
    HostGroup& G = groups[ng++];
    for (int k = 0; k < 2; ++k) {
      const int l = bigs[i + k];
      const int a = add_blk(G, blk_dy(l)), b = add_blk(G, blk_act(l - 1, false, WID));
      add_op(G, a, b, WID, 0, l, l == skip ? 63 : 0, WID);
      add_bias(G, a, 3, l);
    }
  }
  // 3. the odd trunk layer (if any) + the encoding-input products of layer 0 and of the skip layer
  {
    HostGroup& G = groups[ng++];
    if (i < nb) {
      const int l = bigs[i];
      const int a = add_blk(G, blk_dy(l)), b = add_blk(G, blk_act(l - 1, false, WID));
      add_op(G, a, b, WID, 0, l, l == skip ? 63 : 0, WID);
      add_bias(G, a, 3, l);
    }
    const int a0 = add_blk(G, blk_dy(0)), be = add_blk(G, blk_enc());
    add_op(G, a0, be, ENCW, 0, 0, 0, 63);
    add_bias(G, a0, 3, 0);
    if (skip >= 0) {
      // the skip layer's dY block may already be in this group (when the skip layer is the odd one)
      int as = -1;
      for (int b = 0; b < G.n_blk; ++b) if (G.blk[b].flag_slot == skip && G.blk[b].base == chain.dy + (size_t)skip * ACT_BYTES) as = b;
      if (as < 0) as = add_blk(G, blk_dy(skip));
      add_op(G, as, be, ENCW, 0, skip, 0, 63);
    }
  }
  // 4. sh.2 / sigma.2 (transposed products: A = activation^T, B = head-gradient tile)
  {
    HostGroup& G = groups[ng++];
    const int a0 = add_blk(G, blk_act(D + 1, true, 0)), b0 = add_blk(G, blk_head(0, 32));
    const int a1 = add_blk(G, blk_act(D, true, 0)), b1 = add_blk(G, blk_head(16, 16));
    add_op(G, a0, b0, 32, 1, D + 2, 0, WID);
    add_op(G, a1, b1, 16, 2, D + 3, 0, WID);
    add_bias(G, b0, 4, D + 2);
    add_bias(G, b1, 5, D + 3);
  }
  MC_ARG(ng <= WGF_MAX_GROUPS);
  for (int k = 0; k < ng; ++k) MC_ARG(groups[k].cols <= 512 && groups[k].bytes * 2 <= (uint32_t)WGF_RING_BYTES);

  // ---- split of the SMs.  Cost of a group per stage in cycles: its MMAs, or its bytes at what one SM sustains from
  // L2 / HBM (env MCNERF_FUSED_BPC bytes per cycle per SM); the chain's cost per tile comes from measurement
  // (MCNERF_FUSED_CHAIN_PAIRS overrides the split).
  int max_clusters = sms / 2;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1); cfg.blockDim = dim3(BWD_THREADS); cfg.dynamicSmemBytes = SMEM_BWD;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, mlp_tc_bwd_fused_k, &cfg) == cudaSuccess && nc > 0 && nc < max_clusters)
      max_clusters = nc;
    cudaGetLastError();
  }
  const double bpc = getenv("MCNERF_FUSED_BPC") ? atof(getenv("MCNERF_FUSED_BPC")) : 48.0;
  double cost[WGF_MAX_GROUPS], cost_sum = 0;
  for (int k = 0; k < ng; ++k) { cost[k] = fmax(groups[k].mma_cycles, groups[k].bytes / bpc); cost_sum += cost[k]; }
  const int force_chain = getenv("MCNERF_FUSED_CHAIN_PAIRS") ? atoi(getenv("MCNERF_FUSED_CHAIN_PAIRS")) : 0;
  const double chain_cyc = getenv("MCNERF_FUSED_CHAIN_CYC") ? atof(getenv("MCNERF_FUSED_CHAIN_CYC")) : 51000.0;
  int best_pc = 0, best_cnt[WGF_MAX_GROUPS] = {0};
  double best_t = 1e30;
  for (int pc = (force_chain ? force_chain : ng); pc <= (force_chain ? force_chain : max_clusters - ng); ++pc) {
    const int wp = max_clusters - pc;
    if (wp < ng || pc < 1) continue;
    int cnt[WGF_MAX_GROUPS];
    for (int k = 0; k < ng; ++k) cnt[k] = 1;
    for (int used = ng; used < wp; ++used) {
      int b = 0;
      for (int k = 1; k < ng; ++k) if (cost[k] / cnt[k] > cost[b] / cnt[b]) b = k;
      ++cnt[b];
    }
    double tw = 0;
    for (int k = 0; k < ng; ++k) tw = fmax(tw, 2.0 * cost[k] / cnt[k]);       // cycles per tile (two stages)
    const double tchain = chain_cyc / pc;                                     // a chain pair = 2 SMs
    // the weight-gradient side trails the chain by up to one chain iteration: weigh it slightly higher
    const double t = fmax(tchain, 1.05 * tw);
    if (t < best_t) { best_t = t; best_pc = pc; for (int k = 0; k < ng; ++k) best_cnt[k] = cnt[k]; }
  }
  const int exp_mode_early = getenv("MCNERF_FUSED_MODE") ? atoi(getenv("MCNERF_FUSED_MODE")) : 0;
  if (exp_mode_early == 1) best_pc = force_chain ? force_chain : max_clusters;
  MC_ARG(best_pc > 0 || exp_mode_early == 2);
  // measurement modes: 1 = chain pairs only (with readiness signalling, no consumers); 2 = weight-gradient pairs only
  // (every pair of the grid, dY read from the buffer a previous launch left: timing only)
  const int exp_mode = getenv("MCNERF_FUSED_MODE") ? atoi(getenv("MCNERF_FUSED_MODE")) : 0;
  if (exp_mode == 2) {
    best_pc = 0;
    const int wp = max_clusters;
    for (int k = 0; k < ng; ++k) best_cnt[k] = 1;
    for (int used = ng; used < wp; ++used) {
      int b = 0;
      for (int k = 1; k < ng; ++k) if (cost[k] / best_cnt[k] > cost[b] / best_cnt[b]) b = k;
      ++best_cnt[b];
    }
    for (int k = 0; k < ng; ++k) for (int b = 0; b < groups[k].n_blk; ++b) groups[k].blk[b].flag_need = 0;
  }
  if (exp_mode == 1) for (int k = 0; k < ng; ++k) best_cnt[k] = 0;

  FusedArgs fa;
  fa.chain = chain;
  fa.n_chain_ctas = 2 * best_pc;
  fa.flags = (uint32_t*)extra;
  fa.fstr = n_slots + 1;
  fa.n_groups = ng;
  const int discard = getenv("MCNERF_FUSED_DISCARD") ? atoi(getenv("MCNERF_FUSED_DISCARD")) : 0;
  fa.do_discard = discard;
  fa.pf_dist = getenv("MCNERF_FUSED_PF") ? atoi(getenv("MCNERF_FUSED_PF")) : 0;
  fa.scratch = scratch;
  const size_t flag_bytes = (tiles * (size_t)(n_slots + 1) * sizeof(uint32_t) + 255) & ~(size_t)255;
  fa.bias_scratch = (float*)((uint8_t*)extra + flag_bytes);
  FRedArgs ra;
  ra.n_ops = 0;
  ra.scratch = scratch;
  ra.bias_scratch = fa.bias_scratch;
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
  auto bptr = [&](int which) -> float* {
    if (which < D) return g->b[which];
    if (which == D) return g->b_sigma0;
    if (which == D + 1) return g->b_sh0;
    if (which == D + 2) return g->b_sh2;
    return g->b_sigma2;
  };
  int pair0 = 0;
  for (int k = 0; k < ng; ++k) {
    const HostGroup& H = groups[k];
    WGroupD& G = fa.grp[k];
    G.n_blk = H.n_blk; G.n_op = H.n_op; G.n_bias = H.n_bias;
    G.stage_bytes = H.bytes;
    int ns = WGF_RING_BYTES / (int)H.bytes;
    G.n_stage = ns > WGF_MAX_STAGES ? WGF_MAX_STAGES : ns;
    G.used_cols = (H.cols + 31) & ~31;
    G.pair_begin = pair0; G.pair_end = pair0 + best_cnt[k];
    for (int b = 0; b < H.n_blk; ++b) G.blk[b] = H.blk[b];
    int col = 0;
    for (int o = 0; o < H.n_op; ++o) {
      G.op[o].a_off = H.blk[H.op[o].a_blk].smem_off; G.op[o].b_off = H.blk[H.op[o].b_blk].smem_off;
      G.op[o].N = H.op[o].N; G.op[o].tmem_col = col;
      FRedOp& r = ra.op[ra.n_ops++];
      r.pair_begin = G.pair_begin; r.pair_end = G.pair_end; r.tmem_col = col; r.N = H.op[o].N; r.kind = H.op[o].kind;
      r.bias_task = -1; r.dst = wptr(H.op[o].which, &r.ld); r.col_off = H.op[o].col_off; r.n_valid = H.op[o].n_valid;
      col += H.op[o].N;
    }
    for (int t = 0; t < H.n_bias; ++t) {
      G.bias[t].smem_off = H.blk[H.bias_blk[t]].smem_off;
      G.bias[t].n_planes = (int)(H.blk[H.bias_blk[t]].bytes / WG_PL);
      FRedOp& r = ra.op[ra.n_ops++];
      r.pair_begin = G.pair_begin; r.pair_end = G.pair_end; r.tmem_col = 0; r.N = 0; r.kind = H.bias_kind[t];
      r.bias_task = t; r.dst = bptr(H.bias_which[t]); r.ld = 0; r.col_off = 0; r.n_valid = 0;
    }
    pair0 += best_cnt[k];
  }
  MC_ARG(ra.n_ops <= 48 && 2 * (best_pc + pair0) <= 160);

  fa.stats = nullptr;
  static long long* stats_buf = nullptr;
  if (getenv("MCNERF_FUSED_STATS") && atoi(getenv("MCNERF_FUSED_STATS"))) {
    if (!stats_buf) cudaMalloc(&stats_buf, 160 * 8 * sizeof(long long));
    cudaMemsetAsync(stats_buf, 0, 160 * 8 * sizeof(long long), st);
    fa.stats = stats_buf;
  }
  MC_CUDA(cudaMemsetAsync(fa.flags, 0, flag_bytes, st));
  MC_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_fused_k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (best_pc + pair0));
  cfg.blockDim = dim3(BWD_THREADS);
  cfg.dynamicSmemBytes = SMEM_BWD;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_bwd_fused_k, fa));
  MC_LAUNCHED();
  if (exp_mode != 1) {
    fused_reduce_k<<<dim3(256, ra.n_ops), 256, 0, st>>>(ra);
    MC_LAUNCHED();
  }
  if (fa.stats) {        // debug: per-CTA timeline (ns relative to the first start; wait times in SM cycles)
    cudaStreamSynchronize(st);
    static long long h[160 * 8];
    cudaMemcpy(h, fa.stats, sizeof(h), cudaMemcpyDeviceToHost);
    const int n = 2 * (best_pc + pair0);
    long long t0 = h[0];
    for (int c = 0; c < n; ++c) if (h[c * 8] < t0) t0 = h[c * 8];
    long long chain_end = 0;
    for (int c = 0; c < fa.n_chain_ctas; ++c) if (h[c * 8 + 1] - t0 > chain_end) chain_end = h[c * 8 + 1] - t0;
    printf("fused stats: %d chain CTAs end at %lld ns (max)\n", fa.n_chain_ctas, chain_end);
    for (int c = fa.n_chain_ctas; c < n; c += 2)
      printf("  wgrad pair %2d group %lld: start %6lld end %7lld ns  stages %4lld  producer poll %9lld empty-wait %9lld  mma full-wait %9lld cycles\n",
             (c - fa.n_chain_ctas) / 2, h[c * 8 + 6], h[c * 8] - t0, h[c * 8 + 1] - t0, h[c * 8 + 5], h[c * 8 + 2], h[c * 8 + 3], h[c * 8 + 4]);
  }
  return 0;
}

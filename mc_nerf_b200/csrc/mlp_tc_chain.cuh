// bf16 tcgen05 NeRF MLP, backward chain (data gradients) as a device function shared by the stand-alone kernel
// (mlp_tc_bwd.cu) and the fused chain + weight-gradient kernel (mlp_tc_bwd_fused.cu).  See mlp_tc_bwd.cu for the design.
#pragma once
#include "mlp_tc.cuh"

namespace mlptc {

struct BwdArgs {
  BPlan plan;
  const uint8_t* wb;        // transposed weight images
  const float* bias;        // bias block (w_sigma2 at sig2_off)
  int sig2_off;
  int n_slots;              // D + 2
  const uint8_t* stash_bits;  // forward ReLU gate bits [tile][n_slots][BITS_BYTES]
  const float* stash_sh;    // [tile][SH_LD/4 float4 groups][128 rows] float4
  uint8_t* dy;              // [tile][n_slots][ACT_BYTES]
  uint8_t* dy_head;         // [tile][HEAD_BYTES]
  const float* g_out4;      // [rows,4]
  const float* out4;        // [rows,4]
  const float *rays_o, *rays_d, *jitter;
  mcnerf_sampling smp;
  const int32_t* sel_idx;
  int n_rows;
  const int32_t* n_rows_dev;
  const float* x_enc;       // explicit mode (module API)
  int ld_enc;
  const float* dirs_rows;
  float *g_rays_o, *g_rays_d;     // rays mode: accumulated [n_rays,3]
  float* seg_part;                // rays mode, ordered sums: [rows][9] partial (d o, d d, d d via the SH head) of the
                                  // 32-row segment that STARTS at that row; null: fp32 atomics into g_rays_*
  float* g_x_enc;                 // explicit mode: [rows, ld_enc] overwritten
  float* g_dirs_rows;             // explicit mode: [rows,3] overwritten
};

// CTA pair (tcgen05 cta_group::2) like the forward kernel: M = 256 rows per MMA (slot t of both CTAs), each CTA
// stages half (N/2 rows) of every transposed weight chunk, the peer relays "my half landed" to the leader.
constexpr int BSTAGE = 5;
struct __align__(16) BwdBars {
  uint64_t w_full[BSTAGE], w_empty[BSTAGE], a_ready[2], acc_full[2];
  uint64_t st_full[2], st_done[2];      // dY tile of slot t is complete in shared memory / has been copied to the dY stash
  uint32_t tmem_base;
};
constexpr int SMEM_BWD = 2 * ACT_BYTES + 2 * HEAD_BYTES + BSTAGE * STAGE_BYTES + 1024 + 256;
// 20 warps: 16 epilogue warps (TMEM lane quarter = warp % 4, accumulator column quarter = warp / 4) serving both tile
// slots in turn, 1 weight producer, 1 MMA issuer (leader) / relay (peer), 2 stash warps that copy every finished dY
// tile shared memory -> HBM while the next job's MMAs read it (as in the forward kernel).
constexpr int BWD_THREADS = 640;
constexpr int BW_PROD = 16, BW_MMA = 17, BW_STASH0 = 18;

__constant__ float bC0 = 0.28209479177387814f;
__constant__ float bC1 = 0.4886025119029199f;
__constant__ float bC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};

// release-increment of a readiness counter: everything this warp wrote before (made visible to lane 0 by the preceding
// __syncwarp) is visible to whoever observes the new value with an acquire load
__device__ __forceinline__ void signal_ready(uint32_t* flag) {
  __threadfence();
  atomicAdd(flag, 1u);
}
// arrivals per tile image (the consumer waits for this many): every image is signalled by the two stash warps
constexpr int READY_STASH = 2, READY_SIGMA = 2, READY_HEAD = 2;

__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// sum `v` over lanes that share `key` (keys sorted within the warp); true on the first lane of each run
template <int NV>
__device__ __forceinline__ bool seg_reduce(int key, float (&v)[NV], int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int k2 = __shfl_down_sync(0xffffffffu, key, o);
    bool take = (lane + o < 32) && (k2 == key);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float x2 = __shfl_down_sync(0xffffffffu, v[i], o);
      if (take) v[i] += x2;
    }
  }
  int kp = __shfl_up_sync(0xffffffffu, key, 1);
  return lane == 0 || kp != key;
}

__device__ __forceinline__ uint32_t relu_gate2(uint32_t bits, int pos, float lo, float hi) {
  // bits: forward ReLU gate word of a 32-column block; element e = 2k+h (k = pair index, h = 0 lo / 1 hi) is at
  // bit k + 16h (tc::gate_bits).  `pos` = index of the lo element within the block (even).
  // ((bits >> k) & 0x10001) * 0xFFFF expands the two gate bits into a bf16x2 AND-mask (no carries between halves).
  const uint32_t mask = ((bits >> (pos >> 1)) & 0x10001u) * 0xFFFFu;
  return tc::pack_bf16(lo, hi) & mask;
}

// chunk geometry of a job: reduction elements per ring stage (a K=32 job is a single half-size chunk)
__device__ __forceinline__ int job_kc(int n_chunks) { return n_chunks * KC >= KC2 ? KC2 : n_chunks * KC; }

// The whole backward chain of one CTA (see the file header of mlp_tc_bwd.cu).  n_ctas: CTAs running this role (the
// first n_ctas blocks of the grid; whole clusters).  FUSED: the weight-gradient CTA pairs of the same launch
// (mlp_tc_bwd_fused.cu) consume every dY / head tile as soon as it is complete: after a tile image has been written,
// its readiness counter flags[tile * fstr + slot] is incremented (slot n_slots = the head tile) with release
// semantics by each warp that wrote a part of it.
template <bool FUSED>
__device__ __forceinline__ void chain_role(const BwdArgs& a, uint8_t* smem, const int n_ctas, uint32_t* flags,
                                           const int fstr) {
  uint8_t* bufX = smem;                                    // [2][ACT_BYTES]  dY tile (A operand)
  uint8_t* small = smem + 2 * ACT_BYTES;                   // [2][HEAD_BYTES] head-gradient tile
  uint8_t* wst = small + 2 * HEAD_BYTES;                   // [BSTAGE][STAGE_BYTES]
  float* w2s = reinterpret_cast<float*>(wst + BSTAGE * STAGE_BYTES);     // w_sigma2 [256] (no L1 left: keep it on chip)
  BwdBars* bars = reinterpret_cast<BwdBars*>(w2s + 256);
  static_assert(sizeof(BwdBars) <= 192, "BwdBars must leave room for the band weights");
  float* bw_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 192);      // 10 BARF band weights
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int n_tiles = (rows + TM - 1) / TM;
  const int n_pairs = (n_tiles + 1) / 2;
  const int n_jobs = a.plan.n_jobs;
  const uint32_t crank = tc::cluster_ctarank();
  const int n_iter = (n_pairs + n_ctas - 1) / n_ctas;     // same trip count for both CTAs of a pair

  if (tid == 0) {
    for (int i = 0; i < BSTAGE; ++i) { tc::mbar_init(&bars->w_full[i], crank == 0 ? 2 : 1); tc::mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->a_ready[i], 32); /* epilogue warps of both CTAs */ tc::mbar_init(&bars->acc_full[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->st_full[i], 16); tc::mbar_init(&bars->st_done[i], 2); }
    tc::mbar_init_fence();
  }
  if (warp == BW_MMA) tc::tmem_alloc2(&bars->tmem_base, 512);
  if (tid < 256) w2s[tid] = a.bias[a.sig2_off + tid];
  if (tid < 10) bw_s[tid] = a.smp.band_w_dev ? a.smp.band_w_dev[tid] : a.smp.band_w[tid];   // see mlp_tc_fwd.cu
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == BW_PROD) {
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      for (int it = 0; it < n_iter; ++it)
        for (int jn = 0; jn < n_jobs; ++jn) {
          const BJob& jb = a.plan.j[jn];
          const int kc = job_kc(jb.n_chunks), nc2 = jb.n_chunks * KC / kc;
          const uint32_t half = (uint32_t)(jb.N / 2) * kc * 2;
          for (int t = 0; t < 2; ++t)
            for (int c = 0; c < nc2; ++c) {
              tc::mbar_wait(&bars->w_empty[stage], par ^ 1);
              tc::mbar_arrive_expect_tx(&bars->w_full[stage], half);
              tc::bulk_g2s(wst + stage * STAGE_BYTES, a.wb + jb.w_off + (size_t)(c * 2 + crank) * half, half,
                           &bars->w_full[stage]);
              if (++stage == BSTAGE) { stage = 0; par ^= 1; }
            }
        }
    }
  } else if (warp == BW_MMA) {
    // MMA issuer (leader): per-job constants hoisted, descriptors advance by 32-bit adds (the single issuing thread
    // has to stay below the 128 cycles one 128x256x16 MMA takes per SM, or the tensor pipe idles).
    if (lane == 0 && crank == 0) {
      int stage = 0;
      uint32_t par = 0, apar = 0;
      const uint32_t hi = tc::umma_desc_hi(128);
      const uint32_t bufX_lo = tc::umma_desc_lo(tc::smem_u32(bufX), PLANE), small_lo = tc::umma_desc_lo(tc::smem_u32(small), PLANE);
      const uint32_t wst_addr = tc::smem_u32(wst);
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]), empty0 = tc::smem_u32(&bars->w_empty[0]);
      const uint32_t ardy0 = tc::smem_u32(&bars->a_ready[0]), accf0 = tc::smem_u32(&bars->acc_full[0]);
      for (int it = 0; it < n_iter; ++it)
        for (int jn = 0; jn < n_jobs; ++jn) {
          const int N = a.plan.j[jn].N, kc = job_kc(a.plan.j[jn].n_chunks), nc2 = a.plan.j[jn].n_chunks * KC / kc;
          const bool a_small = a.plan.j[jn].a_small != 0, acc0 = a.plan.j[jn].accumulate != 0;
          const uint32_t idesc = tc::umma_idesc_bf16(2 * TM, N);
          const uint32_t b_lo0 = tc::umma_desc_lo(wst_addr, (N / 2) * 16), b_inc = (2u * (N / 2) * 16) >> 4;
          const int n16 = kc / 16;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            tc::mbar_wait_addr(ardy0 + t * 8, (apar >> t) & 1);
            apar ^= 1u << t;
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem + t * 256;
            uint32_t a_lo = a_small ? small_lo + t * (HEAD_BYTES >> 4) : bufX_lo + t * (ACT_BYTES >> 4);
            for (int c = 0; c < nc2; ++c) {
              tc::mbar_wait_addr(full0 + stage * 8, par);
              tc::tcgen05_fence_after();
              const uint32_t b_lo = b_lo0 + stage * (STAGE_BYTES >> 4);
              tc::umma2_bf16_w(d_tmem, a_lo, hi, b_lo, hi, idesc, acc0 || c != 0);
              tc::umma2_bf16_w(d_tmem, a_lo + ((2 * PLANE) >> 4), hi, b_lo + b_inc, hi, idesc, true);
              if (n16 == 4) {
                tc::umma2_bf16_w(d_tmem, a_lo + 2 * ((2 * PLANE) >> 4), hi, b_lo + 2 * b_inc, hi, idesc, true);
                tc::umma2_bf16_w(d_tmem, a_lo + 3 * ((2 * PLANE) >> 4), hi, b_lo + 3 * b_inc, hi, idesc, true);
              }
              tc::umma2_commit_multicast_addr(empty0 + stage * 8, (uint16_t)3);
              a_lo += ((KC2 / 8) * PLANE) >> 4;
              if (++stage == BSTAGE) { stage = 0; par ^= 1; }
            }
            tc::umma2_commit_multicast_addr(accf0 + t * 8, (uint16_t)3);
          }
        }
    } else if (lane == 0) {
      // peer CTA: tell the leader when this CTA's half of each stage has landed
      int stage = 0;
      uint32_t par = 0;
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]);
      const uint32_t leader_full0 = tc::mapa(full0, 0);
      for (int it = 0; it < n_iter; ++it)
        for (int jn = 0; jn < n_jobs; ++jn) {
          const int n = 2 * (a.plan.j[jn].n_chunks * KC / job_kc(a.plan.j[jn].n_chunks));
          for (int c = 0; c < n; ++c) {
            tc::mbar_wait_addr(full0 + stage * 8, par);
            tc::mbar_arrive_remote(leader_full0 + stage * 8);
            if (++stage == BSTAGE) { stage = 0; par ^= 1; }
          }
        }
    }
  } else if (warp >= BW_STASH0) {
    // ---- stash warps: dY tile shared memory -> HBM.  warp-item = (plane p, 32-row block): 512 contiguous bytes on
    // both sides; the two warps split the 128 items of a tile.
    const int sw = warp - BW_STASH0;
    uint32_t spar[2] = {0, 0};
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * n_ctas;
      for (int jn = 0; jn < n_jobs; ++jn) {
        const BJob& jb = a.plan.j[jn];
        if (jb.kind != BK_MASK_STORE) continue;
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t;
          tc::mbar_wait(&bars->st_full[t], spar[t]);
          spar[t] ^= 1;
          if (tile < n_tiles) {
            uint8_t* dst = a.dy + ((size_t)tile * a.n_slots + jb.dy_slot) * ACT_BYTES;
            const uint32_t src = tc::smem_u32(bufX + t * ACT_BYTES);
#pragma unroll 1
            for (int i0 = sw * 64; i0 < sw * 64 + 64; i0 += 8) {
              uint4 v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int item = i0 + u, p = item >> 2, row = (item & 3) * 32 + lane;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                             : "r"(src + p * PLANE + row * 16));
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int item = i0 + u, p = item >> 2, row = (item & 3) * 32 + lane;
                *reinterpret_cast<uint4*>(dst + stash_off(row, p, 32)) = v[u];
              }
            }
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->st_done[t]);
        }
        if (FUSED) {
          // Readiness of what is now complete in HBM, for the weight-gradient pairs of the same launch: ONE release
          // fence per job for both slots' tiles.  Besides this job's dY tiles that covers tiles the EPILOGUE warps wrote
          // directly - the head-gradient tile (prologue) with the first job, the sigma.0 dY tile with the job after
          // BK_SIGMA_INJECT: all 16 epilogue warps arrived on st_full after those stores, this warp waited on it, so
          // the fence is cumulative over them.  Each of the two stash warps signals; consumers wait for READY_STASH.
          __syncwarp();
          if (lane == 0) {
            __threadfence();
            const bool first = jn == 0, after_sigma = jn > 0 && a.plan.j[jn - 1].kind == BK_SIGMA_INJECT;
            for (int t = 0; t < 2; ++t) {
              const int tile = 2 * pair + t;
              if (tile >= n_tiles) continue;
              uint32_t* f = flags + (size_t)tile * fstr;
              atomicAdd(f + jb.dy_slot, 1u);
              if (first) atomicAdd(f + a.n_slots, 1u);
              if (after_sigma) atomicAdd(f + a.plan.j[jn - 1].dy_slot, 1u);
            }
          }
        }
      }
    }
  } else {
    const int lq = warp & 3, cq = warp >> 2;           // TMEM lane quarter, accumulator column quarter
    const int q = lq * 32 + lane;                      // row in tile == TMEM lane
    const uint32_t bufX0 = tc::smem_u32(bufX), small0 = tc::smem_u32(small);
    const uint32_t a_ready_leader = tc::mapa(tc::smem_u32(&bars->a_ready[0]), 0);
    const float* w2 = w2s;
    uint32_t par = 0, stpar = 0, st_pending = 0;       // stash handshake: parity / "a copy of slot t's tile is in flight"
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * n_ctas;
      // per-slot row state of this thread (row q of slot t): needed by the column-split epilogues below
      int ray[2];
      float z[2], g_sigma[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int row_g = (2 * pair + t) * TM + q;
        ray[t] = -1 - lane; z[t] = 0.f; g_sigma[t] = 0.f;
        if (row_g < rows) {
          if (!a.x_enc) {
            int flat = a.sel_idx ? a.sel_idx[row_g] : row_g;
            ray[t] = flat / a.smp.S;
            z[t] = linspace_f(a.smp.near_, a.smp.far_, a.smp.S, flat - ray[t] * a.smp.S) + (a.jitter ? a.jitter[ray[t]] : 0.f);
          }
          g_sigma[t] = a.g_out4[4 * (size_t)row_g];
        }
      }
      // ---------------- head backward: eval_sh + sigmoid, builds the [128 x 32] head-gradient tile.
      // warps 8-11 own the rows of slot 0, warps 12-15 those of slot 1: column quarters 2 and 3 have nothing to do in
      // the previous iteration's last epilogue (the encoding backward is done by column quarter 0 alone), so this
      // prologue of the next tiles overlaps it instead of following it (the head tile buffer was last read by the
      // first job's MMAs).
      if (cq >= 2) {
        const int t = cq - 2, tile = 2 * pair + t, row_g = tile * TM + q;
        const bool valid = row_g < rows, tile_ok = tile < n_tiles;
        float hv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) hv[i] = 0.f;
        float gd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const float4 g = reinterpret_cast<const float4*>(a.g_out4)[row_g];
          const float4 o = reinterpret_cast<const float4*>(a.out4)[row_g];
          const float* dp = a.x_enc ? a.dirs_rows + (size_t)row_g * 3 : a.rays_d + 3 * (size_t)ray[t];
          const float x = dp[0], y = dp[1], zz = dp[2];
          const float Y[9] = {bC0, -bC1 * y, bC1 * zz, -bC1 * x, bC2[0] * x * y, bC2[1] * y * zz,
                              bC2[2] * (2.f * zz * zz - x * x - y * y), bC2[3] * x * zz, bC2[4] * (x * x - y * y)};
          const float gc[3] = {g.y * o.y * (1.f - o.y), g.z * o.z * (1.f - o.z), g.w * o.w * (1.f - o.w)};
          const float4* shp = reinterpret_cast<const float4*>(a.stash_sh) + (size_t)(row_g >> 7) * (TM * SH_LD / 4) + (row_g & (TM - 1));
          float sh[28];
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            float4 v = shp[i * TM];            // [tile][float4 group][row] (mlp_tc_fwd.cu)
            sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
          }
          float comb[9];
#pragma unroll
          for (int b = 0; b < 9; ++b) comb[b] = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int b = 0; b < 9; ++b) {
              hv[9 * ch + b] = gc[ch] * Y[b];
              comb[b] += gc[ch] * sh[9 * ch + b];
            }
          gd[0] = -bC1 * comb[3] + bC2[0] * y * comb[4] - 2.f * bC2[2] * x * comb[6] + bC2[3] * zz * comb[7] +
                  2.f * bC2[4] * x * comb[8];
          gd[1] = -bC1 * comb[1] + bC2[0] * x * comb[4] + bC2[1] * zz * comb[5] - 2.f * bC2[2] * y * comb[6] -
                  2.f * bC2[4] * y * comb[8];
          gd[2] = bC1 * comb[2] + bC2[1] * y * comb[5] + 4.f * bC2[2] * zz * comb[6] + bC2[3] * x * comb[7];
          hv[31] = g.x;
        }
        const uint32_t small_t = small0 + t * HEAD_BYTES;
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          uint4 v = make_uint4(tc::pack_bf16(hv[kg * 8], hv[kg * 8 + 1]), tc::pack_bf16(hv[kg * 8 + 2], hv[kg * 8 + 3]),
                               tc::pack_bf16(hv[kg * 8 + 4], hv[kg * 8 + 5]), tc::pack_bf16(hv[kg * 8 + 6], hv[kg * 8 + 7]));
          sts_v4(small_t + kg * PLANE + q * 16, v);
          if (tile_ok) *reinterpret_cast<uint4*>(a.dy_head + (size_t)tile * HEAD_BYTES + stash_off(q, kg, 4)) = v;
        }
        if (a.x_enc) {
          if (valid) { a.g_dirs_rows[3 * (size_t)row_g] = gd[0]; a.g_dirs_rows[3 * (size_t)row_g + 1] = gd[1]; a.g_dirs_rows[3 * (size_t)row_g + 2] = gd[2]; }
        } else {
          bool head = seg_reduce<3>(ray[t], gd, lane);
          if (valid && head) {
            if (a.seg_part) { float* sp = a.seg_part + (size_t)row_g * 9 + 6; sp[0] = gd[0]; sp[1] = gd[1]; sp[2] = gd[2]; }
            else { atomicAdd(a.g_rays_d + 3 * ray[t], gd[0]); atomicAdd(a.g_rays_d + 3 * ray[t] + 1, gd[1]); atomicAdd(a.g_rays_d + 3 * ray[t] + 2, gd[2]); }
          }
        }
      }
      tc::fence_proxy_async();
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive_remote(a_ready_leader); tc::mbar_arrive_remote(a_ready_leader + 8); }

      for (int jn = 0; jn < n_jobs; ++jn) {
        const BJob& jb = a.plan.j[jn];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t, row_g = tile * TM + q;
          const bool valid = row_g < rows, tile_ok = tile < n_tiles;
          const uint32_t bufX_t = bufX0 + t * ACT_BYTES;
          const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + t * 256;
          uint8_t* dy_tile = a.dy + (size_t)tile * a.n_slots * ACT_BYTES;
          // forward ReLU gate bits of this row for this warp's two 32-column blocks, fetched BEFORE waiting
          uint32_t gate0 = 0, gate1 = 0;
          if (tile_ok && jb.mask_slot >= 0) {
            const uint32_t* gb = reinterpret_cast<const uint32_t*>(a.stash_bits + ((size_t)tile * a.n_slots + jb.mask_slot) * BITS_BYTES) + q;
            gate0 = gb[(2 * cq) * TM]; gate1 = gb[(2 * cq + 1) * TM];      // [32-column block][row] (mlp_tc_fwd.cu)
          }
          tc::mbar_wait(&bars->acc_full[t], par);
          tc::tcgen05_fence_after();
          if ((st_pending >> t & 1u) && (jb.kind == BK_MASK_STORE || jb.kind == BK_SIGMA_INJECT || jb.kind == BK_RELOAD_SKIP)) {
            tc::mbar_wait(&bars->st_done[t], stpar >> t & 1u);     // the stash warps are done reading the tile about to
            stpar ^= 1u << t;                                      // be overwritten
            st_pending &= ~(1u << t);
          }
          if (jb.kind == BK_MASK_STORE) {
            // this warp owns accumulator columns [64 cq, 64 cq + 64), walked in four 16-column halves with the TMEM
            // load of the next half in flight while one is gated, packed, stored (next A operand + dY stash)
            auto half = [&](const uint32_t (&v)[16], int h) {
              const uint32_t gb = h < 2 ? gate0 : gate1;
              const int pos0 = (h & 1) * 16, kg0 = cq * 8 + h * 2;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                uint4 o;
                o.x = relu_gate2(gb, pos0 + j * 8 + 0, __uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                o.y = relu_gate2(gb, pos0 + j * 8 + 2, __uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                o.z = relu_gate2(gb, pos0 + j * 8 + 4, __uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                o.w = relu_gate2(gb, pos0 + j * 8 + 6, __uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                sts_v4(bufX_t + (kg0 + j) * PLANE + q * 16, o);       // next A operand; the stash warps copy it to HBM
              }
            };
            uint32_t va[16], vb[16];
            const uint32_t tcol = taddr + cq * 64;
            tc::tmem_ld16(tcol, va);
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 16, vb);
            half(va, 0);
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 32, va);
            half(vb, 1);
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 48, vb);
            half(va, 2);
            tc::tmem_ld_wait();
            half(vb, 3);
          } else if (jb.kind == BK_SIGMA_INJECT) {
            // d relu(sigma.0) pre-activation = g_sigma * w_sigma2 gated by the forward gate bits; the accumulator
            // (gradient that arrived through sh.0) stays in TMEM and the next job accumulates onto it.
            uint8_t* dyo = dy_tile + (size_t)jb.dy_slot * ACT_BYTES;
            const float gs = g_sigma[t];
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const int kg = cq * 8 + k8;
              const float4 s0 = *reinterpret_cast<const float4*>(w2 + kg * 8);
              const float4 s1 = *reinterpret_cast<const float4*>(w2 + kg * 8 + 4);
              const uint32_t gb = k8 < 4 ? gate0 : gate1;
              const int pos = (kg & 3) * 8;
              uint4 o;
              o.x = relu_gate2(gb, pos + 0, gs * s0.x, gs * s0.y);
              o.y = relu_gate2(gb, pos + 2, gs * s0.z, gs * s0.w);
              o.z = relu_gate2(gb, pos + 4, gs * s1.x, gs * s1.y);
              o.w = relu_gate2(gb, pos + 6, gs * s1.z, gs * s1.w);
              sts_v4(bufX_t + kg * PLANE + q * 16, o);
              if (tile_ok) *reinterpret_cast<uint4*>(dyo + stash_off(q, kg, 32)) = o;
            }

          } else if (jb.kind == BK_RELOAD_SKIP) {
            // bring the skip layer's dY tile (written a few jobs ago by this very thread: same row, same columns)
            // back as the A operand
            const uint8_t* src = dy_tile + (size_t)a.plan.skip_dy_slot * ACT_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const int kg = cq * 8 + k8;
              const uint4 v = tile_ok ? *reinterpret_cast<const uint4*>(src + stash_off(q, kg, 32)) : make_uint4(0, 0, 0, 0);
              sts_v4(bufX_t + kg * PLANE + q * 16, v);
            }
          } else if (cq == 0) {   // BK_ENC_OUT: 64 accumulator columns, one thread per row
            float d[64];
            {
              uint32_t v[32];
              tc::tmem_ld32(taddr, v);
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[i]);
              tc::tmem_ld32(taddr + 32, v);
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) d[32 + i] = __uint_as_float(v[i]);
            }
            if (a.x_enc) {
              if (valid) {
                float* dst = a.g_x_enc + (size_t)row_g * a.ld_enc;
#pragma unroll
                for (int i = 0; i < 63; ++i) dst[i] = d[i];
              }
            } else {
              float gv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              const int ry = ray[t];
              if (valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const float xc = a.rays_o[3 * ry + c] + a.rays_d[3 * ry + c] * z[t];
                  float sn, cs;
                  sincosf(xc, &sn, &cs);
                  float g = d[c], f = 1.f;
#pragma unroll
                  for (int kf = 0; kf < 10; ++kf) {
                    const float bw = bw_s[kf];
                    g += bw * f * (d[3 + c * 20 + kf] * cs - d[3 + c * 20 + 10 + kf] * sn);
                    const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
                    sn = s2; cs = c2; f *= 2.f;
                  }
                  gv[c] = g;
                  gv[3 + c] = g * z[t];
                }
              }
              bool head = seg_reduce<6>(ry, gv, lane);
              if (valid && head) {
                if (a.seg_part) {
                  float* sp = a.seg_part + (size_t)row_g * 9;
#pragma unroll
                  for (int c = 0; c < 6; ++c) sp[c] = gv[c];
                } else {
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    atomicAdd(a.g_rays_o + 3 * ry + c, gv[c]);
                    atomicAdd(a.g_rays_d + 3 * ry + c, gv[3 + c]);
                  }
                }
              }
            }
          }
          if (jn + 1 < n_jobs) {
            tc::fence_proxy_async();
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_remote(a_ready_leader + t * 8);
          }
          if (jb.kind == BK_MASK_STORE) {
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->st_full[t]);
            st_pending |= 1u << t;
          }
        }
        par ^= 1;
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();            // no CTA leaves while its peer may still arrive on its barriers / read its operands
  if (warp == BW_MMA) tc::tmem_dealloc2(tmem, 512);
}

}  // namespace mlptc

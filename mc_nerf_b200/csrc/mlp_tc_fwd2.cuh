// Forward kernel, second generation: the A operand of every hidden layer lives in TENSOR MEMORY.
//
// Why: the first-generation kernel (mlp_tc_fwd_k) keeps the activation tile in shared memory, so per 128x256x256
// tile-layer a CTA moves 256 KB (inference) / 320 KB (training) through its 128 B/clk shared-memory pipe against 2048
// cycles of MMA time - that pipe, not the tensor core, set its rate.  Here the epilogue writes the next
// layer's bf16 A operand straight back into tensor memory (tcgen05.st) and the MMAs read it from there
// (tcgen05.mma [a_tmem] form), the two tile slots of a CTA share every weight block, and the training stash is stored
// from the epilogue's registers: what is left in the shared-memory pipe is the B operand (64 KB read per tile-layer,
// 32 KB filled).
//
// Tensor memory (512 columns) per CTA:  slot t in {0,1}:  A_t = columns [256t, 256t+128)  (128 rows x 256 bf16),
//                                                        acc_t = columns [256t+128, 256t+256) (128 fp32 columns).
// A 256-wide layer runs as two 128-column PASSES h = 0, 1 per slot; the issue order  (h0,t0) (h0,t1) (h1,t0) (h1,t1)
// lets the epilogue of one pass overlap the MMAs of the next one on the other slot.  The epilogue of pass 0 keeps its
// 64 packed bf16 pairs in registers (A_t is still being read by pass 1); the epilogue of pass 1 writes both halves.
// The encoding tile stays in shared memory (layer 0 and the skip layer read it with the ordinary descriptor form).
//
// CTA = 20 warps (+ 2 in training):
//                  0-7 / 8-15  epilogues of slot 0 / 1 (TMEM lane quarter = warp % 4, column half = warp / 4 % 2)
//                  16-17       input stage (sampling + encoding of the NEXT tile of both slots, off the critical path)
//                  18          weight producer (bulk async copies of whole pass blocks into a 3/4-deep ring)
//                  19          leader CTA: MMA issuer; peer CTA: relays "my half has landed"
//                  20-21       training: copy the stash staging buffers to HBM
#pragma once

namespace mlptc {

constexpr int F2_STAGES_MAX = 4;
constexpr int F2_STAGE_BYTES = 64 * (ENCW + WID + BIAS_K) * 2;     // 43 008: widest pass block (skip layer), 64 rows per CTA
constexpr int F2_COPY = 16384;                                      // bulk-copy piece
struct __align__(16) SmemBars2 {
  uint64_t w_full[F2_STAGES_MAX], w_empty[F2_STAGES_MAX];
  uint64_t acc_full[2];       // pass of slot t complete in tensor memory (own copy in each CTA)
  uint64_t epi_done[2];       // leader's copy: the 8 epilogue warps of slot t of BOTH CTAs are done with the pass
  uint64_t enc_ready[2];      // leader's copy: encoding tile of slot t written in both CTAs
  uint64_t enc_free[2];       // last MMA reading the encoding tile of slot t has completed (own copy in each CTA)
  uint64_t st_full[2], st_done[2];   // training: staging buffer of slot t holds a pass / has been copied to the stash
  uint32_t tmem_base;
};
// inference: 4 ring stages; training: 3 stages + one 32 KB stash staging buffer per slot (see the epilogue)
constexpr int F2_STG_BYTES = TM * 128 * 2;                          // one pass of one slot: [row half 2][k-group 16][64 rows][16 B]
#ifndef MCNERF_F2_STASH_WARPS
#define MCNERF_F2_STASH_WARPS 1
#endif
// 1: the training stash goes through a shared-memory staging buffer that two extra warps copy to HBM (the stores'
//    back-pressure - 4.5 TB/s of stash writes - then stalls those warps, not the epilogue); 0: 16-byte stores from the
//    epilogue's registers (measured: 2000-2500 instead of 800 cycles per pass-epilogue)
constexpr bool F2_STASH_WARPS = MCNERF_F2_STASH_WARPS;
__host__ __device__ constexpr int f2_stages(bool train) { return train && F2_STASH_WARPS ? 3 : 4; }
__host__ __device__ constexpr int smem_fwd2(bool train) {
  return f2_stages(train) * F2_STAGE_BYTES + (train && F2_STASH_WARPS ? 2 * F2_STG_BYTES : 0) + 2 * ENC_BYTES + ONES_BYTES + W2_FLOATS * 4 +
         2 * TM * 4 + 256;
}
static_assert(sizeof(SmemBars2) <= 256 && smem_fwd2(true) <= 232448 && smem_fwd2(false) <= 232448, "shared memory budget");
// Inference: 20 warps, 96 registers per thread.  Training adds two stash warps: 22 warps leave 80 registers, which is why
// the epilogue then parks the pass-0 half of the next A operand in the staging buffer instead of in 32 registers.
constexpr int F2_THREADS = 640, F2_THREADS_TRAIN = F2_STASH_WARPS ? 704 : 640;
constexpr int F2_W_ENC0 = 16, F2_W_PROD = 18, F2_W_MMA = 19, F2_W_STASH0 = 20;

__device__ __forceinline__ void tmem_st8p(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <bool TRAIN>
__global__ void __launch_bounds__(TRAIN ? F2_THREADS_TRAIN : F2_THREADS, 1) mlp_tc_fwd2_k(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int F2_STAGES = f2_stages(TRAIN);
  uint8_t* wst = smem;                                      // [F2_STAGES][F2_STAGE_BYTES]
  uint8_t* stg = wst + F2_STAGES * F2_STAGE_BYTES;          // training: [2][F2_STG_BYTES] stash staging
  uint8_t* enc = stg + (TRAIN && F2_STASH_WARPS ? 2 * F2_STG_BYTES : 0);      // [2][ENC_BYTES]
  uint8_t* ones = enc + 2 * ENC_BYTES;                      // broadcast ones operand of the bias MMAs
  float* w2s = reinterpret_cast<float*>(ones + ONES_BYTES); // w_sigma2[256], b_sigma2, BARF band weights at 264
  float* part = w2s + W2_FLOATS;                            // [2][128] sigma partial sums of the upper column halves
  SmemBars2* bars = reinterpret_cast<SmemBars2*>(part + 2 * TM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int n_tiles = (rows + TM - 1) / TM;
  const int n_pairs = (n_tiles + 1) / 2;
  const int n_steps = a.plan.n_steps;
  const uint32_t crank = tc::cluster_ctarank();
  const int n_iter = (n_pairs + (int)gridDim.x - 1) / (int)gridDim.x;
  int last_enc_step = 0;
  for (int s = 0; s < n_steps; ++s)
    if (a.plan.s[s].a_src == A_ENC_ACT) last_enc_step = s;

  if (tid == 0) {
    for (int i = 0; i < F2_STAGES; ++i) { tc::mbar_init(&bars->w_full[i], crank == 0 ? 2 : 1); tc::mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->acc_full[i], 1);
      tc::mbar_init(&bars->epi_done[i], 16);
      tc::mbar_init(&bars->enc_ready[i], 4);
      tc::mbar_init(&bars->enc_free[i], 1);
      tc::mbar_init(&bars->st_full[i], 8);
      tc::mbar_init(&bars->st_done[i], 2);
    }
    tc::mbar_init_fence();
  }
  if (warp == F2_W_MMA) tc::tmem_alloc2(&bars->tmem_base, 512);
  for (int i = tid; i < 257; i += blockDim.x) w2s[i] = a.bias[a.sig2_off + i];
  if (tid < 10) w2s[264 + tid] = a.smp.band_w_dev ? a.smp.band_w_dev[tid] : a.smp.band_w[tid];
  if (tid < 128)
    reinterpret_cast<__nv_bfloat16*>(ones)[tid] = __float2bfloat16((tid < 64 && (tid & 7) < 2) ? 1.f : 0.f);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (TRAIN && F2_STASH_WARPS && warp >= F2_W_STASH0) {
    // ------------------------------------------------------------------ stash warps: staging buffer -> HBM stash
    // a pass of slot t is [row half][k-group 16][64 rows][16 B] on both sides: warp sw copies row half sw (16 KB)
    const int sw = warp - F2_W_STASH0;
    uint32_t spar = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      for (int s = 0; s < n_steps; ++s) {
        const Step& st = a.plan.s[s];
        if (st.epi == EPI_OUT || st.stash_slot < 0) continue;
        for (int h = 0; h < 2; ++h)
          for (int t = 0; t < 2; ++t) {
            const int tile = 2 * pair + t;
            tc::mbar_wait(&bars->st_full[t], (spar >> t) & 1);
            spar ^= 1u << t;
            if (tile < n_tiles) {
              uint8_t* dst = a.stash + ((size_t)tile * a.n_slots + st.stash_slot) * ACT_BYTES + sw * 32768 + h * 16384 + lane * 16;
              const uint32_t src = tc::smem_u32(stg) + t * F2_STG_BYTES + sw * 16384 + lane * 16;
#pragma unroll 1
              for (int i0 = 0; i0 < 32; i0 += 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                               : "r"(src + (i0 + u) * 512));
#pragma unroll
                for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(dst + (i0 + u) * 512) = v[u];
              }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->st_done[t]);
          }
      }
    }
  } else if (warp == F2_W_PROD) {
    if (lane == 0) {
    // ------------------------------------------------------------------ weight producer: one block per (step, pass)
    // A block = this CTA's rows of one pass, whole K plus the bias planes.
    // (Requested by the consuming thread itself, or by a second lane of the MMA warp, the forward got 20 % slower.)
    int stage = 0;
    uint32_t par = 0;
    for (int it = 0; it < n_iter; ++it)
      for (int s = 0; s < n_steps; ++s) {
        const Step& st = a.plan.s[s];
        const int n_pass = st.N == WID ? 2 : 1;
        const uint32_t bytes = (uint32_t)(st.N / (2 * n_pass)) * (uint32_t)(st.n_chunks * KC + BIAS_K) * 2;
        for (int h = 0; h < n_pass; ++h) {
          const uint8_t* src = a.wpack + st.w_off + (size_t)(h * 2 + (int)crank) * bytes;
          uint8_t* dst = wst + stage * F2_STAGE_BYTES;
          tc::mbar_wait(&bars->w_empty[stage], par ^ 1);
          tc::mbar_arrive_expect_tx(&bars->w_full[stage], bytes);
          for (uint32_t off = 0; off < bytes; off += F2_COPY)
            tc::bulk_g2s(dst + off, src + off, min((uint32_t)F2_COPY, bytes - off), &bars->w_full[stage]);
          if (++stage == F2_STAGES) { stage = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == F2_W_MMA) {
    if (lane == 0 && crank == 0) {
      // ------------------------------------------------------------------ MMA issuer
      // Issue-loop economy: tcgen05.mma takes its operands from UNIFORM registers, and a 256x128x16 MMA executes in 64
      // cycles - an IMAD + R2UR pair per operand per MMA (what the compiler emits when an address depends on a loop-carried
      // variable such as the ring stage) makes the issue loop the bottleneck (~85 cycles per MMA measured).  So: tensor
      // memory addresses are literal (the 512-column allocation starts at column 0, checked below), the ring stage is
      // turned into a literal by a switch, and everything else is an immediate offset from a shared-memory base.
      if (tmem != 0) { printf("mcnerf: tensor memory base %u != 0\n", tmem); __trap(); }
      int stage = 0;
      uint32_t par = 0, epar = 0;
      const uint32_t hi = tc::umma_desc_hi(128), ones_hi = tc::umma_desc_hi(0);
      // descriptor (64-row blocks) of the current ring stage: a counter of its own, used by nothing but the MMAs
      const uint32_t b_ring0 = tc::umma_desc_lo(tc::smem_u32(wst), 64 * 16);
      uint32_t b_ring = b_ring0, ring_i = 0;
      const uint32_t enc_lo = tc::umma_desc_lo(tc::smem_u32(enc), PLANE);
      const uint32_t ones_lo = tc::umma_desc_lo(tc::smem_u32(ones), 128);
      const uint32_t wst_addr = tc::smem_u32(wst);
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]), empty0 = tc::smem_u32(&bars->w_empty[0]);
      const uint32_t accf0 = tc::smem_u32(&bars->acc_full[0]), edone0 = tc::smem_u32(&bars->epi_done[0]);
      const uint32_t erdy0 = tc::smem_u32(&bars->enc_ready[0]), efree0 = tc::smem_u32(&bars->enc_free[0]);
      MC_TRACE(long long tr_enc = 0; long long tr_epi = 0; long long tr_w = 0; long long tr_n = 0; long long tr_t0 = 0; long long tr_t1 = 0;)
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < n_steps; ++s) {
          const int N = a.plan.s[s].N, a_src = a.plan.s[s].a_src;
          const int n_pass = N == WID ? 2 : 1, NP = N / (2 * n_pass);
          const uint32_t idesc = tc::umma_idesc_bf16(2 * TM, N / n_pass);
          const uint32_t b_inc = (2u * NP * 16) >> 4;      // narrow steps (sh.2) always read the activation operand
          const bool use_enc = a_src != A_ACT, use_act = a_src != A_ENC;
          for (int h = 0; h < n_pass; ++h) {
            const uint32_t b_lo0 = tc::umma_desc_lo(wst_addr + stage * F2_STAGE_BYTES, NP * 16);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              MC_TRACE(long long c0 = clock64();)
              if (s == 0 && h == 0) tc::mbar_wait_addr(erdy0 + t * 8, it & 1);
              MC_TRACE(long long c1 = clock64();)
              if (it | s | h) {
                tc::mbar_wait_addr(edone0 + t * 8, (epar >> t) & 1);      // accumulator drained (h = 0: and A_t rewritten)
                epar ^= 1u << t;
              }
              MC_TRACE(long long c2 = clock64();)
              if (t == 0) tc::mbar_wait_addr(full0 + stage * 8, par);
              MC_TRACE(if (it >= 1) {
                const long long c3 = clock64();
                tr_enc += c1 - c0; tr_epi += c2 - c1; tr_w += c3 - c2; tr_n += 1;
                if (tr_t0 == 0) tr_t0 = c0;
                tr_t1 = c3;
              })
              tc::tcgen05_fence_after();
              const uint32_t d_tmem = t * 256 + 128, a_tmem = t * 256;
              if (N == WID) {
                constexpr uint32_t BI = (2u * 64 * 16) >> 4;              // descriptor units per K = 16 step (64 rows x 16 B x 2 planes)
                constexpr uint32_t IDESC = tc::umma_idesc_bf16(2 * TM, WID / 2);
                uint32_t b1 = b_ring;
                if (use_enc) {
                  uint32_t e_lo = enc_lo + t * (ENC_BYTES >> 4);
                  tc::umma2_bf16_w(d_tmem, e_lo, hi, b1, hi, IDESC, false);
#pragma unroll
                  for (int j = 1; j < ENCW / 16; ++j) {
                    e_lo += (2 * PLANE) >> 4;
                    b1 += BI;
                    tc::umma2_bf16_w(d_tmem, e_lo, hi, b1, hi, IDESC, true);
                  }
                  b1 += BI;
                }
                if (use_act) {
                  uint32_t a_t = a_tmem;
                  tc::umma2_bf16_ts(d_tmem, a_t, b1, hi, IDESC, use_enc);
#pragma unroll
                  for (int j = 1; j < WID / 16; ++j) {
                    a_t += 8;
                    b1 += BI;
                    tc::umma2_bf16_ts(d_tmem, a_t, b1, hi, IDESC, true);
                  }
                  b1 += BI;
                }
                tc::umma2_bf16_w(d_tmem, ones_lo, ones_hi, b1, hi, IDESC, true);      // + bias
              } else {
                uint32_t b_lo = b_lo0;
#pragma unroll 4
                for (int j = 0; j < WID / 16; ++j) {
                  tc::umma2_bf16_ts(d_tmem, a_tmem + 8 * j, b_lo, hi, idesc, j != 0);
                  b_lo += b_inc;
                }
                tc::umma2_bf16_w(d_tmem, ones_lo, ones_hi, b_lo, hi, idesc, true);
              }
              tc::umma2_commit_multicast_addr(accf0 + t * 8, (uint16_t)3);
              if (s == last_enc_step && h == n_pass - 1) tc::umma2_commit_multicast_addr(efree0 + t * 8, (uint16_t)3);
              if (t == 1) {
                tc::umma2_commit_multicast_addr(empty0 + stage * 8, (uint16_t)3);
                if (++stage == F2_STAGES) { stage = 0; par ^= 1; }
                b_ring += F2_STAGE_BYTES >> 4;
                if (++ring_i == F2_STAGES) { ring_i = 0; b_ring = b_ring0; }
              }
            }
          }
        }
      MC_TRACE(if (a.dbg && blockIdx.x == 0) {
        a.dbg[0] = tr_enc; a.dbg[1] = tr_epi; a.dbg[2] = tr_w; a.dbg[3] = tr_n; a.dbg[4] = tr_t0; a.dbg[5] = tr_t1;
      })
    } else if (lane == 0) {
      // peer CTA: relay "my half of the block has landed" to the leader's full barrier
      int stage = 0;
      uint32_t par = 0;
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]);
      const uint32_t leader_full0 = tc::mapa(full0, 0);
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < n_steps; ++s) {
          const int n_pass = a.plan.s[s].N == WID ? 2 : 1;
          for (int h = 0; h < n_pass; ++h) {
            tc::mbar_wait_addr(full0 + stage * 8, par);
            tc::mbar_arrive_remote(leader_full0 + stage * 8);
            if (++stage == F2_STAGES) { stage = 0; par ^= 1; }
          }
        }
    }
  } else if (warp >= F2_W_ENC0 && warp < F2_W_PROD) {
    // ------------------------------------------------------------------ input stage: 2 warps, 2 rows per thread and slot
    const int e = warp - F2_W_ENC0;
    const uint32_t enc0 = tc::smem_u32(enc);
    const uint32_t erdy_leader = tc::mapa(tc::smem_u32(&bars->enc_ready[0]), 0);
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      for (int t = 0; t < 2; ++t) {
        if (it > 0) tc::mbar_wait(&bars->enc_free[t], (it - 1) & 1);
        const int tile = 2 * pair + t;
        uint8_t* st_enc = (TRAIN && tile < n_tiles) ? a.stash_enc + (size_t)tile * ENC_BYTES : nullptr;
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
          const int q = e * 64 + r * 32 + lane, row_g = tile * TM + q;
          encode_row(a, row_g, tile < n_tiles && row_g < rows, enc0 + t * ENC_BYTES, q, st_enc, w2s + 264);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_remote(erdy_leader + t * 8);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogues of slot t
    const int t = warp >> 3, lq = warp & 3, ch = (warp >> 2) & 1;
    const int q = lq * 32 + lane;                                  // row in tile == TMEM lane
    const uint32_t edone_leader = tc::mapa(tc::smem_u32(&bars->epi_done[t]), 0);
    const uint32_t t_acc = tmem + ((uint32_t)(lq * 32) << 16) + t * 256 + 128 + ch * 64;   // this thread's 64 accumulator columns
    const uint32_t t_a = tmem + ((uint32_t)(lq * 32) << 16) + t * 256 + ch * 32;           // its 32 A columns of pass 0 (+64: pass 1)
    uint32_t par = 0, stpar = 0;
    uint32_t held[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) held[i] = 0;
    MC_TRACE(long long tr_e = 0; long long tr_en = 0;)
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      const int tile = 2 * pair + t, row_g = tile * TM + q;
      const bool valid = tile < n_tiles && row_g < rows;
      float sig_dot = 0.f;
      for (int s = 0; s < n_steps; ++s) {
        const Step& st = a.plan.s[s];
        const int n_pass = st.N == WID ? 2 : 1;
        uint8_t* st_tile = (TRAIN && st.stash_slot >= 0 && tile < n_tiles)
                               ? a.stash + ((size_t)tile * a.n_slots + st.stash_slot) * ACT_BYTES
                               : nullptr;
        uint32_t* gate_out = st_tile ? reinterpret_cast<uint32_t*>(a.stash_bits + ((size_t)tile * a.n_slots + st.stash_slot) * BITS_BYTES) + q
                                     : nullptr;
        for (int h = 0; h < n_pass; ++h) {
          tc::mbar_wait(&bars->acc_full[t], par);
          par ^= 1;
          MC_TRACE(const long long e0 = clock64();)
          tc::tcgen05_fence_after();
          if (st.epi == EPI_OUT) {
            if (ch) part[t * TM + q] = sig_dot;
            named_bar_sync(1 + t * 4 + lq, 64);                    // the two column-half warps of this lane quarter
            if (!ch) {
              uint32_t v[32];
              tc::tmem_ld32(t_acc, v);
              tc::tmem_ld_wait();
              const float sigma_raw = sig_dot + part[t * TM + q] + w2s[256];
              if (valid) {
                float sh[27];
#pragma unroll
                for (int i = 0; i < 27; ++i) sh[i] = __uint_as_float(v[i]);
                const float* dp;
                if (a.x_enc) dp = a.dirs_rows + (size_t)row_g * 3;
                else {
                  int flat = a.sel_idx ? a.sel_idx[row_g] : row_g;
                  dp = a.rays_d + 3 * (size_t)(flat / a.smp.S);
                }
                float x = dp[0], y = dp[1], z = dp[2];
                float Y[9] = {cC0, -cC1 * y, cC1 * z, -cC1 * x, cC2[0] * x * y, cC2[1] * y * z,
                              cC2[2] * (2.f * z * z - x * x - y * y), cC2[3] * x * z, cC2[4] * (x * x - y * y)};
                float c[3];
#pragma unroll
                for (int chn = 0; chn < 3; ++chn) {
                  float acc = 0.f;
#pragma unroll
                  for (int b = 0; b < 9; ++b) acc += Y[b] * sh[9 * chn + b];
                  c[chn] = sigmoid_f(acc);
                }
                reinterpret_cast<float4*>(a.out4)[row_g] = make_float4(sigma_raw, c[0], c[1], c[2]);
                if (TRAIN) {
                  float4* dst = reinterpret_cast<float4*>(a.stash_sh) + (size_t)(row_g >> 7) * (TM * SH_LD / 4) + (row_g & (TM - 1));
#pragma unroll
                  for (int i = 0; i < 7; ++i)
                    dst[i * TM] = make_float4(sh[4 * i], sh[4 * i + 1], sh[4 * i + 2], i < 6 ? sh[4 * i + 3] : 0.f);
                }
              }
            }
          } else {
            const bool relu_step = st.epi == EPI_RELU;                 // else sigma.0: fp32 activations feed the sigma.2 dot
            const bool stash_step = TRAIN && st.stash_slot >= 0;
            // Training: the bf16 tile goes to the HBM stash through a shared-memory staging buffer (see F2_STASH_WARPS); the
            // staging image of a pass is the HBM image of its 16 k-groups: [row half][k-group][64 rows][16 B].
            const uint32_t stg_t = tc::smem_u32(stg) + t * F2_STG_BYTES + (q >> 6) * 16384 + (q & 63) * 16 + ch * 8192;
            constexpr bool PARK = TRAIN && F2_STASH_WARPS;             // pass-0 half parked in the staging buffer, not in registers
            if (relu_step && h == 1) {
              // pass 1 has completed: nothing reads the old A_t any more - first the half kept from pass 0
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (PARK) {
                  uint32_t w[8];
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(stg_t + (2 * g) * 1024));
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(stg_t + (2 * g + 1) * 1024));
                  tmem_st8p(t_a + 8 * g, w);
                } else {
                  tmem_st8p(t_a + 8 * g, held + 8 * g);
                }
              }
            }
            if (F2_STASH_WARPS && stash_step) {                       // the stash warps have copied the previous pass
              tc::mbar_wait(&bars->st_done[t], (stpar & 1) ^ 1);
              stpar ^= 1;
            }
            float dot = 0.f;
            uint32_t sbits = 0;
            const int col0 = h * 128 + ch * 64;                        // first output feature of this thread's 64 columns
            // Two 32-column loads: a TMEM load takes 150-300 cycles while the MMAs of the other slot keep the port busy,
            // and four dependent 16-column loads were most of the epilogue's latency (both at once would need 64 registers
            // next to the 32 held ones).
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t v[16], p[8];
              tc::tmem_ld16(t_acc + 16 * g, v);
              tc::tmem_ld_wait();
              if (relu_step) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = tc::pack_bf16_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
              } else {
                const float4* w4 = reinterpret_cast<const float4*>(w2s + col0 + 16 * g);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const float4 w = w4[i4];
                  const float x0 = fmaxf(__uint_as_float(v[4 * i4]), 0.f), x1 = fmaxf(__uint_as_float(v[4 * i4 + 1]), 0.f);
                  const float x2 = fmaxf(__uint_as_float(v[4 * i4 + 2]), 0.f), x3 = fmaxf(__uint_as_float(v[4 * i4 + 3]), 0.f);
                  dot += x0 * w.x + x1 * w.y + x2 * w.z + x3 * w.w;
                  p[2 * i4] = tc::pack_bf16(x0, x1);
                  p[2 * i4 + 1] = tc::pack_bf16(x2, x3);
                }
              }
              if (relu_step) {
                if (h == 0) {
                  if (!PARK) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) held[8 * g + i] = p[i];
                  }
                } else {
                  tmem_st8p(t_a + 64 + 8 * g, p);
                }
              }
              if (TRAIN) {
                if (stash_step) {
                  if (F2_STASH_WARPS) {
                    st_shared_v4(stg_t + (2 * g) * 1024, p[0], p[1], p[2], p[3]);
                    st_shared_v4(stg_t + (2 * g + 1) * 1024, p[4], p[5], p[6], p[7]);
                  } else if (st_tile) {
                    const int kg = col0 / 8 + 2 * g;
                    *reinterpret_cast<uint4*>(st_tile + stash_off(q, kg, 32)) = make_uint4(p[0], p[1], p[2], p[3]);
                    *reinterpret_cast<uint4*>(st_tile + stash_off(q, kg + 1, 32)) = make_uint4(p[4], p[5], p[6], p[7]);
                  }
                }
                sbits |= tc::sign_bits16(v, g & 1);
                if (g & 1) {
                  if (gate_out) gate_out[(col0 / 32 + (g >> 1)) * TM] = ~sbits;      // gate = accumulator > 0
                  sbits = 0;
                }
              }
            }
            if (!relu_step) sig_dot += dot;
            else if (h == 1) tc::tmem_st_wait();
          }
          tc::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive_remote(edone_leader);
            if (F2_STASH_WARPS && TRAIN && st.epi != EPI_OUT && st.stash_slot >= 0) tc::mbar_arrive(&bars->st_full[t]);
          }
          MC_TRACE(if (it >= 1) { tr_e += clock64() - e0; tr_en += 1; })
          MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2 && s == 2 && lane == 0 && (warp & 7) == 0) a.dbg[128 + (warp >> 3) * 32 + h * 16 + 8] = clock64() - e0;)
        }
      }
    }
    MC_TRACE(if (a.dbg && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15)) { a.dbg[6 + 2 * (warp != 0)] = tr_e; a.dbg[7 + 2 * (warp != 0)] = tr_en; })
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();
  if (warp == F2_W_MMA) tc::tmem_dealloc2(tmem, 512);
}

}  // namespace mlptc

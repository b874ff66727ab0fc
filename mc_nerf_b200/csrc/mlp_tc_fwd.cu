// bf16 tcgen05 NeRF MLP, forward: ONE persistent kernel per network evaluation that fuses
//   ray sampling + sin/cos (BARF-weighted) encoding  ->  8x256 trunk with input skip  ->  sigma head
//   -> SH-27 colour head -> eval_sh + sigmoid,
// keeping every activation on-chip (shared memory as the UMMA A operand, fp32 accumulators in TMEM).
// ref: model/mc_nerf.py:599-602/633-635 (sampling), model/net_block.py:20-35 (encoding), :67-78 (MLP),
//      model/net_utils.py:154-169 (eval_sh).
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) works on four 128-row tiles: two slots per CTA, and every MMA
// covers M = 256 rows (slot t of both CTAs) x N columns.  Each CTA stages only HALF of every weight chunk (its N/2
// output features); the tensor cores exchange the halves, so the shared-memory traffic per MMA that bounded the
// single-CTA version (A 4 KB + B 8 KB read + 8 KB of B filled per 128x256x16) drops to 4 + 4 + 4 KB.
// CTA = 18 warps, two 128-row tiles in flight (TMEM: 2 x 256 fp32 columns):
//   warps 0-7 / 8-15 : input stage + epilogues of tile slot 0 / 1 (TMEM lane quarter = warp % 4, column half = warp / 4 % 2)
//   warp 16          : weight producer - 1-D bulk async copies (TMA engine) of pre-packed UMMA-ready half chunks
//                      (64 reduction columns x N/2 rows = 16 KB) into a 3-deep ring, mbarrier full/empty
//   warp 17          : leader CTA: MMA issuer - one thread issues tcgen05.mma.cta_group::2 (M=256, N<=256, K=16) for
//                      both slots, alternating slots per layer so one slot's epilogue overlaps the other slot's MMAs;
//                      peer CTA: relays "my half of the chunk has landed" to the leader's full barrier
// Weights are re-streamed from L2 per tile (1.26 MB per net: L2 resident); activations never touch HBM in
// inference; in training each layer's bf16 activation tile is also written to the stash for the backward pass.
#include <stdlib.h>
#include "mlp_tc.cuh"

// clock64 tracing of CTA 0 (tools/trace_fwd.py) is compiled in only with -DMCNERF_TC_TRACE: the probes sit in the
// MMA issue loop, whose instruction count bounds the tensor-pipe utilisation.
#ifdef MCNERF_TC_TRACE
#define MC_TRACE(x) x
#else
#define MC_TRACE(x)
#endif
#define EPI_TRACE(k) MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2 && s == 2 && lane == 0 && (warp == 0 || warp == 15)) a.dbg[8 * 32 + (warp ? 16 : 0) + t * 8 + (k)] = clock64();)

namespace mlptc {

// ----------------------------------------------------------------------------------------- packing
// dst chunk image for a [N x K] K-major B operand: chunk c (32 k), plane kg (8 k), row n: ((c*4+kg)*N + n)*16 B.
// value(n,k) = src[n*sn + kk*sk] with the 63->64 pad remap applied to k (pad_k) or n (pad_n): see pack_all_k.
// Forward images use the CTA-pair layout instead: chunk C (KC2 = 64 k), half h (rows h*N/2 .. of the N output
// features), plane kg (8 k), row nn: (((C*2+h)*8+kg)*(N/2) + nn)*16 B - each CTA's half chunk is contiguous.

// ----------------------------------------------------------------------------------------- plan
int build_layout(const mcnerf_mlp_params* p, PackLayout* L) {
  MC_ARG(p && p->width == WID && p->in_ch == 63 && p->sh_dim == 27 && p->depth >= 2 && p->depth <= 12);
  int n_skip = __builtin_popcount(p->skip_mask);
  MC_ARG(n_skip <= 1 && (p->skip_mask & 1u) == 0 && (p->skip_mask >> p->depth) == 0);
  const int D = p->depth;
  L->depth = D;
  L->skip_mask = p->skip_mask;
  Plan& P = L->fwd;
  uint32_t off = 0, boff = 0;
  int ns = 0, bias = 0;
  auto add = [&](int a_src, int K, int N, int epi, int slot) {
    Step& s = P.s[ns];
    s.a_src = a_src; s.n_chunks = K / KC; s.N = N; s.epi = epi; s.w_off = off; s.bias_off = bias; s.stash_slot = slot;
    off += (uint32_t)N * K * 2 + (uint32_t)N * BIAS_K * 2;     // weight chunks, then the bias block (see pack_all_k)
    // transposed images: rows = inputs, reduction = outputs (N).  Encoding inputs get their own 64-row image.
    L->wb_enc[ns] = L->wb_main[ns] = 0;
    if (a_src == A_ENC) { L->wb_main[ns] = boff; boff += (uint32_t)ENCW * N * 2; }
    else if (a_src == A_ENC_ACT) {
      L->wb_enc[ns] = boff; boff += (uint32_t)ENCW * N * 2;
      L->wb_main[ns] = boff; boff += (uint32_t)WID * N * 2;
    } else { L->wb_main[ns] = boff; boff += (uint32_t)WID * N * 2; }
    bias += 256;
    ++ns;
  };
  int skip_layer = -1;
  for (int i = 0; i < D; ++i) {
    if (i == 0) add(A_ENC, ENCW, WID, EPI_RELU, i);
    else if (p->skip_mask >> i & 1u) { add(A_ENC_ACT, ENCW + WID, WID, EPI_RELU, i); skip_layer = i; }
    else add(A_ACT, WID, WID, EPI_RELU, i);
  }
  add(A_ACT, WID, WID, EPI_SIGMA, D);           // sigma.0 (sigma.2 is a dot product in its epilogue)
  add(A_ACT, WID, WID, EPI_RELU, D + 1);        // sh.0
  add(A_ACT, WID, 32, EPI_OUT, -1);             // sh.2 (27 -> 32 columns) + eval_sh + sigmoid
  P.n_steps = ns;
  L->wf_bytes = off;
  L->wb_bytes = boff;
  L->sig2_off = bias;
  L->bias_floats = bias + 256 + 8;

  // backward chain.  Forward-stash slots: l -> output of trunk layer l (h_{l+1}), D -> relu(sigma.0), D+1 -> relu(sh.0).
  // dY-stash slots:  l -> grad wrt pre-activation of trunk layer l, D -> sigma.0, D+1 -> sh.0.
  BPlan& B = L->bwd;
  int nb = 0;
  auto badd = [&](int a_small, int K, int N, int acc, int kind, uint32_t w_off, int mask_slot, int dy_slot) {
    BJob& j = B.j[nb++];
    j.a_small = a_small; j.n_chunks = K / KC; j.N = N; j.accumulate = acc; j.kind = kind; j.w_off = w_off;
    j.mask_slot = mask_slot; j.dy_slot = dy_slot;
  };
  badd(1, 32, WID, 0, BK_MASK_STORE, L->wb_main[D + 2], D + 1, D + 1);     // through sh.2 -> d relu(sh.0) pre-act
  badd(0, WID, WID, 0, BK_SIGMA_INJECT, L->wb_main[D + 1], D, D);          // through sh.0 (acc kept) ; inject sigma
  badd(0, WID, WID, 1, BK_MASK_STORE, L->wb_main[D], D - 1, D - 1);        // + through sigma.0 -> dY of layer D-1
  for (int l = D - 1; l >= 1; --l) badd(0, WID, WID, 0, BK_MASK_STORE, L->wb_main[l], l - 1, l - 1);
  if (skip_layer >= 0) {
    badd(0, WID, ENCW, 0, BK_RELOAD_SKIP, L->wb_main[0], -1, -1);          // layer 0 -> d enc (partial)
    badd(0, WID, ENCW, 1, BK_ENC_OUT, L->wb_enc[skip_layer], -1, -1);      // + skip layer's encoding part
  } else {
    badd(0, WID, ENCW, 0, BK_ENC_OUT, L->wb_main[0], -1, -1);
  }
  B.n_jobs = nb;
  B.skip_dy_slot = skip_layer;

  // weight-gradient jobs
  WPlan& Wp = L->wg;
  int nw = 0;
  auto wadd = [&](int dy, int x, int N, int tr, int hc, int which, int col_off, int n_valid, int bias_mode) {
    WJob& j = Wp.j[nw++];
    j.dy_slot = dy; j.x_slot = x; j.N = N; j.transposed = tr; j.head_col0 = hc; j.which = which; j.col_off = col_off;
    j.n_valid = n_valid; j.bias_mode = bias_mode;
  };
  for (int l = 0; l < D; ++l) {
    if (l == 0) wadd(0, -1, ENCW, 0, 0, 0, 0, 63, 1);
    else if (l == skip_layer) { wadd(l, -1, ENCW, 0, 0, l, 0, 63, 0); wadd(l, l - 1, WID, 0, 0, l, 63, WID, 1); }
    else wadd(l, l - 1, WID, 0, 0, l, 0, WID, 1);
  }
  wadd(D, D - 1, WID, 0, 0, D, 0, WID, 1);          // sigma.0
  wadd(D + 1, D - 1, WID, 0, 0, D + 1, 0, WID, 1);  // sh.0
  wadd(-1, D + 1, 32, 1, 0, D + 2, 0, WID, 2);      // sh.2 (transposed: A = relu(sh.0)^T, B = head tile cols 0..31)
  wadd(-1, D, 16, 1, 16, D + 3, 0, WID, 3);         // sigma.2 (transposed: B = head tile cols 16..31, g_sigma at col 31)
  Wp.n_jobs = nw;
  return 0;
}

// One launch packs every matrix of a network: a table of jobs, each thread resolves its job by a linear scan
// over the (<= 48) element-count prefixes.
constexpr int MAX_PACK_JOBS = 96;
struct PackJob {
  const float* src;
  int64_t sn, sk;          // source strides of the (n, k) indices (floats); bias jobs: unused
  int N, K;                // packed extents (bias jobs: N = padded length, K = 1)
  int pad_k, pad_n, n_valid, k_valid;
  int pair;                // 1: CTA-pair layout (forward images); 2: bias block of a forward image (src = bias[n]):
                           //    rows n, 16 k: k=0 bf16(b), k=1 bf16(b - bf16(b)), rest 0 - multiplied by the ones operand
                           // 3 / 4: the same two for the second-generation forward kernel (mlp_tc_fwd2.cuh): per step
                           //    [pass h][CTA rank][K/8 weight planes + 2 bias planes][NP rows][16 B], NP = rows per CTA and pass
  int kw;                  // pair 4: reduction length K of the step's weight matrix (the bias planes follow its K/8 planes)
  int is_bias;             // 1: fp32 copy of n_valid floats padded with zeros to N
  size_t dst_off;          // byte offset in wf / wb (is_bias: float offset in the bias block)
  int dst_sel;             // 0: wf, 1: wb, 2: bias block
};
struct PackArgs {
  int n_jobs;
  int64_t prefix[MAX_PACK_JOBS + 1];
  PackJob j[MAX_PACK_JOBS];
  uint8_t *wf, *wb;
  float* bias;
};

__global__ void __launch_bounds__(256) pack_all_k(const __grid_constant__ PackArgs a) {
  // work unit: one 16-byte group of 8 consecutive k (matrix jobs) or one float (bias jobs)
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= a.prefix[a.n_jobs]) return;
  int job = 0;
  while (gi >= a.prefix[job + 1]) ++job;
  const PackJob& pj = a.j[job];
  const int64_t t = gi - a.prefix[job];
  if (pj.is_bias) {
    a.bias[pj.dst_off + t] = t < pj.n_valid ? pj.src[t] : 0.f;
    return;
  }
  const int N = pj.N;
  const int n = (int)(t % N), kgi = (int)(t / N);
  uint4* dst = reinterpret_cast<uint4*>((pj.dst_sel == 0 ? a.wf : a.wb) + pj.dst_off);
  float v[8];
  int64_t dg = t;                                                    // destination group index (16-byte units)
  if (pj.pair == 2 || pj.pair == 4) {
    const int NH = N / 2;                                            // K = 16: kgi in {0, 1}
    dg = ((int64_t)(n / NH) * 2 + kgi) * NH + n % NH;
    if (pj.pair == 4) {
      const int NPP = N == WID ? WID / 2 : N, NP = NPP / 2, planes = pj.kw / 8 + 2;
      dg = ((int64_t)((n / NPP) * 2 + (n % NPP) / NP) * planes + pj.kw / 8 + kgi) * NP + n % NP;
    }
    const float b = n < pj.n_valid ? pj.src[n] : 0.f;
    const float hi = __bfloat162float(__float2bfloat16(b));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (kgi == 0) { v[0] = hi; v[1] = b - hi; }
  } else {
    if (pj.pair == 3) {
      const int NPP = N == WID ? WID / 2 : N, NP = NPP / 2, planes = pj.K / 8 + 2;
      dg = ((int64_t)((n / NPP) * 2 + (n % NPP) / NP) * planes + kgi) * NP + n % NP;
    } else if (pj.pair) {
      const int NH = N / 2, kgc = pj.K >= KC2 ? KC2 / 8 : pj.K / 8;   // k-groups per chunk
      dg = (((int64_t)(kgi / kgc) * 2 + n / NH) * kgc + (kgi % kgc)) * NH + n % NH;
    }
    int ns = n;
    bool n_ok = true;
    if (pj.pad_n) { if (n == 63) n_ok = false; else if (n > 63) ns = n - 1; }
    if (ns >= pj.n_valid) n_ok = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kgi * 8 + j;
      int ks = k;
      bool ok = n_ok;
      if (pj.pad_k) { if (k == 63) ok = false; else if (k > 63) ks = k - 1; }
      if (ks >= pj.k_valid) ok = false;
      v[j] = ok ? pj.src[ns * pj.sn + ks * pj.sk] : 0.f;
    }
  }
  uint4 w;
  w.x = tc::pack_bf16(v[0], v[1]);
  w.y = tc::pack_bf16(v[2], v[3]);
  w.z = tc::pack_bf16(v[4], v[5]);
  w.w = tc::pack_bf16(v[6], v[7]);
  dst[dg] = w;
}

// ----------------------------------------------------------------------------------------- forward kernel
struct FwdArgs {
  Plan plan;
  const uint8_t* wpack;
  const float* bias;
  int sig2_off;
  int bias_floats;
  long long* dbg;        // MCNERF_TC_DEBUG=7: clock64 trace of CTA 0 [event][step] (see tools/trace_fwd.py)
  const float *rays_o, *rays_d, *jitter;
  mcnerf_sampling smp;
  const int32_t* sel_idx;
  int n_rows;
  const int32_t* n_rows_dev;
  const float* x_enc;
  int ld_enc;
  const float* dirs_rows;
  float* out4;
  uint8_t* stash;        // [tile][n_slots][ACT_BYTES] or null
  uint8_t* stash_enc;    // [tile][ENC_BYTES]
  float* stash_sh;       // [tile][SH_LD/4 float4 groups][128 rows] float4
  uint8_t* stash_bits;   // [tile][n_slots][BITS_BYTES] ReLU gate bits
  int n_slots;
};

// Biases ride in the accumulator: every step ends with one extra K=16 MMA whose A operand is a broadcast "ones"
// block (2 core matrices; stride-byte offset 0 makes all sixteen 8-row groups read the same one) and whose B operand
// is the step's bias block (bf16 hi + lo) streamed through the ring like a small weight chunk.  That removes the
// bias loads from the epilogue - through shared memory they competed with the tensor core's operand reads for the
// data pipe, through the constant cache they cost 0.25 ms per 786k rows - and the 15 KB they used to occupy pay
// for a 4th ring stage.
constexpr int FSTAGE = 4;
struct __align__(16) SmemBars {
  uint64_t w_full[FSTAGE], w_empty[FSTAGE], a_ready[2], acc_full[2];
  uint64_t st_full[2], st_done[2];      // training: activation tile of slot t is complete / has been copied to the stash
  uint32_t tmem_base;
};
constexpr int ONES_BYTES = 256;
constexpr int W2_FLOATS = 280;         // w_sigma2[256], b_sigma2, pad to 264, then the 10 BARF band weights (+ pad)
constexpr int SMEM_FWD = 2 * ACT_BYTES + 2 * ENC_BYTES + FSTAGE * STAGE_BYTES + ONES_BYTES + W2_FLOATS * 4 + 256;
static_assert(sizeof(SmemBars) <= 256 && SMEM_FWD <= 232448, "shared memory budget");
// 18 warps: 16 epilogue warps (TMEM lane quarter = warp % 4, column quarter = warp / 4) that serve BOTH tile slots in
// turn - the slot whose accumulator just completed gets all 16, four per scheduler, which is what hides the
// TMEM-load / pack / store latency of one warp - plus 1 weight producer and 1 MMA issuer / relay.
// Training adds 2 stash warps: they copy each finished activation tile shared memory -> HBM stash while the next
// layer's MMAs read the same tile, so the 64 KB of stores per tile-layer leave the epilogue's critical path (issued by
// the epilogue warps themselves they stretched it from ~2900 to ~4700 cycles).  20 warps keep the 96-register budget.
constexpr int FWD_THREADS = 576, FWD_THREADS_TRAIN = 640;
constexpr int W_PROD = 16, W_MMA = 17, W_STASH0 = 18;

__constant__ float cC0 = 0.28209479177387814f;
__constant__ float cC1 = 0.4886025119029199f;
__constant__ float cC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// Encode one row (sample) into 64 bf16 features and store them as 8 planes of the enc tile image
// (and optionally to the global stash image).  L = 10 frequencies.
__device__ __forceinline__ void encode_row(const FwdArgs& a, int row_g, bool valid, uint32_t enc_smem, int q,
                                           uint8_t* stash_enc_tile, const float* bw_s) {
  float f[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) f[i] = 0.f;
  if (valid) {
    if (a.x_enc) {
      const float* src = a.x_enc + (size_t)row_g * a.ld_enc;
#pragma unroll
      for (int i = 0; i < 63; ++i) f[i] = src[i];
    } else {
      int flat = a.sel_idx ? a.sel_idx[row_g] : row_g;
      int ray = flat / a.smp.S, k = flat - ray * a.smp.S;
      float z = linspace_f(a.smp.near_, a.smp.far_, a.smp.S, k) + (a.jitter ? a.jitter[ray] : 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float xc = a.rays_o[3 * ray + c] + a.rays_d[3 * ray + c] * z;
        f[c] = xc;
        float sn, cs;
        sincosf(xc, &sn, &cs);
#pragma unroll
        for (int kf = 0; kf < 10; ++kf) {
          const float bw = bw_s[kf];
          f[3 + c * 20 + kf] = sn * bw;
          f[3 + c * 20 + 10 + kf] = cs * bw;
          float s2 = 2.f * sn * cs;            // angle doubling: error grows 2x per octave (<= 5e-5 at 2^9),
          float c2 = 1.f - 2.f * sn * sn;      // far below the bf16 rounding of the feature
          sn = s2;
          cs = c2;
        }
      }
    }
  }
#pragma unroll
  for (int kg = 0; kg < 8; ++kg) {
    uint32_t w0 = tc::pack_bf16(f[kg * 8 + 0], f[kg * 8 + 1]), w1 = tc::pack_bf16(f[kg * 8 + 2], f[kg * 8 + 3]);
    uint32_t w2 = tc::pack_bf16(f[kg * 8 + 4], f[kg * 8 + 5]), w3 = tc::pack_bf16(f[kg * 8 + 6], f[kg * 8 + 7]);
    st_shared_v4(enc_smem + kg * PLANE + q * 16, w0, w1, w2, w3);
    if (stash_enc_tile) *reinterpret_cast<uint4*>(stash_enc_tile + stash_off(q, kg, 8)) = make_uint4(w0, w1, w2, w3);
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(TRAIN ? FWD_THREADS_TRAIN : FWD_THREADS, 1) mlp_tc_fwd_k(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                                   // [2][ACT_BYTES]
  uint8_t* enc = smem + 2 * ACT_BYTES;                   // [2][ENC_BYTES]
  uint8_t* wst = enc + 2 * ENC_BYTES;                    // [FSTAGE][STAGE_BYTES]
  uint8_t* ones = wst + FSTAGE * STAGE_BYTES;            // broadcast ones operand of the bias MMAs
  float* w2s = reinterpret_cast<float*>(ones + ONES_BYTES);   // w_sigma2[256], b_sigma2
  SmemBars* bars = reinterpret_cast<SmemBars*>(w2s + W2_FLOATS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int n_tiles = (rows + TM - 1) / TM;
  const int n_pairs = (n_tiles + 1) / 2;
  const int n_steps = a.plan.n_steps;
  // Both CTAs of a pair run the same number of tile-pair iterations (surplus iterations work on an all-invalid tile).
  const uint32_t crank = tc::cluster_ctarank();
  const int n_iter = (n_pairs + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    // leader: a stage is full when its own half has landed (expect_tx arrive) AND the peer relayed the same for its half
    for (int i = 0; i < FSTAGE; ++i) { tc::mbar_init(&bars->w_full[i], crank == 0 ? 2 : 1); tc::mbar_init(&bars->w_empty[i], 1); }
    // a_ready (leader's copy is the one waited on): one arrive per epilogue warp of BOTH CTAs
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->a_ready[i], 32); tc::mbar_init(&bars->acc_full[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->st_full[i], 16); tc::mbar_init(&bars->st_done[i], 2); }
    tc::mbar_init_fence();
  }
  if (warp == W_MMA) tc::tmem_alloc2(&bars->tmem_base, 512);
  for (int i = tid; i < 257; i += blockDim.x) w2s[i] = a.bias[a.sig2_off + i];
  // BARF band weights: from the device buffer when given (CUDA-graph replays follow a moving window), else the
  // launch-time values; kept in shared memory so that the input stage pays no global round trip per row
  if (tid < 10) w2s[264 + tid] = a.smp.band_w_dev ? a.smp.band_w_dev[tid] : a.smp.band_w[tid];
  if (tid < 128)   // core matrix 0: 8 rows x (1, 1, 0, 0, 0, 0, 0, 0); core matrix 1: zeros
    reinterpret_cast<__nv_bfloat16*>(ones)[tid] = __float2bfloat16((tid < 64 && (tid & 7) < 2) ? 1.f : 0.f);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();            // peer barriers are initialised before anything arrives on them
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ weight producer (this CTA's half chunks)
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < n_steps; ++s) {
          const Step& st = a.plan.s[s];
          const uint32_t half = (uint32_t)(st.N / 2) * KC2 * 2, bias_half = (uint32_t)(st.N / 2) * BIAS_K * 2;
          const int nc2 = st.n_chunks * KC / KC2;
          const uint8_t* src = a.wpack + st.w_off;
          for (int t = 0; t < 2; ++t)
            for (int c = 0; c <= nc2; ++c) {                        // c == nc2: the bias block
              const uint32_t bytes = c < nc2 ? half : bias_half;
              tc::mbar_wait(&bars->w_empty[stage], par ^ 1);       // the pair's MMAs are done with this stage
              tc::mbar_arrive_expect_tx(&bars->w_full[stage], bytes);
              tc::bulk_g2s(wst + stage * STAGE_BYTES, src + (size_t)c * 2 * half + (size_t)crank * bytes, bytes,
                           &bars->w_full[stage]);
              if (++stage == FSTAGE) { stage = 0; par ^= 1; }
            }
        }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (leader) / relay (peer)
    // tcgen05.mma is asynchronous: the tensor pipe stays busy only if this one thread needs fewer cycles per MMA
    // than the MMA takes (128 for 128x256x16 per SM).  So everything per step is hoisted, descriptors advance by
    // one 32-bit add, and the loop body is: wait, 4 MMAs, commit.
    if (lane == 0 && crank == 0) {
      int stage = 0;
      uint32_t par = 0, apar = 0;
      const uint32_t hi = tc::umma_desc_hi(128), ones_hi = tc::umma_desc_hi(0);
      const uint32_t act_lo = tc::umma_desc_lo(tc::smem_u32(act), PLANE), enc_lo = tc::umma_desc_lo(tc::smem_u32(enc), PLANE);
      const uint32_t ones_lo = tc::umma_desc_lo(tc::smem_u32(ones), 128);
      const uint32_t wst_addr = tc::smem_u32(wst);
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]), empty0 = tc::smem_u32(&bars->w_empty[0]);
      const uint32_t ardy0 = tc::smem_u32(&bars->a_ready[0]), accf0 = tc::smem_u32(&bars->acc_full[0]);
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < n_steps; ++s) {
          const int N = a.plan.s[s].N, nc2 = a.plan.s[s].n_chunks * KC / KC2, a_src = a.plan.s[s].a_src;
          const uint32_t idesc = tc::umma_idesc_bf16(2 * TM, N);
          const uint32_t b_lo0 = tc::umma_desc_lo(wst_addr, (N / 2) * 16), b_inc = (2u * (N / 2) * 16) >> 4;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            tc::mbar_wait_addr(ardy0 + t * 8, (apar >> t) & 1);        // slot t of both CTAs holds the A operand
            MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2) a.dbg[(0 + t) * 32 + s] = clock64();)
            apar ^= 1u << t;
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem + t * 256;
            const uint32_t act_t = act_lo + t * (ACT_BYTES >> 4), enc_t = enc_lo + t * (ENC_BYTES >> 4);
            uint32_t a_lo = (a_src == A_ACT) ? act_t : enc_t;
            // the skip layer reads the encoding tile first (exactly one 64-column chunk), then the activation tile
            const int switch_c = (a_src == A_ENC_ACT) ? ENCW / KC2 : -1;
            for (int c = 0; c < nc2; ++c) {
              MC_TRACE(long long tw0 = a.dbg ? clock64() : 0;)
              tc::mbar_wait_addr(full0 + stage * 8, par);
              MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2 && t == 0) a.dbg[7 * 32 + 16 + s] += clock64() - tw0;)
              tc::tcgen05_fence_after();
              if (c == switch_c) a_lo = act_t;
              const uint32_t b_lo = b_lo0 + stage * (STAGE_BYTES >> 4);
#pragma unroll
              for (int j = 0; j < KC2 / 16; ++j)
                tc::umma2_bf16_w(d_tmem, a_lo + j * ((2 * PLANE) >> 4), hi, b_lo + j * b_inc, hi, idesc, (c | j) != 0);
              tc::umma2_commit_multicast_addr(empty0 + stage * 8, (uint16_t)3);
              a_lo += ((KC2 / 8) * PLANE) >> 4;
              if (++stage == FSTAGE) { stage = 0; par ^= 1; }
            }
            // + bias: ones[128 x 16] * bias_block[N x 16]^T
            tc::mbar_wait_addr(full0 + stage * 8, par);
            tc::tcgen05_fence_after();
            tc::umma2_bf16_w(d_tmem, ones_lo, ones_hi, b_lo0 + stage * (STAGE_BYTES >> 4), hi, idesc, true);
            tc::umma2_commit_multicast_addr(empty0 + stage * 8, (uint16_t)3);
            if (++stage == FSTAGE) { stage = 0; par ^= 1; }
            tc::umma2_commit_multicast_addr(accf0 + t * 8, (uint16_t)3);
            MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2) a.dbg[(2 + t) * 32 + s] = clock64();)
          }
        }
    } else if (lane == 0) {
      // peer CTA: tell the leader when this CTA's half of each stage has landed
      int stage = 0;
      uint32_t par = 0;
      const uint32_t full0 = tc::smem_u32(&bars->w_full[0]);
      const uint32_t leader_full0 = tc::mapa(full0, 0);
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < n_steps; ++s) {
          const int n = 2 * (a.plan.s[s].n_chunks * KC / KC2 + 1);
          for (int c = 0; c < n; ++c) {
            tc::mbar_wait_addr(full0 + stage * 8, par);
            tc::mbar_arrive_remote(leader_full0 + stage * 8);
            if (++stage == FSTAGE) { stage = 0; par ^= 1; }
          }
        }
    }
  } else if (TRAIN && warp >= W_STASH0) {
    // ------------------------------------------------------------------ stash warps: activation tile smem -> HBM
    // warp-item = (plane p, 32-row block rb): 512 contiguous bytes on both sides; the two warps split the 128 items
    const int sw = warp - W_STASH0;
    uint32_t spar[2] = {0, 0};
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      for (int s = 0; s < n_steps; ++s) {
        const Step& st = a.plan.s[s];
        if (st.epi != EPI_RELU || st.stash_slot < 0) continue;
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t;
          tc::mbar_wait(&bars->st_full[t], spar[t]);
          spar[t] ^= 1;
          if (tile < n_tiles) {
            uint8_t* dst = a.stash + ((size_t)tile * a.n_slots + st.stash_slot) * ACT_BYTES;
            const uint32_t src = tc::smem_u32(act + t * ACT_BYTES);
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
      }
    }
  } else {
    // ------------------------------------------------------------------ input stage + epilogues (both slots)
    const int lq = warp & 3, cq = warp >> 2;          // TMEM lane quarter, accumulator column quarter
    const int q = lq * 32 + lane;                     // row in tile == TMEM lane
    const uint32_t act0 = tc::smem_u32(act), enc0 = tc::smem_u32(enc);
    const uint32_t a_ready_leader = tc::mapa(tc::smem_u32(&bars->a_ready[0]), 0);
    uint32_t par = 0, stpar = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      // Warps 8-11 encode the rows of slot 0, warps 12-15 those of slot 1: these column quarters have nothing to do in
      // the previous iteration's last epilogue (sh.2 -> out4 is done by column quarter 0 alone), so they run ahead and
      // the encoding (index -> ray -> sin/cos, ~3 k cycles) overlaps it instead of following it.  The encoding tile
      // was last read by the skip layer's MMAs, long complete.
      if (cq >= 2) {
        const int es = cq - 2;
        const int tile = 2 * pair + es, row_g = tile * TM + q;
        uint8_t* st_enc = (TRAIN && tile < n_tiles) ? a.stash_enc + (size_t)tile * ENC_BYTES : nullptr;
        encode_row(a, row_g, tile < n_tiles && row_g < rows, enc0 + es * ENC_BYTES, q, st_enc, w2s + 264);
      }
      tc::fence_proxy_async();
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive_remote(a_ready_leader); tc::mbar_arrive_remote(a_ready_leader + 8); }
      float sig_dot[2] = {0.f, 0.f};                  // this warp's part of the sigma.2 dot product, per slot
      for (int s = 0; s < n_steps; ++s) {
        const Step& st = a.plan.s[s];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t, row_g = tile * TM + q;
          const bool valid = tile < n_tiles && row_g < rows;
          const uint32_t act_t = act0 + t * ACT_BYTES;
          const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + t * 256;
          tc::mbar_wait(&bars->acc_full[t], par);
          MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2 && warp == 0 && lane == 0) a.dbg[(4 + t) * 32 + s] = clock64();)
          tc::tcgen05_fence_after();
          if (st.epi == EPI_OUT) {
            // sigma.2: gather the four column-quarter partial sums through the (now idle) activation tile.
            // Training: the tile still holds relu(sh.0) until the stash warps have copied it - without this wait the
            // partial sums could land in the first 1.5 KB of the stashed tile (features 0..7 of rows 0..95: found with
            // compute-sanitizer, whose slowdown of the stash warps made the race visible as NaN sh.2 weight gradients).
            // Same parity as the next ReLU step's wait for this copy; not toggled here.
            if (TRAIN) tc::mbar_wait(&bars->st_done[t], (stpar >> t & 1) ^ 1);
            float* part = reinterpret_cast<float*>(act + t * ACT_BYTES);
            if (cq != 0) part[(cq - 1) * 128 + q] = sig_dot[t];
            named_bar_sync(1 + t * 4 + lq, 128);       // the 4 warps of this lane quarter
            if (cq != 0) continue;                     // last step: nothing to arrive on
            uint32_t v[32];
            tc::tmem_ld32(taddr, v);
            tc::tmem_ld_wait();
            const float sigma_raw = sig_dot[t] + part[q] + part[128 + q] + part[256 + q] + w2s[256];
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
              for (int ch = 0; ch < 3; ++ch) {
                float acc = 0.f;
#pragma unroll
                for (int b = 0; b < 9; ++b) acc += Y[b] * sh[9 * ch + b];
                c[ch] = sigmoid_f(acc);
              }
              reinterpret_cast<float4*>(a.out4)[row_g] = make_float4(sigma_raw, c[0], c[1], c[2]);
              if (TRAIN) {
                // per tile [float4 group i][row]: a warp's 32 rows write 512 contiguous bytes per store
                float4* dst = reinterpret_cast<float4*>(a.stash_sh) + (size_t)(row_g >> 7) * (TM * SH_LD / 4) + (row_g & (TM - 1));
#pragma unroll
                for (int i = 0; i < 7; ++i)
                  dst[i * TM] = make_float4(sh[4 * i], sh[4 * i + 1], sh[4 * i + 2], i < 6 ? sh[4 * i + 3] : 0.f);
              }
            }
            continue;
          }
          uint8_t* st_tile = (TRAIN && st.stash_slot >= 0 && tile < n_tiles)
                                 ? a.stash + ((size_t)tile * a.n_slots + st.stash_slot) * ACT_BYTES
                                 : nullptr;
          const bool to_smem = st.epi == EPI_RELU;
          const bool direct_stash = TRAIN && !to_smem;          // sigma.0: its tile never exists in shared memory
          if (TRAIN && to_smem) {                                // the stash warps are done copying the previous tile
            tc::mbar_wait(&bars->st_done[t], (stpar >> t & 1) ^ 1);
            stpar ^= 1u << t;
          }
          float dot = 0.f;
          // gate words are stored [32-column block w][row]: a warp's 32 rows write one contiguous 128-byte line
          uint32_t* gate_out = st_tile ? reinterpret_cast<uint32_t*>(a.stash_bits + ((size_t)tile * a.n_slots + st.stash_slot) * BITS_BYTES) + q
                                       : nullptr;
          // 16 accumulator columns (bias already in them): ReLU, bf16 -> next A operand (smem) / stash / sigma dot.
          // This warp owns columns [64 cq, 64 cq + 64) = 32-column blocks 2cq and 2cq+1, walked in four halves with
          // the TMEM load of the next half in flight while one is processed.
          uint32_t sbits = 0;
          auto half = [&](const uint32_t (&v)[16], int h) {
            const int col0 = cq * 64 + h * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint32_t w[4];
              if (!to_smem) {       // sigma.0: the fp32 activations feed the sigma.2 dot product
                const float4 s0 = *reinterpret_cast<const float4*>(w2s + col0 + j * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(w2s + col0 + j * 8 + 4);
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[j * 8 + i]), 0.f);
                dot += x[0] * s0.x + x[1] * s0.y + x[2] * s0.z + x[3] * s0.w + x[4] * s1.x + x[5] * s1.y + x[6] * s1.z +
                       x[7] * s1.w;
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = tc::pack_bf16(x[2 * i], x[2 * i + 1]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  w[i] = tc::pack_bf16_relu(__uint_as_float(v[j * 8 + 2 * i]), __uint_as_float(v[j * 8 + 2 * i + 1]));
              }
              const int kg = col0 / 8 + j;
              if (to_smem) st_shared_v4(act_t + kg * PLANE + q * 16, w[0], w[1], w[2], w[3]);
              if (direct_stash && st_tile) *reinterpret_cast<uint4*>(st_tile + stash_off(q, kg, 32)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (TRAIN) {
              sbits |= tc::sign_bits16(v, h & 1);
              if (h & 1) {
                if (gate_out) gate_out[(2 * cq + (h >> 1)) * TM] = ~sbits;     // gate = accumulator > 0
                sbits = 0;
              }
            }
          };
          EPI_TRACE(0)
          {
            uint32_t va[16], vb[16];
            const uint32_t tcol = taddr + cq * 64;
            tc::tmem_ld16(tcol, va);
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 16, vb);
            EPI_TRACE(1)
            half(va, 0);
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 32, va);
            half(vb, 1);
            EPI_TRACE(2)
            tc::tmem_ld_wait();
            tc::tmem_ld16(tcol + 48, vb);
            half(va, 2);
            EPI_TRACE(3)
            tc::tmem_ld_wait();
            half(vb, 3);
            EPI_TRACE(4)
          }
          if (!to_smem) sig_dot[t] = dot;
          tc::fence_proxy_async();
          EPI_TRACE(5)
          tc::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive_remote(a_ready_leader + t * 8);
            if (TRAIN && to_smem) tc::mbar_arrive(&bars->st_full[t]);
          }
          EPI_TRACE(6)
          MC_TRACE(if (a.dbg && blockIdx.x == 0 && it == 2 && warp == 0 && lane == 0) a.dbg[(6 + t) * 32 + s] = clock64();)
        }
        par ^= 1;
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::cluster_sync();            // no CTA leaves while its peer may still arrive on its barriers / read its operands
  if (warp == W_MMA) tc::tmem_dealloc2(tmem, 512);
}

}  // namespace mlptc

#include "mlp_tc_fwd2.cuh"

using namespace mlptc;

extern "C" int mcnerf_mlp_tc_supported(const mcnerf_mlp_params* p) {
  if (!p) return 0;
  if (p->width != WID || p->in_ch != 63 || p->sh_dim != 27 || p->depth < 2 || p->depth > 12) return 0;
  if (__builtin_popcount(p->skip_mask) > 1 || (p->skip_mask & 1u) || (p->skip_mask >> p->depth)) return 0;
  return 1;
}

extern "C" int mcnerf_mlp_tc_pack_sizes(const mcnerf_mlp_params* p, size_t* wf_bytes, size_t* wb_bytes,
                                        size_t* bias_bytes) {
  PackLayout L;
  if (int e = build_layout(p, &L)) return e;
  if (wf_bytes) *wf_bytes = (fwd_v2_mode() > 0 ? 2 : 1) * L.wf_bytes;      // second layout only when that kernel is enabled
  if (wb_bytes) *wb_bytes = L.wb_bytes;
  if (bias_bytes) *bias_bytes = (size_t)L.bias_floats * sizeof(float);
  return 0;
}

// fp32 reference parameters -> bf16 UMMA-ready images (forward: B(n=out,k=in); dgrad: B(n=in,k=out)) + fp32 bias block
extern "C" int mcnerf_mlp_tc_pack(const mcnerf_mlp_params* p, void* wf, void* wb, float* bias, void* stream) {
  PackLayout L;
  if (int e = build_layout(p, &L)) return e;
  MC_ARG(wf && bias);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = p->depth;
  PackArgs a;
  a.wf = (uint8_t*)wf; a.wb = (uint8_t*)wb; a.bias = bias;
  int nj = 0;
  int64_t tot = 0;
  auto add = [&](const float* src, int64_t sn, int64_t sk, int N, int K, int pad_k, int pad_n, int n_valid, int k_valid,
                 int is_bias, size_t dst_off, int dst_sel) {
    PackJob& j = a.j[nj];
    j.src = src; j.sn = sn; j.sk = sk; j.N = N; j.K = K; j.pad_k = pad_k; j.pad_n = pad_n; j.n_valid = n_valid;
    j.k_valid = k_valid; j.is_bias = is_bias; j.dst_off = dst_off; j.dst_sel = dst_sel; j.pair = 1; j.kw = 0;
    if (dst_sel == 3) { j.dst_sel = 0; j.pair = 2; }
    if (dst_sel == 4) { j.dst_sel = 0; j.pair = 3; }
    if (dst_sel == 5) { j.dst_sel = 0; j.pair = 4; }
    a.prefix[nj] = tot;
    tot += is_bias ? (int64_t)N * K : (int64_t)N * K / 8;      // work units (see pack_all_k)
    ++nj;
  };
  for (int s = 0; s < L.fwd.n_steps; ++s) {
    const Step& sp = L.fwd.s[s];
    const float* W;
    const float* bs;
    int n_out = WID, k_in, ld, pad = 0;
    if (s < D) {
      W = p->W[s]; bs = p->b[s];
      if (s == 0) { k_in = 63; ld = 63; pad = 1; }
      else if (p->skip_mask >> s & 1u) { k_in = 63 + WID; ld = 63 + WID; pad = 1; }
      else { k_in = WID; ld = WID; }
    } else if (s == D) { W = p->W_sigma0; bs = p->b_sigma0; k_in = WID; ld = WID; }
    else if (s == D + 1) { W = p->W_sh0; bs = p->b_sh0; k_in = WID; ld = WID; }
    else { W = p->W_sh2; bs = p->b_sh2; k_in = WID; ld = WID; n_out = 27; }
    const int K = sp.n_chunks * KC;
    // forward image: rows n = output feature, reduction k = input feature (padded)
    add(W, ld, 1, sp.N, K, pad, 0, n_out, k_in, 0, sp.w_off, 0);
    // dgrad images: rows n = input feature, reduction k = output feature (sp.N, zero beyond n_out)
    if (wb) {
      if (sp.a_src == A_ENC) add(W, 1, ld, ENCW, sp.N, 0, 1, 63, n_out, 0, L.wb_main[s], 1);
      else if (sp.a_src == A_ENC_ACT) {
        add(W, 1, ld, ENCW, sp.N, 0, 1, 63, n_out, 0, L.wb_enc[s], 1);
        add(W + 63, 1, ld, WID, sp.N, 0, 0, WID, n_out, 0, L.wb_main[s], 1);
      } else add(W, 1, ld, WID, sp.N, 0, 0, WID, n_out, 0, L.wb_main[s], 1);
    }
    add(bs, 0, 0, 256, 1, 0, 0, n_out, 0, 1, sp.bias_off, 2);
    add(bs, 0, 0, sp.N, BIAS_K, 0, 0, n_out, BIAS_K, 0, sp.w_off + (size_t)sp.N * K * 2, 3);   // bias block of the forward image
    if (fwd_v2_mode() > 0) {
      // the same matrix and bias block in the layout of the second-generation kernel, behind the first image
      add(W, ld, 1, sp.N, K, pad, 0, n_out, k_in, 0, L.wf_bytes + sp.w_off, 4);
      add(bs, 0, 0, sp.N, BIAS_K, 0, 0, n_out, BIAS_K, 0, L.wf_bytes + sp.w_off, 5);
      a.j[nj - 1].kw = K;
    }
  }
  add(p->W_sigma2, 0, 0, 256, 1, 0, 0, WID, 0, 1, L.sig2_off, 2);
  add(p->b_sigma2, 0, 0, 8, 1, 0, 0, 1, 0, 1, L.sig2_off + 256, 2);
  MC_ARG(nj <= MAX_PACK_JOBS);
  a.prefix[nj] = tot;
  a.n_jobs = nj;
  pack_all_k<<<cdiv(tot, 256), 256, 0, st>>>(a);
  MC_LAUNCHED();
  return 0;
}

extern "C" size_t mcnerf_mlp_tc_stash_bytes(const mcnerf_mlp_params* p, int n_rows) {
  if (!mcnerf_mlp_tc_supported(p) || n_rows <= 0) return 0;
  size_t tiles = (size_t)(n_rows + TM - 1) / TM;
  tiles += tiles & 1;   // tiles are processed in pairs
  return tiles * ((size_t)(p->depth + 2) * (ACT_BYTES + BITS_BYTES) + ENC_BYTES + (size_t)TM * SH_LD * sizeof(float));
}

extern "C" int mcnerf_mlp_tc_fwd(const mcnerf_mlp_params* p, const void* wf, const float* bias,
                                 const mcnerf_tc_input* in, float* out4, void* stash, void* stream) {
  PackLayout L;
  if (int e = build_layout(p, &L)) return e;
  MC_ARG(wf && bias && in && out4 && in->n_rows >= 0 && ((uintptr_t)out4 & 15) == 0);
  if (in->n_rows == 0) return 0;
  MC_ARG((in->x_enc && in->dirs_rows && in->ld_enc >= 63) ||
         (in->rays_o && in->rays_d && in->smp.S >= 2 && in->smp.n_freqs == 10));
  FwdArgs a;
  a.plan = L.fwd;
  a.wpack = (const uint8_t*)wf;
  a.bias = bias;
  a.sig2_off = L.sig2_off;
  a.bias_floats = L.bias_floats;
  a.dbg = nullptr;
#ifdef MCNERF_TC_TRACE
  if (const char* dbg = getenv("MCNERF_TC_DEBUG")) {
    static long long* dbg_buf = nullptr;
    if (atoi(dbg) == 7) {
      if (!dbg_buf) cudaMalloc(&dbg_buf, 10 * 32 * sizeof(long long));
      cudaStreamSynchronize((cudaStream_t)stream);
      cudaMemsetAsync(dbg_buf, 0, 10 * 32 * sizeof(long long), (cudaStream_t)stream);
      a.dbg = dbg_buf;
    }
  }
#endif
  a.rays_o = in->rays_o; a.rays_d = in->rays_d; a.jitter = in->jitter; a.smp = in->smp;
  a.sel_idx = in->sample_idx; a.n_rows = in->n_rows; a.n_rows_dev = in->n_rows_dev;
  a.x_enc = in->x_enc; a.ld_enc = in->ld_enc; a.dirs_rows = in->dirs_rows;
  a.out4 = out4;
  a.n_slots = p->depth + 2;
  if (stash) {
    size_t tiles = (size_t)(in->n_rows + TM - 1) / TM;
    tiles += tiles & 1;
    a.stash = (uint8_t*)stash;
    a.stash_enc = a.stash + tiles * (size_t)a.n_slots * ACT_BYTES;
    a.stash_sh = (float*)(a.stash_enc + tiles * ENC_BYTES);
    a.stash_bits = (uint8_t*)(a.stash_sh + tiles * (size_t)TM * SH_LD);
  } else {
    a.stash = nullptr; a.stash_enc = nullptr; a.stash_sh = nullptr; a.stash_bits = nullptr;
  }
  // per device / context attribute: set on every call (cheap), not once per process
  const bool v2 = fwd_v2_mode() >= (stash ? 1 : 2);
  const int smem_bytes = v2 ? smem_fwd2(stash != nullptr) : SMEM_FWD;
  if (v2) {
    a.wpack += L.wf_bytes;
    MC_CUDA(cudaFuncSetAttribute(mlp_tc_fwd2_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd2(true)));
    MC_CUDA(cudaFuncSetAttribute(mlp_tc_fwd2_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd2(false)));
  } else {
    MC_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
    MC_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
  }
  int n_pairs = ((in->n_rows + TM - 1) / TM + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = n_pairs < sms ? n_pairs : sms;
  grid = (grid + 1) & ~1;                                  // whole clusters of 2
  if (grid > sms) grid = sms & ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(v2 ? (stash ? F2_THREADS_TRAIN : F2_THREADS) : stash ? FWD_THREADS_TRAIN : FWD_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (v2) {
    if (stash) MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_fwd2_k<true>, a));
    else MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_fwd2_k<false>, a));
  } else if (stash) MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_fwd_k<true>, a));
  else MC_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_fwd_k<false>, a));
  MC_LAUNCHED();
  if (a.dbg && v2) {     // MMA issuer of CTA 0, iterations >= 1: where it waits
    cudaStreamSynchronize((cudaStream_t)stream);
    long long h[10 * 32];
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const double np = (double)h[3];
    printf("fwd2 trace (CTA 0, iterations >= 1): passes %lld  span %lld cycles = %.0f per pass;  issuer waits per pass: enc %.0f  "
           "epilogue %.0f  weights %.0f;  epilogue duration warp 0: %.0f  warp 15: %.0f cycles per pass\n", h[3], h[5] - h[4],
           (h[5] - h[4]) / np, h[0] / np, h[1] / np, h[2] / np, h[7] ? (double)h[6] / h[7] : 0., h[9] ? (double)h[8] / h[9] : 0.);
    for (int w = 0; w < 2; ++w)
      for (int hh = 0; hh < 2; ++hh) {
        printf("  step 2 slot %d pass %d epilogue (cycles after wake: before ld / after wait, x4 groups; arrived):", w, hh);
        for (int k = 0; k < 9; ++k) printf(" %lld", h[128 + w * 32 + hh * 16 + k]);
        printf("\n");
      }
  } else if (a.dbg) {     // debug trace: clock64 deltas of CTA 0's third tile pair (steady state)
    cudaStreamSynchronize((cudaStream_t)stream);
    long long h[10 * 32];
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[8] = {"mma:a_rdy0", "mma:a_rdy1", "mma:iss0", "mma:iss1", "epi:full0", "epi:full1",
                            "epi:arr0", "epi:arr1"};
    long long t0 = h[0];
    for (int s = 0; s < L.fwd.n_steps; ++s) {
      printf("step %2d:", s);
      for (int e = 0; e < 8; ++e) printf(" %s=%lld", names[e], h[e * 32 + s] ? h[e * 32 + s] - t0 : -1);
      printf(" wait_w0=%lld\n", h[7 * 32 + 16 + s]);
    }
    for (int w = 0; w < 2; ++w)
      for (int t = 0; t < 2; ++t) {
        printf("step 2 epilogue warp %d slot %d (woke, ld0 done, blk0 done, ld1 done, blk1 done, proxy fence, arrived):", w ? 15 : 0, t);
        for (int k = 0; k < 7; ++k) printf(" %lld", h[8 * 32 + w * 16 + t * 8 + k] ? h[8 * 32 + w * 16 + t * 8 + k] - t0 : -1);
        printf("\n");
      }
  }
  return 0;
}


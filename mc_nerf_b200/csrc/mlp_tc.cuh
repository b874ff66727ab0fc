// Shared definitions of the bf16 tcgen05 NeRF-MLP kernels (forward, backward chain, weight gradients).
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace mlptc {

constexpr int TM = 128;                      // samples (rows) per tile = UMMA M
constexpr int WID = 256;                     // hidden width this path is specialised for
constexpr int ENCW = 64;                     // encoding width padded 63 -> 64 (column 63 is zero)
constexpr int KC = 32;                       // reduction elements per streamed weight chunk
constexpr int NSTAGE = 4;                    // weight ring depth
constexpr int STAGE_BYTES = WID * KC * 2;    // 16 KB
constexpr int BIAS_K = 16;                   // forward images: reduction depth of the bias block appended to every step
constexpr int KC2 = 64;                      // forward (CTA-pair) kernel: reduction elements per half chunk (N/2 rows x 64 k)
constexpr int ACT_BYTES = TM * WID * 2;      // 64 KB: one activation tile image
constexpr int ENC_BYTES = TM * ENCW * 2;     // 16 KB
constexpr int PLANE = TM * 16;               // 2 KB: one k-group plane [128 rows x 16 B] of a tile image
constexpr int MAX_STEPS = 20;
constexpr int SH_LD = 32;                    // stashed raw SH coefficients per row (27 used)

// Tile image ("canonical SWIZZLE_NONE layout"): element (row r, feature k) of a [128 x K] bf16 tile lives at
//   (k/8)*PLANE + r*16 + (k%8)*2.
// As a K-major UMMA operand:  LBO (next k-group) = PLANE, SBO (next 8 rows) = 128.
// As an MN-major operand (reduction over rows): LBO (next 8 rows) = 128, SBO (next 8 features) = PLANE.

enum { A_ENC = 0, A_ACT = 1, A_ENC_ACT = 2, A_SMALL = 3 };
enum { EPI_RELU = 0, EPI_SIGMA = 1, EPI_OUT = 2 };

struct Step {
  int a_src;          // which smem tile(s) feed the A operand
  int n_chunks;       // K / KC
  int N;              // output width of this GEMM (multiple of 16, <= 256)
  int epi;            // epilogue kind
  uint32_t w_off;     // byte offset of chunk 0 in the packed weight image (chunk c at w_off + c*N*KC*2)
  int bias_off;       // float offset into the bias block
  int stash_slot;     // activation stash slot written by the epilogue (-1: none)
};

struct Plan {
  int n_steps;
  Step s[MAX_STEPS];
};

// ---- backward chain (dgrad) jobs: dX = dY * W, chained from the heads down to the encoding
enum { BK_MASK_STORE = 0, BK_SIGMA_INJECT = 1, BK_RELOAD_SKIP = 2, BK_ENC_OUT = 3 };
struct BJob {
  int a_small;        // 1: A operand is the 128x32 head-gradient tile, 0: the 128x256 dY tile
  int n_chunks;       // K / KC (K = number of output features of the layer being back-propagated through)
  int N;              // width of dX (256, or 64 for the encoding part)
  int accumulate;     // 1: add onto the accumulator left by the previous job
  int kind;           // epilogue kind
  uint32_t w_off;     // byte offset in the transposed weight image
  int mask_slot;      // forward-stash slot whose ">0" pattern gates dX (ReLU backward)
  int dy_slot;        // dY-stash slot the epilogue writes (-1: none)
};
struct BPlan {
  int n_jobs;
  int skip_dy_slot;   // dY slot of the skip layer (reloaded for the encoding part), -1 if no skip
  BJob j[MAX_STEPS];
};

// ---- weight-gradient jobs: dW = dY^T X over all rows
struct WJob {
  int dy_slot;        // dY-stash slot (A operand, M = output features, two halves of 128) ; -1: head tile
  int x_slot;         // forward-stash slot of X (B operand, N = input features); -1: encoding stash
  int N;              // input features of this job (256, or 64 for the encoding)
  int transposed;     // 1: roles swapped (A = X^T, B = head dY tile): computes dW^T (sh.2 / sigma.2)
  int head_col0;      // transposed jobs: first column of the head tile used as B, and N columns from there
  int which;          // parameter id: 0..depth-1 trunk, depth sigma.0, depth+1 sh.0, depth+2 sh.2, depth+3 sigma.2
  int col_off;        // column offset inside the parameter's [out,in] matrix (63 for the h part of a skip layer)
  int n_valid;        // valid input columns (63 for encoding parts, else 256)
  int bias_mode;      // 0: none; 1: column sums of the A operand (dY) -> bias grad of `which`;
                      // 2: B operand = head tile cols 0..26 -> sh.2 bias; 3: head tile col 31 -> sigma.2 bias
};
struct WPlan {
  int n_jobs;
  WJob j[MAX_STEPS];
};

// host-side description of one packed network (built by mcnerf_mlp_tc_pack)
struct PackLayout {
  int depth;
  uint32_t skip_mask;
  Plan fwd;
  BPlan bwd;
  WPlan wg;
  size_t wf_bytes;          // forward weight image
  size_t wb_bytes;          // transposed (dgrad) weight image
  int bias_floats;          // bias block incl. w_sigma2 / b_sigma2
  int sig2_off;             // float offset of w_sigma2[256] (b_sigma2 follows)
  uint32_t wb_main[MAX_STEPS];  // dgrad image of fwd step s: rows = hidden inputs (256) [or enc inputs for step 0]
  uint32_t wb_enc[MAX_STEPS];   // skip step: rows = encoding inputs (64)
};

// stash geometry (per tile; tiles are processed in pairs so the tile count is rounded up to even)
__host__ __device__ inline size_t stash_tiles(int n_rows) {
  size_t t = (size_t)(n_rows + TM - 1) / TM;
  return t + (t & 1);
}
constexpr int HEAD_BYTES = TM * 32 * 2;      // 8 KB: head-gradient tile [128 x 32] bf16 (g_sh 0..26, g_sigma at 31)
constexpr int BITS_BYTES = TM * 32;          // 4 KB: ReLU gate bits of one tile-layer [8 x 32-column block][128 rows] u32

// Tile images in HBM (activation stash, dY stash, head tile) are stored as two 64-row halves, each a contiguous
// [k-group plane][64 rows][16 B] block, so that the weight-gradient kernel fetches one half of ALL planes with a
// single large bulk copy.  (In shared memory the forward / backward-chain kernels keep whole 128-row planes.)
__host__ __device__ inline uint32_t stash_off(int row, int kg, int n_planes) {
  return (uint32_t)(row >> 6) * (uint32_t)(n_planes * 1024) + (uint32_t)kg * 1024u + (uint32_t)(row & 63) * 16u;
}

int build_layout(const mcnerf_mlp_params* p, PackLayout* L);

// Which forward kernel runs: the first generation (activation tile in shared memory, mlp_tc_fwd_k) or the second (A operand
// in tensor memory, mlp_tc_fwd2.cuh).  The packed forward image holds both layouts back to back (wf_bytes each).
// MCNERF_FWD_V2: 0 = never, 1 = training launches only, 2 = all launches.  Read once.
inline int fwd_v2_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCNERF_FWD_V2");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace mlptc

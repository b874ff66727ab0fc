// Shared definitions of the bf16 tcgen05 NeRF-MLP kernels (forward, backward chain, weight gradients).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace mlptc {

constexpr int TM = 128;                      // samples (rows) per tile = UMMA M
constexpr int WID = 256;                     // hidden width this path is specialised for
constexpr int ENCW = 64;                     // encoding width padded 63 -> 64 (column 63 is zero)
constexpr int KC = 32;                       // reduction elements per streamed weight chunk
constexpr int NSTAGE = 4;                    // weight ring depth
constexpr int STAGE_BYTES = WID * KC * 2;    // 16 KB
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

// host-side description of one packed network (built by mcnerf_mlp_tc_pack)
struct PackLayout {
  int depth;
  uint32_t skip_mask;
  Plan fwd;
  size_t wf_bytes;          // forward weight image
  size_t wb_bytes;          // transposed (dgrad) weight image
  int bias_floats;          // bias block incl. w_sigma2 / b_sigma2
  int sig2_off;             // float offset of w_sigma2[256] (b_sigma2 follows)
  uint32_t wb_off[MAX_STEPS];   // dgrad image offsets, indexed like fwd steps
};

}  // namespace mlptc

// sm_100a building blocks written as inline PTX: mbarrier, bulk async copy (TMA engine, UBLKCP),
// tensor memory (TMEM) management, tcgen05.mma (UMMA) descriptors and issue, tcgen05.ld.
// No CUTLASS/CuTe dependency.  Every spin-wait is bounded and traps instead of hanging the GPU.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
// MCNERF_CHAOS build (python -m mc_nerf_b200.build with MCNERF_CHAOS=1): a pseudo-random delay of up to 4 us after one
// in eight successful mbarrier waits and before one in eight arrivals, identical for the lanes of a warp.  The kernels'
// results must not depend on the relative timing of their warp roles; the parity and determinism tests are run on this
// build to check exactly that (profiles/r02_compute_sanitizer.txt).
#ifdef MCNERF_CHAOS
__device__ __forceinline__ void chaos_delay(uint32_t salt) {
  uint32_t h = ((uint32_t)clock64() * 2654435761u) ^ (salt * 0x9E3779B1u) ^ (blockIdx.x * 7919u + (threadIdx.x >> 5) * 104729u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  if ((h & 7u) == 0) __nanosleep(((h >> 8) & 0x7FFu) * 2);
}
#else
__device__ __forceinline__ void chaos_delay(uint32_t) {}
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  chaos_delay(1);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  chaos_delay(2);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
#ifdef MCNERF_SPIN_TEST_WAIT
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking phase test (polling loops that serve more than one barrier)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Address-based forms for the MMA issue loops (barrier address = base + stage * 8, no pointer re-conversion).
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  if (mbar_try_wait_addr(bar_addr, parity)) { chaos_delay(7 + parity); return; }
  long long t0 = clock64();
  while (!mbar_try_wait_addr(bar_addr, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mcnerf: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar_addr, parity);
      __trap();
    }
  }
  chaos_delay(9 + parity);
}
// Wait with acquire at cluster scope: the phase was (partly) completed by arrives from the peer CTA.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar_addr, uint32_t parity) {
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) {
      printf("mcnerf: cluster mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar_addr, parity);
      __trap();
    }
  }
}
// Bounded wait: ~2 s at 2 GHz, then trap (a wrong phase or a lost arrive must not hang the box).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) { chaos_delay(3 + parity); return; }
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mcnerf: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
  chaos_delay(5 + parity);
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ bulk async copies (1-D TMA)
// global -> shared, completion signalled on an mbarrier as transaction bytes.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// global -> the SAME shared-memory offset of every CTA in `cta_mask` (cluster multicast); each destination CTA's
// mbarrier at the same offset receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared -> global (bulk group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ TMEM
// One full warp calls alloc; the base address lands in *dst_smem.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// ---- CTA-pair (cta_group::2) forms.  One warp of EACH CTA of the pair executes alloc/dealloc; the leader CTA
// (cluster rank 0) issues the MMAs, which read A (128 rows) and half of B (N/2 rows) from each CTA's shared memory
// at the same offsets and write 128 accumulator lanes x N columns into each CTA's tensor memory.
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive (once) on the barrier at this shared-memory offset in every CTA of cta_mask when all MMAs issued so far
// by this thread have completed in both CTAs
__device__ __forceinline__ void umma2_commit_multicast_addr(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar_addr), "h"(cta_mask)
               : "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on a barrier that lives in another CTA of the cluster.  Default (CTA-scope) semantics on purpose: a
// cluster-scope acquire on the waiting side measured ~850 cycles per wait, and nothing here needs it - what the
// arrive publishes is shared memory of the ARRIVING CTA, read only by that CTA's own tensor core (async proxy, made
// visible by fence.proxy.async before the arrive) once the leader issues the pair MMA.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
  chaos_delay(11);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved") canonical layout: the operand is a grid of
// core matrices, each 8 rows x 16 bytes stored as 128 contiguous bytes.
//   lead_bytes   (LBO) = byte distance between core matrices adjacent in the K (reduction) direction
//   stride_bytes (SBO) = byte distance between core matrices adjacent in the M/N direction
// for K-major AND MN-major operands alike (pinned on hardware by tests/test_tc_gpu.py).  A core matrix holds
// 8 rows of the NON-contiguous index x 16 bytes of the contiguous one: 8(mn) x 8(k) elements K-major,
// 8(k) x 8(mn) elements MN-major - the same 128 bytes, which is why one activation tile image serves as the
// A operand of a forward/dgrad GEMM (K-major) and as an operand of the weight-gradient GEMM (MN-major).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lead_bytes, uint32_t stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lead_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((stride_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
  return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
// a_mn / b_mn: 1 when the operand is MN-major (transposed), 0 when K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4)                 // D format: F32
         | (1u << 7)               // A format: BF16
         | (1u << 10)              // B format: BF16
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the 64-bit shared-memory descriptors given as (lo, hi) words: hi (stride-byte offset + version) is
// constant per operand, lo = (addr >> 4) | (lead-byte offset >> 4) << 16 advances by a plain 32-bit add per K step.
// Keeping the issue loop to a handful of instructions per MMA matters: tcgen05.mma is asynchronous, so the tensor
// pipe only stays busy if the single issuing thread needs fewer cycles per MMA than the MMA takes to execute.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lead_bytes) {
  return ((smem_addr >> 4) & 0x3FFF) | (((lead_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t stride_bytes) { return ((stride_bytes >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ void umma_bf16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `cta_mask` (frees a multicast-filled ring stage)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_addr(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar_addr), "h"(cta_mask)
               : "memory");
}

// ------------------------------------------------------------------ TMEM -> registers
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), columns col..col+31.
// The warp may only touch lanes [32*(warp_id%4), +32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ A operand from TENSOR MEMORY (".ts" form)
// D[tmem] (+)= A[tmem] * B[smem]: A is K-major only, row = TMEM lane, one 32-bit column holds the bf16 pair
// (k even in the low half, k odd in the high half) - a K=16 step reads 8 columns.  cta_group::2: each CTA of the pair
// holds its own 128 rows of A in its own tensor memory at the same address (pinned by tests/test_tc_gpu.py).
__device__ __forceinline__ void umma2_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// ---- "whole warp runs the issue loop" forms: every lane executes the call (so that the compiler sees warp-uniform
// operands and keeps them in uniform registers), one elected lane issues.
__device__ __forceinline__ void umma2_bf16_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_w_e(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast_addr_e(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar_addr), "h"(cta_mask)
      : "memory");
}
// registers -> TMEM: thread i of the warp writes lane (base_lane + i), 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest bf16 pack in ONE instruction (negative / NaN inputs clamp to +0)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// ReLU gate flags of a packed non-negative bf16 pair, word index k (0..15) of a 32-column block:
// flag(lo) -> bit k, flag(hi) -> bit 16+k.  (x != 0  <=>  x + 0x7FFF carries into bit 15 of its half.)
__device__ __forceinline__ uint32_t gate_bits(uint32_t packed, int k) {
  return ((packed + 0x7FFF7FFFu) >> (15 - k)) & (0x00010001u << k);
}
// ReLU gate flags straight from 16 fp32 accumulators (sign bits): element 2k -> bit k, element 2k+1 -> bit 16+k of
// the half word `half` (0: columns 0-15 of the block -> k = 0..7, 1: columns 16-31 -> k = 8..15).  One funnel shift
// per element; the caller ORs the halves and inverts once ("positive" = sign bit clear).
__device__ __forceinline__ uint32_t sign_bits16(const uint32_t (&v)[16], int half) {
  uint32_t hi = 0, lo = 0;
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    hi = __funnelshift_l(v[2 * i + 1], hi, 1);     // ends with element 2i+1 at bit i
    lo = __funnelshift_l(v[2 * i], lo, 1);         // element 2i at bit i
  }
  return ((hi << 16) | lo) << (8 * half);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace tc

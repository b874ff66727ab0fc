"""Cycles per tcgen05.mma (M=128, bf16, SWIZZLE_NONE operands resident in shared memory), one CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from mc_nerf_b200._lib import lib
DEV = "cuda"
for mn in (0, 1, 4):
    for N, K in ((256, 256), (256, 64), (128, 256), (64, 256), (32, 256)):
        A = torch.randn(128, K, device=DEV).bfloat16() if mn != 1 else torch.randn(K, 128, device=DEV).bfloat16()
        B = torch.randn(N, K, device=DEV).bfloat16() if mn != 1 else torch.randn(K, N, device=DEV).bfloat16()
        D = torch.empty(128, N, device=DEV)
        cyc = torch.zeros(2, dtype=torch.int64, device=DEV)
        reps = 64
        lib().call("mcnerf_tc_mma_rate", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, mn, reps,
                   ops._p(cyc, torch.int64), ops._stream())
        torch.cuda.synchronize()
        n_mma = reps * K // 16
        print(f"variant={mn} N={N} K={K}: issue {cyc[0].item()/n_mma:.1f} cyc/MMA, complete {cyc[1].item()/n_mma:.1f} cyc/MMA "
              f"(floor {128*N/256:.0f})")
for N, K in ((256, 256), (128, 256), (32, 256)):
    A = torch.randn(256, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    D = torch.empty(256, N, device=DEV)
    cyc = torch.zeros(2, dtype=torch.int64, device=DEV)
    reps = 64
    lib().call("mcnerf_tc_selftest2", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, reps,
               ops._p(cyc, torch.int64), None, ops._stream())
    torch.cuda.synchronize()
    n_mma = reps * K // 16
    print(f"cta_group::2 M=256 N={N} K={K}: issue {cyc[0].item()/n_mma:.1f} cyc/MMA, complete {cyc[1].item()/n_mma:.1f} cyc/MMA")

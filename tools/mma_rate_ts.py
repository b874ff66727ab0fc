"""Cycles per tcgen05.mma with the A operand in tensor memory (CTA pair, M=256), with and without background
tcgen05.ld/st traffic from four other warps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mc_nerf_b200 import ops
from mc_nerf_b200._lib import lib
DEV = "cuda"
for N, K, ns in ((256, 256, 1), (256, 256, 2), (256, 256, 2 | 256), (256, 256, 2 | 512), (256, 256, 1 | 512)):
    for bg in (0, 4000):
        A = torch.randn(256, K, device=DEV).bfloat16()
        B = torch.randn(N, K, device=DEV).bfloat16()
        D = torch.empty(256, N, device=DEV)
        cyc = torch.zeros(2, dtype=torch.int64, device=DEV)
        reps = 64
        lib().call("mcnerf_tc_selftest_ts", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, ns, reps, bg,
                   ops._p(cyc, torch.int64), ops._stream())
        torch.cuda.synchronize()
        n_mma = reps * (ns & 3) * K // 16
        print(f"A in TMEM, cta_group::2 M=256 N={N}/{ns & 3} K={K} bg={bg} bg_mode={(ns >> 4) & 3} kernel_layout={(ns >> 8) & 1} commit_each_pass={ns >> 9}: issue {cyc[0].item()/n_mma:.1f} cyc/MMA, "
              f"complete {cyc[1].item()/n_mma:.1f} cyc/MMA")

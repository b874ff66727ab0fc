"""tcgen05 (UMMA) convention pin: one 128xN tile through the same descriptor helpers the fused MLP kernels
use, for both operand orientations, against torch.matmul on the identical bf16 inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(A, B, N, K, variant):
    from mc_nerf_b200 import ops
    from mc_nerf_b200._lib import lib
    D = torch.full((128, N), float("nan"), device=DEV)
    lib().call("mcnerf_tc_selftest", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, variant,
               ops._stream())
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (32, 256), (64, 32), (128, 16)])
def test_k_major_tile(N, K):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    ref = A.float() @ B.float().t()
    D = run(A, B, N, K, 0)
    err = (D - ref).abs().max().item()
    if not err < 1e-2:
        alt = (run(A, B, N, K, 2) - ref).abs().max().item()
        pytest.fail(f"K-major descriptor convention wrong: err={err}, with lead/stride swapped err={alt}")


@pytest.mark.parametrize("N,K", [(256, 128), (256, 64), (32, 128), (128, 16)])
def test_mn_major_tile(N, K):
    """weight-gradient orientation: D[m,n] = sum_k A[k,m] B[k,n]."""
    g = torch.Generator().manual_seed(N * 1000 + K + 1)
    A = torch.randn(K, 128, generator=g).to(DEV).bfloat16()
    B = torch.randn(K, N, generator=g).to(DEV).bfloat16()
    ref = A.float().t() @ B.float()
    D = run(A, B, N, K, 1)
    err = (D - ref).abs().max().item()
    if not err < 1e-2:
        alt = (run(A, B, N, K, 3) - ref).abs().max().item()
        pytest.fail(f"MN-major descriptor convention wrong: err={err}, with lead/stride swapped err={alt}")


@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (32, 256), (64, 32), (128, 16)])
def test_cta_pair_tile(N, K):
    """tcgen05.mma.cta_group::2: M = 256 over a CTA pair, each CTA holding 128 rows of A and N/2 rows of B."""
    from mc_nerf_b200 import ops
    from mc_nerf_b200._lib import lib
    g = torch.Generator().manual_seed(N * 1000 + K + 2)
    A = torch.randn(256, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    D = torch.full((256, N), float("nan"), device=DEV)
    lib().call("mcnerf_tc_selftest2", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, 1, None,
               None, ops._stream())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert (D - ref).abs().max().item() < 1e-2
    # bias folded into the accumulator by one more MMA (broadcast ones operand, bias as bf16 hi + lo)
    bias = torch.randn(N, generator=g).to(DEV) * 3
    D2 = torch.full((256, N), float("nan"), device=DEV)
    lib().call("mcnerf_tc_selftest2", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D2), N, K, 1, None,
               ops._p(bias), ops._stream())
    torch.cuda.synchronize()
    assert (D2 - D - bias[None, :]).abs().max().item() < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,n_split", [(256, 256, 1), (256, 256, 2), (128, 64, 1), (64, 16, 1), (256, 32, 2)])
def test_cta_pair_tile_with_the_a_operand_in_tensor_memory(N, K, n_split):
    """tcgen05.mma with A read from TMEM (row = lane, one 32-bit column = bf16 pair k even | k odd << 16), cta_group::2,
    whole-N and two-half-N passes: D = A B^T."""
    from mc_nerf_b200 import ops
    from mc_nerf_b200._lib import lib
    g = torch.Generator().manual_seed(7 + N + K)
    A = torch.randn(256, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    D = torch.full((256, N), float("nan"), device=DEV)
    lib().call("mcnerf_tc_selftest_ts", ops._p(A, torch.bfloat16), ops._p(B, torch.bfloat16), ops._p(D), N, K, n_split, 1, 0,
               None, ops._stream())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert (D - ref).abs().max().item() < 1e-2

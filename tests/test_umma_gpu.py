"""GPU tier: tcgen05 / TMEM helper self-test (bf16 operands, fp32 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(32, 32), (96, 32), (128, 32), (32, 128), (128, 128), (16, 16), (256, 64)])
def test_umma_gemm_matches_bf16_matmul(dpt, N, K):
    from dpt_b200._lib import check, lib, ptr, stream_ptr
    torch.manual_seed(N * 1000 + K)
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    D = torch.full((128, N), float("nan"), device="cuda")
    check(lib().dpt_debug_umma_gemm(ptr(A), ptr(B), ptr(D), N, K, stream_ptr()), "dpt_debug_umma_gemm")
    torch.cuda.synchronize()
    want = A.bfloat16().double() @ B.bfloat16().double().T          # exact products of the bf16-rounded operands
    err = (D.double() - want).abs().max().item()
    assert err < 1e-4 * max(1.0, want.abs().max().item()), err

"""CTA-pair (tcgen05 cta_group::2) building blocks: a GEMM whose every MMA spans two SMs, against torch fp32 matmul.
Integer operands -> bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(256, 64, 64), (256, 256, 128), (512, 128, 256), (1000, 192, 64), (256 * 150, 64, 576),
                                   (256 * 75 + 128, 256, 1024)])
def test_pair_gemm(M, N, K):
    from incremental_multimodal_medical_learning_ii_b200 import _native as Nn
    lib = Nn.lib()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randint(-2, 3, (M, K), generator=g).float()
    w = torch.randint(-1, 2, (N, K), generator=g).float()
    ad, wd = a.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    Nn.check(lib.bv_pair_gemm_test(Nn.ptr(ad), Nn.ptr(wd), M, N, K, Nn.ptr(out), Nn.current_stream_handle(out.device)))
    torch.cuda.synchronize()
    ref = a @ w.t()
    assert not torch.isnan(out).any()
    assert torch.equal(out.cpu(), ref), f"max abs diff {(out.cpu() - ref).abs().max().item()}"

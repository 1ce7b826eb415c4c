"""The 8-bit stem kernels alone through ``bv_stem_u8_nhwc``: row-streaming kernel (csrc/stem_rows.cuh, variant 0,
what ``bv_forward`` runs) and tile kernel (csrc/stem_fused.cuh, variant 1).

Reference op: conv1 7x7/2 pad 3 -> bn1 (folded) -> relu -> maxpool 3x3/2 pad 1 of ``ResNetHIML.forward``
(health_multimodal/image/model/resnet.py:34-37) on ToTensor'ed 8-bit frames (transforms.py:12-38; the 1/255 and the
channel sum are folded into the weights by packing.py).  Integer weights -> every fp32 sum is exact -> bit-exact;
Gaussian weights -> fp64 reference, one bf16 ulp (fp32 summation order inside the tensor core is not specified)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _pack_w8(w, bias, dev):
    """[64,7,7] weights + [64] bias -> the ``stem_u8_k8`` matrix of packing.py: k = r*8 + s, chunk 7 = bias hi/mid/lo."""
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    w8 = torch.zeros(64, 8, 8, dtype=torch.float64)
    w8[:, :7, :7] = w.double()
    b64 = bias.double()
    hi = b64.float().to(torch.bfloat16)
    mid = (b64 - hi.double()).float().to(torch.bfloat16)
    lo = (b64 - hi.double() - mid.double()).float().to(torch.bfloat16)
    w8[:, 7, 0], w8[:, 7, 1], w8[:, 7, 2] = hi.double(), mid.double(), lo.double()
    wt = w8.reshape(64, 64).float().to(torch.bfloat16).contiguous().to(dev)
    bt = bias.float().contiguous().to(dev)
    b_eff = hi.double() + mid.double() + lo.double()
    return N.BvConv(wt.data_ptr(), bt.data_ptr(), 64, 64, 1, 1, 1, 0), (wt, bt), b_eff


def _reference(frames, w_bf16, b_eff):
    x = frames.to(DEV).double()[:, None]
    y = F.conv2d(x, w_bf16.double().to(DEV)[:, None], b_eff.to(DEV), stride=2, padding=3)
    y = torch.relu(y.float().to(torch.bfloat16).float())
    y = F.max_pool2d(y, 3, 2, 1)
    return y.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _run(frames, conv, variant):
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    lib = N.lib()
    B, H, W = frames.shape
    out = torch.full((B, H // 4, W // 4, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    fr = frames.to(DEV).contiguous()
    N.check(lib.bv_stem_u8_nhwc(N.ptr(fr), B, H, W, ctypes.byref(conv), N.ptr(out), variant,
                                N.current_stream_handle(torch.device(DEV))))
    torch.cuda.synchronize()
    return out


SHAPES = [(1, 32, 32), (2, 64, 96), (3, 96, 160), (2, 480, 480), (1, 512, 512), (3, 480, 256), (80, 480, 480)]


@pytest.mark.parametrize("variant", [0, 2, 1], ids=["rows", "rows16", "tile"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_stem_integer_bit_exact(shape, variant, monkeypatch):
    B, H, W = shape
    if variant == 2:        # row-streaming kernel with 16 epilogue warps of 16 channels instead of 8 of 32
        monkeypatch.setenv("BV_SR_CG4", "1")
        variant = 0
    g = torch.Generator().manual_seed(B * 7 + H + W)
    frames = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8)
    frames[0, : H // 2] = 255                      # saturated region: large sums, borders included
    w = torch.randint(-1, 2, (64, 7, 7), generator=g).float()
    bias = torch.randint(-300, 300, (64,), generator=g).float()
    conv, keep, b_eff = _pack_w8(w, bias, DEV)
    ref = _reference(frames, w.to(torch.bfloat16), b_eff)
    out = _run(frames, conv, variant)
    assert not torch.isnan(out.float()).any()
    assert torch.equal(out, ref), f"max diff {(out.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("shape", [(2, 64, 96), (2, 480, 480), (1, 512, 512)], ids=lambda s: "x".join(map(str, s)))
def test_stem_gaussian_one_ulp_and_variants_agree(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 3 + W)
    frames = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8)
    w = (torch.randn(64, 7, 7, generator=g) * (49 ** -0.5) / 255.0).to(torch.bfloat16)
    bias = torch.randn(64, generator=g) * 0.3
    conv, keep, b_eff = _pack_w8(w.float(), bias, DEV)
    ref = _reference(frames, w, b_eff).float()
    outs = [_run(frames, conv, v).float() for v in (0, 1)]
    for o in outs:
        # one bf16 ulp = 2^-8 relative (round to nearest: 2^-9, doubled when the fp32 sum lands on the other side)
        assert ((o - ref).abs() <= ref.abs() * 2.0 ** -7 + 1e-6).all(), (o - ref).abs().max().item()
    mism = (outs[0] != outs[1]).float().mean().item()
    print(f"rows vs tile kernel {shape}: {mism:.2e} of the outputs differ (by one bf16 ulp)")
    assert mism < 1e-2

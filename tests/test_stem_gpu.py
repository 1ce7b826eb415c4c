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


@pytest.mark.parametrize("cg4", [False, True], ids=["epi8", "epi16"])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (3, 96, 160), (2, 480, 480), (1, 512, 512), (80, 480, 480)],
                         ids=lambda s: "x".join(map(str, s)))
def test_stem_with_fused_conv1(shape, integer, cg4, monkeypatch):
    """Row-streaming stem + layer1.0 conv1 (1x1 64->64, +bias, ReLU) in one kernel: the max-pool output must equal the
    un-fused kernel's bit for bit, and the conv1 output must equal fp32 conv + bias + ReLU + bf16 rounding of it
    (integer operands: exact; Gaussian: one bf16 ulp)."""
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    if cg4:
        monkeypatch.setenv("BV_SR_CG4", "1")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W = shape
    g = torch.Generator().manual_seed(B + H * 5 + W)
    frames = torch.randint(0, 16 if integer else 256, (B, H, W), generator=g, dtype=torch.uint8)
    if integer:
        w = torch.randint(-1, 2, (64, 7, 7), generator=g).float() * (torch.rand(64, 7, 7, generator=g) < 0.1)
        bias = torch.randint(-60, 60, (64,), generator=g).float()
        w1 = (torch.randint(-1, 2, (64, 64, 1, 1), generator=g).float() * (torch.rand(64, 64, 1, 1, generator=g) < 0.2))
        b1 = torch.randint(-3, 4, (64,), generator=g).float()
    else:
        w = (torch.randn(64, 7, 7, generator=g) * (49 ** -0.5) / 255.0)
        bias = torch.randn(64, generator=g) * 0.3
        w1 = torch.randn(64, 64, 1, 1, generator=g) * 64 ** -0.5
        b1 = torch.randn(64, generator=g) * 0.3
    w, w1 = w.to(torch.bfloat16), w1.to(torch.bfloat16)
    conv, keep, b_eff = _pack_w8(w.float(), bias, DEV)
    c1 = packing.pack_single_conv(w1, b1, 1, 0, torch.device(DEV))
    plain = _run(frames, conv, 0)
    lib = N.lib()
    out = torch.full((B, H // 4, W // 4, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    out1 = torch.full((B, H // 4, W // 4, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    fr = frames.to(DEV).contiguous()
    N.check(lib.bv_stem_conv1_u8_nhwc(N.ptr(fr), B, H, W, ctypes.byref(conv), ctypes.byref(c1[0]), N.ptr(out), N.ptr(out1),
                                      N.current_stream_handle(torch.device(DEV))))
    torch.cuda.synchronize()
    assert torch.equal(out, plain)
    if integer:   # pooled values must be exact bf16 integers for the second conv to be exact
        assert out.float().abs().max() <= 256
    ref1 = torch.relu(F.conv2d(out.double().permute(0, 3, 1, 2), w1.double().to(DEV), b1.double().to(DEV)))
    ref1 = ref1.permute(0, 2, 3, 1).float()
    assert not torch.isnan(out1.float()).any()
    if integer:
        assert torch.equal(out1, ref1.to(torch.bfloat16)), (out1.float() - ref1).abs().max().item()
    else:
        assert ((out1.float() - ref1).abs() <= ref1.abs() * 2.0 ** -7 + 1e-3).all(), (out1.float() - ref1).abs().max().item()

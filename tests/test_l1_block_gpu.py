"""Fused layer1 Bottleneck tail on CTA pairs (csrc/l1_block.cuh) through ``bv_l1_block_nhwc``.

Reference op: conv2(3x3)+bn2+relu -> conv3(1x1)+bn3+identity+relu -> next conv1(1x1)+bn1+relu of torchvision's
Bottleneck (health_multimodal/image/model/resnet.py:39), fp32 ``F.conv2d`` on the same bf16-rounded operands with the
two intermediate tensors rounded to bf16 exactly where the kernel rounds them.  Integer operands -> bit-exact."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x_nhwc, w, b, pad):
    return F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=pad).permute(0, 2, 3, 1)


# the kernel needs an image width that is a multiple of 30 (lane quarters aligned to image lines)
@pytest.mark.parametrize("B,H", [(1, 30), (3, 30), (2, 60), (2, 120), (20, 60), (5, 90)],
                         ids=["tiny", "b3h30", "h60", "h120", "persistent", "h90"])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
@pytest.mark.parametrize("form", ["tap192_n64", "shifted_n64", "shifted_n128"])
def test_l1_block(B, H, integer, form, monkeypatch):
    """form: the tap-fused N = 192 3x3 GEMM (default), the shifted-tap form of it (nine N = 64 MMAs, BV_L1_SH=1), and the
    shifted-tap kernel with a 128-wide chained conv1 (the layer's last block, whose successor is layer2's conv1)."""
    monkeypatch.setenv("BV_L1_SH", "1" if form == "shifted_n64" else "0")
    n2 = 128 if form == "shifted_n128" else 64
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    lib = N.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + H)

    def rnd(shape, lo, hi, scale):
        return torch.randint(lo, hi, shape, generator=g).float() if integer else torch.randn(shape, generator=g) * scale

    t1 = rnd((B, H, H, 64), 0, 3, 1.0).to(torch.bfloat16).to(dev)
    w2 = rnd((64, 64, 3, 3), -1, 2, 576 ** -0.5).to(torch.bfloat16)
    if integer:   # keep |conv2| small enough that every intermediate is an exact bf16 integer
        w2 = (w2 * (torch.rand(w2.shape, generator=g) < 0.25)).to(torch.bfloat16)
    b2 = rnd((64,), -2, 3, 1.0)
    w3 = rnd((256, 64, 1, 1), -1, 2, 64 ** -0.5).to(torch.bfloat16)
    if integer:
        w3 = (w3 * (torch.rand(w3.shape, generator=g) < 0.25)).to(torch.bfloat16)
    b3 = rnd((256,), -2, 3, 1.0)
    w1 = rnd((n2, 256, 1, 1), -1, 2, 256 ** -0.5).to(torch.bfloat16)
    b1 = rnd((n2,), -2, 3, 1.0)
    res = rnd((B, H, H, 256), -2, 3, 1.0).to(torch.bfloat16).to(dev)
    c2 = packing.pack_single_conv(w2, b2, 1, 1, dev)
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    c1 = packing.pack_single_conv(w1, b1, 1, 0, dev)

    t2 = torch.relu(_ref(t1, w2.to(dev), b2.to(dev), 1)).to(torch.bfloat16)
    y = torch.relu(_ref(t2, w3.to(dev), b3.to(dev), 0) + res.float()).to(torch.bfloat16)
    t1n = torch.relu(_ref(y, w1.to(dev), b1.to(dev), 0)).to(torch.bfloat16)
    if integer:
        assert t2.float().abs().max() <= 256 and y.float().abs().max() <= 256, "test data must stay exact in bf16"

    out1 = torch.full((B, H, H, 256), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((B, H, H, n2), float("nan"), device=dev, dtype=torch.bfloat16)
    N.check(lib.bv_l1_block_nhwc(N.ptr(t1), B, H, H, ctypes.byref(c2[0]), ctypes.byref(c3[0]), N.ptr(res), N.ptr(out1),
                                 ctypes.byref(c1[0]), N.ptr(out2), N.current_stream_handle(dev)))
    torch.cuda.synchronize()
    assert not torch.isnan(out1.float()).any() and not torch.isnan(out2.float()).any(), "unwritten output rows"
    if integer:
        assert torch.equal(out1, y), f"out1 max abs diff {(out1.float() - y.float()).abs().max().item()}"
        assert torch.equal(out2, t1n), f"out2 max abs diff {(out2.float() - t1n.float()).abs().max().item()}"
    else:
        torch.testing.assert_close(out1.float(), y.float(), rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(out2.float(), t1n.float(), rtol=3e-2, atol=3e-2)


def test_l1_block_rejects_other_widths():
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    lib = N.lib()
    dev = torch.device("cuda:0")
    z = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)  # noqa: E731
    c2 = packing.pack_single_conv(z(64, 64, 3, 3), torch.zeros(64), 1, 1, dev)
    c3 = packing.pack_single_conv(z(256, 64, 1, 1), torch.zeros(256), 1, 0, dev)
    c1 = packing.pack_single_conv(z(64, 256, 1, 1), torch.zeros(64), 1, 0, dev)
    t1, res = z(1, 16, 16, 64).to(dev), z(1, 16, 16, 256).to(dev)
    out1, out2 = z(1, 16, 16, 256).to(dev), z(1, 16, 16, 64).to(dev)
    rc = lib.bv_l1_block_nhwc(N.ptr(t1), 1, 16, 16, ctypes.byref(c2[0]), ctypes.byref(c3[0]), N.ptr(res), N.ptr(out1),
                              ctypes.byref(c1[0]), N.ptr(out2), N.current_stream_handle(dev))
    assert rc == N.BV_ERR_INVALID


@pytest.mark.parametrize("B,H", [(1, 30), (3, 30), (2, 60), (2, 120), (20, 60), (5, 90)],
                         ids=["tiny", "b3h30", "h60", "h120", "persistent", "h90"])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
@pytest.mark.parametrize("shifted", [False, True], ids=["tap192", "shifted"])
def test_l1_block_downsample_form(B, H, integer, shifted, monkeypatch):
    """First block of layer1 (Bottleneck with a downsample branch): conv2 -> conv3 + downsample_1x1(x0) -> next conv1 in
    one CTA-pair kernel; the downsample is a second K segment of the conv3 accumulation, no identity tensor is read."""
    monkeypatch.setenv("BV_L1_SH", "1" if shifted else "0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    lib = N.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + H + 7)

    def rnd(shape, lo, hi, scale):
        return torch.randint(lo, hi, shape, generator=g).float() if integer else torch.randn(shape, generator=g) * scale

    def sparse(w):
        return (w * (torch.rand(w.shape, generator=g) < 0.25)).to(torch.bfloat16) if integer else w

    t1 = rnd((B, H, H, 64), 0, 3, 1.0).to(torch.bfloat16).to(dev)
    x0 = rnd((B, H, H, 64), 0, 3, 1.0).to(torch.bfloat16).to(dev)
    w2 = sparse(rnd((64, 64, 3, 3), -1, 2, 576 ** -0.5).to(torch.bfloat16))
    b2 = rnd((64,), -2, 3, 1.0)
    w3 = sparse(rnd((256, 64, 1, 1), -1, 2, 64 ** -0.5).to(torch.bfloat16))
    b3 = rnd((256,), -2, 3, 1.0)
    wd = sparse(rnd((256, 64, 1, 1), -1, 2, 64 ** -0.5).to(torch.bfloat16))
    bd = rnd((256,), -2, 3, 1.0)
    w1 = rnd((64, 256, 1, 1), -1, 2, 256 ** -0.5).to(torch.bfloat16)
    b1 = rnd((64,), -2, 3, 1.0)
    c2 = packing.pack_single_conv(w2, b2, 1, 1, dev)
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    cd = packing.pack_single_conv(wd, bd, 1, 0, dev)
    c1 = packing.pack_single_conv(w1, b1, 1, 0, dev)

    t2 = torch.relu(_ref(t1, w2.to(dev), b2.to(dev), 1)).to(torch.bfloat16)
    y = torch.relu(_ref(t2, w3.to(dev), b3.to(dev), 0) + _ref(x0, wd.to(dev), bd.to(dev), 0)).to(torch.bfloat16)
    t1n = torch.relu(_ref(y, w1.to(dev), b1.to(dev), 0)).to(torch.bfloat16)
    if integer:
        assert t2.float().abs().max() <= 256 and y.float().abs().max() <= 256, "test data must stay exact in bf16"
    out1 = torch.full((B, H, H, 256), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((B, H, H, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    N.check(lib.bv_l1_block_ds_nhwc(N.ptr(t1), B, H, H, ctypes.byref(c2[0]), ctypes.byref(c3[0]), N.ptr(x0),
                                    ctypes.byref(cd[0]), N.ptr(out1), ctypes.byref(c1[0]), N.ptr(out2),
                                    N.current_stream_handle(dev)))
    torch.cuda.synchronize()
    assert not torch.isnan(out1.float()).any() and not torch.isnan(out2.float()).any(), "unwritten output rows"
    if integer:
        assert torch.equal(out1, y), f"out1 max abs diff {(out1.float() - y.float()).abs().max().item()}"
        assert torch.equal(out2, t1n), f"out2 max abs diff {(out2.float() - t1n.float()).abs().max().item()}"
    else:
        torch.testing.assert_close(out1.float(), y.float(), rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(out2.float(), t1n.float(), rtol=3e-2, atol=3e-2)

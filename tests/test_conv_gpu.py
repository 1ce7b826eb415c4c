"""Parity of the tcgen05 implicit-GEMM convolution (csrc/conv_gemm.cuh) through the C ABI ``bv_conv2d_nhwc``.

Reference op: torch ``F.conv2d`` in fp32 (TF32 off) on the same bf16-rounded operands, i.e. Conv2d + folded BatchNorm
(+ residual)(+ ReLU) as torchvision's Bottleneck runs them under ``health_multimodal/image/model/resnet.py:39-42``.
Integer-valued operands make every product and partial sum exact in fp32, so those cases are compared BIT-EXACTLY
(any indexing / swizzle / descriptor mistake shows up as a wrong integer); gaussian operands use a tolerance that
only allows fp32 summation-order noise (fp32 output) or one bf16 rounding (bf16 output).
"""
import ctypes
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _conv_native(lib, N, x_nhwc, conv, x2=None, conv2=None, residual=None, relu=True, out_fp32=False):
    B, H, W, _ = x_nhwc.shape
    c0 = conv[0]
    Ho = (H + 2 * c0.pad - c0.r) // c0.stride + 1
    Wo = (W + 2 * c0.pad - c0.s) // c0.stride + 1
    out = torch.full((B, Ho, Wo, c0.cout), float("nan"), device=x_nhwc.device,
                     dtype=torch.float32 if out_fp32 else torch.bfloat16)
    H2 = W2 = 0
    c2p = None
    if x2 is not None:
        H2, W2 = x2.shape[1], x2.shape[2]
        c2p = ctypes.byref(conv2[0])
    stream = N.current_stream_handle(x_nhwc.device)
    N.check(lib.bv_conv2d_nhwc(N.ptr(x_nhwc), B, H, W, ctypes.byref(c0), N.ptr(x2), H2, W2, c2p, N.ptr(residual),
                               1 if relu else 0, N.ptr(out), 1 if out_fp32 else 0, stream))
    torch.cuda.synchronize()
    return out


def _ref(x_nhwc, w, b, stride, pad):
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), b.float(), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def _make(gen, shape, integer, scale=1.0):
    if integer:
        return torch.randint(-2, 3, shape, generator=gen).float()
    return torch.randn(shape, generator=gen) * scale


CASES = [
    # name,            B,  H,  Cin, Cout, k, stride, pad
    ("1x1_tiled_64",   2, 16,   64,   64, 1, 1, 0),
    ("1x1_tail",       1, 15,  128,  128, 1, 1, 0),
    ("1x1_bn256",      2, 16,  256,  256, 1, 1, 0),
    ("1x1_bigK",       2, 15, 2048,  128, 1, 1, 0),
    ("1x1_wideN",      2, 15,  512, 2048, 1, 1, 0),
    ("3x3_s1",         2, 24,   64,   64, 3, 1, 1),
    ("3x3_s1_w120",    3, 120,  64,   64, 3, 1, 1),
    ("3x3_s1_w15",     5, 15,   64,   64, 3, 1, 1),
    ("3x3_s1_c256",    3, 30,  256,  256, 3, 1, 1),
    ("3x3_s2_tail",    3, 30,  128,  128, 3, 2, 1),
    ("3x3_s2_big",     2, 120, 128,  128, 3, 2, 1),
    ("1x1_s2",         2, 30,  256,  512, 1, 2, 0),
]


@pytest.fixture(scope="module")
def env():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    return N, N.lib(), packing


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
def test_single_conv(env, case, integer):
    N, lib, packing = env
    name, B, H, cin, cout, k, stride, pad = case
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    dev = torch.device("cuda:0")
    x = _make(gen, (B, H, H, cin), integer).to(torch.bfloat16).to(dev)
    w = _make(gen, (cout, cin, k, k), integer, scale=(cin * k * k) ** -0.5)
    if integer:
        w = torch.randint(-1, 2, (cout, cin, k, k), generator=gen).float()
    w = w.to(torch.bfloat16)
    b = _make(gen, (cout,), integer)
    conv = packing.pack_single_conv(w, b, stride, pad, dev)
    ref = _ref(x, w.to(dev), b.to(dev), stride, pad)
    out = _conv_native(lib, N, x, conv, relu=False, out_fp32=True)
    assert not torch.isnan(out).any(), "kernel left output rows unwritten"
    if integer:
        assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max().item()}"
    else:
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    # bf16 output + ReLU
    out16 = _conv_native(lib, N, x, conv, relu=True, out_fp32=False)
    ref16 = torch.relu(ref).to(torch.bfloat16)
    if integer:
        assert torch.equal(out16, ref16)
    else:
        torch.testing.assert_close(out16.float(), ref16.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("shape", [(2, 15, 512, 2048), (3, 30, 64, 256), (2, 30, 128, 512), (5, 30, 256, 1024)],
                         ids=["k512n2048", "k64n256", "k128n512", "k256n1024"])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
def test_residual_relu(env, integer, shape):
    """conv3 + identity + ReLU: the residual tile is TMA-loaded into the staging buffer and updated in place.
    The shapes select both residual kernel configurations (long-K 256-wide and short-K 128-wide tiles)."""
    N, lib, packing = env
    gen = torch.Generator().manual_seed(5)
    dev = torch.device("cuda:0")
    B, H, cin, cout = shape
    x = _make(gen, (B, H, H, cin), integer).to(torch.bfloat16).to(dev)
    w = (torch.randint(-1, 2, (cout, cin, 1, 1), generator=gen).float() if integer
         else torch.randn(cout, cin, 1, 1, generator=gen) * cin ** -0.5).to(torch.bfloat16)
    b = _make(gen, (cout,), integer)
    res = _make(gen, (B, H, H, cout), integer).to(torch.bfloat16).to(dev)
    conv = packing.pack_single_conv(w, b, 1, 0, dev)
    ref = torch.relu(_ref(x, w.to(dev), b.to(dev), 1, 0) + res.float())
    out = _conv_native(lib, N, x, conv, residual=res, relu=True, out_fp32=True)
    if integer:
        assert torch.equal(out, ref)
    else:
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    out16 = _conv_native(lib, N, x, conv, residual=res, relu=True, out_fp32=False)
    if integer:
        assert torch.equal(out16, ref.to(torch.bfloat16))
    else:
        torch.testing.assert_close(out16.float(), ref.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])  # 5 (weight-resident, N == 64 only) is what the N=64 cases run
def test_every_kernel_configuration(env, cfg, monkeypatch):
    """Force each <BN, STAGES, NBUF> instantiation on one shape (N=256 fits all tile widths), with a residual."""
    N, lib, packing = env
    monkeypatch.setenv("BV_FORCE_CFG", str(cfg))
    gen = torch.Generator().manual_seed(13)
    dev = torch.device("cuda:0")
    B, H, cin, cout = 4, 30, 192, 256
    x = torch.randint(-2, 3, (B, H, H, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    w = torch.randint(-1, 2, (cout, cin, 3, 3), generator=gen).float().to(torch.bfloat16)
    b = torch.randint(-2, 3, (cout,), generator=gen).float()
    res = torch.randint(-2, 3, (B, H, H, cout), generator=gen).float().to(torch.bfloat16).to(dev)
    conv = packing.pack_single_conv(w, b, 1, 1, dev)
    ref = torch.relu(_ref(x, w.to(dev), b.to(dev), 1, 1) + res.float()).to(torch.bfloat16)
    out = _conv_native(lib, N, x, conv, residual=res, relu=True, out_fp32=False)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
def test_fused_downsample(env, stride, integer):
    """conv3 (1x1 over t2) + downsample (strided 1x1 over the block input) accumulated in one TMEM tile."""
    N, lib, packing = env
    gen = torch.Generator().manual_seed(7 + stride)
    dev = torch.device("cuda:0")
    B, Hin, cin, mid = 3, 30, 256, 128
    cout = mid * 4
    Ho = Hin // stride
    x = _make(gen, (B, Hin, Hin, cin), integer).to(torch.bfloat16).to(dev)
    t2 = _make(gen, (B, Ho, Ho, mid), integer).to(torch.bfloat16).to(dev)
    if integer:
        w3 = torch.randint(-1, 2, (cout, mid, 1, 1), generator=gen).float()
        wd = torch.randint(-1, 2, (cout, cin, 1, 1), generator=gen).float()
    else:
        w3 = torch.randn(cout, mid, 1, 1, generator=gen) * mid ** -0.5
        wd = torch.randn(cout, cin, 1, 1, generator=gen) * cin ** -0.5
    w3, wd = w3.to(torch.bfloat16), wd.to(torch.bfloat16)
    b3, bd = _make(gen, (cout,), integer), _make(gen, (cout,), integer)
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    cd = packing.pack_single_conv(wd, bd, stride, 0, dev)
    ref = torch.relu(_ref(t2, w3.to(dev), b3.to(dev), 1, 0) + _ref(x, wd.to(dev), bd.to(dev), stride, 0))
    out = _conv_native(lib, N, t2, c3, x2=x, conv2=cd, relu=True, out_fp32=True)
    if integer:
        assert torch.equal(out, ref)
    else:
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_many_tiles_persistent(env):
    """More tiles than SMs, so every CTA loops over several tiles and both TMEM accumulators are reused."""
    N, lib, packing = env
    gen = torch.Generator().manual_seed(11)
    dev = torch.device("cuda:0")
    B, H, cin, cout = 8, 60, 128, 512         # M = 28800 -> 225 m-blocks x 2 n-blocks
    x = torch.randint(-2, 3, (B, H, H, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    w = torch.randint(-1, 2, (cout, cin, 1, 1), generator=gen).float().to(torch.bfloat16)
    b = torch.randint(-2, 3, (cout,), generator=gen).float()
    conv = packing.pack_single_conv(w, b, 1, 0, dev)
    ref = _ref(x, w.to(dev), b.to(dev), 1, 0)
    out = _conv_native(lib, N, x, conv, relu=False, out_fp32=True)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("cfg", [4, 5, 6, 7], ids=["stream", "weights_resident", "wide_rows", "tap3_n192"])
@pytest.mark.parametrize("H", [15, 60, 37])
def test_narrow_3x3_configurations(env, cfg, H, monkeypatch):
    """The four N=64 kernels on the layer1 3x3 shape: B tiles streamed, weights resident in smem, the wide-row mode
    (one 130-pixel load per filter row, horizontal taps through row-shifted smem descriptors) and the tap-fused kernel
    (conv3x3_tap3.cuh: one N=192 MMA per filter row, taps recombined across rows in the epilogue)."""
    N, lib, packing = env
    monkeypatch.setenv("BV_FORCE_CFG", str(cfg))
    gen = torch.Generator().manual_seed(17 + H)
    dev = torch.device("cuda:0")
    B, cin, cout = 3, 64, 64
    x = torch.randint(-2, 3, (B, H, H, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    w = torch.randint(-1, 2, (cout, cin, 3, 3), generator=gen).float().to(torch.bfloat16)
    b = torch.randint(-2, 3, (cout,), generator=gen).float()
    conv = packing.pack_single_conv(w, b, 1, 1, dev)
    ref = torch.relu(_ref(x, w.to(dev), b.to(dev), 1, 1)).to(torch.bfloat16)
    out = _conv_native(lib, N, x, conv, relu=True, out_fp32=False)
    assert not torch.isnan(out.float()).any()
    assert torch.equal(out, ref), f"max abs diff {(out.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("case", [(1, 15, 128, 128, 3, 1, 1), (3, 30, 128, 128, 3, 2, 1), (5, 60, 128, 128, 3, 1, 1),
                                  (2, 15, 1024, 128, 1, 1, 0), (7, 30, 128, 128, 3, 1, 1)],
                         ids=["one_block", "s2_ragged", "many_tiles", "1x1_longK", "odd_blocks"])
def test_two_m_tiles_per_stage(env, case, monkeypatch):
    """kCfg128M2: two 128-row m-tiles share every weight stage (MT = 2), against the single-tile configuration and
    the fp32 reference; block counts that are odd exercise the half-empty last stage."""
    N, lib, packing = env
    B, H, cin, cout, k, stride, pad = case
    gen = torch.Generator().manual_seed(B * 100 + H)
    dev = torch.device("cuda:0")
    x = torch.randint(-2, 3, (B, H, H, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    w = torch.randint(-1, 2, (cout, cin, k, k), generator=gen).float().to(torch.bfloat16)
    b = torch.randint(-2, 3, (cout,), generator=gen).float()
    conv = packing.pack_single_conv(w, b, stride, pad, dev)
    # fp64 reference: cuDNN may pick a Winograd/FFT algorithm for fp32 3x3 convolutions, which is not exact on integers
    ref64 = F.conv2d(x.double().permute(0, 3, 1, 2), w.to(dev).double(), b.to(dev).double(), stride=stride, padding=pad)
    ref = torch.relu(ref64.permute(0, 2, 3, 1)).float().to(torch.bfloat16)
    monkeypatch.setenv("BV_FORCE_CFG", "8")
    out = _conv_native(lib, N, x, conv, relu=True, out_fp32=False)
    assert not torch.isnan(out.float()).any(), "unwritten rows"
    assert torch.equal(out, ref), f"max abs diff {(out.float() - ref.float()).abs().max().item()}"
    monkeypatch.setenv("BV_FORCE_CFG", "3")
    assert torch.equal(_conv_native(lib, N, x, conv, relu=True, out_fp32=False), ref)
    # kCfg128TR: the same stage computed transposed (weights as the A operand, N = 256 pixels, channel-major accumulator
    # transposed by the epilogue's 2-byte stores); also without ReLU (negative outputs survive)
    monkeypatch.setenv("BV_FORCE_CFG", "9")
    out = _conv_native(lib, N, x, conv, relu=True, out_fp32=False)
    assert not torch.isnan(out.float()).any(), "unwritten rows"
    assert torch.equal(out, ref), f"transposed tile: max abs diff {(out.float() - ref.float()).abs().max().item()}"
    ref_lin = ref64.permute(0, 2, 3, 1).float().to(torch.bfloat16)
    assert torch.equal(_conv_native(lib, N, x, conv, relu=False, out_fp32=False), ref_lin)


PAIR_CASES = [
    # name,               B,  H,  Cin, Cout, k, stride, pad, residual, cfg
    ("1x1_k1024_n256",     2, 30, 1024,  256, 1, 1, 0, False, 10),
    ("1x1_odd_blocks",     5, 15, 2048,  512, 1, 1, 0, False, 10),      # 1125 rows = 9 m-blocks: the last pair is half empty
    ("3x3_c256",           3, 30,  256,  256, 3, 1, 1, False, 10),
    ("3x3_s2_c256",        3, 60,  256,  256, 3, 2, 1, False, 10),
    ("3x3_c512_longK",     2, 15,  512,  512, 3, 1, 1, False, 10),
    ("many_tiles",         8, 60,  128,  512, 1, 1, 0, False, 10),      # 225 m-blocks x 2 n-blocks = 226 pair tiles > 74 pairs
    ("res_k256_n1024",     5, 30,  256, 1024, 1, 1, 0, True, 11),
    ("res_k512_n2048",     2, 15,  512, 2048, 1, 1, 0, True, 11),
    ("res_odd_blocks",     3, 15,  512,  256, 1, 1, 0, True, 11),       # 675 rows = 6 blocks; B=1 below gives 2 blocks
    ("res_many_tiles",     8, 60,  128,  512, 1, 1, 0, True, 11),
    ("res_3x3",            4, 30,  192,  256, 3, 1, 1, True, 11),
    ("deep_with_res",      4, 30,  192,  256, 3, 1, 1, True, 10),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_cta_pair_configurations(env, case, monkeypatch):
    """The 256-wide tile on CTA pairs (tcgen05 cta_group::2; kCfg256PairDeep = 10, kCfg256PairRes = 11): every shape is
    computed by the forced pair configuration AND by the single-CTA configuration it replaces, both bit-exact against
    the fp64 reference on integer operands."""
    N, lib, packing = env
    name, B, H, cin, cout, k, stride, pad, with_res, cfg = case
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    dev = torch.device("cuda:0")
    x = torch.randint(-2, 3, (B, H, H, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    w = torch.randint(-1, 2, (cout, cin, k, k), generator=gen).float().to(torch.bfloat16)
    b = torch.randint(-2, 3, (cout,), generator=gen).float()
    conv = packing.pack_single_conv(w, b, stride, pad, dev)
    ref64 = F.conv2d(x.double().permute(0, 3, 1, 2), w.to(dev).double(), b.to(dev).double(), stride=stride, padding=pad)
    ref = ref64.permute(0, 2, 3, 1)
    res = None
    if with_res:
        res = torch.randint(-2, 3, tuple(ref.shape), generator=gen).float().to(torch.bfloat16).to(dev)
        ref = ref + res.double()
    ref = torch.relu(ref).float().to(torch.bfloat16)
    for forced in (cfg, 0 if cfg == 10 else 1):
        monkeypatch.setenv("BV_FORCE_CFG", str(forced))
        out = _conv_native(lib, N, x, conv, residual=res, relu=True, out_fp32=False)
        assert not torch.isnan(out.float()).any(), f"cfg {forced}: unwritten rows"
        assert torch.equal(out, ref), f"cfg {forced}: max abs diff {(out.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("stride", [1, 2])
def test_cta_pair_fused_downsample(env, stride, monkeypatch):
    """Two K segments (conv3 over t2 + strided downsample over the block input) accumulated by the pair kernel."""
    N, lib, packing = env
    gen = torch.Generator().manual_seed(70 + stride)
    dev = torch.device("cuda:0")
    B, Hin, cin, mid = 3, 30, 512, 256
    cout = mid * 4
    Ho = Hin // stride
    x = torch.randint(-2, 3, (B, Hin, Hin, cin), generator=gen).float().to(torch.bfloat16).to(dev)
    t2 = torch.randint(-2, 3, (B, Ho, Ho, mid), generator=gen).float().to(torch.bfloat16).to(dev)
    w3 = torch.randint(-1, 2, (cout, mid, 1, 1), generator=gen).float().to(torch.bfloat16)
    wd = torch.randint(-1, 2, (cout, cin, 1, 1), generator=gen).float().to(torch.bfloat16)
    b3 = torch.randint(-2, 3, (cout,), generator=gen).float()
    bd = torch.randint(-2, 3, (cout,), generator=gen).float()
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    cd = packing.pack_single_conv(wd, bd, stride, 0, dev)
    ref = torch.relu(_ref(t2, w3.to(dev), b3.to(dev), 1, 0) + _ref(x, wd.to(dev), bd.to(dev), stride, 0)).to(torch.bfloat16)
    for forced in (10, 0):
        monkeypatch.setenv("BV_FORCE_CFG", str(forced))
        out = _conv_native(lib, N, t2, c3, x2=x, conv2=cd, relu=True, out_fp32=False)
        assert torch.equal(out, ref), f"cfg {forced}"

"""Parity of the chained conv3 -> next-conv1 kernel (csrc/chain_gemm.cuh) through the C ABI ``bv_conv_chain_nhwc``.

Reference op: the tail of one torchvision Bottleneck (conv3 + bn3 + identity|downsample + ReLU) followed by the head of
the next (conv1 + bn1 + ReLU), as driven by ``health_multimodal/image/model/resnet.py:38-42``, computed with fp32
``F.conv2d`` (TF32 off) on the same bf16-rounded operands; the block output is rounded to bf16 between the two
convolutions exactly as the kernel does.  Integer operands make every partial sum exact, so those cases are BIT-EXACT.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    from incremental_multimodal_medical_learning_ii_b200 import packing
    return N, N.lib(), packing


def _ref(x_nhwc, w, b, stride, pad):
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), b.float(), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def _rand(gen, shape, integer, lo=-2, hi=3, scale=1.0):
    if integer:
        return torch.randint(lo, hi, shape, generator=gen).float()
    return torch.randn(shape, generator=gen) * scale


CASES = [
    # name,          B,  H, mid,  N1,  N2, ds_cin, ds_stride   (ds_cin == 0 -> identity residual)
    ("l1_res",       3, 24,  64, 256,  64,   0, 1),
    ("l1_ds",        3, 24,  64, 256,  64,  64, 1),
    ("l1_to_l2",     2, 24,  64, 256, 128,   0, 1),
    ("l2_res",       3, 15, 128, 512, 128,   0, 1),
    ("l2_ds_s2",     2, 15, 128, 512, 128, 256, 2),
    ("l2_to_l3",     2, 15, 128, 512, 256,   0, 1),
    ("persistent",   8, 60,  64, 256,  64,   0, 1),   # 225 tiles > 148 SMs: every barrier ring wraps
    ("one_tile",     1,  8,  64, 256,  64,   0, 1),   # M = 64: a single, half-empty tile
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
def test_chain(env, case, integer):
    N, lib, packing = env
    name, B, H, mid, n1, n2, ds_cin, ds_stride = case
    gen = torch.Generator().manual_seed(101 + len(name) + B * H)
    dev = torch.device("cuda:0")
    t2 = _rand(gen, (B, H, H, mid), integer).to(torch.bfloat16).to(dev)
    w3 = (_rand(gen, (n1, mid, 1, 1), True, -1, 2) if integer else _rand(gen, (n1, mid, 1, 1), False, scale=mid ** -0.5))
    w3 = w3.to(torch.bfloat16)
    b3 = _rand(gen, (n1,), integer)
    w1 = (_rand(gen, (n2, n1, 1, 1), True, -1, 2) if integer else _rand(gen, (n2, n1, 1, 1), False, scale=n1 ** -0.5))
    w1 = w1.to(torch.bfloat16)
    b1 = _rand(gen, (n2,), integer)
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    c1 = packing.pack_single_conv(w1, b1, 1, 0, dev)
    y = _ref(t2, w3.to(dev), b3.to(dev), 1, 0)
    x2 = res = cd = None
    H2 = 0
    if ds_cin:
        H2 = H * ds_stride
        x2 = _rand(gen, (B, H2, H2, ds_cin), integer).to(torch.bfloat16).to(dev)
        wd = (_rand(gen, (n1, ds_cin, 1, 1), True, -1, 2) if integer
              else _rand(gen, (n1, ds_cin, 1, 1), False, scale=ds_cin ** -0.5)).to(torch.bfloat16)
        bd = _rand(gen, (n1,), integer)
        cd = packing.pack_single_conv(wd, bd, ds_stride, 0, dev)
        y = y + _ref(x2, wd.to(dev), bd.to(dev), ds_stride, 0)
    else:
        res = _rand(gen, (B, H, H, n1), integer).to(torch.bfloat16).to(dev)
        y = y + res.float()
    ref1 = torch.relu(y).to(torch.bfloat16)
    ref2 = torch.relu(_ref(ref1, w1.to(dev), b1.to(dev), 1, 0)).to(torch.bfloat16)

    out1 = torch.full((B, H, H, n1), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((B, H, H, n2), float("nan"), device=dev, dtype=torch.bfloat16)
    N.check(lib.bv_conv_chain_nhwc(N.ptr(t2), B, H, H, ctypes.byref(c3[0]), N.ptr(x2), H2, H2,
                                   ctypes.byref(cd[0]) if cd is not None else None, N.ptr(res), N.ptr(out1),
                                   ctypes.byref(c1[0]), N.ptr(out2), N.current_stream_handle(dev)))
    torch.cuda.synchronize()
    assert not torch.isnan(out1.float()).any() and not torch.isnan(out2.float()).any(), "unwritten output rows"
    if integer:
        assert torch.equal(out1, ref1), f"out1 max abs diff {(out1.float() - ref1.float()).abs().max().item()}"
        assert torch.equal(out2, ref2), f"out2 max abs diff {(out2.float() - ref2.float()).abs().max().item()}"
    else:
        torch.testing.assert_close(out1.float(), ref1.float(), rtol=1e-2, atol=1e-2)
        # out2 is computed from the kernel's own bf16 out1, which may differ from ref1 by one bf16 ulp per element
        ref2k = torch.relu(_ref(out1, w1.to(dev), b1.to(dev), 1, 0)).to(torch.bfloat16)
        torch.testing.assert_close(out2.float(), ref2k.float(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(out2.float(), ref2.float(), rtol=3e-2, atol=3e-2)


def test_chain_rejects_unsupported_shapes(env):
    N, lib, packing = env
    dev = torch.device("cuda:0")
    t2 = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=dev)
    c3 = packing.pack_single_conv(torch.zeros(256, 64, 1, 1, dtype=torch.bfloat16), torch.zeros(256), 1, 0, dev)
    bad = packing.pack_single_conv(torch.zeros(64, 256, 3, 3, dtype=torch.bfloat16), torch.zeros(64), 1, 1, dev)
    out1 = torch.zeros(1, 8, 8, 256, dtype=torch.bfloat16, device=dev)
    out2 = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=dev)
    rc = lib.bv_conv_chain_nhwc(N.ptr(t2), 1, 8, 8, ctypes.byref(c3[0]), None, 0, 0, None, None, N.ptr(out1),
                                ctypes.byref(bad[0]), N.ptr(out2), N.current_stream_handle(dev))
    assert rc == N.BV_ERR_INVALID


PAIR_CHAIN_CASES = [
    # name,             B,  H, mid,   N1,  N2
    ("l3_res",          3, 30, 256, 1024, 256),     # 2700 rows = 22 m-blocks = 11 pair tiles
    ("l3_odd_blocks",   1, 15, 256, 1024, 256),     # 225 rows = 2 blocks, the second mostly empty
    ("l3_one_block",    1,  8, 256, 1024, 256),     # 64 rows: the peer CTA's whole tile is out of bounds
    ("l3_three_blocks", 1, 19, 256, 1024, 256),     # 361 rows = 3 blocks: the last pair has an out-of-bounds peer tile
    ("l3_persistent",  24, 30, 256, 1024, 256),     # 21600 rows = 169 blocks = 85 pair tiles > 74 pairs: every ring wraps
    ("l2_to_l3",        2, 30, 128,  512, 256),
    ("l2_res",          3, 30, 128,  512, 128),
    ("l2_persistent",  12, 60, 128,  512, 128),     # 43200 rows = 338 blocks = 169 pair tiles
]


@pytest.mark.parametrize("case", PAIR_CHAIN_CASES, ids=[c[0] for c in PAIR_CHAIN_CASES])
@pytest.mark.parametrize("integer", [True, False], ids=["int", "gauss"])
def test_pair_chain(env, case, integer):
    """conv3 + identity + ReLU chained with the next conv1 + ReLU on CTA pairs (csrc/pair_chain.cuh) through
    ``bv_pair_chain_nhwc``: resident conv3 input tile, half of every weight stage per CTA, y sub-tiles feed both the TMA
    store and the second GEMM.  Integer operands -> bit-exact against fp32 ``F.conv2d``."""
    N, lib, packing = env
    name, B, H, mid, n1, n2 = case
    gen = torch.Generator().manual_seed(211 + len(name) + B * H)
    dev = torch.device("cuda:0")
    t2 = _rand(gen, (B, H, H, mid), integer, 0, 3).to(torch.bfloat16).to(dev)
    if integer:   # sparse ternary weights keep |y| small enough that y and every partial sum are exact bf16 / fp32 integers
        w3 = _rand(gen, (n1, mid, 1, 1), True, -1, 2) * (torch.rand(n1, mid, 1, 1, generator=gen) < 0.1)
        w1 = _rand(gen, (n2, n1, 1, 1), True, -1, 2) * (torch.rand(n2, n1, 1, 1, generator=gen) < 0.05)
    else:
        w3 = _rand(gen, (n1, mid, 1, 1), False, scale=mid ** -0.5)
        w1 = _rand(gen, (n2, n1, 1, 1), False, scale=n1 ** -0.5)
    w3, w1 = w3.to(torch.bfloat16), w1.to(torch.bfloat16)
    b3, b1 = _rand(gen, (n1,), integer), _rand(gen, (n2,), integer)
    res = _rand(gen, (B, H, H, n1), integer).to(torch.bfloat16).to(dev)
    c3 = packing.pack_single_conv(w3, b3, 1, 0, dev)
    c1 = packing.pack_single_conv(w1, b1, 1, 0, dev)
    ref1 = torch.relu(_ref(t2, w3.to(dev), b3.to(dev), 1, 0) + res.float())
    if integer:
        assert ref1.abs().max() <= 256, "test data must stay exact in bf16"
    ref1 = ref1.to(torch.bfloat16)
    ref2 = torch.relu(_ref(ref1, w1.to(dev), b1.to(dev), 1, 0)).to(torch.bfloat16)
    out1 = torch.full((B, H, H, n1), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((B, H, H, n2), float("nan"), device=dev, dtype=torch.bfloat16)
    N.check(lib.bv_pair_chain_nhwc(N.ptr(t2), B, H, H, ctypes.byref(c3[0]), N.ptr(res), N.ptr(out1), ctypes.byref(c1[0]),
                                   N.ptr(out2), N.current_stream_handle(dev)))
    torch.cuda.synchronize()
    assert not torch.isnan(out1.float()).any() and not torch.isnan(out2.float()).any(), "unwritten output rows"
    if integer:
        assert torch.equal(out1, ref1), f"out1 max abs diff {(out1.float() - ref1.float()).abs().max().item()}"
        assert torch.equal(out2, ref2), f"out2 max abs diff {(out2.float() - ref2.float()).abs().max().item()}"
    else:
        torch.testing.assert_close(out1.float(), ref1.float(), rtol=1e-2, atol=1e-2)
        ref2k = torch.relu(_ref(out1, w1.to(dev), b1.to(dev), 1, 0)).to(torch.bfloat16)
        torch.testing.assert_close(out2.float(), ref2k.float(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(out2.float(), ref2.float(), rtol=3e-2, atol=3e-2)


def test_pair_chain_rejects_unsupported_shapes(env):
    N, lib, packing = env
    dev = torch.device("cuda:0")
    z = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)  # noqa: E731
    t2, res = z(1, 8, 8, 64).to(dev), z(1, 8, 8, 256).to(dev)
    c3 = packing.pack_single_conv(z(256, 64, 1, 1), torch.zeros(256), 1, 0, dev)      # K1 = 64: layer1 shape, not supported
    c1 = packing.pack_single_conv(z(64, 256, 1, 1), torch.zeros(64), 1, 0, dev)
    out1, out2 = z(1, 8, 8, 256).to(dev), z(1, 8, 8, 64).to(dev)
    rc = lib.bv_pair_chain_nhwc(N.ptr(t2), 1, 8, 8, ctypes.byref(c3[0]), N.ptr(res), N.ptr(out1), ctypes.byref(c1[0]),
                                N.ptr(out2), N.current_stream_handle(dev))
    assert rc == N.BV_ERR_INVALID

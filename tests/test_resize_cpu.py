"""The resize oracle (oracle/pil_resize_oracle.py) pinned against the real thing: torchvision Resize + CenterCrop on
PIL images with the Pillow installed here, bit-exact (reference pipeline: DataRetrieval.py:175-180)."""
import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms

import pil_resize_oracle as R

SIZES = [(320, 390), (390, 320), (320, 320), (512, 512), (97, 131), (1024, 833), (600, 2000), (2320, 2828 // 4)]


def _frame(h, w, seed):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, size=(h // 8 + 2, w // 8 + 2)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8), dtype=np.float32))[:h, :w] * 0.7 + g.integers(0, 77, size=(h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("hw", SIZES, ids=[f"{h}x{w}" for h, w in SIZES])
@pytest.mark.parametrize("size,crop", [(512, 480), (512, 512), (256, 224)])
def test_oracle_equals_pil(hw, size, crop):
    h, w = hw
    img = _frame(h, w, h * 10007 + w)
    pipeline = transforms.Compose([transforms.ToPILImage(), transforms.Resize(size), transforms.CenterCrop(crop)])
    ref = np.asarray(pipeline(torch.from_numpy(img).unsqueeze(0)))          # the reference's own transform objects
    out = R.resize_center_crop(img, size, crop)
    assert out.shape == (crop, crop) and out.dtype == np.uint8
    assert np.array_equal(out, ref), f"max abs diff {np.abs(out.astype(int) - ref.astype(int)).max()}"


def test_resized_size_and_crop_origin_follow_torchvision():
    assert R.resized_size(390, 320, 512) == (624, 512)
    assert R.resized_size(320, 390, 512) == (512, 624)
    assert R.resized_size(2828, 2320, 512) == (int(512 * 2828 / 2320), 512)
    assert R.crop_origin(624, 512, 480) == (72, 16)
    assert R.crop_origin(625, 512, 480) == (72, 16)          # round(72.5) -> 72 (banker's rounding, like torchvision)
    with pytest.raises(ValueError):
        R.resize_center_crop(np.zeros((100, 100), np.uint8), 64, 96)


def test_coefficients_are_normalised_fixed_point():
    for n_in, n_out in [(390, 624), (2828, 624), (512, 512), (97, 512)]:
        bounds, kk, ksize = R.precompute_coeffs(n_in, n_out)
        assert kk.shape == (n_out, ksize)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
        s = kk.sum(axis=1)
        assert np.abs(s - (1 << R.PRECISION_BITS)).max() <= ksize            # rows sum to 1.0 up to rounding


def test_identity_and_constant_images():
    img = _frame(480, 480, 5)
    assert np.array_equal(R.resize_center_crop(img, 480, 480), img)           # no pass runs when sizes match
    const = np.full((333, 444), 201, np.uint8)
    assert (R.resize_center_crop(const, 512, 480) == 201).all()


def test_native_workspace_helper_matches_the_plan(native_lib):
    """Host-only ABI helper (no GPU): the temp image holds exactly the source rows the cropped vertical pass touches."""
    lib = native_lib
    assert lib.bv_resize_workspace_bytes(4, 320, 390, 512, 480) > 0
    assert lib.bv_resize_workspace_bytes(4, 100, 100, 64, 96) == 0           # crop exceeds the resized frame
    assert lib.bv_resize_workspace_bytes(0, 320, 390, 512, 480) == 0
    bounds, _, _ = R.precompute_coeffs(390, 624)                              # 390x320 -> 624x512, crop rows 72..551
    rows = bounds[72 + 479, 0] + bounds[72 + 479, 1] - bounds[72, 0]
    need = lib.bv_resize_workspace_bytes(1, 390, 320, 512, 480)
    assert rows * 480 <= need <= rows * 480 + 512

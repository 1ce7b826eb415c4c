"""GPU Resize + CenterCrop (csrc/resize.cuh through ``bv_resize_center_crop_u8``) against PIL itself and the oracle:
bit-exact, since the path is integer arithmetic on bytes (DataRetrieval.py:175-180)."""
import numpy as np
import pytest
import torch
from torchvision import transforms

import pil_resize_oracle as R

pytestmark = pytest.mark.gpu

SIZES = [(320, 390), (390, 320), (320, 320), (97, 131), (1024, 833), (600, 2000), (480, 480), (512, 700), (700, 512)]


def _frames(n, h, w, seed):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, size=(n, h // 8 + 2, w // 8 + 2)).astype(np.float32)
    img = np.kron(base, np.ones((1, 8, 8), dtype=np.float32))[:, :h, :w] * 0.7 + g.integers(0, 77, size=(n, h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("hw", SIZES, ids=[f"{h}x{w}" for h, w in SIZES])
@pytest.mark.parametrize("size,crop", [(512, 480), (512, 512), (480, 480)])
def test_resize_crop_bit_exact_vs_pil(hw, size, crop):
    from incremental_multimodal_medical_learning_ii_b200.image.data.gpu_transforms import GpuResizeCenterCrop
    h, w = hw
    fr = _frames(3, h, w, h * 31 + w)
    pipeline = transforms.Compose([transforms.ToPILImage(), transforms.Resize(size), transforms.CenterCrop(crop)])
    ref = np.stack([np.asarray(pipeline(torch.from_numpy(f).unsqueeze(0))) for f in fr])
    assert np.array_equal(ref[0], R.resize_center_crop(fr[0], size, crop))          # oracle == PIL on this case too
    out = GpuResizeCenterCrop(size, crop)(torch.from_numpy(fr).cuda())
    assert out.shape == (3, 1, crop, crop) and out.dtype == torch.uint8
    got = out[:, 0].cpu().numpy()
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got.astype(int) - ref.astype(int)).max()}"


def test_mixed_sizes_keep_order_and_feed_the_model_dtype():
    from incremental_multimodal_medical_learning_ii_b200.image.data.gpu_transforms import GpuResizeCenterCrop
    sizes = [(320, 390), (390, 320), (320, 390), (333, 333)]
    frames = [_frames(1, h, w, i)[0] for i, (h, w) in enumerate(sizes)]
    out = GpuResizeCenterCrop(512, 480)([torch.from_numpy(f).cuda() for f in frames])
    for i, f in enumerate(frames):
        assert np.array_equal(out[i, 0].cpu().numpy(), R.resize_center_crop(f, 512, 480))


def test_checksum_at_batch_scale():
    """A CheXpert-small sized batch (256 frames of 320x390): every output byte equals the oracle's on a sample, and
    the whole batch is deterministic (two runs, identical bytes)."""
    from incremental_multimodal_medical_learning_ii_b200.image.data.gpu_transforms import GpuResizeCenterCrop
    fr = torch.from_numpy(_frames(256, 320, 390, 9)).cuda()
    t = GpuResizeCenterCrop(512, 480)
    a, b = t(fr), t(fr)
    assert torch.equal(a, b)
    for i in (0, 100, 255):
        assert np.array_equal(a[i, 0].cpu().numpy(), R.resize_center_crop(fr[i].cpu().numpy(), 512, 480))


def test_rejects_bad_input():
    from incremental_multimodal_medical_learning_ii_b200.image.data.gpu_transforms import GpuResizeCenterCrop
    t = GpuResizeCenterCrop(64, 96)
    with pytest.raises(ValueError):
        t(torch.zeros(1, 100, 100, dtype=torch.uint8, device="cuda"))     # crop larger than the resized frame
    with pytest.raises(ValueError):
        GpuResizeCenterCrop(512, 480)(torch.zeros(1, 100, 100, dtype=torch.uint8))   # CPU tensor

"""Input contract of the BioViL image model (reference ``health_multimodal/image/data/transforms.py``):
``Resize -> CenterCrop -> ToTensor -> ExpandChannels``, i.e. float32 in [0,1] (= uint8 / 255), no mean/std
normalisation, one grey channel copied three times.  Host-side (PIL / torchvision); not on the GPU hot path."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
from torchvision.transforms import CenterCrop, Compose, Resize, ToTensor


class ExpandChannels:
    """[1, H, W] -> [3, H, W] by copying the channel (reference transforms.py:12-25)."""

    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        if data.shape[0] != 1:
            raise ValueError(f"Expected input of shape [1, H, W], found {data.shape}")
        return torch.repeat_interleave(data, 3, dim=0)


def create_chest_xray_transform_for_inference(resize: int, center_crop_size: int) -> Compose:
    """Reference transforms.py:28-38."""
    return Compose([Resize(resize), CenterCrop(center_crop_size), ToTensor(), ExpandChannels()])


def infer_resize_params(val_img_transforms: Sequence[Callable]) -> Tuple[Optional[int], Optional[int]]:
    """Sizes the pipeline resizes / crops to (reference transforms.py:41-70); ``ValueError`` for unknown transforms."""
    resize_size = None
    crop_size = None
    supported = (Resize, CenterCrop, ToTensor, ExpandChannels)
    for t in val_img_transforms:
        if type(t) not in supported:
            raise ValueError(f"Unsupported transform type {type(t)}. Supported types are {supported}")
        if isinstance(t, Resize):
            if resize_size is not None or crop_size is not None:
                raise ValueError("Expected Resize to be the first transform if present in val_img_transforms")
            assert t.max_size is None
            assert isinstance(t.size, int), f"Expected int, got {t.size}"
            resize_size = t.size
        elif isinstance(t, CenterCrop):
            if crop_size is not None:
                raise ValueError(f"Crop size has already been set to {crop_size} in a previous transform")
            assert len(t.size) == 2 and t.size[0] == t.size[1], "Only square center crop supported"
            crop_size = t.size[0]
    return resize_size, crop_size

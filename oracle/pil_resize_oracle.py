"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's frame pre-processing for 8-bit grayscale images,
``transforms.Resize(size)`` -> ``transforms.CenterCrop(crop)`` on a PIL image (reference: ``DataRetrieval.py:175-180``
``get_bio_vil_pipeline``; ``health_multimodal/image/data/transforms.py:30-41``), before ``ToTensor`` divides by 255.

The arithmetic lives in third-party code that is not under /root/reference:
  * torchvision (pinned 0.10.0, requirements.txt:112) ``Resize`` with an int size: the SHORT side becomes ``size`` and
    the long side ``int(size * long / short)``; ``CenterCrop``: ``top = int(round((h - crop) / 2.0))`` (Python round);
  * Pillow (pinned 9.3.0, requirements.txt:64) ``Image.resize(..., BILINEAR)`` for mode "L":
    ``src/libImaging/Resample.c`` - ``precompute_coeffs`` (double), ``normalize_coeffs_8bpc`` (22-bit fixed point),
    ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc`` (int32 accumulate, +0.5 rounding, clip).
Pinned by ``tests/test_resize_cpu.py`` against the Pillow installed in this image (bit-exact on random sizes).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision ``_compute_resized_output_size`` for an int ``size`` (no max_size): returns (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_origin(h: int, w: int, crop: int) -> Tuple[int, int]:
    """torchvision ``center_crop`` offsets (top, left) for an image at least ``crop`` on both sides."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c ``precompute_coeffs`` (box = whole axis, BILINEAR: support 1.0) + ``normalize_coeffs_8bpc``.
    Returns (bounds[out,2] = (xmin, count), kk[out, ksize] int32, ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            wgt = 1.0 - a if a < 1.0 else 0.0
            k[x] = wgt
            ww += wgt
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(ss: np.ndarray) -> np.ndarray:
    return np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8bpc pass of Resample.c along ``axis`` (1 = horizontal, 0 = vertical)."""
    src = img if axis == 1 else img.T
    bounds, kk, _ = precompute_coeffs(src.shape[1], out_size)
    out = np.empty((src.shape[0], out_size), dtype=np.uint8)
    s64 = src.astype(np.int64)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (s64[:, xmin:xmin + cnt] * kk[xx, :cnt].astype(np.int64)).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        out[:, xx] = _clip8(acc)
    return out if axis == 1 else out.T


def pil_resize_bilinear(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """``Image.fromarray(img, 'L').resize((new_w, new_h), BILINEAR)``: horizontal pass first, then vertical; a pass
    whose size does not change is skipped (Resample.c ``ImagingResample``)."""
    out = img
    if new_w != img.shape[1]:
        out = resample_axis(out, new_w, axis=1)
    if new_h != img.shape[0]:
        out = resample_axis(out, new_h, axis=0)
    return np.ascontiguousarray(out)


def resize_center_crop(img: np.ndarray, size: int, crop: int) -> np.ndarray:
    """The reference's Resize(size) -> CenterCrop(crop) on one [H, W] uint8 frame -> [crop, crop] uint8."""
    new_h, new_w = resized_size(img.shape[0], img.shape[1], size)
    if new_h < crop or new_w < crop:
        raise ValueError("crop larger than the resized image (torchvision would zero-pad; not used by the reference)")
    r = pil_resize_bilinear(img, new_h, new_w)
    top, left = crop_origin(new_h, new_w, crop)
    return np.ascontiguousarray(r[top:top + crop, left:left + crop])

"""Synthetic chest-X-ray-shaped frames and prompt embeddings (there is no network for CheXpert / CXR-BERT).

Frames follow the reference's input contract (``health_multimodal/image/data/transforms.py:28-38``:
``ToTensor`` of an 8-bit grey image then ``ExpandChannels``): values are ``k/255`` with ``k`` a uint8, one
channel replicated three times.  They are produced by a counter-based *integer* hash so that the CPU oracle
and the GPU produce bit-identical frames from ``(seed, frame index)`` alone, whatever the device, batch split
or rank - which is what lets a sharded run be compared with a sequential one.

Two distributions:

* ``iid``        - every pixel an independent uniform byte.  Embeddings of different frames are nearly
                   colinear under random-init weights, so this is the *weak* parity case.
* ``structured`` - a coarse 8x8 field, bilinearly upsampled in fixed point, times a per-frame contrast, plus
                   10 % pixel noise.  Gives embeddings that actually differ between frames.
"""
from __future__ import annotations

import torch

_M32 = 0xFFFFFFFF


def _mix32(x: torch.Tensor) -> torch.Tensor:
    """lowbias32-style avalanche on int64 tensors holding uint32 values (identical on CPU and CUDA)."""
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def _hash3(seed: int, a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    h = _mix32(a + 0x9E3779B9 * (seed + 1))
    h = _mix32(h ^ (b * 0x85EBCA6B & _M32))
    h = _mix32(h ^ (c * 0xC2B2AE35 & _M32))
    return h


def synthetic_frames_u8(first: int, count: int, size: int = 480, kind: str = "iid", seed: int = 0,
                        device="cpu") -> torch.Tensor:
    """Frames ``first .. first+count-1`` as uint8 ``[count, 1, size, size]`` on ``device``."""
    dev = torch.device(device)
    idx = torch.arange(first, first + count, device=dev, dtype=torch.int64).view(-1, 1, 1)
    yy = torch.arange(size, device=dev, dtype=torch.int64).view(1, -1, 1)
    xx = torch.arange(size, device=dev, dtype=torch.int64).view(1, 1, -1)
    noise = _hash3(seed, idx, yy, xx) & 0xFF                                   # [count,size,size]
    if kind == "iid":
        k = noise
    elif kind == "structured":
        cells = 8
        assert size % cells == 0, "structured frames need size % 8 == 0"
        step = size // cells
        gy, gx = yy // step, xx // step
        fy, fx = yy - gy * step, xx - gx * step                                # 0..step-1

        def field(a, b):
            return _hash3(seed + 101, idx, a, b) & 0xFF

        f00, f01 = field(gy, gx), field(gy, gx + 1)
        f10, f11 = field(gy + 1, gx), field(gy + 1, gx + 1)
        low = (f00 * (step - fy) * (step - fx) + f01 * (step - fy) * fx
               + f10 * fy * (step - fx) + f11 * fy * fx) // (step * step)      # 0..255
        contrast = _hash3(seed + 202, idx, torch.zeros_like(idx), torch.zeros_like(idx)) & 0xFF
        k = (low * contrast // 255) * 9 // 10 + noise // 10
        k = torch.clamp(k, 0, 255)
    else:
        raise ValueError(f"unknown frame kind {kind!r}")
    return k.to(torch.uint8).unsqueeze(1)


def frames_as_reference_input(frames_u8: torch.Tensor) -> torch.Tensor:
    """uint8 ``[B,1,H,W]`` -> float32 ``[B,3,H,W]`` in [0,1]: what ``ToTensor`` + ``ExpandChannels``
    (``transforms.py:12-25,37``) hand to the reference model."""
    x = frames_u8.to(torch.float32) / 255.0
    return torch.repeat_interleave(x, 3, dim=1)


def synthetic_prompt_embeddings(num_labels: int = 14, prompts_per_polarity: int = 1, dim: int = 128,
                                seed: int = 29, min_margin_against: torch.Tensor | None = None,
                                min_margin: float = 0.0) -> torch.Tensor:
    """Stand-in for CXR-BERT output: ``[L, 2, P, D]`` ~ N(0,1), index 0 = positive, 1 = negative prompts
    (``Trainer.bert_forward_mean`` returns un-normalised ``[P,128]`` per polarity, ``Trainer.py:1657-1680``).

    If ``min_margin_against`` (oracle embeddings ``[B,D]``) is given, labels are re-drawn until every image's
    fp32 margin ``|cos_pos - cos_neg|`` exceeds ``min_margin``; used to build a prompt set whose decisions are
    above the bf16 noise floor."""
    g = torch.Generator().manual_seed(seed)
    out = torch.randn(num_labels, 2, prompts_per_polarity, dim, generator=g)
    if min_margin_against is None or min_margin <= 0:
        return out
    e = torch.nn.functional.normalize(min_margin_against.float(), dim=-1)
    for l in range(num_labels):
        for _ in range(10000):
            t = torch.nn.functional.normalize(out[l].mean(dim=1), dim=-1)      # [2,D]
            s = e @ t.T
            if (s[:, 0] - s[:, 1]).abs().min().item() > min_margin:
                break
            out[l] = torch.randn(2, prompts_per_polarity, dim, generator=g)
        else:
            raise RuntimeError("could not build a prompt set with the requested margin")
    return out

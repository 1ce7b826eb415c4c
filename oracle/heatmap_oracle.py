"""TEST INFRASTRUCTURE (oracle): CPU restatement of the arithmetic ``heat_to_image_kernel`` / ``bv_heatmaps_to_image_size``
implements - ``ImageTextInferenceEngine.convert_similarity_to_image_size(..., interpolation="nearest")`` of the reference
(health_multimodal/vlp/inference_engine.py:113-155) - in plain numpy loops, index by index, without torch.

Pinned: ``tests/test_scorer_pins_cpu.py::test_heatmap_oracle_matches_reference_outputs`` checks it bit for bit against the
outputs the REFERENCE's own function produced in the build container (``tests/golden/vlp_golden.npz``, written by
``oracle/make_golden_scorer.py``).  Only tests may import this module.

Arithmetic restated:
* the patch grid covers the centre-crop square of ``int(crop_size * min(height, width) / resize_size)`` original pixels
  (``crop_size`` pixels without a resize, the whole image without a crop)              reference :133-137, :148-153
* ``F.interpolate(mode="nearest")``: source index of destination index ``d`` is ``d`` when the sizes are equal, ``d >> 1`` when
  the destination is exactly twice the source, else ``min(int(floorf(d * (float(src) / dst))), src - 1)`` in fp32
  (ATen ``nearest_idx`` / ``compute_scales_value``)                                     reference :138-143
* ``F.pad(value=NaN)`` with margins ``(floor(mw / 2), ceil(mw / 2), floor(mh / 2), ceil(mh / 2))``; negative margins crop
                                                                                        reference :144-146
"""
import math

import numpy as np


def nearest_src_index(dst: int, src_size: int, dst_size: int) -> int:
    if dst_size == src_size:
        return dst
    if dst_size == 2 * src_size:
        return dst >> 1
    scale = np.float32(src_size) / np.float32(dst_size)
    return min(int(np.floor(np.float32(dst) * scale)), src_size - 1)


def heatmap_to_image_size(grid: np.ndarray, width: int, height: int, resize_size, crop_size) -> np.ndarray:
    """``grid`` [gh, gw] -> [height, width] float32, NaN outside the centre-crop square."""
    gh, gw = grid.shape
    side_h, side_w, top, left = height, width, 0, 0
    if crop_size:
        side = int(crop_size * min(height, width) / resize_size) if resize_size else int(crop_size)
        side_h = side_w = side
        left = math.floor((width - side) / 2)
        top = math.floor((height - side) / 2)
    out = np.full((height, width), np.nan, dtype=np.float32)
    sx = [nearest_src_index(x - left, gw, side_w) if 0 <= x - left < side_w else -1 for x in range(width)]
    for y in range(height):
        yy = y - top
        if not 0 <= yy < side_h:
            continue
        row = grid[nearest_src_index(yy, gh, side_h)]
        for x in range(width):
            if sx[x] >= 0:
                out[y, x] = row[sx[x]]
    return out


def heatmaps_to_image_size(heat: np.ndarray, width: int, height: int, resize_size, crop_size) -> np.ndarray:
    """``heat`` [B, gh, gw, L] (channel-last, as the CUDA path writes it) -> [B, L, height, width]."""
    B, gh, gw, L = heat.shape
    out = np.empty((B, L, height, width), dtype=np.float32)
    for b in range(B):
        for l in range(L):
            out[b, l] = heatmap_to_image_size(heat[b, :, :, l], width, height, resize_size, crop_size)
    return out

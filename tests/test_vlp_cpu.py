"""``ImageTextInferenceEngine.convert_similarity_to_image_size`` (reference health_multimodal/vlp/inference_engine.py:113-155)
against an independent index-arithmetic restatement: nearest-neighbour upsampling of the patch grid to the un-cropped
square of the ORIGINAL image, NaN in the margins the centre crop removed."""
import math

import numpy as np
import pytest
import torch

from incremental_multimodal_medical_learning_ii_b200.vlp.inference_engine import ImageTextInferenceEngine as E


def _expected(grid, width, height, resize, crop):
    g = grid.numpy()
    if crop is None:
        ys = np.minimum((np.arange(height) * (g.shape[0] / height)).astype(np.int64), g.shape[0] - 1)
        xs = np.minimum((np.arange(width) * (g.shape[1] / width)).astype(np.int64), g.shape[1] - 1)
        return g[np.ix_(ys, xs)]
    side = int(crop * min(height, width) / resize) if resize is not None else crop
    idx_y = np.minimum((np.arange(side) * (g.shape[0] / side)).astype(np.int64), g.shape[0] - 1)
    idx_x = np.minimum((np.arange(side) * (g.shape[1] / side)).astype(np.int64), g.shape[1] - 1)
    up = g[np.ix_(idx_y, idx_x)]
    out = np.full((height, width), np.nan, dtype=np.float32)
    top, left = math.floor((height - side) / 2), math.floor((width - side) / 2)
    out[top:top + side, left:left + side] = up
    return out


@pytest.mark.parametrize("w,h,resize,crop", [(390, 320, 512, 480), (320, 390, 512, 480), (2828, 2320, 512, 480),
                                             (333, 333, None, 300), (640, 480, None, None), (512, 512, 512, 512)])
def test_convert_similarity_to_image_size(w, h, resize, crop):
    grid = torch.arange(15 * 15, dtype=torch.float32).reshape(15, 15) / 7.0
    got = E.convert_similarity_to_image_size(grid, width=w, height=h, resize_size=resize, crop_size=crop)
    exp = _expected(grid, w, h, resize, crop)
    assert got.shape == (h, w)
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(exp))


def test_alias_package_exposes_vlp():
    import health_multimodal.vlp as vlp
    from health_multimodal.vlp.inference_engine import ImageTextInferenceEngine
    assert vlp.ImageTextInferenceEngine is ImageTextInferenceEngine is E


def test_similarity_map_refuses_cpu_embeddings():
    with pytest.raises(RuntimeError):
        E._get_similarity_map_from_embeddings(torch.zeros(15, 15, 128), torch.zeros(1, 128))

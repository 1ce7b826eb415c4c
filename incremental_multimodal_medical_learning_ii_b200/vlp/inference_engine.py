"""``ImageTextInferenceEngine`` (reference ``health_multimodal/vlp/inference_engine.py:20-158``) on the B200 image path.

Same constructor, method names, argument meaning and assertions as the reference.  What changes underneath: the image
embeddings come from the sm_100a kernels, and the Gaussian smoothing of the patch similarity map
(``ndimage.gaussian_filter(map, sigma=(1.5, 1.5), order=0)``, reference :107-109) runs on the GPU
(``ImageModel.smooth_heatmaps`` -> ``bv_smooth_heatmaps``).  The text side is any object with the reference
``TextInferenceEngine`` interface (``.model.training`` and ``get_embeddings_from_prompt(prompts, normalize=...)``); CXR-BERT
itself is outside this repository's scope.  ``get_similarity_maps_from_tensor`` is the batched addition used for
config 5 (heat-maps for many frames x many labels in one pass).
"""
from __future__ import annotations

from math import ceil, floor
from pathlib import Path
from typing import Callable, List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from ..image.inference_engine import ImageInferenceEngine


class ImageTextInferenceEngine:
    """Functions related to joint image / text inference."""

    def __init__(self, image_inference_engine: ImageInferenceEngine, text_inference_engine) -> None:
        self.image_inference_engine = image_inference_engine
        self.text_inference_engine = text_inference_engine

    def _check_eval(self) -> None:
        assert not self.image_inference_engine.model.training
        assert not self.text_inference_engine.model.training

    @torch.no_grad()
    def get_similarity_score_from_raw_data(self, image_path: Path, query_text: Union[List[str], str]) -> float:
        """Cosine similarity between one image and one or more phrases; several phrases are averaged before the
        L2 normalisation (reference :31-57)."""
        self._check_eval()
        prompts = [query_text] if isinstance(query_text, str) else query_text
        image_embedding = self.image_inference_engine.get_projected_global_embedding(image_path)
        text_embedding = self.text_inference_engine.get_embeddings_from_prompt(prompts, normalize=False)
        assert text_embedding.shape[0] == len(prompts)
        text_embedding = F.normalize(text_embedding.mean(dim=0), dim=0, p=2).to(image_embedding.device)
        return (image_embedding @ text_embedding.t()).item()

    def get_similarity_map_from_raw_data(self, image_path: Path, query_text: str,
                                         interpolation: str = "nearest") -> np.ndarray:
        """Heat-map of patch x text similarities with the shape of the input image (reference :59-91)."""
        self._check_eval()
        assert isinstance(query_text, str)
        image_embedding, (width, height) = self.image_inference_engine.get_projected_patch_embeddings(image_path)
        text_embedding = self.text_inference_engine.get_embeddings_from_prompt(query_text)
        sim = self._get_similarity_map_from_embeddings(image_embedding, text_embedding.to(image_embedding.device))
        return self.convert_similarity_to_image_size(
            sim, width=width, height=height, resize_size=self.image_inference_engine.resize_size,
            crop_size=self.image_inference_engine.crop_size, val_img_transform=self.image_inference_engine.transform,
            interpolation=interpolation)

    @staticmethod
    def _get_similarity_map_from_embeddings(projected_patch_embeddings: torch.Tensor,
                                            projected_text_embeddings: torch.Tensor, sigma: float = 1.5) -> torch.Tensor:
        """Smoothed similarity map ``[H', W']`` of ``[H', W', D]`` patch embeddings against a ``[1, D]`` text embedding
        (reference :93-111).  CUDA inputs are smoothed on the GPU; CPU inputs raise (no CPU path in this package)."""
        n_h, n_w, feature_size = projected_patch_embeddings.shape
        assert feature_size == projected_text_embeddings.shape[1]
        assert projected_text_embeddings.shape[0] == 1
        assert projected_text_embeddings.dim() == 2
        if not projected_patch_embeddings.is_cuda:
            raise RuntimeError("the similarity map is computed on the GPU; pass CUDA embeddings")
        from ..image.model.model import ImageModel
        sim = (projected_patch_embeddings.reshape(-1, feature_size).float() @ projected_text_embeddings.float().t())
        sim = sim.reshape(1, n_h, n_w, 1).contiguous()
        return ImageModel.smooth_heatmaps(None, sim, sigma)[0, :, :, 0].cpu()

    @torch.no_grad()
    def get_similarity_maps_from_tensor(self, frames: torch.Tensor, text_embeddings: torch.Tensor,
                                        sigma: Optional[float] = 1.5,
                                        image_size: Optional[tuple] = None) -> torch.Tensor:
        """Batched: frames ``[B,1|3,H,W]`` on the model's device, ``text_embeddings`` ``[L, D]`` (un-normalised) ->
        similarity maps ``[B, H', W', L]``, smoothed with ``sigma`` (``None``: raw).  With ``image_size = (width, height)``
        of the ORIGINAL images the maps come back as ``[B, L, height, width]`` in their pixels, NaN outside the centre crop:
        ``convert_similarity_to_image_size`` (nearest) on the GPU for the whole batch (reference :113-155)."""
        model = self.image_inference_engine.model
        patches = model.get_patchwise_projected_embeddings(frames, normalize=True)
        t = F.normalize(text_embeddings.to(patches.device).float(), dim=-1)
        heat = (patches @ t.t()).contiguous()
        if sigma is not None:
            heat = model.smooth_heatmaps(heat, sigma)
        if image_size is None:
            return heat
        width, height = image_size
        return model.heatmaps_to_image_size(heat, width, height, self.image_inference_engine.resize_size,
                                            self.image_inference_engine.crop_size)

    @staticmethod
    def convert_similarity_to_image_size(similarity_map: torch.Tensor, width: int, height: int,
                                         resize_size: Optional[int], crop_size: Optional[int],
                                         val_img_transform: Optional[Callable] = None,
                                         interpolation: str = "nearest") -> np.ndarray:
        """Patch grid -> original image size (reference :113-155): undo the centre crop in the resized image's
        coordinates (``int(crop * min(h, w) / resize)`` pixels of the original), interpolate the grid to that square,
        and pad the uncovered margins with NaN; without a crop the grid is stretched over the whole image."""
        grid = similarity_map.reshape(1, 1, similarity_map.shape[0], similarity_map.shape[1])
        align = False if interpolation in ("linear", "bilinear", "bicubic", "trilinear") else None
        if crop_size is None:
            return F.interpolate(grid, size=(height, width), mode=interpolation, align_corners=align)[0, 0].numpy()
        side = int(crop_size * min(height, width) / resize_size) if resize_size is not None else crop_size
        out = F.interpolate(grid, size=(side, side), mode=interpolation, align_corners=align)[0, 0]
        margin_w, margin_h = width - side, height - side
        pads = (floor(margin_w / 2), ceil(margin_w / 2), floor(margin_h / 2), ceil(margin_h / 2))
        return F.pad(out, pads, value=float("NaN")).numpy()

"""Reference ``health_multimodal/image/utils.py``: default transform sizes and the engine factory.  The fork's
``get_biovil_resnet_inference()`` calls ``get_biovil_resnet()`` without its now-mandatory argument (utils.py:21) and
fails; here ``pretrained`` is an optional checkpoint path (there is no network to download from)."""
from __future__ import annotations

from .data.transforms import create_chest_xray_transform_for_inference
from .inference_engine import ImageInferenceEngine
from .model import get_biovil_resnet

TRANSFORM_RESIZE = 512
TRANSFORM_CENTER_CROP_SIZE = 480


def get_biovil_resnet_inference(pretrained=None) -> ImageInferenceEngine:
    image_model = get_biovil_resnet(pretrained)
    transform = create_chest_xray_transform_for_inference(resize=TRANSFORM_RESIZE,
                                                          center_crop_size=TRANSFORM_CENTER_CROP_SIZE)
    return ImageInferenceEngine(image_model=image_model, transform=transform)

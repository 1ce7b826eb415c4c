"""B200-native mirror of the reference's ``health_multimodal.image`` package (same public names)."""
from .inference_engine import ImageInferenceEngine
from .model import ImageModel, ImageModelOutput, ResnetType, get_biovil_resnet
from .utils import get_biovil_resnet_inference

__all__ = ["ImageModel", "ImageModelOutput", "ResnetType", "ImageInferenceEngine", "get_biovil_resnet",
           "get_biovil_resnet_inference"]

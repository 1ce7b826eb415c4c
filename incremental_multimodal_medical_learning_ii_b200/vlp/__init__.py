"""Joint image/text inference (reference ``health_multimodal/vlp``): only the image half and the similarity arithmetic run
here; the text engine is consumed through ``get_embeddings_from_prompt``."""
from .inference_engine import ImageTextInferenceEngine

__all__ = ["ImageTextInferenceEngine"]

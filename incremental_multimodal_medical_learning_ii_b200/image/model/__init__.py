from .model import BIOMED_VLP_CXR_BERT_SPECIALIZED, CXR_BERT_COMMIT_TAG, ImageModel, ImageModelOutput, ResnetType
from .model import get_biovil_resnet

__all__ = ["ImageModel", "ImageModelOutput", "ResnetType", "get_biovil_resnet", "CXR_BERT_COMMIT_TAG",
           "BIOMED_VLP_CXR_BERT_SPECIALIZED"]

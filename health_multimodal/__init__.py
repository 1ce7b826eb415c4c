"""Drop-in alias: ``health_multimodal.image`` resolves to the B200-native implementation, so the reference's scripts
(``chexpert-get-embedding.py:7``, ``test_first_emb.py:14``, ``trash/lower_bound_mcs.py:13``: ``from
health_multimodal.image import get_biovil_resnet``) run unchanged with this repository on ``PYTHONPATH``.  The image side and the image/text similarity arithmetic (``health_multimodal.vlp``) are provided: the text encoder (CXR-BERT) is outside the hot path and is consumed through its ``[P,128]`` outputs."""
import importlib
import pkgutil
import sys

__version__ = "0.1.3+b200"

# Everything that is NOT replaced here (``health_multimodal.text`` with CXR-BERT, ``health_multimodal.common``) keeps
# coming from the reference checkout further down ``sys.path``: its ``health_multimodal`` directory is appended to this
# package's search path, while ``.image`` and ``.vlp`` are pinned to the B200 implementation in ``sys.modules`` below.
__path__ = pkgutil.extend_path(__path__, __name__)

_IMPL = "incremental_multimodal_medical_learning_ii_b200.image"
_SUBMODULES = ("", ".model", ".model.model", ".model.resnet", ".model.modules", ".inference_engine", ".utils",
               ".data", ".data.transforms", ".data.io")
for _sub in _SUBMODULES:
    sys.modules[f"{__name__}.image{_sub}"] = importlib.import_module(_IMPL + _sub)
image = sys.modules[f"{__name__}.image"]
# joint image/text inference (reference health_multimodal/vlp): the image half + similarity arithmetic
for _sub in ("", ".inference_engine"):
    sys.modules[f"{__name__}.vlp{_sub}"] = importlib.import_module("incremental_multimodal_medical_learning_ii_b200.vlp" + _sub)
vlp = sys.modules[f"{__name__}.vlp"]

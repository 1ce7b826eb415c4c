"""Import the reference's OWN ``health_multimodal.image`` from /root/reference (TEST INFRASTRUCTURE ONLY).

Used in the build container to (1) prove ``oracle/biovil_oracle.py`` equals the reference and (2) generate
the golden vectors under ``tests/golden`` (``oracle/make_golden.py``).  ``/root/reference`` does not exist on
the GPU box, so nothing that runs there may call :func:`load_reference_image_model`; use
:func:`reference_available` to gate.

Nothing under /root/reference is modified.  Three in-memory patches make the fork importable with the
container's torchvision 0.26 and without network access:

1. ``torchvision.models.resnet.model_urls`` was removed in torchvision 0.15 but is imported at
   ``image/model/resnet.py:10`` -> provide a dummy dict.
2. ``pydicom`` / ``SimpleITK`` / ``skimage`` are imported by ``image/data/io.py:10-13`` but only used inside
   ``load_image`` -> register empty stub modules.
3. ``ImageEncoder._create_encoder`` hard-wires ``pretrained=True`` (``model.py:194``), which downloads ImageNet
   weights (``resnet.py:57-59``) -> make the download and the ``load_state_dict(None)`` it feeds no-ops.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BIOVIL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "health_multimodal", "image", "model", "model.py"))


def _import_reference_image_package():
    import torchvision.models.resnet as tvr

    if not hasattr(tvr, "model_urls"):
        tvr.model_urls = {"resnet18": "unused://", "resnet50": "unused://"}
    for name in ("pydicom", "SimpleITK", "skimage", "skimage.io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "skimage" in sys.modules and "skimage.io" in sys.modules:
        setattr(sys.modules["skimage"], "io", sys.modules["skimage.io"])

    # The repo ships its own drop-in ``health_multimodal`` alias package; make sure the name resolves to the
    # reference here by importing it under a private sys.modules snapshot.
    saved = {k: v for k, v in sys.modules.items() if k == "health_multimodal" or k.startswith("health_multimodal.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import health_multimodal.image.model.resnet as hres
        import health_multimodal.image as himage

        hres.load_state_dict_from_url = lambda url, progress=True: None
        _orig = hres.ResNetHIML.load_state_dict

        def _tolerant(self, state_dict, *a, **k):
            if state_dict is None:
                return None
            return _orig(self, state_dict, *a, **k)

        hres.ResNetHIML.load_state_dict = _tolerant
        ref_modules = {k: v for k, v in sys.modules.items()
                       if k == "health_multimodal" or k.startswith("health_multimodal.")}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "health_multimodal" or k.startswith("health_multimodal."):
                del sys.modules[k]
        sys.modules.update(saved)
    return himage, ref_modules


def load_reference_image_model(seed: int = 27):
    """Build the reference ``ImageModel`` exactly as ``chexpert-get-embedding.py:30-45`` does (seed 27,
    ``get_biovil_resnet`` -> ``.train(mode=False, my_freeze=True)`` -> ``.eval()``), with random init because
    the BioViL checkpoint cannot be downloaded offline."""
    import contextlib
    import io

    import torch

    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    himage, _ = _import_reference_image_package()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):       # the fork prints "joint_feature_size 128"
        model = himage.get_biovil_resnet(pretrained=None)
        model.train(mode=False, my_freeze=True)
        model.eval()
    return model

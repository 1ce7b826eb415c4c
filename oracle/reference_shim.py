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


class _AnyStub(types.ModuleType):
    """A module whose every attribute is a callable that returns another stub (matplotlib / playsound stand-in:
    ``plt.ion()``, ``plt.figure()``, ``fig.add_subplot()`` ... all become no-ops)."""

    class _Obj:
        def __call__(self, *a, **k):
            return _AnyStub._Obj()

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _AnyStub._Obj()

        def __iter__(self):
            return iter(())

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyStub._Obj()


def _torchmetrics_stub():
    """``torchmetrics`` is imported by Trainer.py:9 but is not installed and not pinned in requirements.txt.  The stub
    provides ``functional.pairwise_cosine_similarity`` with the published algorithm (rows divided by their L2 norm, plain
    division, ``x @ y.T``, no diagonal zeroing when ``y`` is given) - this one function is the only part of the
    Trainer scoring path that is NOT the reference's own code when the fixtures are generated."""
    import torch

    def pairwise_cosine_similarity(x, y=None, reduction=None, zero_diagonal=None):
        zero_diag = zero_diagonal if zero_diagonal is not None else (y is None)
        if y is None:
            y = x
        xn = x / torch.linalg.norm(x, ord=2, dim=1, keepdim=True)
        yn = y / torch.linalg.norm(y, ord=2, dim=1, keepdim=True)
        d = xn @ yn.T
        if zero_diag:
            d.fill_diagonal_(0)
        return d

    tm = types.ModuleType("torchmetrics")
    fn = types.ModuleType("torchmetrics.functional")
    fn.pairwise_cosine_similarity = pairwise_cosine_similarity
    tm.functional = fn
    return tm, fn


class _ReferenceModules:
    """Context manager: while active, ``health_multimodal`` / ``Trainer`` / ``DataRetrieval`` ... resolve to the
    REFERENCE's files (the repo ships its own drop-in ``health_multimodal`` alias, which is hidden meanwhile), with the
    in-memory patches of this module applied and the missing third-party packages stubbed.  On exit ``sys.modules`` and
    ``sys.path`` are restored, so the product package and the reference never mix."""

    OWN = ("health_multimodal", "Trainer", "DataRetrieval", "HeatMapPlotter", "models", "new_texts_prompts")
    STUBS = ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "playsound", "torchmetrics",
             "torchmetrics.functional")

    def _mine(self, k):
        return any(k == n or k.startswith(n + ".") for n in self.OWN + self.STUBS)

    def __enter__(self):
        import torchvision.models.resnet as tvr
        if not hasattr(tvr, "model_urls"):
            tvr.model_urls = {"resnet18": "unused://", "resnet50": "unused://"}
        self.saved = {k: v for k, v in sys.modules.items() if self._mine(k)}
        for k in self.saved:
            del sys.modules[k]
        for name in ("pydicom", "SimpleITK", "skimage", "skimage.io"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        setattr(sys.modules["skimage"], "io", sys.modules["skimage.io"])
        for name in self.STUBS[:5]:
            sys.modules[name] = _AnyStub(name)
        tm, fn = _torchmetrics_stub()
        sys.modules["torchmetrics"], sys.modules["torchmetrics.functional"] = tm, fn
        sys.path.insert(0, REFERENCE_ROOT)
        import health_multimodal.image.model.resnet as hres
        hres.load_state_dict_from_url = lambda url, progress=True: None
        if not getattr(hres.ResNetHIML.load_state_dict, "_tolerant", False):
            _orig = hres.ResNetHIML.load_state_dict

            def _tolerant(self_, state_dict, *a, **k):
                return None if state_dict is None else _orig(self_, state_dict, *a, **k)

            _tolerant._tolerant = True
            hres.ResNetHIML.load_state_dict = _tolerant
        return self

    def __exit__(self, *exc):
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if self._mine(k):
                del sys.modules[k]
        sys.modules.update(self.saved)
        return False


def reference_modules() -> "_ReferenceModules":
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    return _ReferenceModules()


class FakeTextEngine:
    """Stands in for the reference's ``TextInferenceEngine`` (CXR-BERT weights need the network): a fixed
    ``{prompt string: [128] embedding}`` table behind ``get_embeddings_from_prompt(prompts, normalize)``
    (text/inference_engine.py:51-70: un-normalised projected embeddings, optionally L2-normalised over dim 1)."""

    class _M:
        training = False

    def __init__(self, table):
        self.table = table
        self.model = self._M()
        self.calls = 0

    def get_embeddings_from_prompt(self, prompts, normalize=True, verbose=True):
        import torch
        import torch.nn.functional as F
        self.calls += 1
        prompts = [prompts] if isinstance(prompts, str) else list(prompts)
        e = torch.stack([self.table[p] for p in prompts])
        return F.normalize(e, dim=1) if normalize else e

    def to(self, device):
        return self


def load_reference_vlp_engine_class():
    """The reference's own ``health_multimodal.vlp.inference_engine.ImageTextInferenceEngine`` class object
    (vlp/inference_engine.py:21-158).  Its methods only touch the two engines handed to the constructor, so the class
    stays usable after the import context has been left."""
    with reference_modules():
        from health_multimodal.vlp.inference_engine import ImageTextInferenceEngine
    return ImageTextInferenceEngine


def load_reference_trainer_module(text_table, image_model=False, text_model=False, max_emb=False,
                                  train_logit_diff=True, pred_logit_diff=False):
    """Import the reference's ``Trainer.py`` (module object) with its module-level switches set (Trainer.py:41-56) and
    ``get_cxr_bert_inference`` replaced by a :class:`FakeTextEngine` over ``text_table``.  Zero-shot evaluation is the
    configuration without adapters (``IMAGE_MODEL = TEXT_MODEL = False``, comment at Trainer.py:39-40)."""
    with reference_modules():
        import Trainer as T
    T.IMAGE_MODEL, T.TEXT_MODEL, T.SHARED = image_model, text_model, False
    T.MAX_EMB, T.TRAIN_LOGIT_DIFF, T.PRED_LOGIT_DIFF = max_emb, train_logit_diff, pred_logit_diff
    T.get_cxr_bert_inference = lambda: FakeTextEngine(text_table)
    T.tqdm = lambda it, **k: it
    return T


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

"""B200-native ``ImageModel``: the reference's class surface (``health_multimodal/image/model/model.py``) in front of
the sm_100a kernels.

What is kept from the reference (file:line are the reference's):

* ``get_biovil_resnet(pretrained)`` (model.py:61-70) - ``str``/``Path``/``None``; ``TypeError`` otherwise (model.py:115-116).
* ``ImageModel`` attributes ``encoder`` (``.encoder`` = ResNet-50 trunk), ``projector`` (``.model`` indices 0,1,3),
  ``feature_size`` 2048, ``freeze_encoder``, ``classifier`` (None), a ``state_dict`` with the checkpoint's 328 keys,
  ``train(mode, my_freeze)`` (model.py:131-139), construction leaves the module in train mode (model.py:112).
* ``forward(x)`` returns the un-normalised ``[B,128]`` projected global embedding, usable as a plain tensor
  (``torch.cat`` in chexpert-get-embedding.py:74-79) - the fork's behaviour (model.py:154) - **and** it carries
  the upstream ``ImageModelOutput`` fields (model.py:79-85) as attributes, because ``ImageInferenceEngine``
  (inference_engine.py:81) and ``get_patchwise_projected_embeddings`` (model.py:171) read them.  The fork broke
  those two callers; here both styles work.
* ``get_patchwise_projected_embeddings(x, normalize)`` -> ``[B,H',W',128]`` (model.py:161-175).

What is different: the arithmetic runs in hand-written CUDA (``csrc/``) on folded/packed weights; the module is
inference-only (it raises in training mode and on CPU tensors - there is no eager fallback), nothing is
downloaded, and batched helpers (``set_prompts`` / ``embed_and_score``) expose the fused scorer.
"""
from __future__ import annotations

import ctypes
import enum
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from ... import _native as N
from ...packing import PackedWeights
from .modules import MLP
from .resnet import resnet50

MODEL_TYPE = "resnet50"
JOINT_FEATURE_SIZE = 128

BIOMED_VLP_CXR_BERT_SPECIALIZED = "microsoft/BiomedVLP-CXR-BERT-specialized"
CXR_BERT_COMMIT_TAG = "v1.1"
BIOVIL_IMAGE_WEIGHTS_NAME = "biovil_image_resnet50_proj_size_128.pt"

TypeImageEncoder = Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]


@enum.unique
class ResnetType(str, enum.Enum):
    RESNET18 = "resnet18"
    RESNET50 = "resnet50"


class ImageModelOutput(torch.Tensor):
    """The ``[B,128]`` projected global embedding *as a tensor*, plus the upstream output fields as attributes.

    ``projected_patch_embeddings`` ``[B,128,H',W']``, ``img_embedding`` ``[B,2048]`` and ``patch_embedding``
    ``[B,2048,H',W']`` are produced on first access by one more pass over the retained input (the default
    forward only materialises what the fork's callers use).  Any torch op on this object returns a plain tensor.
    """

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def _wrap(global_emb: torch.Tensor, model: "ImageModel", frames: Optional[torch.Tensor],
              extras: Optional[Dict[str, torch.Tensor]] = None) -> "ImageModelOutput":
        out = torch.Tensor._make_subclass(ImageModelOutput, global_emb, False)
        out._model = model
        out._frames = frames                       # None when the extras were computed eagerly
        out._frames_version = None if frames is None else frames._version
        out._extras = extras
        return out

    def _full(self) -> Dict[str, torch.Tensor]:
        if getattr(self, "_extras", None) is None:
            model, frames = getattr(self, "_model", None), getattr(self, "_frames", None)
            if model is None or frames is None:
                raise AttributeError("this tensor no longer carries ImageModel outputs")
            if frames._version != self._frames_version:
                # a staging / DataLoader buffer that was refilled: the extras would describe OTHER frames than the
                # embedding this tensor holds.  Fail instead of returning inconsistent fields.
                raise RuntimeError("the input batch was modified in place after forward(); read the extra outputs "
                                   "before reusing the buffer, or set model.eager_outputs = True")
            self._extras = model._run(frames, patch=True, pooled=True, trunk=True)
            self._frames = None                    # the input is not needed (and not kept alive) any longer
        return self._extras

    def release(self) -> "ImageModelOutput":
        """Drop the reference to the input batch (callers that keep many outputs, e.g. in a list)."""
        self._frames = None
        return self

    @property
    def projected_global_embedding(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)

    @property
    def projected_patch_embeddings(self) -> torch.Tensor:
        return self._full()["patch"].permute(0, 3, 1, 2)          # stored channel-last; reference layout B D H W

    @property
    def img_embedding(self) -> torch.Tensor:
        return self._full()["pooled"]

    @property
    def patch_embedding(self) -> torch.Tensor:
        return self._full()["trunk"].permute(0, 3, 1, 2).float()

    @property
    def class_logits(self):
        return None


class _Engine:
    """Native handle + packed weights + workspaces for one device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("the BioViL B200 path needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device
        self.lib = N.lib()
        self.weights = PackedWeights(state_dict, device)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            N.check(self.lib.bv_create(ctypes.byref(self.handle), self.weights.pointer(), device.index or 0))
        self._workspaces: Dict[Tuple[int, int, int, int], torch.Tensor] = {}
        self._static: Dict[tuple, dict] = {}          # forward_graphed: static buffers per call signature
        self._bad_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.num_labels = 0
        self._prompt_keep = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.bv_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def workspace(self, B: int, C: int, H: int, W: int) -> torch.Tensor:
        key = (B, C, H, W)
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = self.lib.bv_workspace_bytes(B, C, H, W)
            if nbytes == 0:
                raise ValueError(f"unsupported frame batch shape {(B, C, H, W)}: H and W must be multiples of 32")
            self._workspaces.clear()                   # keep one workspace alive at a time
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._workspaces[key] = ws
        return ws

    def set_prompts(self, prompts: torch.Tensor, heat_text: Optional[torch.Tensor]) -> None:
        L, two, P, D = prompts.shape
        assert two == 2 and D == JOINT_FEATURE_SIZE
        p = prompts.to(self.device, torch.float32).contiguous()
        ht = None if heat_text is None else heat_text.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_set_prompts(self.handle, N.ptr(p), L, P, N.ptr(ht),
                                            N.current_stream_handle(self.device)))
        self._prompt_keep = (p, ht)
        self.num_labels = L

    def forward(self, frames: torch.Tensor, outs: N.BvOutputs) -> None:
        B, C, H, W = frames.shape
        dtype = N.BV_DTYPE_U8 if frames.dtype == torch.uint8 else N.BV_DTYPE_F32
        ws = self.workspace(B, C, H, W)
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_forward(self.handle, N.ptr(frames), dtype, B, C, H, W, N.ptr(ws), ws.numel(),
                                        ctypes.byref(outs), N.current_stream_handle(self.device)))

    def quantize(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """fp32 frames [B,C,H,W] -> uint8 [B,1,H,W] when they are 8-bit data (k/255, identical channels), else None.
        One kernel + one 4-byte flag read (the only host synchronisation of a float-input call)."""
        B, C, H, W = x.shape
        u8 = torch.empty(B, 1, H, W, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_quantize_frames_f32(N.ptr(x), B, C, H, W, N.ptr(u8), N.ptr(self._bad_flag),
                                                    N.current_stream_handle(self.device)))
        return u8 if int(self._bad_flag.item()) == 0 else None

    def forward_graphed(self, frames: torch.Tensor, shapes, normalize_patch: bool) -> Dict[str, torch.Tensor]:
        """Small-batch path: static input / output buffers per call signature, the whole forward replayed as ONE CUDA
        graph launch (``bv_forward_graph``), results returned as copies of the static outputs (one packed buffer, one
        copy).  The reference's extraction loop calls the model with batch size 1 (chexpert-get-embedding.py:47-49):
        about 40 launches of a few microseconds each are launch-bound, a graph replay is not."""
        B, C, H, W = frames.shape
        key = (B, C, H, W, frames.dtype, tuple(sorted(shapes)), bool(normalize_patch), self.num_labels)
        st = self._static.get(key)
        if st is None:
            if len(self._static) >= 4:
                self._static.pop(next(iter(self._static)))
            offs, total = {}, 0
            for k, (shape, dt) in shapes.items():
                nbytes = int(torch.empty(0, dtype=dt).element_size())
                for d in shape:
                    nbytes *= d
                offs[k] = (total, nbytes)
                total += (nbytes + 255) // 256 * 256
            packed = torch.empty(max(total, 256), dtype=torch.uint8, device=self.device)
            views = {k: packed[o:o + n].view(shapes[k][1]).view(shapes[k][0]) for k, (o, n) in offs.items()}
            outs = N.BvOutputs()
            outs.normalize_patch = 1 if normalize_patch else 0
            for k, t in views.items():
                setattr(outs, _OUT_FIELD[k], t.data_ptr())
            st = {"frames": torch.empty_like(frames), "packed": packed, "offs": offs, "outs": outs,
                  "ws": torch.empty(self.lib.bv_workspace_bytes(B, C, H, W), dtype=torch.uint8, device=self.device)}
            self._static[key] = st
        st["frames"].copy_(frames)
        dtype = N.BV_DTYPE_U8 if frames.dtype == torch.uint8 else N.BV_DTYPE_F32
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_forward_graph(self.handle, N.ptr(st["frames"]), dtype, B, C, H, W, N.ptr(st["ws"]),
                                              st["ws"].numel(), ctypes.byref(st["outs"]),
                                              N.current_stream_handle(self.device)))
        result = st["packed"].clone()
        return {k: result[o:o + n].view(shapes[k][1]).view(shapes[k][0]) for k, (o, n) in st["offs"].items()}

    def score(self, emb: torch.Tensor) -> Dict[str, torch.Tensor]:
        B = emb.shape[0]
        L = self.num_labels
        sim = torch.empty(B, L, 2, dtype=torch.float32, device=self.device)
        prob = torch.empty(B, L, dtype=torch.float32, device=self.device)
        pred = torch.empty(B, L, dtype=torch.uint8, device=self.device)
        score = torch.empty(B, L, dtype=torch.float32, device=self.device)
        if B == 0:          # empty shard / empty batch: empty results, no launch (the C ABI rejects B <= 0)
            return {"sim": sim, "prob": prob, "pred": pred, "score": score}
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_score(self.handle, N.ptr(emb), B, N.ptr(sim), N.ptr(prob), N.ptr(pred), N.ptr(score),
                                      N.current_stream_handle(self.device)))
        return {"sim": sim, "prob": prob, "pred": pred, "score": score}

    def launches(self) -> int:
        return int(self.lib.bv_last_forward_launches(self.handle))

    def set_profile(self, enable: bool) -> None:
        N.check(self.lib.bv_set_profile(self.handle, 1 if enable else 0))

    def get_profile(self):
        """Per-launch ``(name, flops, bytes, ms)`` of the most recent forward (needs ``set_profile(True)``)."""
        buf = (N.BvLaunchInfo * 128)()
        n = self.lib.bv_get_profile(self.handle, buf, 128)
        if n < 0:
            N.check(n)
        return [(buf[i].name.decode(), buf[i].flops, buf[i].bytes, buf[i].ms) for i in range(n)]


class ImageEncoder(nn.Module):
    """Image encoder trunk (reference model.py:178-228).  ``forward`` returns the pooled embedding, or
    ``(patch_embeddings [B,2048,H',W'], pooled [B,2048])`` with ``return_patch_embeddings=True`` (model.py:197-205)."""

    def __init__(self, img_model_type: str):
        super().__init__()
        self.img_model_type = img_model_type
        self.encoder = self._create_encoder()
        self._owner = None                      # set by ImageModel (not a registered submodule)

    def _create_encoder(self, **kwargs: Any) -> nn.Module:
        supported = ResnetType.RESNET18, ResnetType.RESNET50
        if self.img_model_type not in supported:
            raise NotImplementedError(f"Image model type \"{self.img_model_type}\" must be in {supported}")
        if self.img_model_type == ResnetType.RESNET18:
            raise NotImplementedError("BioViL uses resnet50; the B200 path implements the Bottleneck trunk only")
        return resnet50(pretrained=True, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_owner"] = None                  # weak reference to the owning ImageModel: re-made by its __setstate__
        return state

    def forward(self, x: torch.Tensor, return_patch_embeddings: bool = False) -> TypeImageEncoder:
        owner = object.__getattribute__(self, "_owner")
        if owner is None:
            raise RuntimeError("ImageEncoder must be used through an ImageModel on the B200 path")
        res = owner()._run(x, pooled=True, trunk=return_patch_embeddings)
        if return_patch_embeddings:
            return res["trunk"].permute(0, 3, 1, 2).float(), res["pooled"]
        return res["pooled"]

    def reload_encoder_with_dilation(self, replace_stride_with_dilation=None) -> None:
        raise NotImplementedError("dilated trunk variants are not part of the BioViL hot path")


def get_encoder_output_dim(module: nn.Module) -> int:
    """The reference probes with a 1x3x32x32 forward (model.py:231-247); for the ResNet-50 trunk it is 512 * 4."""
    return 2048


class ImageModel(nn.Module):
    """Image encoder module (reference model.py:88-175), inference-only, executed by sm_100a kernels."""

    MAX_BATCH = 512          # frames per native call; larger batches are processed in chunks

    def __init__(self,
                 img_model_type: str,
                 joint_feature_size: int,
                 freeze_encoder: bool = False,
                 pretrained_model_path: Optional[Union[str, Path]] = None,
                 **downstream_classifier_kwargs: Any):
        super().__init__()
        if joint_feature_size != JOINT_FEATURE_SIZE:
            raise NotImplementedError("the B200 kernels are specialised for the 128-d BioViL joint space")
        self.encoder = ImageEncoder(img_model_type)
        self.feature_size = get_encoder_output_dim(self.encoder)
        self.projector = MLP(input_dim=self.feature_size, output_dim=joint_feature_size,
                             hidden_dim=joint_feature_size, use_1x1_convs=True)
        self.downstream_classifier_kwargs = downstream_classifier_kwargs
        if downstream_classifier_kwargs:
            raise NotImplementedError("downstream classifiers are outside the BioViL embedding hot path")
        self.classifier = None
        self.freeze_encoder = freeze_encoder
        import weakref
        self.encoder._owner = weakref.ref(self)
        self._engine: Optional[_Engine] = None
        self._engine_version = None
        self._version_tensors = None
        self._prompts = None
        self.eager_outputs = False      # forward() computes patch / pooled / trunk outputs eagerly instead of on access
        self.cuda_graphs = "auto"       # replay small batches as one CUDA graph launch: True, False or "auto" (B <= 32)
        self.train()

        if pretrained_model_path is not None:
            if not isinstance(pretrained_model_path, (str, Path)):
                raise TypeError(f"Expected a string or Path, got {type(pretrained_model_path)}")
            state_dict = torch.load(pretrained_model_path, map_location="cpu")
            self.load_state_dict(state_dict)

    # ---- reference surface --------------------------------------------------------------------------------
    def train(self, mode: bool = True, my_freeze: bool = False) -> Any:
        """Same signature as the fork (model.py:131-139)."""
        super().train(mode=mode)
        if my_freeze:
            self.encoder.train(mode=False)
            self.projector.train(mode=False)
        return self

    def forward(self, x: torch.Tensor) -> ImageModelOutput:
        if self.eager_outputs:                     # every upstream field in the same pass; the input is not retained
            res = self._run(x, patch=True, pooled=True, trunk=True)
            return ImageModelOutput._wrap(res["global"], self, None, res)
        res = self._run(x)
        return ImageModelOutput._wrap(res["global"], self, x)

    def create_downstream_classifier(self, **kwargs: Any):
        raise NotImplementedError("downstream classifiers are outside the BioViL embedding hot path")

    @torch.no_grad()
    def get_patchwise_projected_embeddings(self, input_img: torch.Tensor, normalize: bool) -> torch.Tensor:
        """``[B, H', W', 128]`` projected patch embeddings, L2-normalised over the last dim if ``normalize``."""
        assert not self.training, "This function is only implemented for evaluation mode"
        return self._run(input_img, patch=True, normalize_patch=bool(normalize), want_global=False)["patch"]

    # ---- batched scoring API (fused projector + L2 norm + cosine + pos/neg softmax) -------------------------
    def set_prompts(self, prompts: torch.Tensor, reduce: str = "mean",
                    heat_text: Optional[torch.Tensor] = None) -> None:
        """Install text-prompt embeddings ``[L, 2, P, 128]`` (0 = positive, 1 = negative, un-normalised: what
        ``Trainer.bert_forward_mean`` gets from CXR-BERT, Trainer.py:1657-1680).  ``reduce='mean'`` averages the P
        prompts first (Trainer.py:1665-1666); ``'max'`` keeps them and takes the max cosine (Trainer.py:1691-1694)."""
        if prompts.dim() == 3:
            prompts = prompts.unsqueeze(2)
        if reduce == "mean":
            reduced = prompts.float().mean(dim=2, keepdim=True)
        elif reduce == "max":
            reduced = prompts.float()
        else:
            raise ValueError(f"reduce must be 'mean' or 'max', got {reduce!r}")
        if heat_text is None:
            heat_text = prompts.float()[:, 0].mean(dim=1)          # vlp/inference_engine.py:52 - mean, then normalise
        self._prompts = (reduced, heat_text)
        if self._engine is not None:
            self._engine.set_prompts(reduced, heat_text)

    @torch.no_grad()
    def embed_and_score(self, frames: torch.Tensor, heat: bool = False, patch: bool = False) -> Dict[str, torch.Tensor]:
        """One pass: global embeddings + zero-shot scores (and optionally normalised patch embeddings / heat-maps)."""
        if self._prompts is None:
            raise RuntimeError("call set_prompts() first")
        return self._run(frames, score=True, heat=heat, patch=patch, normalize_patch=True)

    @torch.no_grad()
    def smooth_heatmaps(self, heat: torch.Tensor, sigma: float = 1.5) -> torch.Tensor:  # `self` is not used
        """Gaussian smoothing of ``[B,H',W',L]`` similarity maps on the GPU with the semantics of
        ``ndimage.gaussian_filter(map, sigma=(sigma, sigma), order=0)`` (vlp/inference_engine.py:107-109)."""
        if heat.dim() != 4 or not heat.is_cuda:
            raise ValueError("expected CUDA similarity maps [B, H', W', L]")
        heat = heat.float().contiguous()
        out = torch.empty_like(heat)
        B, gh, gw, L = heat.shape
        with torch.cuda.device(heat.device):
            N.check(N.lib().bv_smooth_heatmaps(N.ptr(heat), B, gh, gw, L, float(sigma), N.ptr(out),
                                               N.current_stream_handle(heat.device)))
        return out

    def heatmaps_to_image_size(self, heat: torch.Tensor, width: int, height: int, resize_size: Optional[int],
                               crop_size: Optional[int]) -> torch.Tensor:  # `self` is not used
        """``[B,H',W',L]`` similarity maps -> ``[B,L,height,width]`` in the original image's pixels on the GPU:
        ``ImageTextInferenceEngine.convert_similarity_to_image_size`` with its default ``interpolation="nearest"``
        (vlp/inference_engine.py:113-155; nearest upsampling over the centre-crop square, NaN outside it), batched."""
        if heat.dim() != 4 or not heat.is_cuda:
            raise ValueError("expected CUDA similarity maps [B, H', W', L]")
        heat = heat.float().contiguous()
        B, gh, gw, L = heat.shape
        out = torch.empty(B, L, int(height), int(width), dtype=torch.float32, device=heat.device)
        if B == 0 or L == 0:
            return out
        with torch.cuda.device(heat.device):
            N.check(N.lib().bv_heatmaps_to_image_size(N.ptr(heat), B, gh, gw, L, int(height), int(width), int(resize_size or 0),
                                                      int(crop_size or 0), N.ptr(out), N.current_stream_handle(heat.device)))
        return out

    @torch.no_grad()
    def score_embeddings(self, emb: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Score cached ``[B,128]`` embeddings (the ``Trainer.val/test`` path) against the installed prompts."""
        if self._prompts is None:
            raise RuntimeError("call set_prompts() first")
        eng = self._get_engine()
        return eng.score(emb.to(eng.device, torch.float32).contiguous())

    # ---- machinery ----------------------------------------------------------------------------------------
    GRAPH_MAX_BATCH = 32     # "auto": batches up to this size are launch-bound (about 40 launches of a few microseconds each)

    def _apply(self, fn, *args, **kwargs):
        self._engine = None
        self._version_tensors = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._engine = None
        self._version_tensors = None
        return super().load_state_dict(*args, **kwargs)

    def __getstate__(self):
        # the native handle (ctypes) and the packed weights are a cache derived from the parameters: never pickled / copied
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_version"] = None
        state["_version_tensors"] = None
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        import weakref
        self.encoder._owner = weakref.ref(self)

    def _param_version(self):
        # in-place updates of any parameter / buffer bump its version counter; the tensor list itself is cached (walking
        # the module tree costs ~0.2 ms, which is most of a batch-1 call) and rebuilt when the module is moved / reloaded
        if self._version_tensors is None:
            self._version_tensors = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in self._version_tensors)

    def _get_engine(self) -> _Engine:
        version = self._param_version()
        device = self._version_tensors[0].device
        if self._engine is None or self._engine.device != device or self._engine_version != version:
            self._version_tensors = None
            version = self._param_version()
            self._engine = _Engine(self.state_dict(), device)
            self._engine_version = version
            if self._prompts is not None:
                self._engine.set_prompts(*self._prompts)
        return self._engine

    def _prepare_frames(self, x: torch.Tensor, eng: "_Engine") -> torch.Tensor:
        device = eng.device
        if x.dim() != 4 or x.shape[1] not in (1, 3):
            raise ValueError(f"expected frames [B, 1|3, H, W], got {tuple(x.shape)}")
        if x.device != device:
            raise RuntimeError(f"input is on {x.device} but the model is on {device}")
        if x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError(f"frame size {x.shape[2]}x{x.shape[3]} must be a multiple of 32")
        if x.shape[0] == 0:
            return x[:, :1].contiguous()
        if x.dtype == torch.uint8:
            if x.shape[1] == 3:
                if not bool((x[:, :1] == x).all()):
                    raise ValueError("uint8 frames with three different channels are not a BioViL input")
                x = x[:, :1]
            return x.contiguous()
        x = x.float().contiguous()
        # Frames made by ToTensor + ExpandChannels (transforms.py:12-38) are k/255 with identical channels: send the
        # exact 8-bit integers through the single-channel stem instead of rounding k/255 to bf16.  One kernel converts
        # and validates (|x*255 - k| <= 1e-3 grey levels, 0 <= k <= 255, channels equal); one flag read decides.
        u8 = eng.quantize(x)
        if u8 is not None:
            return u8
        if x.shape[1] == 3 and bool((x[:, :1] == x).all()):
            return x[:, :1].contiguous()
        return x

    def _run(self, x: torch.Tensor, want_global: bool = True, patch: bool = False, normalize_patch: bool = False,
             pooled: bool = False, trunk: bool = False, score: bool = False, heat: bool = False):
        if self.training:
            raise RuntimeError("ImageModel (B200) is inference-only: call .eval() first (the reference's BatchNorm "
                               "statistics are folded into the convolutions)")
        eng = self._get_engine()
        dev = eng.device
        with torch.no_grad():
            frames = self._prepare_frames(x, eng)
            B, C, H, W = frames.shape
            gh, gw = H // 32, W // 32
            L = eng.num_labels
            shapes = {}
            if want_global or score:
                shapes["global"] = ((B, JOINT_FEATURE_SIZE), torch.float32)
            if patch:
                shapes["patch"] = ((B, gh, gw, JOINT_FEATURE_SIZE), torch.float32)
            if pooled:
                shapes["pooled"] = ((B, 2048), torch.float32)
            if trunk:
                shapes["trunk"] = ((B, gh, gw, 2048), torch.bfloat16)
            if score:
                shapes["sim"] = ((B, L, 2), torch.float32)
                shapes["prob"] = ((B, L), torch.float32)
                shapes["pred"] = ((B, L), torch.uint8)
                shapes["score"] = ((B, L), torch.float32)
            if heat:
                shapes["heat"] = ((B, gh, gw, L), torch.float32)
            if B == 0:      # an empty batch gives empty outputs (what the reference's torch ops do), no launch
                return {k: torch.empty(shape, dtype=dt, device=dev) for k, (shape, dt) in shapes.items()}
            graphs = self.cuda_graphs
            if (graphs is True or (graphs == "auto" and B <= self.GRAPH_MAX_BATCH)) and B <= self.MAX_BATCH:
                return eng.forward_graphed(frames, shapes, normalize_patch)
            res = {k: torch.empty(shape, dtype=dt, device=dev) for k, (shape, dt) in shapes.items()}
            for b0 in range(0, B, self.MAX_BATCH):
                b1 = min(B, b0 + self.MAX_BATCH)
                outs = N.BvOutputs()
                outs.normalize_patch = 1 if normalize_patch else 0
                for k, t in res.items():
                    setattr(outs, _OUT_FIELD[k], t[b0:b1].data_ptr())
                eng.forward(frames[b0:b1], outs)
        return res


_OUT_FIELD = {"global": "global_emb", "patch": "patch_emb", "pooled": "pooled", "trunk": "trunk_nhwc_bf16",
              "sim": "sim", "prob": "prob", "pred": "pred", "score": "score", "heat": "heat"}


def get_biovil_resnet(pretrained) -> ImageModel:
    """Instantiate the BioViL image model from a local checkpoint path (or ``None`` for random init); the fork's
    signature (model.py:61-70).  Nothing is downloaded."""
    return ImageModel(img_model_type=MODEL_TYPE, joint_feature_size=JOINT_FEATURE_SIZE,
                      pretrained_model_path=pretrained)

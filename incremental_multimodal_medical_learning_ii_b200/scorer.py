"""Zero-shot scorer: the cosine / pos-neg decision of ``Trainer.val`` / ``Trainer.test`` on the fused CUDA kernel.

Reference semantics (Trainer.py): per label, ``bert_forward_mean`` gives the mean (un-normalised) CXR-BERT embedding of
the positive and of the negative prompts (:1657-1680); ``myCosineSimilarity`` is torchmetrics'
``pairwise_cosine_similarity`` of the image embeddings against it, or the max over per-prompt cosines when ``MAX_EMB``
(:1682-1704); then ``predicted = argmax([neg, pos])``, ``score = (pos+1)/2``, ``logit = pos - neg`` fed to
``BCEWithLogitsLoss`` (:805-844), i.e. P(pos) = sigmoid(pos - neg).

Only the no-grad evaluation path is replaced; training calls (``use_grad=True``, Trainer.py:569-572) need autograd and
stay on torch.  The text encoder is not part of this path: prompts arrive as ``[L, 2, P, 128]`` tensors.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _native as N


class ZeroShotScorer:
    """Scores cached ``[B,128]`` image embeddings on the GPU without needing the image model's weights."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ZeroShotScorer runs on a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = N.lib()
        self.handle = ctypes.c_void_p()
        self._weights = N.BvWeights()            # all-NULL: this handle can only score
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_create(ctypes.byref(self.handle), ctypes.byref(self._weights), self.device.index or 0))
        self.num_labels = 0
        self._keep = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.bv_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def set_prompts(self, prompts: torch.Tensor, reduce: str = "mean") -> None:
        if prompts.dim() == 3:
            prompts = prompts.unsqueeze(2)
        p = prompts.float()
        if reduce == "mean":
            p = p.mean(dim=2, keepdim=True)
        elif reduce != "max":
            raise ValueError(f"reduce must be 'mean' or 'max', got {reduce!r}")
        p = p.to(self.device).contiguous()
        L, two, P, D = p.shape
        assert two == 2 and D == 128
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_set_prompts(self.handle, N.ptr(p), L, P, None, N.current_stream_handle(self.device)))
        self._keep = p
        self.num_labels = L

    @torch.no_grad()
    def score(self, emb: torch.Tensor) -> Dict[str, torch.Tensor]:
        if self.num_labels == 0:
            raise RuntimeError("call set_prompts() first")
        emb = emb.to(self.device, torch.float32).contiguous()
        B, L = emb.shape[0], self.num_labels
        out = {"sim": torch.empty(B, L, 2, dtype=torch.float32, device=self.device),
               "prob": torch.empty(B, L, dtype=torch.float32, device=self.device),
               "pred": torch.empty(B, L, dtype=torch.uint8, device=self.device),
               "score": torch.empty(B, L, dtype=torch.float32, device=self.device)}
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_score(self.handle, N.ptr(emb), B, N.ptr(out["sim"]), N.ptr(out["prob"]),
                                      N.ptr(out["pred"]), N.ptr(out["score"]), N.current_stream_handle(self.device)))
        out["logit"] = out["sim"][..., 0] - out["sim"][..., 1]
        return out


_default_scorer: Optional[ZeroShotScorer] = None


@torch.no_grad()
def my_cosine_similarity(x: torch.Tensor, y: torch.Tensor, use_grad: bool = False, to_plot: bool = False,
                         train: bool = False, pos: bool = True, max_emb: bool = False) -> torch.Tensor:
    """Drop-in for ``Trainer.myCosineSimilarity(x, y, ...)`` (Trainer.py:1682-1704) in no-grad mode.

    ``x`` [B,128] image embeddings, ``y`` [1,128] (mean prompt embedding) or [P,128] with ``max_emb`` (max over the P
    per-prompt cosines).  Returns ``[B,1]`` like the reference."""
    global _default_scorer
    if use_grad:
        raise RuntimeError("the CUDA scorer is inference-only; keep training-time cosines on torch autograd")
    dev = x.device
    if _default_scorer is None or _default_scorer.device != dev:
        _default_scorer = ZeroShotScorer(dev)
    y = y.reshape(-1, y.shape[-1])
    if not max_emb and y.shape[0] != 1:
        raise ValueError("without max_emb the reference compares against a single [1,128] embedding")
    prompts = torch.stack([y, y], dim=0).unsqueeze(0)            # [1, 2, P, 128]: same vectors as pos and neg
    _default_scorer.set_prompts(prompts, reduce="max")
    return _default_scorer.score(x)["sim"][:, 0, :1].contiguous()


class TrainerEvalScorer:
    """The whole label loop of ``Trainer.val`` / ``Trainer.test`` (Trainer.py:797-837, 1019-1047) as ONE kernel launch.

    The reference re-runs CXR-BERT on every label's prompts for every batch (``bert_forward_mean`` inside the batch
    loop, Trainer.py:816 / 1030) although the prompt embeddings are constant during evaluation, then launches
    2 x L ``pairwise_cosine_similarity`` calls.  Here the ``[L,2,P,128]`` prompt embeddings are installed once and a
    batch of cached image embeddings gives ``predicted_labels``, ``tmp_score`` and ``logits`` exactly as the loop
    fills them:

    * ``train_logit_diff`` (Trainer.py:52, 809-814): False compares against the positive prompts only (the reference
      passes the positive prompts as "negative" too, so ``logits = pos`` and ``predicted_labels = argmax([pos, pos]) = 0``);
    * ``pred_logit_diff`` (Trainer.py:53, 824-827): ``tmp_score = (pos+1)/2`` or ``(pos-neg+2)/4``;
    * ``max_emb`` (Trainer.py:49, 1691-1694): max over per-prompt cosines instead of the cosine to the mean prompt.
    """

    def __init__(self, prompts: torch.Tensor, device="cuda:0", train_logit_diff: bool = True,
                 pred_logit_diff: bool = False, max_emb: bool = False):
        if prompts.dim() == 3:
            prompts = prompts.unsqueeze(2)
        if not train_logit_diff:
            prompts = torch.stack([prompts[:, 0], prompts[:, 0]], dim=1)   # the "trick" of Trainer.py:814
        self.train_logit_diff, self.pred_logit_diff = train_logit_diff, pred_logit_diff
        self._scorer = ZeroShotScorer(device)
        self._scorer.set_prompts(prompts, reduce="max" if max_emb else "mean")

    @torch.no_grad()
    def __call__(self, embs: torch.Tensor) -> Dict[str, torch.Tensor]:
        r = self._scorer.score(embs)
        pos, neg = r["sim"][..., 0], r["sim"][..., 1]
        tmp_score = (pos - neg + 2) / 4 if self.pred_logit_diff else r["score"]
        logits = r["logit"] if self.train_logit_diff else pos
        return {"predicted_labels": r["pred"].float(), "tmp_score": tmp_score, "logits": logits, "sim": r["sim"]}

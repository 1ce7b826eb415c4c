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
from typing import Dict

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
        if B == 0:
            out = {"sim": emb.new_empty(0, L, 2), "prob": emb.new_empty(0, L), "pred": torch.empty(0, L, dtype=torch.uint8, device=self.device),
                   "score": emb.new_empty(0, L)}
            out["logit"] = out["sim"][..., 0] - out["sim"][..., 1]
            return out
        out = {"sim": torch.empty(B, L, 2, dtype=torch.float32, device=self.device),
               "prob": torch.empty(B, L, dtype=torch.float32, device=self.device),
               "pred": torch.empty(B, L, dtype=torch.uint8, device=self.device),
               "score": torch.empty(B, L, dtype=torch.float32, device=self.device)}
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_score(self.handle, N.ptr(emb), B, N.ptr(out["sim"]), N.ptr(out["prob"]),
                                      N.ptr(out["pred"]), N.ptr(out["score"]), N.current_stream_handle(self.device)))
        out["logit"] = out["sim"][..., 0] - out["sim"][..., 1]
        return out


@torch.no_grad()
def my_cosine_similarity(x: torch.Tensor, y: torch.Tensor, use_grad: bool = False, to_plot: bool = False,
                         train: bool = False, pos: bool = True, max_emb: bool = False) -> torch.Tensor:
    """Drop-in for ``Trainer.myCosineSimilarity(x, y, ...)`` (Trainer.py:1682-1704) in no-grad mode.

    ``x`` [B,128] image embeddings, ``y`` [128] / [1,128] (mean prompt embedding) or [P,128] with ``max_emb`` (max over
    the P per-prompt cosines, :1691-1694).  Returns ``[B,1]`` (``[B]`` with ``max_emb``, as ``torch.max(res, dim=1)``
    does in the reference); ``to_plot`` compares one vector with one vector (:1687-1688) -> ``[1,1]``.
    One stateless kernel launch (``bv_pairwise_cosine``): no prompt installation, no allocation besides the result."""
    if use_grad:
        raise RuntimeError("the CUDA scorer is inference-only; keep training-time cosines on torch autograd")
    if not x.is_cuda:
        raise RuntimeError("my_cosine_similarity runs on a CUDA device (sm_100a); there is no CPU fallback")
    dev = x.device
    if to_plot:
        x = x.reshape(1, -1)
        y = y.reshape(1, -1)
    else:
        y = y.reshape(1, -1) if not max_emb else y.reshape(-1, y.shape[-1])
    x = x.to(torch.float32).contiguous()
    y = y.to(dev, torch.float32).contiguous()
    if x.dim() != 2 or x.shape[1] != 128 or y.shape[1] != 128:
        raise ValueError(f"expected [B,128] embeddings against [P,128] prompts, got {tuple(x.shape)} and {tuple(y.shape)}")
    B, P = x.shape[0], y.shape[0]
    reduce_max = bool(max_emb and not to_plot)
    out = torch.empty((B,) if reduce_max else (B, P), dtype=torch.float32, device=dev)
    if B == 0:
        return out
    with torch.cuda.device(dev):
        N.check(N.lib().bv_pairwise_cosine(N.ptr(x), N.ptr(y), B, P, 1 if reduce_max else 0, N.ptr(out),
                                           N.current_stream_handle(dev)))
    return out


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


def _trainer_switches(trainer):
    """The module-level switches of the reference's ``Trainer.py`` (:41-56) as the trainer's own module holds them."""
    mod = type(trainer).__init__.__globals__             # the globals of the module that defines the class
    return mod, {k: bool(mod[k]) for k in ("IMAGE_MODEL", "TEXT_MODEL", "MAX_EMB", "TRAIN_LOGIT_DIFF", "PRED_LOGIT_DIFF")}


@torch.no_grad()
def trainer_prompt_tensor(trainer) -> torch.Tensor:
    """``[L,2,P,128]`` prompt embeddings of a reference ``Trainer``: ``bert_forward_mean`` (Trainer.py:1657-1680, text
    adapter included) called ONCE per label instead of once per label and batch (:816, :1030) - in evaluation the
    prompts and both adapters are constant.  Mean-reduced prompts arrive as ``[128]`` (P = 1); with ``MAX_EMB`` the
    per-prompt rows are kept and ragged prompt lists are padded by repeating a row (the max is unchanged)."""
    _, sw = _trainer_switches(trainer)
    rows = []
    for label_name in trainer.class_names:
        pos_prompt = trainer.prompts[label_name]["positive"]
        neg_prompt = trainer.prompts[label_name]["negative"] if sw["TRAIN_LOGIT_DIFF"] else pos_prompt   # :809-814
        pos_e, neg_e = trainer.bert_forward_mean(pos_prompt, neg_prompt, use_grad=False)
        rows.append([e.reshape(-1, e.shape[-1]).float() for e in (pos_e, neg_e)])
    P = max(e.shape[0] for pair in rows for e in pair)
    pad = lambda e: torch.cat([e, e[:1].expand(P - e.shape[0], -1)]) if e.shape[0] < P else e   # noqa: E731
    return torch.stack([torch.stack([pad(p), pad(n)]) for p, n in rows])


def patch_trainer_eval(trainer, scorer_factory=None):
    """Make a reference ``Trainer`` instance evaluate on the fused scorer: ``trainer.val`` / ``trainer.test`` keep their
    signatures and everything around the label loop (adapters, ``change_labels``, criterion, TensorBoard scalar,
    ``evaluate_model``, the plots after ``test``: Trainer.py:773-866, 989-1072), but the 2 x L ``myCosineSimilarity``
    calls and the 2 x L CXR-BERT forwards per batch become one prompt installation per call and one kernel launch per
    batch (``TrainerEvalScorer``).  Returns the trainer.

    ``scorer_factory(prompts, device, train_logit_diff, pred_logit_diff, max_emb)`` builds the per-batch scorer; the
    default is the CUDA ``TrainerEvalScorer`` (no CPU fallback).  Tests inject the oracle here to check this plumbing
    against the unpatched reference loop on a machine without a GPU."""
    mod, _ = _trainer_switches(trainer)
    factory = scorer_factory or (lambda prompts, device, tld, pld, mx: TrainerEvalScorer(prompts, device, tld, pld, mx))

    def _evaluate(loader, criterion, epoch, split, desc):
        _, sw = _trainer_switches(trainer)              # read at call time: scripts flip the switches between runs
        if sw["IMAGE_MODEL"]:
            trainer.image_adapter.eval()
        if sw["TEXT_MODEL"]:
            trainer.text_adapter.eval()
        scorer = factory(trainer_prompt_tensor(trainer), trainer.device, sw["TRAIN_LOGIT_DIFF"], sw["PRED_LOGIT_DIFF"],
                         sw["MAX_EMB"])
        y_true, y_pred, y_score = [], [], []
        for batch_idx, (embs, labels) in enumerate(mod["tqdm"](loader, desc=desc), start=1):
            embs, labels = embs.to(trainer.device), labels.to(trainer.device)
            new_embs = trainer.image_adapter(embs) if sw["IMAGE_MODEL"] else embs
            r = scorer(new_embs)
            kept = labels
            if criterion is not None:                   # val only (:839-848)
                if trainer.change_labels:
                    labels = mod["change_values"](labels)
                loss = criterion(r["logits"], labels)
                if trainer.writer is not None:
                    trainer.writer.add_scalar(f"{split}/Loss", loss.item(), (epoch - 1) * len(loader) + batch_idx)
            y_true.append(kept.cpu().numpy())
            y_pred.append(r["predicted_labels"].cpu().numpy())
            y_score.append(r["tmp_score"].cpu().numpy())
        import numpy as np
        return np.concatenate(y_true), np.concatenate(y_pred), np.concatenate(y_score), sw

    @torch.no_grad()
    def val(val_loader, criterion, epoch, epochs, mode="joint", tasks_order=None):
        y_true, y_pred, y_score, _ = _evaluate(val_loader, criterion, epoch, "val",
                                               "Validating on chexpert mode: " + mode + ", Epoch " + str(epoch))
        trainer.evaluate_model(y_true, y_pred, y_score, mode, epoch, "val", epochs, tasks_order)

    @torch.no_grad()
    def test(test_loader, criterion, epoch, epochs, mode="joint", tasks_order=None, plot_tsne_array=None):
        y_true, y_pred, y_score, sw = _evaluate(test_loader, None, epoch, "test", "Testing on chexpert mode: " + mode)
        trainer.evaluate_model(y_true, y_pred, y_score, mode, epoch, "test", epochs, tasks_order)
        if sw["TRAIN_LOGIT_DIFF"]:                      # :1062-1072
            trainer.plot_cosine_similarity_text_embs(epoch, epochs)
        else:
            trainer.plot_cosine_similarity_text_embs_only_pos_prompts(epoch, epochs)
        trainer.plot_new_text_embeddings(epoch, epochs)
        if plot_tsne_array is not None:
            trainer.plot_tsne_sani_malati(plot_tsne_array[1], epoch, epochs)
            trainer.plot_tsne_multiclass(plot_tsne_array[0], epoch, epochs)

    trainer.val, trainer.test = val, test
    return trainer

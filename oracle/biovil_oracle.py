"""CPU fp32 restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

This module is the parity oracle for the B200 path.  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package never imports anything under ``oracle/``.

Parity status: **pinned against the reference itself run in the build
container** (``oracle/reference_shim.py`` imports ``/root/reference``'s own
``health_multimodal.image`` and ``oracle/make_golden.py`` asserts that this
restatement reproduces it bit-for-bit on the same state_dict and frames; the
outputs are committed under ``tests/golden/``).  The reference ships no tests,
golden vectors or KATs of its own for this path, so there
is nothing else to pin against.

Everything is plain ``torch.nn.functional`` on CPU tensors in fp32, written
from the reference files cited at each function (paths relative to
``/root/reference``).  The functions take a *state_dict* with the reference's
328 keys, never a module, so that the oracle cannot accidentally share code
with the product's ``ImageModel``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # torchvision BatchNorm2d default used by ResNet.__init__ (norm_layer=None)
LAYER_PLAN = (3, 4, 6, 3)   # health_multimodal/image/model/resnet.py:80  (Bottleneck, [3, 4, 6, 3])
LAYER_WIDTH = (64, 128, 256, 512)
EXPANSION = 4


def _bn(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """Eval-mode BatchNorm2d with running statistics (torch.nn.BatchNorm2d.forward, training=False)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, momentum=0.0, eps=BN_EPS)


def bottleneck(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, stride: int) -> torch.Tensor:
    """torchvision ``Bottleneck.forward`` (v1.5: the stride sits on the 3x3 conv).

    Called through ``ResNetHIML.forward`` -> ``self.layerN`` (image/model/resnet.py:39-42).
    conv1x1 -> BN -> ReLU -> conv3x3(stride) -> BN -> ReLU -> conv1x1 -> BN -> (+identity|downsample) -> ReLU.
    """
    identity = x
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        identity = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + identity)


def resnet_trunk(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "encoder.encoder.",
                 return_intermediate: bool = False):
    """``ResNetHIML.forward`` (image/model/resnet.py:25-47): stem, maxpool, layer1..4, returns x4 (no avgpool/fc)."""
    x0 = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)            # resnet.py:34
    x0 = F.relu(_bn(x0, sd, prefix + "bn1"))                                      # resnet.py:35-36
    x0 = F.max_pool2d(x0, kernel_size=3, stride=2, padding=1)                     # resnet.py:37
    feats = [x0]
    cur = x0
    for li, (nblocks, _w) in enumerate(zip(LAYER_PLAN, LAYER_WIDTH), start=1):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 1) else 1
            cur = bottleneck(cur, sd, f"{prefix}layer{li}.{bi}", stride)
        feats.append(cur)                                                         # resnet.py:39-42
    return tuple(feats) if return_intermediate else cur


def projector(patch_x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "projector.model.") -> torch.Tensor:
    """``MLP.forward`` with ``use_1x1_convs=True`` (image/model/modules.py:29-55).

    Conv2d(2048,128,1,bias=False) -> BatchNorm2d(128) -> ReLU -> Conv2d(128,128,1,bias=True), per patch.
    """
    h = F.conv2d(patch_x, sd[prefix + "0.weight"])
    h = F.relu(_bn(h, sd, prefix + "1"))
    return F.conv2d(h, sd[prefix + "3.weight"], sd[prefix + "3.bias"])


def image_model_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """``ImageModel.forward`` (image/model/model.py:141-154) plus the pieces of the upstream
    ``ImageModelOutput`` (model.py:79-85) that the fork's forward computes but drops.

    ``x`` is float32 [B,3,H,W] in [0,1] (image/data/transforms.py:37 - ToTensor then ExpandChannels).
    """
    assert x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3
    with torch.no_grad():
        x4 = resnet_trunk(x, sd)                                                  # model.py:200
        pooled = torch.flatten(F.adaptive_avg_pool2d(x4, (1, 1)), 1)              # model.py:201
        proj_patch = projector(x4, sd)                                            # model.py:144
        proj_global = torch.mean(proj_patch, dim=(2, 3))                          # model.py:145
    return {
        "projected_global_embedding": proj_global,        # what the fork's forward returns (model.py:154)
        "projected_patch_embeddings": proj_patch,         # [B,128,H',W']
        "img_embedding": pooled,                          # [B,2048]
        "patch_embedding": x4,                            # [B,2048,H',W']
    }


def patchwise_projected_embeddings(sd, x: torch.Tensor, normalize: bool) -> torch.Tensor:
    """``ImageModel.get_patchwise_projected_embeddings`` as upstream intends it (model.py:161-175):
    optional L2 normalisation over D (F.normalize default eps 1e-12), then B D H W -> B H W D."""
    p = image_model_forward(sd, x)["projected_patch_embeddings"]
    if normalize:
        p = F.normalize(p, dim=1)
    return p.permute(0, 2, 3, 1).contiguous()


def normalized_global_embedding(sd, x: torch.Tensor) -> torch.Tensor:
    """``ImageInferenceEngine.get_projected_global_embedding`` after the image is loaded
    (image/inference_engine.py:81-82): forward then ``F.normalize(dim=-1)``."""
    return F.normalize(image_model_forward(sd, x)["projected_global_embedding"], dim=-1)


# ----------------------------------------------------------------------------------------------
# Scorer
# ----------------------------------------------------------------------------------------------

def pairwise_cosine_similarity(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torchmetrics.functional.pairwise_cosine_similarity(x, y) restated (third-party, not vendored
    in /root/reference and not pinned in requirements.txt; call sites Trainer.py:1688-1692).

    Published algorithm: each row of x and y is divided by its L2 norm (plain division, no eps),
    then ``x @ y.T``; no diagonal zeroing when y is given.
    """
    xn = x / torch.linalg.norm(x, ord=2, dim=1, keepdim=True)
    yn = y / torch.linalg.norm(y, ord=2, dim=1, keepdim=True)
    return xn @ yn.T


def reduce_prompts(prompts: torch.Tensor, reduce: str) -> torch.Tensor:
    """``Trainer.bert_forward_mean`` prompt side (Trainer.py:1657-1680): mean over the P un-normalised
    prompt embeddings of one polarity when ``not MAX_EMB`` (Trainer.py:1665-1666,1675-1676).
    ``prompts`` is [L,2,P,D] with index 0 = positive, 1 = negative.  With ``reduce='max'`` the P
    prompts are kept (the max is taken over per-prompt cosines, Trainer.py:1691-1694)."""
    assert prompts.dim() == 4
    if reduce == "mean":
        return prompts.mean(dim=2, keepdim=True)
    assert reduce == "max"
    return prompts


def zero_shot_score(emb: torch.Tensor, prompts: torch.Tensor, reduce: str = "mean") -> Dict[str, torch.Tensor]:
    """Per-label decision of ``Trainer.val/test`` (Trainer.py:805-837, 1019-1047) with
    ``TRAIN_LOGIT_DIFF=True`` (Trainer.py:52) and ``PRED_LOGIT_DIFF=False`` (Trainer.py:53).

    emb [B,D] un-normalised image embeddings (what the extraction driver caches);
    prompts [L,2,P,D] un-normalised text embeddings (0 = positive, 1 = negative).
    Returns sim [B,L,2] (pos,neg cosine), logit = pos-neg, prob = sigmoid(logit) (= softmax over
    {pos,neg} at temperature 1, what BCEWithLogitsLoss sees at Trainer.py:844), score = (pos+1)/2
    (Trainer.py:825), score_diff = (pos-neg+2)/4 (Trainer.py:827), pred = argmax([neg,pos])
    (Trainer.py:836: 1 iff pos > neg, ties -> 0).
    """
    t = reduce_prompts(prompts.float(), reduce)
    B, L = emb.shape[0], t.shape[0]
    sim = torch.empty(B, L, 2, dtype=torch.float32)
    for l in range(L):                                    # the label loop of Trainer.py:806
        for pol in range(2):
            res = pairwise_cosine_similarity(emb.float(), t[l, pol])      # [B,P']
            sim[:, l, pol] = res.max(dim=1).values if reduce == "max" else res[:, 0]
    pos, neg = sim[..., 0], sim[..., 1]
    logit = pos - neg
    pred = torch.argmax(torch.stack([neg, pos], dim=-1), dim=-1).to(torch.uint8)
    return {"sim": sim, "logit": logit, "prob": torch.sigmoid(logit), "score": (pos + 1) / 2,
            "score_diff": (pos - neg + 2) / 4, "pred": pred}


def trainer_val_batch(embs: torch.Tensor, prompts: torch.Tensor, train_logit_diff: bool = True,
                      pred_logit_diff: bool = False, max_emb: bool = False) -> Dict[str, torch.Tensor]:
    """One batch of ``Trainer.val`` / ``Trainer.test`` restated line by line (Trainer.py:797-837, 1019-1047), with the
    module-level switches TRAIN_LOGIT_DIFF (:52), PRED_LOGIT_DIFF (:53) and MAX_EMB (:49) as arguments.
    ``prompts`` [L,2,P,D] stands for what ``get_embeddings_from_prompt(normalize=False)`` returns per label."""
    B, L = embs.shape[0], prompts.shape[0]
    predicted_labels = torch.zeros(B, L)
    tmp_score = torch.zeros(B, L)
    logits = torch.empty(B, L)

    def cos(x, y):                                           # Trainer.myCosineSimilarity, no-grad branch (:1682-1704)
        if not max_emb:
            return pairwise_cosine_similarity(x, y.reshape(1, -1))
        return torch.max(pairwise_cosine_similarity(x, y), dim=1).values

    for i in range(L):
        pos_prompt = prompts[i, 0]
        neg_prompt = prompts[i, 1] if train_logit_diff else prompts[i, 0]          # :809-814
        pos_e = pos_prompt if max_emb else pos_prompt.mean(dim=0)                  # bert_forward_mean :1665-1666
        neg_e = neg_prompt if max_emb else neg_prompt.mean(dim=0)
        pos_s, neg_s = cos(embs.float(), pos_e.float()), cos(embs.float(), neg_e.float())
        if not pred_logit_diff:
            tmp_score[:, i] = (pos_s.flatten() + 1) / 2
        else:
            tmp_score[:, i] = (pos_s.flatten() - neg_s.flatten() + 2) / 4
        logits[:, i] = pos_s.flatten() - neg_s.flatten() if train_logit_diff else pos_s.flatten()
        predicted_labels[:, i] = torch.argmax(torch.cat([neg_s.reshape(-1, 1), pos_s.reshape(-1, 1)], dim=1), dim=1)
    return {"predicted_labels": predicted_labels, "tmp_score": tmp_score, "logits": logits}


def patch_similarity_map(patch_emb_normalized: torch.Tensor, prompts_pos: torch.Tensor) -> torch.Tensor:
    """``ImageTextInferenceEngine._get_similarity_map_from_embeddings`` before smoothing
    (vlp/inference_engine.py:93-108) batched over images and labels.

    patch_emb_normalized [B,H',W',D] (L2-normalised over D); prompts_pos [L,P,D] un-normalised: the
    text side is mean over P then L2-normalise (vlp/inference_engine.py:52-53).  Returns [B,H',W',L]."""
    t = F.normalize(prompts_pos.float().mean(dim=1), dim=-1)             # [L,D]
    return patch_emb_normalized.float() @ t.T


def gaussian_smooth_map(sim_map: torch.Tensor, sigma: float = 1.5) -> torch.Tensor:
    """scipy.ndimage.gaussian_filter(order=0, mode='reflect', truncate=4.0) on the last two spatial dims
    of [.., H', W'] (vlp/inference_engine.py:109).  Used by the 'next' row for heat-map post-processing."""
    from scipy import ndimage
    import numpy as np
    a = np.ascontiguousarray(sim_map.detach().cpu().numpy())     # reshape below must be a view, not a copy
    out = np.empty_like(a)
    flat_in = a.reshape(-1, a.shape[-2], a.shape[-1])
    flat_out = out.reshape(-1, a.shape[-2], a.shape[-1])
    for i in range(flat_in.shape[0]):
        flat_out[i] = ndimage.gaussian_filter(flat_in[i], sigma=(sigma, sigma), order=0)
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------------------------
# Weights used by the golden vectors / parity tests
# ----------------------------------------------------------------------------------------------

def randomize_batchnorm_(sd: Dict[str, torch.Tensor], seed: int = 28) -> Dict[str, torch.Tensor]:
    """BN-randomised variant: a default-init BN is an identity, so a wrong BN fold would
    pass parity.  gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5), generator seed 28."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        if k.endswith(".running_mean"):
            base = k[: -len(".running_mean")]
            n = sd[k].numel()
            sd[base + ".weight"] = torch.rand(n, generator=g) + 0.5
            sd[base + ".bias"] = torch.randn(n, generator=g) * 0.1
            sd[base + ".running_mean"] = torch.randn(n, generator=g) * 0.1
            sd[base + ".running_var"] = torch.rand(n, generator=g) + 0.5
    return sd

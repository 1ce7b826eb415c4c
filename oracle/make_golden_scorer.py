"""Generate the scorer-side fixtures from the reference's OWN code (run HERE, where /root/reference exists):

    python oracle/make_golden_scorer.py          # after oracle/make_golden.py

* ``tests/golden/vlp_golden.npz``  - ``health_multimodal/vlp/inference_engine.py`` of the reference, imported through
  ``oracle/reference_shim.py``: ``get_similarity_score_from_raw_data`` (:31-57), ``_get_similarity_map_from_embeddings``
  (:93-111), ``convert_similarity_to_image_size`` (:113-155) and ``get_similarity_map_from_raw_data`` (:59-91), driven
  with the reference image model's embeddings (``tests/golden/biovil_golden.pt``) and a fixed prompt-embedding table in
  place of CXR-BERT (its weights need the network).
* ``tests/golden/trainer_golden.pt`` - the reference's ``Trainer.val`` (:773-866) and ``Trainer.test`` (:989-1072) run
  UNCHANGED on cached embeddings, for every combination of the module switches the scorer mirrors; what they hand to
  ``evaluate_model`` (``y_pred``, ``y_score``) and to the criterion (``logits``) is recorded.  Stubbed while importing:
  ``matplotlib`` / ``playsound`` (absent, plotting only), ``torchmetrics`` (absent and unpinned: its
  ``pairwise_cosine_similarity`` is restated from the published algorithm - the one residue that is not reference code).

The oracle restatement (``oracle/biovil_oracle.py``) is asserted equal to every recorded result, which pins it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import biovil_oracle as O  # noqa: E402
import reference_shim as RS  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CLASS_NAMES = ["Atelectasis", "Cardiomegaly", "Consolidation", "Edema", "Pleural Effusion"]    # Trainer.py:209
RESIZE_CASES = [(390, 320, 512, 480), (320, 390, 512, 480), (333, 333, None, 300), (200, 150, None, None),
                (512, 512, 512, 512), (301, 257, 256, 224)]


def prompt_table(L: int, P: int, seed: int):
    """``{string: [128] embedding}`` and the ``[L,2,P,128]`` tensor it was cut from (0 = positive, 1 = negative)."""
    t = FR.synthetic_prompt_embeddings(L, P, 128, seed=seed)
    table = {f"label {l} {'pos' if pol == 0 else 'neg'} {j}": t[l, pol, j] for l in range(L) for pol in range(2)
             for j in range(P)}
    return table, t


class _RefImageEngine:
    """What ``ImageInferenceEngine`` (image/inference_engine.py:58-87) hands to the VLP engine, computed from the
    reference image model's outputs: the fork's own engine cannot run (it reads ``.projected_global_embedding`` off the
    bare tensor the fork's ``forward`` returns), so lines :81-86 are applied to the stored outputs."""

    class _M:
        training = False

    def __init__(self, global_emb, patch_norm, size_wh, resize, crop):
        self.model = self._M()
        self.g, self.p, self.size = global_emb, patch_norm, size_wh
        self.resize_size, self.crop_size, self.transform = resize, crop, None

    def get_projected_global_embedding(self, image_path):
        e = F.normalize(self.g[int(image_path)][None], dim=-1)          # :82
        assert e.shape[0] == 1 and e.ndim == 2                          # :84-85
        return e[0]

    def get_projected_patch_embeddings(self, image_path):
        return self.p[int(image_path)], self.size                       # :66-70


def make_vlp(golden) -> None:
    E = RS.load_reference_vlp_engine_class()
    g = golden["bnrand/structured"]
    table, t = prompt_table(14, 5, seed=29)
    text = RS.FakeTextEngine(table)
    img = _RefImageEngine(g["global256"], g["patch_norm_first2"], (390, 320), 512, 480)
    eng = E(img, text)
    out = {}
    # (1) similarity score, one phrase and five phrases (mean before normalisation), 32 frames x 14 labels
    s1 = np.zeros((32, 14), np.float64)
    s5 = np.zeros((32, 14), np.float64)
    for i in range(32):
        for l in range(14):
            s1[i, l] = eng.get_similarity_score_from_raw_data(i, f"label {l} pos 0")
            s5[i, l] = eng.get_similarity_score_from_raw_data(i, [f"label {l} pos {j}" for j in range(5)])
    out["score_p1"], out["score_p5"] = s1, s5
    # the restated scorer's positive cosine is the same number
    zs = O.zero_shot_score(g["global256"][:32], t, "mean")["sim"][..., 0].double().numpy()
    assert np.abs(zs - s5).max() <= 1e-6, np.abs(zs - s5).max()
    zs1 = O.zero_shot_score(g["global256"][:32], t[:, :, :1], "mean")["sim"][..., 0].double().numpy()
    assert np.abs(zs1 - s1).max() <= 1e-6
    # (2) smoothed similarity maps of the two stored patch grids against 14 single prompts
    maps = np.zeros((2, 14, 15, 15), np.float32)
    for i in range(2):
        for l in range(14):
            te = text.get_embeddings_from_prompt(f"label {l} pos 0")                      # [1,128] normalised (:78)
            maps[i, l] = E._get_similarity_map_from_embeddings(g["patch_norm_first2"][i], te).numpy()
    out["smoothed_maps"] = maps
    raw = O.patch_similarity_map(g["patch_norm_first2"], t[:, 0, :1])                      # [2,15,15,14]
    sm = O.gaussian_smooth_map(raw.permute(0, 3, 1, 2), 1.5).numpy()
    assert np.abs(sm - maps).max() <= 1e-6, np.abs(sm - maps).max()
    # (3) patch grid -> image size, nearest and bilinear
    grid = torch.arange(15 * 15, dtype=torch.float32).reshape(15, 15) / 7.0
    for k, (w, h, rs, cs) in enumerate(RESIZE_CASES):
        for interp in ("nearest", "bilinear"):
            r = E.convert_similarity_to_image_size(grid, width=w, height=h, resize_size=rs, crop_size=cs,
                                                   interpolation=interp)
            out[f"resize_{k}_{interp}"] = r.astype(np.float32)
    out["resize_cases"] = np.array([[w, h, rs or 0, cs or 0] for (w, h, rs, cs) in RESIZE_CASES], np.int64)
    # (4) the whole get_similarity_map_from_raw_data flow
    out["full_map_0"] = eng.get_similarity_map_from_raw_data(0, "label 3 pos 0").astype(np.float32)
    out["full_map_1_bilinear"] = eng.get_similarity_map_from_raw_data(1, "label 7 pos 0", interpolation="bilinear").astype(np.float32)
    path = os.path.join(GOLDEN_DIR, "vlp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


class _Recorder:
    def __init__(self):
        self.logits, self.losses = [], []
        self.bce = torch.nn.BCEWithLogitsLoss()

    def __call__(self, logits, labels):
        self.logits.append(logits.detach().clone())
        loss = self.bce(logits, labels)
        self.losses.append(float(loss))
        return loss


def run_reference_trainer(embs, labels, P, single_prompt, max_emb, train_logit_diff, pred_logit_diff, which,
                          batch=256):
    """One pass of the reference's ``Trainer.val`` / ``Trainer.test`` over (embs, labels) in batches of ``batch``."""
    table, t = prompt_table(5, P, seed=37 + P)
    prompts = {name: {"positive": [f"label {l} pos {j}" for j in range(P)],
                      "negative": [f"label {l} neg {j}" for j in range(P)]} for l, name in enumerate(CLASS_NAMES)}
    T = RS.load_reference_trainer_module(table, max_emb=max_emb, train_logit_diff=train_logit_diff,
                                         pred_logit_diff=pred_logit_diff)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        tr = T.Trainer(single_prompt, prompts, CLASS_NAMES, "standard", 1e-3, torch.device("cpu"), None)
    seen = {}
    tr.evaluate_model = lambda y_true, y_pred, y_score, *a, **k: seen.update(y_true=y_true, y_pred=y_pred, y_score=y_score)
    for name in ("plot_cosine_similarity_text_embs", "plot_cosine_similarity_text_embs_only_pos_prompts",
                 "plot_new_text_embeddings"):
        setattr(tr, name, lambda *a, **k: None)
    loader = [(embs[i:i + batch], labels[i:i + batch]) for i in range(0, embs.shape[0], batch)]
    rec = _Recorder()
    if which == "val":
        tr.val(loader, rec, 1, 1)
    else:
        tr.test(loader, rec, 1, 1)
    res = {"y_pred": torch.from_numpy(seen["y_pred"]), "y_score": torch.from_numpy(seen["y_score"]), "prompts": t,
           "bert_calls": tr.bert_encoder.calls}
    if which == "val":
        res["logits"] = torch.cat(rec.logits)
        res["loss"] = torch.tensor(rec.losses)
    return res


def make_trainer(golden) -> None:
    g = golden["bnrand/structured"]
    gen = torch.Generator().manual_seed(5)
    # 256 reference image embeddings + 300 generic ones: 3 batches of 256 / 256 / 44 (ragged tail)
    embs = torch.cat([g["global256"], torch.randn(300, 128, generator=gen) * 3.0])
    labels = (torch.rand(embs.shape[0], 5, generator=gen) > 0.5).float()
    out = {"embs": embs, "labels": labels, "class_names": CLASS_NAMES, "cases": {}}
    cases = [("single", 1, True, False, True, False), ("mean4", 4, False, False, True, False),
             ("max4", 4, False, True, True, False), ("mean4_posonly", 4, False, False, False, False),
             ("mean4_preddiff", 4, False, False, True, True), ("max4_preddiff", 4, False, True, True, True)]
    for name, P, single, max_emb, tld, pld in cases:
        for which in ("val", "test"):
            r = run_reference_trainer(embs, labels, P, single, max_emb, tld, pld, which)
            # pin the line-by-line restatement to what the reference's own loop produced
            o = O.trainer_val_batch(embs, r["prompts"], tld, pld, max_emb)
            assert torch.equal(o["predicted_labels"], r["y_pred"]), (name, which)
            assert (o["tmp_score"] - r["y_score"]).abs().max() <= 1e-6, (name, which)
            if which == "val":
                assert (o["logits"] - r["logits"]).abs().max() <= 1e-6, (name, which)
            r.update(P=P, single_prompt=single, max_emb=max_emb, train_logit_diff=tld, pred_logit_diff=pld)
            out["cases"][f"{name}/{which}"] = r
            print(f"[trainer {name}/{which}] positives {int(r['y_pred'].sum())} of {r['y_pred'].numel()}, "
                  f"BERT forwards in the reference loop: {r['bert_calls']}")
    path = os.path.join(GOLDEN_DIR, "trainer_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path))


def main() -> None:
    torch.manual_seed(0)
    golden = torch.load(os.path.join(GOLDEN_DIR, "biovil_golden.pt"), map_location="cpu")
    make_vlp(golden)
    make_trainer(golden)


if __name__ == "__main__":
    main()

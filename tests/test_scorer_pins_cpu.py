"""Scorer-side parity pins (CPU half).

``tests/golden/vlp_golden.npz`` and ``tests/golden/trainer_golden.pt`` were produced by the REFERENCE's own code
(``health_multimodal/vlp/inference_engine.py`` and ``Trainer.val`` / ``Trainer.test`` run unchanged; see
``oracle/make_golden_scorer.py``).  Here, without a GPU:

* the oracle restatement is checked against those fixtures (so it stays pinned wherever the tests run);
* the host-side pieces of the product that need no GPU (``convert_similarity_to_image_size``, the prompt hoisting and
  the val/test plumbing of ``scorer.patch_trainer_eval``) are checked against the fixtures and - when ``/root/reference``
  is present - against the reference running live.
The CUDA kernels themselves are compared with the same fixtures in ``tests/test_scorer_pins_gpu.py``.
"""
import os

import numpy as np
import pytest
import torch

import biovil_oracle as O
import reference_shim as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vlp():
    return np.load(os.path.join(ROOT, "tests", "golden", "vlp_golden.npz"))


@pytest.fixture(scope="module")
def trainer_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "trainer_golden.pt"), map_location="cpu")


def _prompt_tensor(seed, L, P):
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    return FR.synthetic_prompt_embeddings(L, P, 128, seed=seed)


def test_oracle_scorer_matches_reference_vlp_engine(golden, vlp):
    g = golden["bnrand/structured"]
    t = _prompt_tensor(29, 14, 5)
    s5 = O.zero_shot_score(g["global256"][:32], t, "mean")["sim"][..., 0]
    s1 = O.zero_shot_score(g["global256"][:32], t[:, :, :1], "mean")["sim"][..., 0]
    assert np.abs(s5.double().numpy() - vlp["score_p5"]).max() <= 1e-6       # vlp/inference_engine.py:31-57
    assert np.abs(s1.double().numpy() - vlp["score_p1"]).max() <= 1e-6
    raw = O.patch_similarity_map(g["patch_norm_first2"], t[:, 0, :1])         # [2,15,15,14]
    sm = O.gaussian_smooth_map(raw.permute(0, 3, 1, 2), 1.5).numpy()
    assert np.abs(sm - vlp["smoothed_maps"]).max() <= 1e-6                    # vlp/inference_engine.py:93-111


def test_convert_similarity_to_image_size_matches_reference(vlp):
    from incremental_multimodal_medical_learning_ii_b200.vlp.inference_engine import ImageTextInferenceEngine as E
    grid = torch.arange(15 * 15, dtype=torch.float32).reshape(15, 15) / 7.0
    for k, (w, h, rs, cs) in enumerate(vlp["resize_cases"].tolist()):
        for interp in ("nearest", "bilinear"):
            got = E.convert_similarity_to_image_size(grid, width=w, height=h, resize_size=rs or None,
                                                     crop_size=cs or None, interpolation=interp)
            ref = vlp[f"resize_{k}_{interp}"]
            assert got.shape == ref.shape == (h, w)
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref)), (k, interp)   # vlp/inference_engine.py:113-155


def test_heatmap_oracle_matches_reference_outputs(vlp):
    """oracle/heatmap_oracle.py (the restatement the CUDA heat-map resize kernel is checked against) reproduces the
    reference's own ``convert_similarity_to_image_size`` outputs bit for bit (nearest mode, vlp/inference_engine.py:113-155),
    and agrees with the host function on shapes the fixtures do not hold (identity, exact doubling, a crop larger than the
    image)."""
    import heatmap_oracle as HO
    from incremental_multimodal_medical_learning_ii_b200.vlp.inference_engine import ImageTextInferenceEngine as E
    grid = (torch.arange(15 * 15, dtype=torch.float32).reshape(15, 15) / 7.0).numpy()
    for k, (w, h, rs, cs) in enumerate(vlp["resize_cases"].tolist()):
        got = HO.heatmap_to_image_size(grid, w, h, rs or None, cs or None)
        ref = vlp[f"resize_{k}_nearest"]
        assert got.shape == ref.shape
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(ref)), k
    g = torch.Generator().manual_seed(4)
    for gh, gw, w, h, rs, cs in ((16, 16, 16, 16, None, None), (16, 16, 32, 32, None, None), (7, 9, 133, 101, None, None),
                                 (15, 15, 200, 150, None, 240), (4, 4, 37, 91, 64, 48)):
        hm = torch.randn(gh, gw, generator=g)
        got = HO.heatmap_to_image_size(hm.numpy(), w, h, rs, cs)
        ref = E.convert_similarity_to_image_size(hm, width=w, height=h, resize_size=rs, crop_size=cs)
        assert got.shape == ref.shape
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))


def test_oracle_label_loop_matches_reference_trainer(trainer_golden):
    tg = trainer_golden
    assert len(tg["cases"]) == 12
    for name, c in tg["cases"].items():
        o = O.trainer_val_batch(tg["embs"], c["prompts"], c["train_logit_diff"], c["pred_logit_diff"], c["max_emb"])
        assert torch.equal(o["predicted_labels"], c["y_pred"]), name          # Trainer.py:834-837 / 1044-1047
        assert (o["tmp_score"] - c["y_score"]).abs().max() <= 1e-6, name      # Trainer.py:824-827
        if "logits" in c:
            assert (o["logits"] - c["logits"]).abs().max() <= 1e-6, name      # Trainer.py:829-832
        # the reference re-runs CXR-BERT for every label of every batch: 2 x 5 forwards x 3 batches
        assert c["bert_calls"] == 30


# ---- live against the reference (build container only) ---------------------------------------------------------------
needs_reference = pytest.mark.skipif(not RS.reference_available(), reason="/root/reference is not present on this machine")


def _oracle_factory(prompts, device, tld, pld, mx):
    """TEST-ONLY scorer backend for patch_trainer_eval: the oracle's label loop (there is no CPU path in the product)."""
    return lambda embs: O.trainer_val_batch(embs.cpu(), prompts.cpu(), tld, pld, mx)


@needs_reference
@pytest.mark.parametrize("which", ["val", "test"])
@pytest.mark.parametrize("case", ["single", "mean4", "max4", "mean4_posonly", "max4_preddiff"])
def test_patched_reference_trainer_reproduces_its_own_loop(trainer_golden, case, which):
    """The reference's Trainer, UNCHANGED, with val/test replaced by scorer.patch_trainer_eval: same evaluate_model inputs
    as its own label loop gave (fixtures), 10 CXR-BERT forwards per call instead of 10 per batch."""
    import contextlib
    import io
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_scorer as MG
    from incremental_multimodal_medical_learning_ii_b200.scorer import patch_trainer_eval
    tg = trainer_golden
    c = tg["cases"][f"{case}/{which}"]
    table, t = MG.prompt_table(5, c["P"], seed=37 + c["P"])
    assert torch.equal(t, c["prompts"])
    prompts = {name: {"positive": [f"label {l} pos {j}" for j in range(c["P"])],
                      "negative": [f"label {l} neg {j}" for j in range(c["P"])]} for l, name in enumerate(tg["class_names"])}
    T = RS.load_reference_trainer_module(table, max_emb=c["max_emb"], train_logit_diff=c["train_logit_diff"],
                                         pred_logit_diff=c["pred_logit_diff"])
    with contextlib.redirect_stdout(io.StringIO()):
        tr = T.Trainer(c["single_prompt"], prompts, tg["class_names"], "standard", 1e-3, torch.device("cpu"), None)
    seen = {}
    tr.evaluate_model = lambda y_true, y_pred, y_score, *a, **k: seen.update(y_true=y_true, y_pred=y_pred, y_score=y_score)
    for nm in ("plot_cosine_similarity_text_embs", "plot_cosine_similarity_text_embs_only_pos_prompts", "plot_new_text_embeddings"):
        setattr(tr, nm, lambda *a, **k: None)
    patch_trainer_eval(tr, scorer_factory=_oracle_factory)
    loader = [(tg["embs"][i:i + 256], tg["labels"][i:i + 256]) for i in range(0, tg["embs"].shape[0], 256)]
    rec = MG._Recorder()
    getattr(tr, which)(loader, rec, 1, 1)
    assert np.array_equal(seen["y_pred"], c["y_pred"].numpy())
    assert np.abs(seen["y_score"] - c["y_score"].numpy()).max() <= 1e-6
    assert np.array_equal(seen["y_true"], tg["labels"].numpy())
    if which == "val":
        assert (torch.cat(rec.logits) - c["logits"]).abs().max() <= 1e-6
        assert np.allclose(rec.losses, c["loss"].numpy(), atol=1e-6)
    assert tr.bert_encoder.calls == 10                   # 2 x 5 CXR-BERT forwards per CALL (hoisted); the reference: 30 (per batch)


@needs_reference
def test_reference_vlp_engine_live_matches_fixture(golden, vlp):
    """Regenerates two fixture entries from the reference itself: the committed fixtures are what the reference computes."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_scorer as MG
    E = RS.load_reference_vlp_engine_class()
    g = golden["bnrand/structured"]
    table, _ = MG.prompt_table(14, 5, seed=29)
    eng = E(MG._RefImageEngine(g["global256"], g["patch_norm_first2"], (390, 320), 512, 480), RS.FakeTextEngine(table))
    assert abs(eng.get_similarity_score_from_raw_data(3, [f"label 2 pos {j}" for j in range(5)]) - vlp["score_p5"][3, 2]) <= 1e-7
    full = eng.get_similarity_map_from_raw_data(0, "label 3 pos 0")
    assert np.array_equal(np.nan_to_num(full), np.nan_to_num(vlp["full_map_0"]))

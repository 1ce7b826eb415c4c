"""Scorer-side and config-1 parity pins (GPU half): the CUDA path against fixtures the REFERENCE's own code produced.

* ``biovil_golden.pt``  - ``ImageModel.forward`` of the reference on all 256 frames of BASELINE.json configs[0]
  (4 weight-variant x frame-kind fixtures) + scores of those embeddings + prompt sets whose fp32 margins exceed 5e-3;
* ``vlp_golden.npz``    - the reference's ``ImageTextInferenceEngine`` (similarity score, smoothed similarity maps);
* ``trainer_golden.pt`` - the reference's ``Trainer.val`` / ``Trainer.test`` label loops, run unchanged.

Tolerances (BASELINE.json north_star): per-embedding cosine >= 0.999, probabilities within 1e-3, labels identical -
asserted STRICTLY on the high-margin prompt sets; fp32-only kernels (scorer on given embeddings, smoothing) 1e-5 / 2e-6.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vlp():
    return np.load(os.path.join(ROOT, "tests", "golden", "vlp_golden.npz"))


@pytest.fixture(scope="module")
def trainer_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "trainer_golden.pt"), map_location="cpu")


def _model(randomize_bn):
    from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as SW
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    m = get_biovil_resnet(None)
    m.load_state_dict(SW.make_state_dict(27, randomize_bn=randomize_bn))
    m.train(mode=False, my_freeze=True)
    m.eval().to(DEV)
    return m


@pytest.fixture(scope="module", params=["default", "bnrand"])
def variant_model(request):
    return request.param, _model(request.param == "bnrand")


@pytest.mark.parametrize("kind", ["iid", "structured"])
def test_config1_256_frames_vs_reference(variant_model, golden, kind):
    """BASELINE.json configs[0] at its stated size: 256 frames, every embedding and every label decision."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m = variant_model
    g = golden[f"{variant}/{kind}"]
    fr = FR.synthetic_frames_u8(0, 256, 480, kind=kind, seed=0, device=DEV)
    assert int(fr.long().sum()) == g["frames256_checksum"]
    ref = g["global256"]
    for pname, P in (("p1", 1), ("p5", 5)):
        # (a) high-margin prompt set: labels must be IDENTICAL, strictly
        hm = g[f"margin_{pname}"]
        m.set_prompts(hm["prompts"], reduce="mean")
        res = m.embed_and_score(fr)
        emb = res["global"].float().cpu()
        cos = F.cosine_similarity(emb, ref, dim=-1)
        rel = ((emb - ref).norm(dim=1) / ref.norm(dim=1))
        ec, rc = emb - emb.mean(0, keepdim=True), ref - ref.mean(0, keepdim=True)
        ccos = F.cosine_similarity(ec, rc, dim=-1)
        dprob = (res["prob"].cpu() - hm["prob"]).abs().max().item()
        dsim = (res["sim"].cpu() - hm["sim"]).abs().max().item()
        flips = int((res["pred"].cpu() != hm["pred"]).sum())
        print(f"[{variant}/{kind}/{pname}] 256 frames: cosine min {cos.min():.6f} rel-L2 max {rel.max():.3e} centred-cos min "
              f"{ccos.min():.4f}; margin set: max|dsim| {dsim:.2e} max|dprob| {dprob:.2e} flips {flips}/{hm['pred'].numel()}")
        assert cos.min().item() >= 0.999
        assert rel.max().item() <= 2e-2
        if kind == "structured":
            assert ccos.min().item() >= 0.97
        assert dprob <= 1e-3
        assert flips == 0, "predicted labels must be identical on the high-margin prompt set"
        # (b) plain prompt set: probabilities within 1e-3 everywhere; labels identical wherever fp32 itself decides
        ps = g[f"score256_{pname}_mean"]
        m.set_prompts(FR.synthetic_prompt_embeddings(14, P, 128, seed=29), reduce="mean")
        res = m.embed_and_score(fr)
        assert (res["prob"].cpu() - ps["prob"]).abs().max().item() <= 1e-3
        margin = (ps["sim"][..., 0] - ps["sim"][..., 1]).abs()
        wrong = (res["pred"].cpu() != ps["pred"])
        print(f"[{variant}/{kind}/{pname}] plain set: flips {int(wrong.sum())} (all at fp32 margin <= {margin[wrong].max().item() if wrong.any() else 0:.1e})")
        assert not bool((wrong & (margin > 1e-3)).any())
        # scoring the reference's OWN embeddings reproduces its sims / labels (fp32 kernel, no bf16 involved)
        r2 = m.score_embeddings(ref.to(DEV))
        assert (r2["sim"].cpu() - ps["sim"]).abs().max().item() <= 1e-5
        assert torch.equal(r2["pred"].cpu()[margin > 1e-6], ps["pred"][margin > 1e-6])


def test_trainer_eval_scorer_vs_reference_trainer(trainer_golden):
    """TrainerEvalScorer (one launch per batch) against what the reference's Trainer.val / Trainer.test handed to
    evaluate_model and to the criterion, for every switch combination (Trainer.py:797-837, 1019-1047)."""
    from incremental_multimodal_medical_learning_ii_b200.scorer import TrainerEvalScorer, my_cosine_similarity
    tg = trainer_golden
    embs = tg["embs"].to(DEV)
    for name, c in tg["cases"].items():
        sc = TrainerEvalScorer(c["prompts"], DEV, c["train_logit_diff"], c["pred_logit_diff"], c["max_emb"])
        out = sc(embs)
        assert (out["tmp_score"].cpu() - c["y_score"]).abs().max().item() <= 1e-5, name
        decided = torch.ones_like(c["y_pred"], dtype=torch.bool)
        if "logits" in c:
            assert (out["logits"].cpu() - c["logits"]).abs().max().item() <= 1e-5, name
            if c["train_logit_diff"]:
                decided = c["logits"].abs() > 1e-6                            # fp32 ties are not decisions
        n_undecided = int((~decided).sum())
        assert n_undecided <= 2, name
        assert torch.equal(out["predicted_labels"].cpu()[decided], c["y_pred"][decided]), name
    # Trainer.myCosineSimilarity drop-in (stateless, one launch): pos cosine = 2 * tmp_score - 1 when not PRED_LOGIT_DIFF
    c = tg["cases"]["mean4/val"]
    for l in range(5):
        y = c["prompts"][l, 0].mean(dim=0)                                    # bert_forward_mean, Trainer.py:1665-1666
        got = my_cosine_similarity(embs, y.to(DEV), use_grad=False)
        assert got.shape == (embs.shape[0], 1)
        assert (got[:, 0].cpu() - (2 * c["y_score"][:, l] - 1)).abs().max().item() <= 1e-5
    c = tg["cases"]["max4/val"]
    got = my_cosine_similarity(embs, c["prompts"][2, 0].to(DEV), use_grad=False, max_emb=True)
    assert got.shape == (embs.shape[0],)
    assert (got.cpu() - (2 * c["y_score"][:, 2] - 1)).abs().max().item() <= 1e-5
    one = my_cosine_similarity(embs[7], c["prompts"][2, 0, 1].to(DEV), to_plot=True)
    assert one.shape == (1, 1)
    with pytest.raises(RuntimeError):
        my_cosine_similarity(embs, c["prompts"][2, 0, :1].to(DEV), use_grad=True)


def test_image_text_engine_vs_reference_vlp_engine(golden, vlp):
    """health_multimodal/vlp/inference_engine.py:31-57 and :93-111 of the reference (fixtures) against the B200 path."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    from incremental_multimodal_medical_learning_ii_b200.vlp import ImageTextInferenceEngine as E
    g = golden["bnrand/structured"]
    m = _model(True)
    t = FR.synthetic_prompt_embeddings(14, 5, 128, seed=29)
    fr = FR.synthetic_frames_u8(0, 32, 480, kind="structured", seed=0, device=DEV)
    emb = F.normalize(m(fr).projected_global_embedding, dim=-1)                       # image/inference_engine.py:81-82
    for key, tt in (("score_p5", t[:, 0]), ("score_p1", t[:, 0, :1])):
        text = F.normalize(tt.mean(dim=1), dim=-1).to(DEV)                             # vlp/inference_engine.py:52-53
        got = (emb @ text.t()).double().cpu().numpy()
        d = np.abs(got - vlp[key]).max()
        print(f"similarity score {key}: max abs diff vs the reference engine {d:.2e}")
        assert d <= 2e-3                                                               # bf16 trunk vs fp32 reference
    # smoothing + matmul on the REFERENCE's patch embeddings: fp32 kernel against the reference's scipy call
    for i in range(2):
        for l in range(14):
            te = F.normalize(t[l, 0, :1], dim=1).to(DEV)
            sm = E._get_similarity_map_from_embeddings(g["patch_norm_first2"][i].to(DEV), te)
            assert (sm.numpy() - vlp["smoothed_maps"][i, l]).__abs__().max() <= 2e-6
    # the whole patch path (bf16 trunk) through the batched API
    m.set_prompts(t[:, :, :1], reduce="mean")
    res = m.embed_and_score(fr[:2], heat=True)
    sm = m.smooth_heatmaps(res["heat"], 1.5).permute(0, 3, 1, 2).cpu().numpy()         # [2,14,15,15]
    d = np.abs(sm - vlp["smoothed_maps"]).max()
    print(f"smoothed similarity maps, full path: max abs diff vs the reference engine {d:.2e}")
    assert d <= 5e-3


def test_heatmaps_to_image_size_vs_reference_function(vlp):
    """bv_heatmaps_to_image_size (nearest upsample of the patch grid over the centre-crop square + NaN pad, batched) against
    ``convert_similarity_to_image_size`` - the host function that tests/test_scorer_pins_cpu.py pins bit-equal to the
    reference's (vlp/inference_engine.py:113-155) on the same cases - and against the reference's own recorded outputs."""
    from incremental_multimodal_medical_learning_ii_b200.image.model.model import ImageModel
    from incremental_multimodal_medical_learning_ii_b200.vlp.inference_engine import ImageTextInferenceEngine as E
    grid = torch.arange(15 * 15, dtype=torch.float32).reshape(15, 15) / 7.0          # the grid of the golden cases
    g = torch.Generator().manual_seed(4)
    heat = torch.randn(3, 15, 15, 5, generator=g)
    heat[0, :, :, 0] = grid
    for k, (w, h, rs, cs) in enumerate(vlp["resize_cases"].tolist()):
        got = ImageModel.heatmaps_to_image_size(None, heat.cuda(), w, h, rs or None, cs or None).cpu().numpy()
        assert got.shape == (3, 5, h, w)
        ref0 = vlp[f"resize_{k}_nearest"]                                             # the reference's own output
        assert np.array_equal(np.isnan(got[0, 0]), np.isnan(ref0)) and np.array_equal(np.nan_to_num(got[0, 0]), np.nan_to_num(ref0)), k
        for b, l in ((1, 3), (2, 4)):
            ref = E.convert_similarity_to_image_size(heat[b, :, :, l], width=w, height=h, resize_size=rs or None, crop_size=cs or None)
            assert np.array_equal(np.isnan(got[b, l]), np.isnan(ref)) and np.array_equal(np.nan_to_num(got[b, l]), np.nan_to_num(ref)), (k, b, l)
    # other grid sizes / scale factors (identity, exact doubling, fractional), a crop larger than the image (F.pad crops)
    for gh, gw, w, h, rs, cs in ((16, 16, 16, 16, None, None), (16, 16, 32, 32, None, None), (7, 9, 333, 201, None, None),
                                 (15, 15, 400, 300, None, 480), (15, 15, 390, 320, 512, 480), (4, 4, 37, 91, 64, 48)):
        hm = torch.randn(2, gh, gw, 3, generator=g)
        got = ImageModel.heatmaps_to_image_size(None, hm.cuda(), w, h, rs, cs).cpu().numpy()
        for b in range(2):
            for l in range(3):
                ref = E.convert_similarity_to_image_size(hm[b, :, :, l], width=w, height=h, resize_size=rs, crop_size=cs)
                assert got[b, l].shape == ref.shape
                assert np.array_equal(np.isnan(got[b, l]), np.isnan(ref)) and np.array_equal(np.nan_to_num(got[b, l]), np.nan_to_num(ref)), (gh, gw, w, h, rs, cs)
    import heatmap_oracle as HO                                                   # the numpy restatement (pinned on the CPU side)
    hm = torch.randn(2, 15, 15, 3, generator=g)
    got = ImageModel.heatmaps_to_image_size(None, hm.cuda(), 390, 320, 512, 480).cpu().numpy()
    ref = HO.heatmaps_to_image_size(hm.numpy(), 390, 320, 512, 480)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))
    assert ImageModel.heatmaps_to_image_size(None, torch.empty(0, 15, 15, 14, device="cuda"), 64, 48, 512, 480).shape == (0, 14, 48, 64)
    with pytest.raises(ValueError):
        ImageModel.heatmaps_to_image_size(None, torch.zeros(1, 15, 15, 14), 64, 48, 512, 480)      # CPU tensor: no CPU path

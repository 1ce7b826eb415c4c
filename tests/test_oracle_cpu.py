"""The oracle (oracle/biovil_oracle.py, a CPU restatement of the reference) against the fixtures the reference itself
produced (tests/golden/biovil_golden.pt, written by oracle/make_golden.py with /root/reference imported)."""
import torch
import torch.nn.functional as F

import biovil_oracle as O
import weights as Wt
from incremental_multimodal_medical_learning_ii_b200 import frames as FR


def test_oracle_reproduces_reference_outputs(golden):
    torch.manual_seed(0)
    for variant, rbn in (("default", False), ("bnrand", True)):
        sd = Wt.make_state_dict(27, randomize_bn=rbn)
        for kind in ("iid", "structured"):
            g = golden[f"{variant}/{kind}"]
            assert abs(Wt.state_dict_checksum(sd) - g["weights_checksum"]) <= 1e-6 * g["weights_checksum"]
            fr = FR.synthetic_frames_u8(0, 2, 480, kind=kind, seed=0)
            out = O.image_model_forward(sd, FR.frames_as_reference_input(fr))
            ref = g["global"][:2]
            rel = ((out["projected_global_embedding"] - ref).norm() / ref.norm()).item()
            assert rel <= 1e-5, (variant, kind, rel)          # same algorithm; only thread-count summation order may differ
            relp = ((out["img_embedding"] - g["pooled"][:2]).norm() / g["pooled"][:2].norm()).item()
            assert relp <= 1e-5
            patch = F.normalize(out["projected_patch_embeddings"], dim=1).permute(0, 2, 3, 1)
            assert (patch - g["patch_norm_first2"]).abs().max().item() <= 1e-5
            assert torch.allclose(O.patchwise_projected_embeddings(sd, FR.frames_as_reference_input(fr), True), patch)


def test_frames_checksum_matches_golden(golden):
    for kind in ("iid", "structured"):
        fr = FR.synthetic_frames_u8(0, 8, 480, kind=kind, seed=0)
        assert int(fr.long().sum()) == golden[f"default/{kind}"]["frames_checksum"]


def test_scorer_restatement_vs_golden_and_properties(golden):
    g = golden["bnrand/structured"]
    emb = g["global"]
    for pname, P in (("p1", 1), ("p5", 5)):
        prompts = FR.synthetic_prompt_embeddings(14, P, 128, seed=29)
        for reduce in ("mean", "max"):
            s = O.zero_shot_score(emb, prompts, reduce)
            gs = g[f"score_{pname}_{reduce}"]
            assert torch.allclose(s["sim"], gs["sim"], atol=1e-6)
            assert torch.equal(s["pred"], gs["pred"])
            # decision rule of Trainer.py:836: 1 iff pos > neg; prob = sigmoid(pos - neg) > 0.5 iff pred
            assert torch.equal(s["pred"].bool(), s["sim"][..., 0] > s["sim"][..., 1])
            assert torch.equal(s["pred"].bool(), s["prob"] > 0.5)
            assert (s["sim"].abs() <= 1 + 1e-6).all()
            # cosine is scale invariant in both arguments
            s2 = O.zero_shot_score(emb * 3.7, prompts * 0.2, reduce)
            assert torch.allclose(s2["sim"], s["sim"], atol=1e-6)
    # with one prompt per polarity mean and max coincide
    p1 = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
    assert torch.allclose(O.zero_shot_score(emb, p1, "mean")["sim"], O.zero_shot_score(emb, p1, "max")["sim"])
    # explicit normalise-then-matmul (trash/lower_bound_mcs.py:82,101-111)
    t = F.normalize(p1[:, :, 0], dim=-1)
    alt = torch.einsum("bd,lpd->blp", F.normalize(emb, dim=-1), t)
    assert torch.allclose(alt, O.zero_shot_score(emb, p1, "mean")["sim"], atol=1e-6)


def test_bn_randomisation_changes_outputs():
    """A default-init BatchNorm is an identity; the BN-randomised fixture must actually differ."""
    a = Wt.make_state_dict(27, randomize_bn=False)
    b = Wt.make_state_dict(27, randomize_bn=True)
    k = "encoder.encoder.layer3.2.bn2.running_var"
    assert torch.all(a[k] == 1) and not torch.all(b[k] == 1)
    assert torch.equal(a["encoder.encoder.layer3.2.conv2.weight"], b["encoder.encoder.layer3.2.conv2.weight"])


def test_small_frames_and_margin_prompts():
    sd = Wt.make_state_dict(27, randomize_bn=True)
    fr = FR.synthetic_frames_u8(5, 3, 64, kind="structured", seed=1)
    out = O.image_model_forward(sd, FR.frames_as_reference_input(fr))
    assert out["projected_global_embedding"].shape == (3, 128)
    assert out["projected_patch_embeddings"].shape == (3, 128, 2, 2)
    assert torch.allclose(out["projected_patch_embeddings"].mean(dim=(2, 3)), out["projected_global_embedding"], atol=1e-5)
    prompts = FR.synthetic_prompt_embeddings(4, 2, 128, seed=3, min_margin_against=out["projected_global_embedding"],
                                             min_margin=5e-3)
    s = O.zero_shot_score(out["projected_global_embedding"], prompts, "mean")
    assert ((s["sim"][..., 0] - s["sim"][..., 1]).abs() > 5e-3).all()


def test_trainer_val_batch_restatement_is_consistent_with_zero_shot_score():
    """The line-by-line restatement of the Trainer.val label loop (Trainer.py:797-837) agrees with the vectorised
    scorer oracle for the default switches, and shows the documented degenerate behaviour of TRAIN_LOGIT_DIFF=False."""
    import torch
    import biovil_oracle as O
    g = torch.Generator().manual_seed(3)
    embs, prompts = torch.randn(64, 128, generator=g), torch.randn(5, 2, 4, 128, generator=g)
    for max_emb in (False, True):
        loop = O.trainer_val_batch(embs, prompts, True, False, max_emb)
        vec = O.zero_shot_score(embs, prompts, "max" if max_emb else "mean")
        assert torch.allclose(loop["logits"], vec["logit"], atol=1e-6)
        assert torch.allclose(loop["tmp_score"], vec["score"], atol=1e-6)
        assert torch.equal(loop["predicted_labels"], vec["pred"].float())
        diff = O.trainer_val_batch(embs, prompts, True, True, max_emb)
        assert torch.allclose(diff["tmp_score"], vec["score_diff"], atol=1e-6)
    single = O.trainer_val_batch(embs, prompts, False, False, False)
    assert single["predicted_labels"].sum() == 0                   # argmax([pos, pos]) -> index 0
    assert torch.allclose(single["logits"], O.zero_shot_score(embs, prompts, "mean")["sim"][..., 0], atol=1e-6)


def test_gaussian_smooth_oracle_handles_views_and_matches_direct_scipy():
    import numpy as np
    import torch
    from scipy import ndimage
    import biovil_oracle as O
    heat = torch.rand(2, 15, 15, 3, generator=torch.Generator().manual_seed(1))
    sm = O.gaussian_smooth_map(heat.permute(0, 3, 1, 2), 1.5)              # non-contiguous view in
    direct = ndimage.gaussian_filter(heat[1, :, :, 2].numpy(), sigma=(1.5, 1.5), order=0)
    assert np.array_equal(sm[1, 2].numpy(), direct)
    # smoothing preserves the mean of a constant map and is bounded by the input range
    const = O.gaussian_smooth_map(torch.full((1, 15, 15), 0.25), 1.5)
    assert torch.allclose(const, torch.full((1, 15, 15), 0.25), atol=1e-6)

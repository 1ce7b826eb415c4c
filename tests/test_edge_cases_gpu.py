"""Edge cases of the scorer / model surface against what the reference's torch ops do on the same inputs.

* empty batches (the reference's ops return empty tensors: an empty shard of the extraction loop, an empty DataLoader tail);
* degenerate cosines: a zero image embedding or a zero prompt gives 0/0 = NaN in torchmetrics'
  ``pairwise_cosine_similarity`` (plain division, no eps), and ``torch.argmax(cat([neg, pos]))`` (Trainer.py:836) then
  treats NaN as the maximum with the FIRST maximum winning - the CUDA label must be that one, not just ``pos > neg``;
* exact ties (identical positive and negative prompts -> label 0, the "trick" of Trainer.py:814);
* ``torch.max`` over per-prompt cosines (MAX_EMB, Trainer.py:1691-1694) propagates NaN, ``fmaxf`` would not.
The checker is the oracle's ``zero_shot_score`` / ``pairwise_cosine_similarity`` (torch fp32 on the CPU).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _same_with_nan(a, b, tol=0.0):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.equal(torch.isnan(a), torch.isnan(b)), "NaN pattern differs from the reference's"
    m = ~torch.isnan(a)
    return (a[m] - b[m]).abs().max().item() <= tol if m.any() else True


@pytest.mark.parametrize("reduce,P", [("mean", 1), ("mean", 5), ("max", 5)])
def test_degenerate_cosines_follow_torch_argmax(reduce, P):
    import biovil_oracle as O
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    from incremental_multimodal_medical_learning_ii_b200.scorer import ZeroShotScorer
    L = 14
    prompts = FR.synthetic_prompt_embeddings(L, P, 128, seed=5).clone()          # [L,2,P,128]
    prompts[1, 0] = 0.0                     # zero positive prompts of label 1: pos = NaN -> label 1
    prompts[2, 1] = 0.0                     # zero negative prompts of label 2: neg = NaN -> label 0
    prompts[3] = 0.0                        # both zero -> label 0 (first maximum)
    prompts[4, 1] = prompts[4, 0]           # exact tie -> label 0
    if P > 1:
        prompts[5, 0, 2] = 0.0              # ONE zero prompt among P: the mean survives, the max must be NaN
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(40, 128, generator=g)
    emb[7] = 0.0                            # zero image embedding: every cosine NaN -> every label 0
    emb[8] *= 1e-30                         # squares underflow in fp32: norm 0 in the reference too -> NaN, label 0
    ref = O.zero_shot_score(emb, prompts, reduce=reduce)
    sc = ZeroShotScorer(DEV)
    sc.set_prompts(prompts, reduce=reduce)
    got = sc.score(emb.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(got["pred"].cpu(), ref["pred"]), "labels differ from torch.argmax([neg, pos])"
    assert _same_with_nan(got["sim"], ref["sim"], 2e-6)
    assert _same_with_nan(got["prob"], ref["prob"], 2e-6)
    assert _same_with_nan(got["score"], ref["score"], 2e-6)
    assert ref["pred"][7].sum() == 0 and ref["pred"][:, 1].sum() == 38 and ref["pred"][8].sum() == 0        # the cases above do occur
    if reduce == "max":
        assert torch.isnan(ref["sim"][0, 5, 0]) and torch.isnan(got["sim"][0, 5, 0].cpu())


def test_pairwise_cosine_degenerate_and_empty():
    import biovil_oracle as O
    from incremental_multimodal_medical_learning_ii_b200.scorer import my_cosine_similarity
    g = torch.Generator().manual_seed(3)
    x = torch.randn(9, 128, generator=g)
    x[4] = 0.0
    y = torch.randn(5, 128, generator=g)
    y[2] = 0.0
    ref = O.pairwise_cosine_similarity(x, y)                       # [9,5] with a NaN row and a NaN column
    got_max = my_cosine_similarity(x.to(DEV), y.to(DEV), max_emb=True)
    assert got_max.shape == (9,) and torch.isnan(got_max).all()    # torch.max propagates the NaN column into every row
    got = my_cosine_similarity(x.to(DEV), y[0].to(DEV))
    assert _same_with_nan(got[:, 0], ref[:, 0], 2e-6) and torch.isnan(got[4, 0])
    empty = my_cosine_similarity(torch.empty(0, 128, device=DEV), y[0].to(DEV))
    assert empty.shape == (0, 1)
    assert my_cosine_similarity(torch.empty(0, 128, device=DEV), y.to(DEV), max_emb=True).shape == (0,)


def test_empty_batches_give_empty_outputs():
    import weights as Wt
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    from incremental_multimodal_medical_learning_ii_b200.scorer import TrainerEvalScorer, ZeroShotScorer
    prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
    sc = ZeroShotScorer(DEV)
    sc.set_prompts(prompts)
    out = sc.score(torch.empty(0, 128, device=DEV))
    assert out["sim"].shape == (0, 14, 2) and out["prob"].shape == (0, 14) and out["pred"].dtype == torch.uint8
    assert out["logit"].shape == (0, 14)
    ts = TrainerEvalScorer(prompts, device=DEV)
    tv = ts(torch.empty(0, 128, device=DEV))
    assert all(v.shape[0] == 0 for v in tv.values())
    m = get_biovil_resnet(None)
    m.load_state_dict(Wt.make_state_dict(27))
    m.eval().to(DEV)
    m.set_prompts(prompts)
    res = m.embed_and_score(torch.empty(0, 1, 96, 96, dtype=torch.uint8, device=DEV), heat=True, patch=True)
    assert res["global"].shape == (0, 128) and res["prob"].shape == (0, 14) and res["pred"].shape == (0, 14)
    assert res["patch"].shape == (0, 3, 3, 128) and res["heat"].shape == (0, 3, 3, 14)
    assert m.score_embeddings(torch.empty(0, 128, device=DEV))["sim"].shape == (0, 14, 2)
    assert m(torch.empty(0, 3, 96, 96, device=DEV)).shape == (0, 128)
    one = m.embed_and_score(FR.synthetic_frames_u8(1, 1, 96, kind="iid", seed=2).to(DEV))      # and still works after
    assert one["global"].shape == (1, 128) and torch.isfinite(one["global"]).all()

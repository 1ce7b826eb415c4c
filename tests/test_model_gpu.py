"""End-to-end parity of the B200 ``ImageModel`` against the reference's outputs (tests/golden, produced by the
reference itself in oracle/make_golden.py) and against the oracle at small frame sizes.

Tolerances (BASELINE.json north_star): per-embedding cosine >= 0.999, zero-shot probabilities within 1e-3, predicted
labels identical.  Because random-init embeddings are nearly colinear the cosine bound alone is weak,
so relative L2 and the cosine of batch-centred embeddings are asserted too.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _model(randomize_bn):
    import weights as Wt
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    sd = Wt.make_state_dict(27, randomize_bn=randomize_bn)
    m = get_biovil_resnet(None)
    m.load_state_dict(sd)
    m.train(mode=False, my_freeze=True)
    m.eval()
    m.to(DEV)
    return m, sd


@pytest.fixture(scope="module", params=["default", "bnrand"])
def setup(request, golden):
    m, sd = _model(request.param == "bnrand")
    return request.param, m, sd, golden


def _metrics(a, ref):
    a, ref = a.float().cpu(), ref.float().cpu()
    cos = F.cosine_similarity(a, ref, dim=-1).min().item()
    rel = ((a - ref).norm() / ref.norm()).item()
    ac, rc = a - a.mean(0, keepdim=True), ref - ref.mean(0, keepdim=True)
    ccos = F.cosine_similarity(ac, rc, dim=-1).min().item()
    return cos, rel, ccos


@pytest.mark.parametrize("kind", ["iid", "structured"])
def test_global_embedding_vs_reference(setup, kind):
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    import weights as Wt
    variant, m, sd, golden = setup
    g = golden[f"{variant}/{kind}"]
    assert abs(Wt.state_dict_checksum(sd) - g["weights_checksum"]) <= 1e-6 * g["weights_checksum"]
    fr = FR.synthetic_frames_u8(0, 8, 480, kind=kind, seed=0)
    assert int(fr.long().sum()) == g["frames_checksum"]
    out = m(fr.to(DEV))
    assert out.shape == (8, 128) and out.dtype == torch.float32 and not out.requires_grad
    cos, rel, ccos = _metrics(out, g["global"])
    print(f"[{variant}/{kind}] cosine {cos:.6f} rel-L2 {rel:.3e} centred-cosine {ccos:.4f}")
    assert cos >= 0.999
    assert rel <= 2e-2
    if kind == "structured":
        assert ccos >= 0.98
    # float [B,3,H,W] input as the reference's transforms produce it (exact k/255 detection -> same result)
    out3 = m(FR.frames_as_reference_input(fr).to(DEV))
    assert torch.equal(out3.cpu(), out.cpu())
    # torch.cat usability (chexpert-get-embedding.py:79) and upstream attribute access (inference_engine.py:81)
    cat = torch.cat([torch.empty(0, 128, device=DEV), out])
    assert type(cat) is torch.Tensor and cat.shape == (8, 128)
    assert torch.equal(out.projected_global_embedding, cat)


def test_patch_pooled_and_scores_vs_reference(setup):
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    g = golden[f"{variant}/structured"]
    fr = FR.synthetic_frames_u8(0, 8, 480, kind="structured", seed=0).to(DEV)
    patch = m.get_patchwise_projected_embeddings(fr[:2], normalize=True)
    assert patch.shape == (2, 15, 15, 128)
    pc = F.cosine_similarity(patch.cpu().reshape(-1, 128), g["patch_norm_first2"].reshape(-1, 128), dim=-1)
    print(f"[{variant}] patch cosine min {pc.min().item():.6f}")
    assert pc.min().item() >= 0.999
    assert (patch.norm(dim=-1) - 1).abs().max().item() < 1e-4
    raw = m.get_patchwise_projected_embeddings(fr[:2], normalize=False)
    rel = ((raw.cpu() - g["patch_raw_first2"]).norm() / g["patch_raw_first2"].norm()).item()
    assert rel <= 2e-2
    out = m(fr)
    pooled = out.img_embedding
    rel = ((pooled.cpu() - g["pooled"]).norm() / g["pooled"].norm()).item()
    print(f"[{variant}] pooled rel-L2 {rel:.3e}")
    assert rel <= 2e-2
    assert out.projected_patch_embeddings.shape == (8, 128, 15, 15)
    assert out.patch_embedding.shape == (8, 2048, 15, 15)
    assert out.class_logits is None
    enc_patch, enc_pooled = m.encoder(fr[:2], return_patch_embeddings=True)
    assert enc_patch.shape == (2, 2048, 15, 15) and enc_pooled.shape == (2, 2048)
    # zero-shot scores: 14 labels, 1 and 5 prompts per polarity, mean and max reduction
    for pname, P in (("p1", 1), ("p5", 5)):
        prompts = FR.synthetic_prompt_embeddings(14, P, 128, seed=29)
        for reduce in ("mean", "max"):
            gs = g[f"score_{pname}_{reduce}"]
            m.set_prompts(prompts, reduce=reduce)
            res = m.embed_and_score(fr, heat=(pname == "p5" and reduce == "mean"))
            dsim = (res["sim"].cpu() - gs["sim"]).abs().max().item()
            dprob = (res["prob"].cpu() - gs["prob"]).abs().max().item()
            margin = (gs["sim"][..., 0] - gs["sim"][..., 1]).abs()
            flips = (res["pred"].cpu() != gs["pred"])
            print(f"[{variant}] {pname}/{reduce}: max|dsim| {dsim:.2e} max|dprob| {dprob:.2e} flips {int(flips.sum())} "
                  f"(min margin {margin.min().item():.1e})")
            assert dprob <= 1e-3
            assert not bool((flips & (margin > 1e-3)).any()), "label flipped outside the arithmetic noise floor"
            # scoring the reference's own embeddings must reproduce its labels exactly
            res2 = m.score_embeddings(g["global"].to(DEV))
            assert torch.equal(res2["pred"].cpu(), gs["pred"])
            assert (res2["prob"].cpu() - gs["prob"]).abs().max().item() <= 1e-5
            assert (res2["sim"].cpu() - gs["sim"]).abs().max().item() <= 1e-5
            assert (res2["score"].cpu() - gs["score"]).abs().max().item() <= 1e-5
            if "heat" in res:
                dh = (res["heat"][:2].cpu() - g["heat_first2_p5"]).abs().max().item()
                print(f"[{variant}] heat-map max abs diff {dh:.2e}")
                assert dh <= 1e-3 * 5


def test_small_sizes_and_float_paths_vs_oracle(setup):
    """Ragged / small cases against the CPU oracle: 1 frame, non-square frames, general float inputs."""
    import biovil_oracle as O
    variant, m, sd, golden = setup
    gen = torch.Generator().manual_seed(3)
    for (B, H, W) in ((1, 64, 64), (3, 96, 160), (5, 128, 96)):
        x3 = torch.rand(B, 3, H, W, generator=gen)                      # general float, three different channels
        ref = O.image_model_forward(sd, x3)["projected_global_embedding"]
        out = m(x3.to(DEV))
        cos, rel, _ = _metrics(out, ref)
        print(f"[{variant}] f3 {B}x{H}x{W}: cosine {cos:.6f} rel {rel:.3e}")
        assert cos >= 0.999 and rel <= 3e-2
        x1 = torch.rand(B, 1, H, W, generator=gen)                      # general float, one channel
        ref = O.image_model_forward(sd, x1.repeat(1, 3, 1, 1))["projected_global_embedding"]
        out = m(x1.to(DEV))
        cos, rel, _ = _metrics(out, ref)
        print(f"[{variant}] f1 {B}x{H}x{W}: cosine {cos:.6f} rel {rel:.3e}")
        assert cos >= 0.999 and rel <= 3e-2


def test_full_bench_batch_matches_chunks(setup):
    """BASELINE.json configs[1] at full size (512 frames of 480x480): every kernel runs its full persistent schedule
    (all 148 CTAs, hundreds of tiles each).  Size-independent property: the results equal, bit for bit, those of the
    same frames sent through in chunks of 64."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    m.set_prompts(FR.synthetic_prompt_embeddings(14, 5, 128, seed=31), reduce="mean")
    fr = torch.cat([FR.synthetic_frames_u8(o, 64, 480, kind="structured", seed=5, device=DEV) for o in range(0, 512, 64)])
    full = m.embed_and_score(fr)
    full = {k: full[k].clone() for k in ("global", "prob", "pred", "sim")}
    assert not torch.isnan(full["global"]).any()
    assert full["global"].shape == (512, 128) and full["prob"].shape == (512, 14)
    for o in range(0, 512, 64):
        part = m.embed_and_score(fr[o:o + 64].contiguous())
        for k in full:
            assert torch.equal(part[k], full[k][o:o + 64]), f"[{variant}] chunk at {o} differs in {k}"


def test_512_crop_8bit_frames_vs_oracle(setup):
    """The reference's other crop size (512x512: 16x16 patch grid, 128 pooled columns = two full stem strips + a ragged
    one, layer1 width not a multiple of 30 so the CTA-pair block kernel is not eligible) on 8-bit frames, plus a
    non-square 8-bit case, against the CPU oracle."""
    import biovil_oracle as O
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
    m.set_prompts(prompts, reduce="mean")
    for (B, H, W) in ((2, 512, 512), (3, 160, 352)):
        fr = FR.synthetic_frames_u8(11, B, max(H, W), kind="structured", seed=4)[:, :, :H, :W].contiguous()
        ref = O.image_model_forward(sd, FR.frames_as_reference_input(fr))
        res = m.embed_and_score(fr.to(DEV))
        cos, rel, _ = _metrics(res["global"], ref["projected_global_embedding"])
        sc = O.zero_shot_score(ref["projected_global_embedding"], prompts, "mean")
        dprob = (res["prob"].cpu() - sc["prob"]).abs().max().item()
        print(f"[{variant}] u8 {B}x{H}x{W}: cosine {cos:.6f} rel {rel:.3e} max|dprob| {dprob:.2e}")
        assert cos >= 0.999 and rel <= 3e-2 and dprob <= 1e-3
        grid = m(fr.to(DEV)).projected_patch_embeddings
        assert grid.shape == (B, 128, H // 32, W // 32)


def test_guards(setup):
    variant, m, sd, golden = setup
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 50, 64, device=DEV))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 64, 64, device=DEV))
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64, device=DEV))
    with pytest.raises(AssertionError):
        m.get_patchwise_projected_embeddings(torch.zeros(1, 3, 64, 64, device=DEV), normalize=True)
    m.eval()


def test_fused_stem_matches_unfused_path(setup, monkeypatch):
    """The fused 8-bit stem kernel (conv7x7 + BN + ReLU + max-pool on tcgen05) against the gather + GEMM + max-pool
    path it replaces: same bf16 rounding points, only the fp32 summation order differs."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    for (B, H, W, kind) in ((3, 96, 160, "structured"), (2, 480, 480, "iid"), (1, 32, 32, "iid")):
        fr = FR.synthetic_frames_u8(7, B, max(H, W), kind=kind, seed=2)[:, :, :H, :W].contiguous().to(DEV)
        m._engine = None
        fused = m(fr).projected_global_embedding.clone()
        trunk_f = m(fr).patch_embedding.clone()
        monkeypatch.setenv("BV_NO_FUSED_STEM", "1")
        m._engine = None
        plain = m(fr).projected_global_embedding.clone()
        trunk_p = m(fr).patch_embedding.clone()
        monkeypatch.delenv("BV_NO_FUSED_STEM")
        m._engine = None
        rel = ((fused - plain).norm() / plain.norm()).item()
        relt = ((trunk_f - trunk_p).norm() / trunk_p.norm()).item()
        print(f"[{variant}] fused vs unfused stem {B}x{H}x{W}: embedding rel {rel:.2e}, trunk rel {relt:.2e}")
        assert rel <= 2e-3 and relt <= 1e-2      # two bf16 pipelines; both sit ~4e-3 from the fp32 oracle


def test_batch_invariance_and_ragged_batches(setup):
    """Size-independent property: a frame's embedding and scores do not depend on the batch it travels in (GEMM rows
    are independent; tiles, chunks and persistent-CTA schedules change with the batch).  Covers odd batch sizes whose
    GEMM row counts are not multiples of the 128-row tile at every layer (ragged last tiles in all kernels)."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
    m.set_prompts(prompts, reduce="mean")
    fr = FR.synthetic_frames_u8(100, 37, 480, kind="structured", seed=0).to(DEV)
    full = m.embed_and_score(fr)
    for n in (1, 7, 33):
        part = m.embed_and_score(fr[:n].contiguous())
        for k in ("global", "prob", "pred", "sim"):
            assert torch.equal(part[k], full[k][:n]), f"[{variant}] batch of {n} differs from batch of 37 in {k}"
    small = FR.synthetic_frames_u8(5, 9, 96, kind="iid", seed=1).to(DEV)          # 96x96: 3x3 patch grid, M = 9 * 9 at layer4
    a = m.embed_and_score(small)
    b = m.embed_and_score(small[2:5].contiguous())
    assert torch.equal(a["global"][2:5], b["global"])


def test_cuda_graph_replay_equals_direct_launches(setup):
    """Small batches replay the whole forward as ONE CUDA graph launch (bv_forward_graph): same kernels, same buffers,
    bit-identical results; different frame contents reuse the graph (static input buffer), a new shape records a new one."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    m.set_prompts(FR.synthetic_prompt_embeddings(14, 1, 128, seed=29), reduce="mean")
    for (B, S) in ((1, 512), (3, 96), (8, 480)):
        outs = {}
        for mode in (False, "auto", True):
            m.cuda_graphs = mode
            runs = []
            for rep in range(3):
                fr = FR.synthetic_frames_u8(50 + rep, B, S, kind="structured", seed=7).to(DEV)
                r = m.embed_and_score(fr)
                runs.append({k: r[k].clone() for k in ("global", "prob", "pred", "sim")})
            outs[mode] = runs
        for rep in range(3):
            for k in ("global", "prob", "pred", "sim"):
                assert torch.equal(outs[False][rep][k], outs["auto"][rep][k]), (B, S, rep, k)
                assert torch.equal(outs[False][rep][k], outs[True][rep][k]), (B, S, rep, k)
        assert not torch.equal(outs["auto"][0]["global"], outs["auto"][1]["global"])      # replays saw the new frames
    m.cuda_graphs = "auto"
    # results are copies: a later call must not overwrite an earlier result
    a = m(FR.synthetic_frames_u8(1, 1, 96, kind="iid", seed=1).to(DEV)).clone()
    keep = m(FR.synthetic_frames_u8(1, 1, 96, kind="iid", seed=1).to(DEV))
    m(FR.synthetic_frames_u8(2, 1, 96, kind="iid", seed=1).to(DEV))
    assert torch.equal(keep, a)


def test_float_frames_are_quantised_on_device_only_when_they_are_8bit(setup):
    """ToTensor + ExpandChannels input (k/255, identical channels) takes the exact 8-bit stem through one conversion kernel;
    anything else (off-grid values, out-of-range, NaN, differing channels) takes the float stems."""
    import biovil_oracle as O
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    fr = FR.synthetic_frames_u8(3, 2, 96, kind="structured", seed=9)
    ref = m(fr.to(DEV)).clone()
    x3 = FR.frames_as_reference_input(fr).to(DEV)
    assert torch.equal(m(x3), ref)                                      # [B,3,H,W] float k/255
    assert torch.equal(m(x3[:, :1].contiguous()), ref)                  # [B,1,H,W] float k/255
    eng = m._get_engine()
    assert eng.quantize(x3) is not None
    off = x3.clone(); off[0, :, 5, 7] += 0.3 / 255                      # not 8-bit data any more
    assert eng.quantize(off) is None
    ch = x3.clone(); ch[1, 2, 0, 0] = 0.5                               # channels differ
    assert eng.quantize(ch) is None
    big = x3.clone(); big[0, :, 1, 1] = 256.0 / 255                     # k = 256 out of range
    assert eng.quantize(big) is None
    nan = x3.clone(); nan[0, :, 2, 2] = float("nan")
    assert eng.quantize(nan) is None
    out = m(off)                                                        # general float path still within tolerance
    o = O.image_model_forward(sd, off.cpu())["projected_global_embedding"]
    assert F.cosine_similarity(out.float().cpu(), o, dim=-1).min().item() >= 0.999


def test_output_guards_and_copies(setup):
    """ImageModelOutput: lazily computed fields refuse to describe a refilled input buffer; eager_outputs computes them in
    the same pass; deepcopy / torch.save of a model that has run work (the native handle is a derived cache)."""
    import copy
    import io
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    variant, m, sd, golden = setup
    buf = FR.synthetic_frames_u8(0, 2, 96, kind="iid", seed=3).to(DEV)
    out = m(buf)
    buf.copy_(FR.synthetic_frames_u8(9, 2, 96, kind="iid", seed=3).to(DEV))     # staging buffer refilled
    with pytest.raises(RuntimeError):
        out.img_embedding
    m.eager_outputs = True
    out = m(buf)
    pooled = out.img_embedding.clone()
    buf.zero_()
    assert torch.equal(out.img_embedding, pooled) and out._frames is None
    m.eager_outputs = False
    m2 = copy.deepcopy(m)
    b = io.BytesIO()
    torch.save(m, b)
    b.seek(0)
    m3 = torch.load(b, weights_only=False)
    x = FR.synthetic_frames_u8(4, 2, 96, kind="structured", seed=3).to(DEV)
    assert torch.equal(m2(x), m(x)) and torch.equal(m3.eval()(x), m(x))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """kernel attributes / SM count / cluster capability are per-device state - an ImageModel on cuda:1 and a
    scorer on cuda:0 in the same process must both launch (dynamic shared memory opt-in applies per device)."""
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    from incremental_multimodal_medical_learning_ii_b200.scorer import ZeroShotScorer
    m0, sd = _model(True)
    fr = FR.synthetic_frames_u8(0, 4, 480, kind="structured", seed=0)
    a = m0(fr.to("cuda:0")).cpu()
    m1 = get_model_on("cuda:1", sd)
    b = m1(fr.to("cuda:1")).cpu()
    assert torch.equal(a, b)
    prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
    s0, s1 = ZeroShotScorer("cuda:0"), ZeroShotScorer("cuda:1")
    s0.set_prompts(prompts); s1.set_prompts(prompts)
    assert torch.equal(s0.score(a)["prob"].cpu(), s1.score(a)["prob"].cpu())


def get_model_on(device, sd):
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    m = get_biovil_resnet(None)
    m.load_state_dict(sd)
    m.eval().to(device)
    return m

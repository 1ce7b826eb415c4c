"""Generate ``tests/golden/*.pt`` from the reference itself (run HERE, where /root/reference exists).

    python oracle/make_golden.py

1. builds the seeded state_dicts (``oracle/weights.py``) and loads them into the reference's own ``ImageModel``
   (imported through ``oracle/reference_shim.py``, nothing under /root/reference is modified);
2. runs the reference on seeded synthetic frames and ASSERTS that the restatement ``oracle/biovil_oracle.py``
   reproduces it (this is what pins the oracle);
3. stores the reference outputs as small fixtures, together with everything needed to regenerate the inputs.

The scorer half (Trainer.py) cannot be imported (torchmetrics / matplotlib / HF download in ``Trainer.__init__``), so
its fixtures come from the restatement, cross-checked against the explicit normalise-then-matmul formulation the
reference also uses (``trash/lower_bound_mcs.py:82,101-111``, ``vlp/inference_engine.py:52-55``).
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import biovil_oracle as O  # noqa: E402
import reference_shim as RS  # noqa: E402
import weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
N_FRAMES = 8            # frames whose pooled / patch / trunk outputs are stored
N_GLOBAL = 256          # BASELINE.json configs[0]: 256 frames - global embeddings + scores of ALL of them are stored
ORACLE_CHECK = 16       # frames on which the restatement is asserted equal to the reference (per variant x kind)
SIZE = 480
MARGIN = 5e-3           # prompt set whose fp32 decisions sit above the bf16 noise floor


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = RS.load_reference_image_model(seed=27)
    assert ref.training is False

    out = {"meta": {"n_frames": N_FRAMES, "size": SIZE, "frame_seed": 0, "weight_seed": 27, "bn_seed": 28,
                    "prompt_seed": 29, "torch": str(torch.__version__)}}
    for variant, rbn in (("default", False), ("bnrand", True)):
        sd = Wt.make_state_dict(27, randomize_bn=rbn)
        missing = set(ref.state_dict()) ^ set(sd)
        assert not missing, missing
        ref.load_state_dict(sd)
        ref.eval()
        for kind in ("iid", "structured"):
            fr_all = FR.synthetic_frames_u8(0, N_GLOBAL, SIZE, kind=kind, seed=0)
            fr = fr_all[:N_FRAMES]
            x = FR.frames_as_reference_input(fr)
            with torch.no_grad():
                g_ref = ref(x).detach()                                         # model.py:141-154
                x4_ref, pooled_ref = ref.encoder(x, return_patch_embeddings=True)
                p_ref = ref.projector(x4_ref)
                # configs[0] at its stated size: the reference's own forward on all 256 frames (batches of 16)
                # (.detach(): the fork's forward re-enables grad, model.py:142, so every output would pin its autograd graph)
                g_all = torch.cat([ref(FR.frames_as_reference_input(fr_all[i:i + 16])).detach()
                                   for i in range(0, N_GLOBAL, 16)])
            assert torch.equal(g_all[:N_FRAMES], g_ref) or (g_all[:N_FRAMES] - g_ref).abs().max() <= 1e-6 * g_ref.abs().max()
            o = O.image_model_forward(sd, x)
            o16 = O.image_model_forward(sd, FR.frames_as_reference_input(fr_all[N_GLOBAL - ORACLE_CHECK:]))
            err16 = (o16["projected_global_embedding"] - g_all[N_GLOBAL - ORACLE_CHECK:]).abs().max().item() / g_all.abs().max().item()
            print(f"[{variant}/{kind}] oracle vs reference, last {ORACLE_CHECK} of {N_GLOBAL} frames: {err16:.3e}")
            assert err16 <= 1e-6
            for name, a, b in (("global", g_ref, o["projected_global_embedding"]),
                               ("pooled", pooled_ref, o["img_embedding"]),
                               ("patch", p_ref, o["projected_patch_embeddings"]),
                               ("trunk", x4_ref, o["patch_embedding"])):
                err = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-30)
                print(f"[{variant}/{kind}] oracle vs reference {name}: max rel-to-max err {err:.3e}")
                assert err <= 1e-6, (variant, kind, name, err)
            patch_norm = F.normalize(p_ref, dim=1).permute(0, 2, 3, 1).contiguous()        # model.py:172-174
            key = f"{variant}/{kind}"
            out[key] = {
                "weights_checksum": Wt.state_dict_checksum(sd),
                "frames_checksum": int(fr.long().sum()),
                "global": g_ref.clone(),
                "pooled": pooled_ref.clone(),
                "patch_norm_first2": patch_norm[:2].clone(),
                "patch_raw_first2": p_ref[:2].permute(0, 2, 3, 1).contiguous().clone(),
                "trunk_absmax": float(x4_ref.abs().max()),
                "global256": g_all.clone(),
                "frames256_checksum": int(fr_all.long().sum()),
            }
            # high-margin prompt sets (every |pos - neg| of the REFERENCE embeddings > MARGIN) and their decisions:
            # the GPU path must reproduce these labels strictly
            for pname, P, seed in (("p1", 1, 41), ("p5", 5, 43)):
                hp = FR.synthetic_prompt_embeddings(14, P, 128, seed=seed, min_margin_against=g_all, min_margin=MARGIN)
                s = O.zero_shot_score(g_all, hp, reduce="mean")
                assert ((s["sim"][..., 0] - s["sim"][..., 1]).abs() > MARGIN).all()
                out[key][f"margin_{pname}"] = {"prompts": hp.clone(), "sim": s["sim"].clone(), "prob": s["prob"].clone(),
                                               "pred": s["pred"].clone(), "min_margin": MARGIN}
            # plain (seed 29) prompt sets on all 256 reference embeddings
            for pname, P in (("p1", 1), ("p5", 5)):
                prompts = FR.synthetic_prompt_embeddings(14, P, 128, seed=29)
                s = O.zero_shot_score(g_all, prompts, reduce="mean")
                out[key][f"score256_{pname}_mean"] = {k: s[k].clone() for k in ("sim", "prob", "pred")}
            # zero-shot scores of the reference embeddings against synthetic prompts (restated scorer)
            for pname, P in (("p1", 1), ("p5", 5)):
                prompts = FR.synthetic_prompt_embeddings(14, P, 128, seed=29)
                for reduce in ("mean", "max"):
                    s = O.zero_shot_score(g_ref, prompts, reduce=reduce)
                    out[key][f"score_{pname}_{reduce}"] = {k: v.clone() for k, v in s.items()}
            heat = O.patch_similarity_map(patch_norm[:2], FR.synthetic_prompt_embeddings(14, 5, 128, seed=29)[:, 0])
            out[key]["heat_first2_p5"] = heat.clone()
            # explicit normalise-then-matmul cross-check of the scorer restatement
            prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
            t = F.normalize(prompts[:, :, 0], dim=-1)
            alt = torch.einsum("bd,lpd->blp", F.normalize(g_ref, dim=-1), t)
            assert (alt - out[key]["score_p1_mean"]["sim"]).abs().max() < 1e-6
    torch.save(out, os.path.join(GOLDEN_DIR, "biovil_golden.pt"))
    print("wrote", os.path.join(GOLDEN_DIR, "biovil_golden.pt"), os.path.getsize(os.path.join(GOLDEN_DIR, "biovil_golden.pt")))


if __name__ == "__main__":
    main()

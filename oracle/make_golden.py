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
N_FRAMES = 8
SIZE = 480


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
            fr = FR.synthetic_frames_u8(0, N_FRAMES, SIZE, kind=kind, seed=0)
            x = FR.frames_as_reference_input(fr)
            with torch.no_grad():
                g_ref = ref(x)                                                  # model.py:141-154
                x4_ref, pooled_ref = ref.encoder(x, return_patch_embeddings=True)
                p_ref = ref.projector(x4_ref)
            o = O.image_model_forward(sd, x)
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
            }
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

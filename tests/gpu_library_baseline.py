"""Same-box GPU LIBRARY baseline (SURVEY 8d "Reference beside it (2)"): the reference's encoder arithmetic - the oracle's
plain torch.nn.functional restatement of ImageModel.forward - run on the B200 through cuDNN/cuBLAS in PyTorch eager,
fp32 (TF32 allowed, PyTorch's default for convolutions) and bf16 autocast + channels_last, plus the scorer in torch.
This is the "existing Blackwell library path" the hand-written kernels are compared with; it is NOT product code and
is not used by bench.py's arms.  Run on a GPU box:   python tests/gpu_library_baseline.py [batch] [steps]
Prints one JSON line per variant."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import biovil_oracle as O  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in Wt.make_state_dict(27).items()}
    prompts = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29).to(dev)
    fr = torch.cat([FR.synthetic_frames_u8(o, min(64, batch - o), 480, kind="structured", device=dev)
                    for o in range(0, batch, 64)])
    x = (fr.float() / 255.0).expand(-1, 3, -1, -1).contiguous()          # ToTensor + ExpandChannels

    def score(emb):
        t = torch.nn.functional.normalize(prompts.mean(dim=2), dim=-1)    # [L,2,128]
        e = torch.nn.functional.normalize(emb.float(), dim=-1)
        sim = torch.einsum("bd,lpd->blp", e, t)
        return torch.sigmoid(sim[..., 0] - sim[..., 1])

    variants = {
        "torch_eager_fp32_tf32": dict(autocast=False, channels_last=False),
        "torch_eager_bf16_channels_last": dict(autocast=True, channels_last=True),
    }
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    for name, v in variants.items():
        xin = x.contiguous(memory_format=torch.channels_last) if v["channels_last"] else x
        sdv = {k: (t.contiguous(memory_format=torch.channels_last) if (v["channels_last"] and t.dim() == 4) else t)
               for k, t in sd.items()}

        def step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=v["autocast"]):
                emb = O.image_model_forward(sdv, xin)["projected_global_embedding"]
            return score(emb)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            p = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"baseline": name, "images_per_s": batch / (ms / 1000.0), "batch": batch, "ms_per_step": ms,
                          "torch": torch.__version__, "checksum": float(p.sum())}), flush=True)


if __name__ == "__main__":
    main()

"""Batch-1 latency: the way the UNMODIFIED reference extraction script drives the model (chexpert-get-embedding.py:47-49,
68-80: DataLoader batch_size 1, fp32 [1,3,512,512] frames = ToTensor + ExpandChannels of an 8-bit radiograph, one
``resnet50(images)`` call per frame, ``torch.cat`` of the result).

Measured per call, synchronised after each (the loop's next step depends on nothing, but a per-frame latency is what
"batch size 1" means): (a) this package with CUDA-graph replay (default for B <= 32), (b) the same with direct launches,
(c) the reference's arithmetic through cuDNN/cuBLAS in PyTorch eager, fp32 with TF32 convolutions - the oracle's
functional restatement moved to the GPU, which is what running the reference module there executes.
NOT product code.  Usage (GPU box):  python tests/latency_b1.py [size] [calls]   -> one JSON line per variant."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import biovil_oracle as O  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet  # noqa: E402


def timed(fn, frames, calls):
    for x in frames[:5]:
        fn(x)
    torch.cuda.synchronize()
    lat = []
    for i in range(calls):
        x = frames[i % len(frames)]
        t0 = time.perf_counter()
        out = fn(x)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    return {"ms_median": lat[len(lat) // 2], "ms_p10": lat[len(lat) // 10], "ms_p90": lat[9 * len(lat) // 10],
            "frames_per_s": 1e3 / (sum(lat) / len(lat))}, out


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda:0")
    sd = Wt.make_state_dict(27)
    model = get_biovil_resnet(None)
    model.load_state_dict(sd)
    model.train(mode=False, my_freeze=True)
    model.eval().to(dev)
    frames = [FR.frames_as_reference_input(FR.synthetic_frames_u8(i, 1, size, kind="structured", seed=0)).to(dev)
              for i in range(8)]                                             # [1,3,size,size] fp32, as the DataLoader yields
    results = {}
    with torch.no_grad():
        model.cuda_graphs = "auto"
        r, a = timed(lambda x: model(x), frames, calls)
        results["b200_cuda_graph"] = r
        model.cuda_graphs = False
        r, b = timed(lambda x: model(x), frames, calls)
        results["b200_direct_launches"] = r
        assert torch.equal(a, b), "graph replay and direct launches must give identical embeddings"
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        r, c = timed(lambda x: O.image_model_forward(sd_dev, x)["projected_global_embedding"], frames, min(calls, 50))
        results["torch_eager_cudnn_fp32_tf32"] = r
        cos = torch.nn.functional.cosine_similarity(a.float(), c.float(), dim=-1).min().item()
    for k, v in results.items():
        print(json.dumps({"variant": k, "input": f"[1,3,{size},{size}] fp32", "calls": calls, "launches": model._engine.launches(),
                          **{kk: round(vv, 4) for kk, vv in v.items()}, "cosine_vs_cudnn_eager": round(cos, 6)}))


if __name__ == "__main__":
    main()

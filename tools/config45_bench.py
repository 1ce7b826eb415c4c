"""BASELINE.json configs[3] and configs[4] timed on one B200 (device-resident, CUDA events, batch 512 of 1x480x480 u8):
  configs[3]  multi-prompt scoring: 14 labels x 5 pos/neg prompts (140 prompt embeddings), mean and max reduction
  configs[4]  normalised 15x15 patch embeddings + per-patch prompt heat-maps for 14 labels (+ Gaussian smoothing)
next to the plain configs[1] step.  Usage: python tools/config45_bench.py [steps]   -> one JSON line per variant."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
B = 512
dev = torch.device("cuda:0")
m = get_biovil_resnet(None)
m.load_state_dict(Wt.make_state_dict(27))
m.eval().to(dev)
fr = torch.cat([FR.synthetic_frames_u8(o, 64, 480, kind="structured", device=dev) for o in range(0, B, 64)])
p1 = FR.synthetic_prompt_embeddings(14, 1, 128, seed=29)
p5 = FR.synthetic_prompt_embeddings(14, 5, 128, seed=29)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


variants = []
m.set_prompts(p1, reduce="mean")
variants.append(("configs[1] 28 prompts, global embedding + scores", lambda: m.embed_and_score(fr)))
m2 = m


def with_prompts(p, reduce, **kw):
    def run():
        return m.embed_and_score(fr, **kw)
    return run


for name, p, reduce, kw in (("configs[3] 140 prompts, mean over 5 per polarity", p5, "mean", {}),
                            ("configs[3] 140 prompts, max over 5 per polarity (MAX_EMB)", p5, "max", {}),
                            ("configs[4] patch embeddings [512,15,15,128] + heat-maps [512,15,15,14]", p5, "mean",
                             dict(heat=True, patch=True))):
    m.set_prompts(p, reduce=reduce)
    ms, out = timed(with_prompts(p, reduce, **kw))
    line = {"variant": name, "batch": B, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3)}
    if "heat" in out:
        ms_s, _ = timed(lambda: m.smooth_heatmaps(out["heat"], 1.5))
        line["gaussian_smoothing_ms"] = round(ms_s, 4)
    print(json.dumps(line), flush=True)
m.set_prompts(p1, reduce="mean")
ms, _ = timed(variants[0][1])
print(json.dumps({"variant": variants[0][0], "batch": B, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3)}))

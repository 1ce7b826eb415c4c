"""One warm-up forward + one forward of configs[1] (batch 512, 1x480x480 u8) for use under ncu:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'conv_gemm|chain_gemm|conv3x3_tap3|l1_block|stem_|head' -s 45 -c 45 ... (one forward = 45 launches: stem, 43 conv kernels, head)
Usage: python tools/ncu_step.py [batch] [forwards]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = get_biovil_resnet(None)
m.load_state_dict(Wt.make_state_dict(27))
m.eval().to("cuda:0")
m.set_prompts(FR.synthetic_prompt_embeddings(14, 1, 128, seed=29), reduce="mean")
fr = torch.cat([FR.synthetic_frames_u8(o, min(64, B - o), 480, kind="structured", device="cuda:0")
                for o in range(0, B, 64)])
for _ in range(n):
    res = m.embed_and_score(fr)
torch.cuda.synchronize()
print("ok", float(res["prob"].sum()), "launches_per_forward", m._get_engine().launches())

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt
from incremental_multimodal_medical_learning_ii_b200 import frames as FR
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
m = get_biovil_resnet(None); m.load_state_dict(Wt.make_state_dict(27)); m.eval().to("cuda:0")
fr = torch.cat([FR.synthetic_frames_u8(o, 64, 480, kind="structured", device="cuda:0") for o in range(0, 512, 64)])
m(fr); torch.cuda.synchronize()
os.environ["BV_TIMING"] = "1"
m(fr); torch.cuda.synchronize()

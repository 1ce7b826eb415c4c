import ctypes, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from incremental_multimodal_medical_learning_ii_b200 import _native as N, packing
lib = N.lib(); dev = torch.device("cuda:0")
os.environ["BV_FORCE_CFG"] = sys.argv[1] if len(sys.argv) > 1 else "6"
B, H, C = 1, 15, 64
x = torch.zeros(B, H, H, C)
for y in range(H):
    for xx in range(H):
        x[0, y, xx, 0] = y * 16 + xx + 1          # channel 0 encodes the pixel position
x = x.to(torch.bfloat16).to(dev)
for r in range(3):
    for s in range(3):
        w = torch.zeros(64, 64, 3, 3); w[0, 0, r, s] = 1
        conv = packing.pack_single_conv(w.to(torch.bfloat16), torch.zeros(64), 1, 1, dev)
        out = torch.full((B, H, H, 64), float("nan"), device=dev, dtype=torch.bfloat16)
        N.check(lib.bv_conv2d_nhwc(N.ptr(x), B, H, H, ctypes.byref(conv[0]), None, 0, 0, None, None, 0, N.ptr(out), 0,
                                   N.current_stream_handle(dev)))
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dev), padding=1).permute(0, 2, 3, 1)
        o, rr = out[0, :, :, 0].float().cpu(), ref[0, :, :, 0].cpu()
        bad = (o != rr)
        print(f"tap ({r},{s}): mismatches {int(bad.sum())}/{H*H}  nan {int(torch.isnan(o).sum())}")
        if bad.any() and (r, s) in ((0, 0), (1, 1), (1, 2)):
            print(" got rows 0..3:\n", o[:4, :].int() if not torch.isnan(o).any() else o[:4])
            print(" ref rows 0..3:\n", rr[:4, :].int())

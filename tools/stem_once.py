"""Launch the row-streaming stem kernel a few times at the bench shape (ncu target)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_stem_gpu as T  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import _native as N  # noqa: E402

B, H, W = 512, 480, 480
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8).to(T.DEV)
w = (torch.randn(64, 7, 7, generator=g) / 7 / 255).to(torch.bfloat16)
conv, keep, _ = T._pack_w8(w.float(), torch.randn(64, generator=g), T.DEV)
out = torch.empty((B, H // 4, W // 4, 64), device=T.DEV, dtype=torch.bfloat16)
lib = N.lib()
st = N.current_stream_handle(torch.device(T.DEV))
for i in range(int(os.environ.get("N_LAUNCH", "3"))):
    N.check(lib.bv_stem_u8_nhwc(N.ptr(frames), B, H, W, ctypes.byref(conv), N.ptr(out), 0, st))
torch.cuda.synchronize()
print("ok")

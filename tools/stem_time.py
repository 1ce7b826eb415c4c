"""Time the two 8-bit stem kernels alone (bv_stem_u8_nhwc) at the bench shape: 512 frames of 480x480."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_stem_gpu as T  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import _native as N  # noqa: E402

B, H, W = 512, 480, 480
g = torch.Generator().manual_seed(0)
frames = [torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8).to(T.DEV) for _ in range(2)]
w = (torch.randn(64, 7, 7, generator=g) / 7 / 255).to(torch.bfloat16)
conv, keep, _ = T._pack_w8(w.float(), torch.randn(64, generator=g), T.DEV)
out = torch.empty((B, H // 4, W // 4, 64), device=T.DEV, dtype=torch.bfloat16)
lib = N.lib()
st = N.current_stream_handle(torch.device(T.DEV))
import os
CASES = [(1, "", "", "", "tile (stem_fused)"), (0, "1", "", "", "rows, 8 epilogue warps"), (0, "", "", "", "rows, 16 epilogue warps"),
         (0, "1", "", "", "rows, 8 epilogue warps"), (0, "", "", "", "rows, 16 epilogue warps"),
         (0, "1", "", "2", "rows, 8 epilogue warps, no MMA"), (0, "1", "", "3", "rows, 8 epilogue warps, no loads, no MMA")]
for variant, cg2, raw, dbg, name in CASES:
    os.environ["BV_SR_CG4"] = "" if cg2 else "1"
    os.environ["BV_SR_RAW"] = raw
    os.environ["BV_SR_DEBUG"] = dbg or "0"
    for i in range(3):
        N.check(lib.bv_stem_u8_nhwc(N.ptr(frames[i & 1]), B, H, W, ctypes.byref(conv), N.ptr(out), variant, st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for i in range(n):
        N.check(lib.bv_stem_u8_nhwc(N.ptr(frames[i & 1]), B, H, W, ctypes.byref(conv), N.ptr(out), variant, st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gb = (B * H * W + out.numel() * 2) / 1e9
    print(f"{name}: {ms:.3f} ms per launch, {gb / ms * 1e3:.0f} GB/s algorithmic")

# fused conv1 variant (bv_stem_conv1_u8_nhwc): max-pool output + layer1.0 conv1 output from one kernel
from incremental_multimodal_medical_learning_ii_b200 import packing  # noqa: E402
os.environ["BV_SR_DEBUG"] = "0"
c1 = packing.pack_single_conv((torch.randn(64, 64, 1, 1, generator=g) / 8).to(torch.bfloat16), torch.randn(64, generator=g), 1, 0,
                              torch.device(T.DEV))
out1 = torch.empty_like(out)
for cg4, dbg, name in (("", "0", "rows + conv1, 8 epilogue warps"), ("1", "0", "rows + conv1, 16 epilogue warps"),
                       ("", "4", "rows + conv1, 8 warps, no conv1 stores"), ("", "8", "rows + conv1, 8 warps, no conv1 epilogue math/stores"),
                       ("", "16", "rows + conv1, 8 warps, no A-buffer stores"), ("", "28", "rows + conv1, 8 warps, none of the three")):
    os.environ["BV_SR_CG4"] = cg4
    os.environ["BV_SR_DEBUG"] = dbg
    for i in range(3):
        N.check(lib.bv_stem_conv1_u8_nhwc(N.ptr(frames[i & 1]), B, H, W, ctypes.byref(conv), ctypes.byref(c1[0]), N.ptr(out),
                                          N.ptr(out1), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        N.check(lib.bv_stem_conv1_u8_nhwc(N.ptr(frames[i & 1]), B, H, W, ctypes.byref(conv), ctypes.byref(c1[0]), N.ptr(out),
                                          N.ptr(out1), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = (B * H * W + 2 * out.numel() * 2) / 1e9
    print(f"{name}: {ms:.3f} ms per launch, {gb / ms * 1e3:.0f} GB/s algorithmic")

"""Time single convolutions through bv_conv2d_nhwc with CUDA events (L2-resident or streaming inputs)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from incremental_multimodal_medical_learning_ii_b200 import _native as N  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import packing  # noqa: E402


def run(name, B, H, cin, cout, k, stride, pad, env=None, residual=False, iters=20):
    for kk, vv in (env or {}).items():
        os.environ[kk] = vv
    lib = N.lib()
    dev = torch.device("cuda:0")
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k) * (cin * k * k) ** -0.5).to(torch.bfloat16)
    conv = packing.pack_single_conv(w, torch.zeros(cout), stride, pad, dev)
    Ho = (H + 2 * pad - k) // stride + 1
    out = torch.empty(B, Ho, Ho, cout, device=dev, dtype=torch.bfloat16)
    res = torch.randn(B, Ho, Ho, cout, device=dev).to(torch.bfloat16) if residual else None
    st = N.current_stream_handle(dev)

    def call():
        N.check(lib.bv_conv2d_nhwc(N.ptr(x), B, H, H, ctypes.byref(conv[0]), None, 0, 0, None, N.ptr(res), 1,
                                   N.ptr(out), 0, st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    M = B * Ho * Ho
    flops = 2.0 * M * cout * cin * k * k
    a_bytes = M * cin * k * k * 2
    print(f"{name:42s} M={M:8d} {ms * 1000:8.1f} us  {flops / ms / 1e9:7.1f} TF  A-tile fill {a_bytes / ms / 1e6 / 148:6.1f} GB/s/SM"
          f"  in+out {(x.numel() + out.numel()) * 2 / ms / 1e6:7.0f} GB/s", flush=True)
    for kk in (env or {}):
        os.environ.pop(kk, None)


if __name__ == "__main__":
    # L2-resident inputs (<= ~30 MB): isolates TMA / MMA / epilogue throughput from HBM
    run("3x3 64->64 120^2 B=8 (im2col)", 8, 120, 64, 64, 3, 1, 1)
    run("3x3 64->64 120^2 B=8 no-BRES", 8, 120, 64, 64, 3, 1, 1, env={"BV_NO_BRES": "1"})
    run("3x3 64->64 120^2 B=8 nacc=1", 8, 120, 64, 64, 3, 1, 1, env={"BV_FORCE_NACC": "1"})
    run("1x1 64->64 120^2 B=16 tiled", 16, 120, 64, 64, 1, 1, 0)
    run("1x1 64->64 120^2 B=16 im2col", 16, 120, 64, 64, 1, 1, 0, env={"BV_FORCE_IM2COL": "1"})
    run("1x1 576->64 120^2 B=2 tiled", 2, 120, 576, 64, 1, 1, 0)
    run("1x1 576->64 120^2 B=2 im2col", 2, 120, 576, 64, 1, 1, 0, env={"BV_FORCE_IM2COL": "1"})
    run("1x1 256->64 120^2 B=4 tiled", 4, 120, 256, 64, 1, 1, 0)
    run("1x1 256->64 120^2 B=4 im2col", 4, 120, 256, 64, 1, 1, 0, env={"BV_FORCE_IM2COL": "1"})
    run("3x3 128->128 60^2 B=16", 16, 60, 128, 128, 3, 1, 1)
    run("3x3 128->128 60^2 B=16 nacc=1", 16, 60, 128, 128, 3, 1, 1, env={"BV_FORCE_NACC": "1"})
    run("1x1 1152->128 60^2 B=4 tiled", 4, 60, 1152, 128, 1, 1, 0)
    run("3x3 256->256 30^2 B=32", 32, 30, 256, 256, 3, 1, 1)
    run("1x1 2304->256 30^2 B=8 tiled", 8, 30, 2304, 256, 1, 1, 0)
    # streaming
    run("3x3 64->64 120^2 B=512 (stream)", 512, 120, 64, 64, 3, 1, 1, iters=5)

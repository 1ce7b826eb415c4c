"""Bisect helper for csrc/stem_rows.cuh: runs the smallest stem case in a fresh process per BV_SR_DEBUG mode."""
import os
import subprocess
import sys

CODE = r'''
import ctypes, sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_stem_gpu as T
g = torch.Generator().manual_seed(1)
B, H, W = map(int, sys.argv[1:4])
frames = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8)
w = torch.randint(-1, 2, (64, 7, 7), generator=g).float()
bias = torch.randint(-300, 300, (64,), generator=g).float()
conv, keep, b_eff = T._pack_w8(w, bias, T.DEV)
ref = T._reference(frames, w.to(torch.bfloat16), b_eff)
out = T._run(frames, conv, 0)
d = (out.float() - ref.float())
bad = (out != ref)
print("OK mismatches", bad.float().mean().item(), "nan", torch.isnan(out.float()).float().mean().item())
if bad.any():
    idx = bad.nonzero()
    print("first bad", idx[:5].tolist(), "rows bad", sorted(set(idx[:,1].tolist()))[:20], "cols bad", sorted(set(idx[:,2].tolist()))[:40],
          "ch bad", sorted(set(idx[:,3].tolist()))[:70])
    i = idx[0].tolist()
    print("out", out[i[0], i[1], i[2], :8].tolist(), "ref", ref[i[0], i[1], i[2], :8].tolist())
'''
for shape in sys.argv[1].split(","):
    for mode in sys.argv[2].split(","):
        env = dict(os.environ, BV_SR_DEBUG=mode)
        r = subprocess.run([sys.executable, "-c", CODE, *shape.split("x")], env=env, capture_output=True, text=True, timeout=120)
        tail = (r.stdout.strip().splitlines() or [""])[-4:] if r.returncode == 0 else (r.stderr.strip().splitlines() or [""])[-1:]
        print(f"shape {shape} mode {mode}: rc={r.returncode} :: " + " | ".join(tail), flush=True)

"""Throughput of the GPU input stage: JPEG bytes -> nvJPEG decode -> PIL-exact Resize/CenterCrop on the device -> ImageModel,
next to the reference's host stage (PIL decode + PIL resize, what DataRetrieval.py:70-96, 175-180 does per worker).
CheXpert-small sized grey JPEGs (390x320, quality 90).  Usage: python tools/jpeg_pipeline_bench.py [n]  -> JSON lines."""
import io
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image.data.gpu_decode import GpuJpegPipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
datas = []
for i in range(64):
    base = rng.integers(20, 230, size=(21, 25)).astype(np.float32)
    img = np.clip(np.kron(base, np.ones((16, 16), np.float32))[:320, :390] + rng.normal(0, 5, size=(320, 390)), 0, 255).astype(np.uint8)
    b = io.BytesIO()
    Image.fromarray(img, mode="L").save(b, format="JPEG", quality=90)
    datas.append(b.getvalue())
datas = [datas[i % 64] for i in range(n)]
dev = "cuda:0"
model = get_biovil_resnet(None)
model.load_state_dict(Wt.make_state_dict(27))
model.eval().to(dev)
pipe = GpuJpegPipeline(dev, resize=512, center_crop_size=480)
pipe(datas[:8])
torch.cuda.synchronize()
t0 = time.perf_counter()
frames = pipe(datas)
torch.cuda.synchronize()
t_dec = time.perf_counter() - t0
t0 = time.perf_counter()
emb = model(frames)
torch.cuda.synchronize()
t_emb = time.perf_counter() - t0
print(json.dumps({"stage": "GPU: nvJPEG decode + Resize(512)/CenterCrop(480) on the device (one host thread)", "frames": n,
                  "images_per_s": round(n / t_dec), "ms_per_frame": round(1e3 * t_dec / n, 3),
                  "jpeg_bytes_per_frame": int(np.mean([len(d) for d in datas]))}))
for th in (4, 8, 16):
    pipe_t = GpuJpegPipeline(dev, resize=512, center_crop_size=480, threads=th)
    pipe_t(datas[:32])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fr_t = pipe_t(datas)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert torch.equal(fr_t, frames)
    print(json.dumps({"stage": f"GPU: nvJPEG decode + resize on the device, {th} host threads", "frames": n,
                      "images_per_s": round(n / dt), "host_cores": os.cpu_count()}))
for backend, bname in ((3, "hardware JPEG engines"), (2, "GPU-assisted Huffman")):
    try:
        dec = pipe.decoder
        dec.decode_batched(datas[:64], backend)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = []
        for o in range(0, n, 64):
            outs += dec.decode_batched(datas[o:o + 64], backend)
        fr_b = pipe.transform(outs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        diff = (fr_b.int() - frames.int()).abs().max().item()
        print(json.dumps({"stage": f"GPU: nvjpegDecodeBatched ({bname}), batches of 64, + resize on the device", "frames": n,
                          "images_per_s": round(n / dt), "max_abs_diff_vs_default_backend": diff}))
    except Exception as e:
        print(json.dumps({"stage": f"GPU: nvjpegDecodeBatched ({bname})", "unavailable": str(e)[:160]}))
for th in (2, 4, 8):
    try:
        pipe_b = GpuJpegPipeline(dev, resize=512, center_crop_size=480, threads=th, batched=True)
        pipe_b(datas)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fr_b = pipe_b(datas)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        diff = (fr_b.int() - frames.int()).abs().max().item()
        print(json.dumps({"stage": f"GPU: nvjpegDecodeBatched (GPU-assisted Huffman), {th} host threads x {n // th} frames, + resize on the device",
                          "frames": n, "images_per_s": round(n / dt), "max_abs_diff_vs_default_backend": diff}))
    except Exception as e:
        print(json.dumps({"stage": f"GPU: threaded nvjpegDecodeBatched, {th} threads", "unavailable": str(e)[:160]}))
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.pipeline import JpegBytesPipeline  # noqa: E402
model.set_prompts(FR.synthetic_prompt_embeddings(14, 1, 128, seed=29), reduce="mean")
for _ in range(2):
    model.embed_and_score(frames)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    model.embed_and_score(frames)
torch.cuda.synchronize()
t_emb = (time.perf_counter() - t0) / 4
print(json.dumps({"stage": "GPU: ImageModel.embed_and_score on a decoded batch (warm)", "frames": n, "images_per_s": round(n / t_emb)}))
for th in (2, 4):
    jp = JpegBytesPipeline(model, resize=512, center_crop_size=480, threads=th)
    for _ in jp.run([datas] * 2):
        pass
    torch.cuda.synchronize()
    nb = 10
    t0 = time.perf_counter()
    chk = 0
    for res in jp.run([datas] * nb):
        last = res
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"stage": f"GPU: JPEG bytes -> embeddings + scores, decode of batch i+1 ({th} host threads, GPU-assisted Huffman) under the model on batch i",
                      "frames": nb * n, "images_per_s": round(nb * n / dt), "batched_backend": jp.stage.batched,
                      "h2d_bytes_per_frame": int(np.mean([len(d) for d in datas]))}))
from torchvision import transforms  # noqa: E402
tf = transforms.Compose([transforms.Resize(512), transforms.CenterCrop(480)])
t0 = time.perf_counter()
for d in datas[:128]:
    np.asarray(tf(Image.open(io.BytesIO(d)).convert("L")))
t_host = (time.perf_counter() - t0) / 128
print(json.dumps({"stage": "host reference: PIL decode + Resize(512) + CenterCrop(480), one core", "images_per_s": round(1 / t_host),
                  "ms_per_frame": round(1e3 * t_host, 3)}))

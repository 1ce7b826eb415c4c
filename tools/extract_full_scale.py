"""configs[2] of BASELINE.json at full size: CheXpert-scale extraction of N synthetic frames (default 224,000) sharded
over the ranks of one box, embedded + zero-shot scored in batches of 512 (ragged tail batch included), ONE all-gather of
embeddings / probabilities / labels at the end (chexpert-get-embedding.py:68-113).

    python tools/extract_full_scale.py --out profiles/x.json                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/extract_full_scale.py --out profiles/x.json                           # 2 GPUs

Reports images/s over the model calls alone (CUDA events around every batch, frame synthesis excluded), including
and excluding the gather, and checks size-independent properties of the gathered result:
  * order: row i of the gathered tensors is frame i (re-embedding a sample of frames, alone, reproduces rows bit for bit:
    batch invariance + shard order + ragged tail),
  * every probability is sigmoid(pos - neg) in (0, 1), every label is prob > 0.5, no NaN, and the checksums of the
    gathered tensors are identical on all ranks."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from incremental_multimodal_medical_learning_ii_b200 import synthetic_weights as Wt  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import extraction as EX  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200 import frames as FR  # noqa: E402
from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=224000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = get_biovil_resnet(None)
    model.load_state_dict(Wt.make_state_dict(27))
    model.eval().to(dev)
    model.set_prompts(FR.synthetic_prompt_embeddings(14, 1, 128, seed=29), reduce="mean")

    def source(first, count):
        return FR.synthetic_frames_u8(first, count, 480, kind="structured", seed=0, device=dev)

    events = []

    def embed(frames):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = model.embed_and_score(frames)
        e1.record()
        events.append((e0, e1))
        return res

    embed(source(0, args.batch))          # warm-up: plan, tensor maps, clocks
    events.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    local = EX.extract_shard(embed, source, args.frames, args.batch, rank, world)
    torch.cuda.synchronize()
    t_forward = time.perf_counter() - t0
    model_ms = sum(a.elapsed_time(b) for a, b in events)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    full = EX.gather_shards(local, args.frames, rank, world)
    g1.record()
    torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1)
    t_wall = time.perf_counter() - t0

    # ---- properties of the gathered result ----
    emb, prob, pred = full["global"], full["prob"], full["pred"]
    assert emb.shape == (args.frames, 128) and prob.shape == (args.frames, 14) and pred.shape == (args.frames, 14)
    assert not torch.isnan(emb).any() and not torch.isnan(prob).any()
    assert ((prob > 0) & (prob < 1)).all()
    mism = pred.bool() != (prob > 0.5)            # pred is pos > neg; sigmoid(pos - neg) may round to exactly 0.5
    assert not (mism & ((prob - 0.5).abs() > 1e-6)).any()
    start, end = EX.shard_range(args.frames, rank, world)
    probe = sorted({0, 1, 511, 512, args.frames // 3, args.frames // 2, args.frames - 1, start, end - 1,
                    end - 1 - (end - start) % args.batch})
    for i in probe:
        alone = model.embed_and_score(source(i, 1))
        assert torch.equal(alone["global"][0], emb[i]), f"frame {i}: gathered row differs from the frame embedded alone"
        assert torch.equal(alone["prob"][0], prob[i])
    checks = torch.stack([emb.double().sum(), emb.double().abs().sum(), prob.double().sum(), pred.double().sum()])
    if world > 1:
        lo, hi = checks.clone(), checks.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the gathered tensors"
        stats = torch.tensor([model_ms, gather_ms, t_forward, t_wall], device=dev, dtype=torch.float64)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)       # max over ranks, as bench.py does
        model_ms, gather_ms, t_forward, t_wall = stats.tolist()
    if rank == 0:
        line = {
            "workload": f"configs[2]: {args.frames} synthetic 1x480x480 8-bit frames sharded over {world} B200, batch {args.batch} "
                        f"(+ ragged tail), one all-gather of emb/prob/pred",
            "n_gpus": world, "frames": args.frames, "batches_per_rank": len(events),
            "tail_batch": (end - start) % args.batch,
            "images_per_s_model_only": args.frames / (model_ms / 1e3),
            "images_per_s_model_plus_gather": args.frames / ((model_ms + gather_ms) / 1e3),
            "images_per_s_wall_incl_frame_synthesis": args.frames / t_wall,
            "model_ms_max_over_ranks": model_ms, "gather_ms": gather_ms,
            "gathered_bytes": emb.numel() * 4 + prob.numel() * 4 + pred.numel() * pred.element_size(),
            "checksums": {"emb_sum": checks[0].item(), "emb_abs_sum": checks[1].item(), "prob_sum": checks[2].item(),
                          "positives": int(checks[3].item())},
            "properties": f"{len(probe)} probed rows bit-identical to the frame embedded alone; prob in (0,1); pred == prob > 0.5; "
                          "checksums equal on all ranks",
        }
        s = json.dumps(line)
        print(s)
        if args.out:
            with open(args.out, "w") as f:
                f.write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

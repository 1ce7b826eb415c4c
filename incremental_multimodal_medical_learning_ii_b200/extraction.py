"""Sharded embedding extraction: the B200 version of the reference's ``chexpert-get-embedding.py`` hot loop.

The reference (chexpert-get-embedding.py:68-113) iterates a ``shuffle=False`` DataLoader with batch size 1, calls
``resnet50(images)``, grows a tensor with ``torch.cat`` and every 5000 samples saves
``TensorDataset(embeddings, labels)``.  Here frames are processed in large batches, one process per GPU owns a
CONTIGUOUS index range (so concatenating the ranks' results in rank order reproduces the reference's sequential
order), there is no communication during the forward, and embeddings / probabilities / labels are collected with ONE
``all_gather_into_tensor`` on padded buffers at the end (NCCL over NVLink on GPUs, gloo in the CPU tests).
Patch embeddings and heat-maps are never gathered (25.8 GB for 224k frames).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist

FrameSource = Callable[[int, int], torch.Tensor]        # (first index, count) -> frames [count, 1|3, H, W]
EmbedFn = Callable[[torch.Tensor], Dict[str, torch.Tensor]]

GATHERED_KEYS = ("global", "prob", "pred")


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[start, end)`` of ``range(n_items)`` for ``rank``; earlier ranks take the extra
    item when ``n_items % world_size != 0``.  Shards tile the range in rank order."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def extract_shard(embed_fn: EmbedFn, frame_source: FrameSource, n_frames: int, batch_size: int, rank: int = 0,
                  world_size: int = 1, keys=GATHERED_KEYS) -> Dict[str, torch.Tensor]:
    """Run ``embed_fn`` over this rank's shard in batches of ``batch_size`` (the last batch may be ragged) and return
    the concatenated per-frame results plus ``"range"`` = (start, end)."""
    start, end = shard_range(n_frames, rank, world_size)
    out: Dict[str, torch.Tensor] = {}
    pos = 0
    for first in range(start, end, batch_size):
        count = min(batch_size, end - first)
        res = embed_fn(frame_source(first, count))
        for k in keys:
            if k not in res:
                continue
            t = res[k]
            if k not in out:            # preallocate once: no O(N^2) torch.cat growth (chexpert-get-embedding.py:79)
                out[k] = torch.empty((end - start,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            out[k][pos:pos + count].copy_(t)
        pos += count
    out["range"] = torch.tensor([start, end], dtype=torch.int64)
    return out


def gather_shards(local: Dict[str, torch.Tensor], n_frames: int, rank: int = 0, world_size: int = 1,
                  keys=GATHERED_KEYS) -> Dict[str, torch.Tensor]:
    """All-gather the ranks' shard results into full ``[n_frames, ...]`` tensors in the reference's sequential order.

    Shards differ by at most one row, so every rank pads to the largest shard, ONE collective per key moves the padded
    blocks, and the padding rows are dropped using the shard sizes (which every rank can compute locally)."""
    if world_size == 1:
        return {k: local[k] for k in keys if k in local}
    sizes = [shard_range(n_frames, r, world_size) for r in range(world_size)]
    longest = max(e - s for s, e in sizes)
    full: Dict[str, torch.Tensor] = {}
    for k in keys:
        if k not in local:
            continue
        t = local[k]
        padded = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        padded[: t.shape[0]].copy_(t)
        gathered = torch.empty((world_size * longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(gathered, padded)
        blocks = gathered.view((world_size, longest) + tuple(t.shape[1:]))
        full[k] = torch.cat([blocks[r, : e - s] for r, (s, e) in enumerate(sizes)], dim=0)
    return full


def save_embedding_chunks(embeddings: torch.Tensor, labels: torch.Tensor, out_dir: str, chunk: int = 5000,
                          prefix: str = "embeddings_dataset") -> list:
    """Write the embedding store the reference's Trainer reads: ``TensorDataset(emb[N,128], labels[N,L])`` files named
    ``{prefix}_{k}.pt`` every ``chunk`` samples plus ``{prefix}_final.pt`` for the remainder
    (chexpert-get-embedding.py:86-113; glued later by CSV_reformatting/glue_dataset.py:33-38)."""
    from torch.utils.data import TensorDataset
    os.makedirs(out_dir, exist_ok=True)
    embeddings = embeddings.detach().to("cpu", torch.float32)
    labels = labels.detach().to("cpu")
    paths = []
    n = embeddings.shape[0]
    k = 0
    for first in range(0, n - n % chunk, chunk):
        k += 1
        path = os.path.join(out_dir, f"{prefix}_{k * chunk}.pt")
        torch.save(TensorDataset(embeddings[first:first + chunk].clone(), labels[first:first + chunk].clone()), path)
        paths.append(path)
    if n % chunk:
        path = os.path.join(out_dir, f"{prefix}_final.pt")
        torch.save(TensorDataset(embeddings[n - n % chunk:].clone(), labels[n - n % chunk:].clone()), path)
        paths.append(path)
    return paths


def extract_embeddings(model, frame_source: FrameSource, n_frames: int, batch_size: int = 512,
                       rank: Optional[int] = None, world_size: Optional[int] = None, score: bool = True
                       ) -> Dict[str, torch.Tensor]:
    """Whole pipeline on the B200 model: shard -> batched forward (+ fused zero-shot scoring) -> one all-gather."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world_size > 1 else 0
    if score:
        embed_fn = model.embed_and_score
        keys = GATHERED_KEYS
    else:
        def embed_fn(frames):
            return {"global": model(frames).projected_global_embedding}
        keys = ("global",)
    local = extract_shard(embed_fn, frame_source, n_frames, batch_size, rank, world_size, keys)
    return gather_shards(local, n_frames, rank, world_size, keys)


def extract_to_store(model, raw_frame_source: Callable[[int, int], torch.Tensor],
                     label_source: Callable[[int, int], torch.Tensor], n_frames: int, out_dir: str,
                     resize: int = 512, crop: int = 480, batch_size: int = 512, chunk: int = 5000,
                     rank: int = 0, world_size: int = 1) -> list:
    """The whole ``chexpert-get-embedding.py`` job for this rank's shard, on the device end to end:
    raw 8-bit frames -> GPU Resize/CenterCrop (PIL-exact) -> ``ImageModel`` -> un-normalised ``[n,128]`` embeddings ->
    reference-format chunk files written by a background thread (``embedding_store.AsyncChunkWriter``).

    ``raw_frame_source(first, count)`` returns decoded frames ``[count,h,w]`` uint8 (same size within a call) on the
    model's device or on the host; ``label_source(first, count)`` the ``[count,5]`` labels.  Each rank writes its
    shard under ``out_dir/rank{r}`` (one directory when ``world_size == 1``), so ranks never contend for a file and
    gluing the rank directories in rank order reproduces the reference's sequential order."""
    from .embedding_store import AsyncChunkWriter
    from .image.data.gpu_transforms import GpuResizeCenterCrop
    device = next(model.parameters()).device
    transform = GpuResizeCenterCrop(resize, crop)
    start, end = shard_range(n_frames, rank, world_size)
    directory = out_dir if world_size == 1 else os.path.join(out_dir, f"rank{rank}")
    with AsyncChunkWriter(directory, chunk=chunk) as writer:
        for first in range(start, end, batch_size):
            count = min(batch_size, end - first)
            raw = raw_frame_source(first, count).to(device, non_blocking=True)
            emb = model(transform(raw)).projected_global_embedding
            writer.add(emb, label_source(first, count))
    return writer.paths

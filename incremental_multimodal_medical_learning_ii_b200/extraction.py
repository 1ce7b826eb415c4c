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
    the concatenated per-frame results plus ``"range"`` = (start, end).  A rank whose shard is empty
    (``n_frames < world_size``) returns only ``"range"``; :func:`gather_shards` still makes it take part in the
    collective."""
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


def pack_rows(tensors) -> torch.Tensor:
    """``[n, ...]`` tensors of any dtype -> ONE ``[n, row_bytes]`` uint8 tensor (each row: the rows of the inputs, byte for
    byte, one after the other).  This is the staging buffer of the single all-gather."""
    n = tensors[0].shape[0]
    width = lambda t: int(t[0:1].numel()) if n else int(torch.Size(t.shape[1:]).numel())   # noqa: E731  (elements per row)
    return torch.cat([t.contiguous().view(n, width(t)).view(torch.uint8) for t in tensors], dim=1)


def unpack_rows(packed: torch.Tensor, specs) -> list:
    """Inverse of :func:`pack_rows`; ``specs`` = [(trailing shape, dtype), ...] in the packing order."""
    n, out, col = packed.shape[0], [], 0
    for shape, dtype in specs:
        width = int(torch.empty(0, dtype=dtype).element_size())
        for d in shape:
            width *= d
        out.append(packed[:, col:col + width].reshape(-1).clone().view(dtype).view((n,) + tuple(shape)))
        col += width
    return out


def gather_shards(local: Dict[str, torch.Tensor], n_frames: int, rank: int = 0, world_size: int = 1,
                  keys=GATHERED_KEYS, specs: Optional[Dict[str, tuple]] = None, device=None) -> Dict[str, torch.Tensor]:
    """All-gather the ranks' shard results into full ``[n_frames, ...]`` tensors in the reference's sequential order
    with ONE collective: embeddings, probabilities and labels of a frame are packed into one byte row (582 B for 14
    labels), every rank pads its block to the largest shard (shards differ by at most one row), one
    ``all_gather_into_tensor`` moves the blocks, and the padding rows are dropped using the shard sizes, which every
    rank computes locally.

    ``specs`` ({key: (trailing shape, dtype)}) and ``device`` are only needed by a rank whose shard is EMPTY
    (``n_frames < world_size``): it has no tensors to read them from but must enter the collective like every other
    rank; when omitted they are exchanged with a small ``all_gather_object`` first."""
    present = [k for k in keys if k in local]
    if world_size == 1:
        return {k: local[k] for k in present}
    sizes = [shard_range(n_frames, r, world_size) for r in range(world_size)]
    longest = max(e - s for s, e in sizes)
    if n_frames < world_size and specs is None:
        mine = {k: (tuple(local[k].shape[1:]), local[k].dtype) for k in present}
        everyone = [None] * world_size
        dist.all_gather_object(everyone, mine)
        specs = next(m for m in everyone if m)
    if specs is None:
        specs = {k: (tuple(local[k].shape[1:]), local[k].dtype) for k in present}
    order = [k for k in keys if k in specs]
    if present:
        device = local[present[0]].device
        block = pack_rows([local[k] for k in order])
    else:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        block = pack_rows([torch.empty((0,) + tuple(specs[k][0]), dtype=specs[k][1], device=device) for k in order])
    padded = torch.zeros((longest, block.shape[1]), dtype=torch.uint8, device=device)
    padded[: block.shape[0]].copy_(block)
    gathered = torch.empty((world_size * longest, block.shape[1]), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, padded)
    blocks = gathered.view(world_size, longest, block.shape[1])
    rows = torch.cat([blocks[r, : e - s] for r, (s, e) in enumerate(sizes)], dim=0)
    return dict(zip(order, unpack_rows(rows, [specs[k] for k in order])))


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
                     resize: int = 512, crop: int = 512, batch_size: int = 512, chunk: int = 5000,
                     rank: int = 0, world_size: int = 1) -> list:
    """The whole ``chexpert-get-embedding.py`` job for this rank's shard, on the device end to end:
    raw 8-bit frames -> GPU Resize/CenterCrop (PIL-exact) -> ``ImageModel`` -> un-normalised ``[n,128]`` embeddings ->
    reference-format chunk files written by a background thread (``embedding_store.AsyncChunkWriter``).

    The defaults are the extraction script's transform, ``DataRetrieval(size=512)`` = ``Resize(512)`` then
    ``CenterCrop(512)`` (DataRetrieval.py:175-178; 16x16 patch grid); ``resize=512, crop=480`` is the OTHER transform of the
    code base, ``ImageInferenceEngine``'s (image/utils.py:11-12; 15x15 grid, the benchmark's frame size).

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

"""Embedding store: the on-disk format the reference's trainers read, written without stalling the GPUs.

Reference format:

* ``chexpert-get-embedding.py:86-113`` saves ``TensorDataset(emb[n,128] fp32, labels[n,5] fp32)`` to
  ``embeddings_dataset_{k}.pt`` every ``checkpoint_interval`` = 5000 samples (batch size 1) and the remainder to
  ``embeddings_dataset_final.pt``;
* ``CSV_reformatting/glue_dataset.py:33-38`` glues the numbered chunks into one ``ConcatDataset`` and saves it as
  ``embeddings_dataset_final_old.pt`` - the file ``Trainer.py:221-235`` loads with ``torch.load`` and feeds to a
  ``DataLoader``; ``Trainer.concat_to_tensor_dataloader`` (``Trainer.py:1252-1271``) walks
  ``loader.dataset.datasets[i].tensors``.

``AsyncChunkWriter`` takes finished ``[n,128]`` blocks (device or host tensors) and writes chunk files on a background
thread: the device->host copy goes through a pinned staging buffer on a side stream, so neither the copy nor
``torch.save`` sits between two batches of the extraction loop.  ``load_embedding_chunks`` / ``glue_embedding_chunks``
are the read side and reproduce ``glue_dataset.py`` (numeric chunk order, then the final remainder).
"""
from __future__ import annotations

import os
import queue
import re
import threading
from typing import List, Optional, Sequence

import torch
from torch.utils.data import ConcatDataset, TensorDataset

PREFIX = "embeddings_dataset"
_CHUNK_RE = re.compile(r"_(\d+)\.pt$")


def chunk_paths(directory: str, prefix: str = PREFIX, include_final: bool = True) -> List[str]:
    """Chunk files of a store in the order the samples were produced: numbered chunks ascending, then ``_final``."""
    numbered = []
    for name in os.listdir(directory):
        if not name.startswith(prefix + "_"):
            continue
        m = _CHUNK_RE.search(name)
        if m:
            numbered.append((int(m.group(1)), os.path.join(directory, name)))
    paths = [p for _, p in sorted(numbered)]
    final = os.path.join(directory, f"{prefix}_final.pt")
    if include_final and os.path.exists(final):
        paths.append(final)
    return paths


def load_embedding_chunks(directory: str, prefix: str = PREFIX, include_final: bool = True) -> ConcatDataset:
    """``ConcatDataset`` over the store's ``TensorDataset`` chunks (what ``glue_dataset.py:36`` builds; the reference
    script itself skips the ``_final`` remainder - pass ``include_final=False`` for that exact behaviour)."""
    paths = chunk_paths(directory, prefix, include_final)
    if not paths:
        raise FileNotFoundError(f"no {prefix}_*.pt chunks in {directory}")
    return ConcatDataset([torch.load(p, weights_only=False) for p in paths])


def glue_embedding_chunks(directory: str, out_name: str = "embeddings_dataset_final_old.pt", prefix: str = PREFIX,
                          include_final: bool = True) -> str:
    """Write the glued ``ConcatDataset`` file ``Trainer.py:221-235`` loads (``glue_dataset.py:38``)."""
    out = os.path.join(directory, out_name)
    torch.save(load_embedding_chunks(directory, prefix, include_final), out)
    return out


def concat_to_tensors(dataset) -> TensorDataset:
    """``Trainer.concat_to_tensor_dataloader``'s flattening (``Trainer.py:1252-1271``) for a loaded store."""
    parts: Sequence[TensorDataset] = dataset.datasets if isinstance(dataset, ConcatDataset) else [dataset]
    return TensorDataset(torch.cat([d.tensors[0] for d in parts]), torch.cat([d.tensors[1] for d in parts]))


class AsyncChunkWriter:
    """Accumulates ``(embeddings, labels)`` blocks and writes reference-format chunk files on a background thread.

    ``add`` never blocks on disk: a block is copied into a host buffer (asynchronously on a side stream when it lives
    on a GPU) and handed to the writer thread; files appear as ``{prefix}_{k*chunk}.pt`` and, on ``close``, the
    remainder as ``{prefix}_final.pt`` - identical names and contents to the reference loop with batch size 1.
    """

    def __init__(self, directory: str, chunk: int = 5000, prefix: str = PREFIX, emb_dim: int = 128,
                 num_labels: int = 5, max_pending: int = 8):
        os.makedirs(directory, exist_ok=True)
        self.directory, self.chunk, self.prefix = directory, int(chunk), prefix
        self.emb_dim, self.num_labels = emb_dim, num_labels
        self._emb = torch.empty(self.chunk, emb_dim, dtype=torch.float32)
        self._lab = torch.empty(self.chunk, num_labels, dtype=torch.float32)
        self._fill = 0
        self._written = 0
        self.paths: List[str] = []
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self._err: Optional[BaseException] = None
        self._stream = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    # ---- producer side -------------------------------------------------------------------------------------
    def add(self, embeddings: torch.Tensor, labels: torch.Tensor) -> None:
        if self._err is not None:
            raise RuntimeError("embedding store writer failed") from self._err
        if embeddings.shape[0] != labels.shape[0]:
            raise ValueError("embeddings and labels differ in length")
        if embeddings.shape[1:] != (self.emb_dim,) or labels.shape[1:] != (self.num_labels,):
            raise ValueError(f"expected [n,{self.emb_dim}] embeddings and [n,{self.num_labels}] labels")
        if embeddings.is_cuda:
            # device -> pinned host on a side stream that waits for the producer's stream; the event travels with
            # the block so that the writer thread (not the caller) waits for the copy
            if self._stream is None:
                self._stream = torch.cuda.Stream(embeddings.device)
            self._stream.wait_stream(torch.cuda.current_stream(embeddings.device))
            with torch.cuda.stream(self._stream):
                e = torch.empty(embeddings.shape, dtype=torch.float32, pin_memory=True)
                l = torch.empty(labels.shape, dtype=torch.float32, pin_memory=True)
                e.copy_(embeddings.detach().float(), non_blocking=True)
                l.copy_(labels.detach().float(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            embeddings.record_stream(self._stream)
            if labels.is_cuda:
                labels.record_stream(self._stream)
            self._q.put(("block", e, l, ev))
        else:
            self._q.put(("block", embeddings.detach().float().clone(), labels.detach().float().clone(), None))

    def close(self) -> List[str]:
        self._q.put(("close",))
        self._thread.join()
        if self._err is not None:
            raise RuntimeError("embedding store writer failed") from self._err
        return self.paths

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        else:
            self._q.put(("close",))
            self._thread.join()

    # ---- writer thread -------------------------------------------------------------------------------------
    def _flush(self, final: bool) -> None:
        if self._fill == 0:
            return
        if final:
            path = os.path.join(self.directory, f"{self.prefix}_final.pt")
        else:
            path = os.path.join(self.directory, f"{self.prefix}_{self._written + self._fill}.pt")
        torch.save(TensorDataset(self._emb[: self._fill].clone(), self._lab[: self._fill].clone()), path)
        self.paths.append(path)
        self._written += self._fill
        self._fill = 0

    def _run(self) -> None:
        try:
            while True:
                item = self._q.get()
                if item[0] == "close":
                    self._flush(final=True)
                    return
                _, e, l, ev = item
                if ev is not None:
                    ev.synchronize()
                pos = 0
                while pos < e.shape[0]:
                    n = min(self.chunk - self._fill, e.shape[0] - pos)
                    self._emb[self._fill:self._fill + n].copy_(e[pos:pos + n])
                    self._lab[self._fill:self._fill + n].copy_(l[pos:pos + n])
                    self._fill += n
                    pos += n
                    if self._fill == self.chunk:
                        self._flush(final=False)
        except BaseException as ex:  # surfaced to the producer on the next add()/close()
            self._err = ex
            while True:               # keep draining so that a blocked producer is released
                try:
                    if self._q.get(timeout=0.1)[0] == "close":
                        return
                except queue.Empty:
                    return

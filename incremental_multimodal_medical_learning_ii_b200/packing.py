"""Turn a BioViL ``state_dict`` (the reference's 328 keys) into the device-side weight table the kernels read.

Eval-mode BatchNorm is folded into the preceding convolution in float64
(``w' = w * gamma / sqrt(var + eps)``, ``b' = beta - mean * gamma / sqrt(var + eps)``), weights are re-laid
out ``OIHW -> O(HW)I`` (K-major for the implicit GEMM: k = (r*S + s)*Cin + c) and rounded to bf16 once;
biases stay fp32.  The stem gets three variants (see ``bv_weights`` in ``include/biovil_b200.h``): the three
identical ``ExpandChannels`` input channels (``health_multimodal/image/data/transforms.py:12-25``) are summed
into one, and for 8-bit frames ``ToTensor``'s 1/255 is folded in so pixel integers stay exact in bf16.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import torch

from . import _native as N

BN_EPS = 1e-5                      # nn.BatchNorm2d default used by torchvision ResNet and the projector MLP
LAYER_PLAN = (3, 4, 6, 3)          # health_multimodal/image/model/resnet.py:80
ENC = "encoder.encoder."
PROJ = "projector.model."


def fold_bn(w: torch.Tensor, sd: Dict[str, torch.Tensor], bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(folded weight, folded bias) in float64 for conv weight ``w`` followed by eval-mode BatchNorm ``bn``."""
    g = sd[bn + ".weight"].double()
    b = sd[bn + ".bias"].double()
    m = sd[bn + ".running_mean"].double()
    v = sd[bn + ".running_var"].double()
    scale = g / torch.sqrt(v + BN_EPS)
    return w.double() * scale.view(-1, 1, 1, 1), b - m * scale


def block_names() -> List[Tuple[str, int]]:
    """``(prefix, stride)`` of the 16 Bottleneck blocks in execution order (stride sits on conv2: ResNet v1.5)."""
    out = []
    for li, n in enumerate(LAYER_PLAN, start=1):
        for bi in range(n):
            out.append((f"{ENC}layer{li}.{bi}", 2 if (bi == 0 and li > 1) else 1))
    return out


class PackedWeights:
    """Owns the packed device tensors and the ``bv_weights`` struct that points into them."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        self.device = device
        self._keep: List[torch.Tensor] = []
        self.struct = N.BvWeights()
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}

        # ---- stem -------------------------------------------------------------------------------------------
        w, b = fold_bn(sd[ENC + "conv1.weight"], sd, ENC + "bn1")          # [64,3,7,7]
        w3 = w.reshape(64, 147)
        w1 = w.sum(dim=1).reshape(64, 49)
        self.struct.stem_f3 = self._matrix(torch.nn.functional.pad(w3, (0, 192 - 147)), b, cin=192)
        self.struct.stem_f1 = self._matrix(torch.nn.functional.pad(w1, (0, 64 - 49)), b, cin=64)
        self.struct.stem_u8 = self._matrix(torch.nn.functional.pad(w1 / 255.0, (0, 64 - 49)), b, cin=64)
        # fused stem kernel: k = r*8 + s, window column s = 7 carries zero weights; row r = 7 (k = 56..58) carries the
        # BatchNorm bias split into three bf16 terms (hi + mid + lo reproduces the fp32 bias to 24 bits): the kernel
        # feeds 1.0 in those three A columns, so the accumulator already holds conv + bias
        w8 = torch.zeros(64, 8, 8, dtype=w.dtype)
        w8[:, :7, :7] = w.sum(dim=1) / 255.0
        b64 = b.double()
        b_hi = b64.float().to(torch.bfloat16)
        b_mid = (b64 - b_hi.double()).float().to(torch.bfloat16)
        b_lo = (b64 - b_hi.double() - b_mid.double()).float().to(torch.bfloat16)
        w8[:, 7, 0], w8[:, 7, 1], w8[:, 7, 2] = b_hi.to(w.dtype), b_mid.to(w.dtype), b_lo.to(w.dtype)
        self.struct.stem_u8_k8 = self._matrix(w8.reshape(64, 64), b, cin=64)

        # ---- bottlenecks ------------------------------------------------------------------------------------
        for i, (p, stride) in enumerate(block_names()):
            self.struct.conv1[i] = self._conv(sd, p + ".conv1", p + ".bn1", 1, 0)
            self.struct.conv2[i] = self._conv(sd, p + ".conv2", p + ".bn2", stride, 1)
            self.struct.conv3[i] = self._conv(sd, p + ".conv3", p + ".bn3", 1, 0)
            if (p + ".downsample.0.weight") in sd:
                self.struct.downsample[i] = self._conv(sd, p + ".downsample.0", p + ".downsample.1", stride, 0)

        # ---- projector (modules.py:43-47) -------------------------------------------------------------------
        self.struct.proj0 = self._conv(sd, PROJ + "0", PROJ + "1", 1, 0)
        w2 = sd[PROJ + "3.weight"].reshape(128, 128).float()                 # [d][k]
        self.proj3_wt = self._dev(w2.t().contiguous())                        # [k][d]
        self.proj3_b = self._dev(sd[PROJ + "3.bias"].float().contiguous())
        self.struct.proj3_wt = self.proj3_wt.data_ptr()
        self.struct.proj3_b = self.proj3_b.data_ptr()

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        t = t.to(self.device)
        self._keep.append(t)
        return t

    def _matrix(self, w2d: torch.Tensor, bias: torch.Tensor, cin: int) -> N.BvConv:
        wt = self._dev(w2d.to(torch.float32).to(torch.bfloat16).contiguous())
        bt = self._dev(bias.to(torch.float32).contiguous())
        return N.BvConv(wt.data_ptr(), bt.data_ptr(), cin, w2d.shape[0], 1, 1, 1, 0)

    def _conv(self, sd, conv: str, bn: str, stride: int, pad: int) -> N.BvConv:
        w, b = fold_bn(sd[conv + ".weight"], sd, bn)
        cout, cin, r, s = w.shape
        wt = self._dev(w.permute(0, 2, 3, 1).contiguous().to(torch.float32).to(torch.bfloat16))
        bt = self._dev(b.to(torch.float32).contiguous())
        return N.BvConv(wt.data_ptr(), bt.data_ptr(), cin, cout, r, s, stride, pad)

    def pointer(self):
        return ctypes.byref(self.struct)


def pack_single_conv(w_oihw: torch.Tensor, bias: torch.Tensor, stride: int, pad: int, device):
    """bf16 K-major weight + fp32 bias + ``bv_conv`` for one (already folded) convolution; used by the kernel tests."""
    cout, cin, r, s = w_oihw.shape
    wt = w_oihw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(device)
    bt = bias.to(torch.float32).contiguous().to(device)
    return N.BvConv(wt.data_ptr(), bt.data_ptr(), cin, cout, r, s, stride, pad), (wt, bt)

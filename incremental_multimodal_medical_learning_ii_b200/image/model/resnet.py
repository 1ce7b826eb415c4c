"""Parameter container for the BioViL ResNet-50 trunk.

Mirrors the module *names* of torchvision's ``ResNet``/``Bottleneck`` as used by ``ResNetHIML``
(``health_multimodal/image/model/resnet.py:15-47,73-80`` in the reference) so that ``state_dict()`` is
key-compatible with ``biovil_image_resnet50_proj_size_128.pt`` (``encoder.encoder.conv1.weight`` ...
``encoder.encoder.fc.bias``).  It only *holds* parameters: the arithmetic is done by the sm_100a kernels
(``csrc/``) on a folded, packed copy, so none of these modules has a usable ``forward`` - there is no eager
fallback.  Unlike the reference (``resnet.py:57-59``) nothing is downloaded.
"""
from __future__ import annotations

import torch
from torch import nn

LAYER_PLAN = (3, 4, 6, 3)        # resnet50(): Bottleneck, [3, 4, 6, 3]  (reference resnet.py:80)
LAYER_WIDTH = (64, 128, 256, 512)
EXPANSION = 4


class _NoEager(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} only holds parameters; run the enclosing ImageModel "
                           "(sm_100a kernels) - there is no eager/CPU fallback")


class Bottleneck(_NoEager):
    """conv1x1 -> bn -> relu -> conv3x3(stride) -> bn -> relu -> conv1x1 -> bn -> (+identity | downsample) -> relu."""

    def __init__(self, inplanes: int, planes: int, stride: int, downsample: bool):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * EXPANSION, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * EXPANSION)
        self.stride = stride
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * EXPANSION, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes * EXPANSION))
        else:
            self.downsample = None


class ResNetHIML(_NoEager):
    """ResNet-50 trunk parameters; the (unused) ``fc`` head is kept because the checkpoint carries it."""

    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for li, (n, width) in enumerate(zip(LAYER_PLAN, LAYER_WIDTH), start=1):
            blocks = []
            for bi in range(n):
                stride = 2 if (bi == 0 and li > 1) else 1
                blocks.append(Bottleneck(inplanes, width, stride, downsample=(bi == 0)))
                inplanes = width * EXPANSION
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.fc = nn.Linear(512 * EXPANSION, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


def resnet50(**kwargs) -> ResNetHIML:
    kwargs.pop("pretrained", None)      # the reference hard-wires pretrained=True (model.py:194); nothing to fetch here
    kwargs.pop("progress", None)
    return ResNetHIML(**kwargs)

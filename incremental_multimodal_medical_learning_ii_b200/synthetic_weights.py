"""Deterministic random-init BioViL ``state_dict`` (synthetic stand-in for ``biovil_image_resnet50_proj_size_128.pt``).

The real BioViL checkpoint cannot be downloaded (no network), so the benchmark, the tools and the parity fixtures use
seeded random weights with the distributions the reference's constructors use: torchvision ResNet convs
``kaiming_normal_(fan_out, relu)``, BatchNorm identity (gamma 1, beta 0, mean 0, var 1), projector convs / fc
PyTorch-default ``kaiming_uniform_(a=sqrt(5))``.  Built from an explicit ``torch.Generator`` on CPU so that every
machine regenerates the same tensors.  Keys, shapes and dtypes equal the reference's 328-key state_dict
(``health_multimodal/image/model/model.py:100-118``; checked against the reference itself in ``oracle/make_golden.py``).

This is input DATA (like ``frames.py``), not arithmetic of the path: it lives in the package so that nothing on the
product side imports ``oracle/``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

LAYER_PLAN = (3, 4, 6, 3)       # health_multimodal/image/model/resnet.py:80  (Bottleneck, [3, 4, 6, 3])
LAYER_WIDTH = (64, 128, 256, 512)
EXPANSION = 4


def _kaiming_normal(g, cout, cin, k):
    std = math.sqrt(2.0 / (cout * k * k))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _uniform(g, shape, bound):
    return (torch.rand(*shape, generator=g) * 2.0 - 1.0) * bound


def _bn(sd, name, n):
    sd[name + ".weight"] = torch.ones(n)
    sd[name + ".bias"] = torch.zeros(n)
    sd[name + ".running_mean"] = torch.zeros(n)
    sd[name + ".running_var"] = torch.ones(n)
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def randomize_batchnorm_(sd: Dict[str, torch.Tensor], seed: int = 28) -> Dict[str, torch.Tensor]:
    """BN-randomised variant: a default-init BN is an identity, so a wrong BN fold would
    pass parity.  gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5), generator seed 28."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        if k.endswith(".running_mean"):
            base = k[: -len(".running_mean")]
            n = sd[k].numel()
            sd[base + ".weight"] = torch.rand(n, generator=g) + 0.5
            sd[base + ".bias"] = torch.randn(n, generator=g) * 0.1
            sd[base + ".running_mean"] = torch.randn(n, generator=g) * 0.1
            sd[base + ".running_var"] = torch.rand(n, generator=g) + 0.5
    return sd


def make_state_dict(seed: int = 27, randomize_bn: bool = False, bn_seed: int = 28) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    e = "encoder.encoder."
    sd[e + "conv1.weight"] = _kaiming_normal(g, 64, 3, 7)
    _bn(sd, e + "bn1", 64)
    inplanes = 64
    for li, (n, width) in enumerate(zip(LAYER_PLAN, LAYER_WIDTH), start=1):
        for bi in range(n):
            p = f"{e}layer{li}.{bi}"
            sd[p + ".conv1.weight"] = _kaiming_normal(g, width, inplanes, 1)
            _bn(sd, p + ".bn1", width)
            sd[p + ".conv2.weight"] = _kaiming_normal(g, width, width, 3)
            _bn(sd, p + ".bn2", width)
            sd[p + ".conv3.weight"] = _kaiming_normal(g, width * EXPANSION, width, 1)
            _bn(sd, p + ".bn3", width * EXPANSION)
            if bi == 0:
                sd[p + ".downsample.0.weight"] = _kaiming_normal(g, width * EXPANSION, inplanes, 1)
                _bn(sd, p + ".downsample.1", width * EXPANSION)
            inplanes = width * EXPANSION
    sd[e + "fc.weight"] = _uniform(g, (1000, 2048), 1.0 / math.sqrt(2048))
    sd[e + "fc.bias"] = _uniform(g, (1000,), 1.0 / math.sqrt(2048))
    pm = "projector.model."
    sd[pm + "0.weight"] = _uniform(g, (128, 2048, 1, 1), 1.0 / math.sqrt(2048))
    _bn(sd, pm + "1", 128)
    sd[pm + "3.weight"] = _uniform(g, (128, 128, 1, 1), 1.0 / math.sqrt(128))
    sd[pm + "3.bias"] = _uniform(g, (128,), 1.0 / math.sqrt(128))
    if randomize_bn:
        randomize_batchnorm_(sd, seed=bn_seed)
    return sd


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint (sum of |w| in float64 over floating tensors)."""
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))

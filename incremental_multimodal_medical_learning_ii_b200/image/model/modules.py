"""Projector parameters: ``MLP`` with ``use_1x1_convs=True`` (reference ``health_multimodal/image/model/modules.py:12-55``):
``Conv2d(in,hidden,1,bias=False) -> BatchNorm2d -> ReLU -> Conv2d(hidden,out,1,bias=True)`` under ``.model`` indices
0, 1, 2, 3, so the state_dict keys are ``projector.model.{0,1,3}.*``.  Parameters only; the sm_100a kernels compute."""
from __future__ import annotations

from typing import Optional

from torch import nn


class MLP(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, hidden_dim: Optional[int] = None,
                 use_1x1_convs: bool = True) -> None:
        super().__init__()
        if not use_1x1_convs:
            raise NotImplementedError("the BioViL projector uses 1x1 convolutions (use_1x1_convs=True)")
        if hidden_dim is None:
            raise NotImplementedError("the BioViL projector has a hidden layer (hidden_dim=joint_feature_size)")
        self.output_dim = output_dim
        self.input_dim = input_dim
        self.model = nn.Sequential(
            nn.Conv2d(input_dim, hidden_dim, kernel_size=1, bias=False),
            nn.BatchNorm2d(hidden_dim),
            nn.ReLU(inplace=True),
            nn.Conv2d(hidden_dim, output_dim, kernel_size=1, bias=True))

    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError("MLP only holds parameters; run the enclosing ImageModel (no eager/CPU fallback)")

"""Projection heads -- same names, constructor arguments, parameter names (state-dict keys) and forward semantics as
the reference's ``mmgclip/networks/projection.py``, computing through the sm_100a kernels.

The ``nn.Linear`` / ``nn.LayerNorm`` sub-modules are kept purely as parameter containers (identical default
initialisation and ``state_dict`` layout, so a reference ``model.pth`` loads unchanged); their ``forward`` is never
called -- every contraction goes through :func:`mmgclip_b200.ops.linear`.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class LinearProjectionLayer(nn.Module):
    """One bias-free ``Linear(embedding_dim, projection_dim)`` (reference: projection.py:4-33).

    ``dropout`` is accepted and ignored, exactly like the reference (projection.py:15).  State-dict key: ``layer.weight``.
    """

    def __init__(self, embedding_dim, projection_dim=512, dropout=0, precision=None):
        super().__init__()
        self.layer = nn.Linear(embedding_dim, projection_dim, bias=False)
        self.precision = precision
        for p in self.layer.parameters():
            p.requires_grad = True

    def forward(self, x):
        return ops.linear(x, self.layer.weight, None, prec=self.precision)

    def forward_normalized(self, x):
        """Projection + row L2-normalisation as one fused operator (what MMGCLIP.forward needs, mmgclip_model.py:124-129)."""
        return ops.project_normalize(x, self.layer.weight, prec=self.precision)


class MultiLinearHead(nn.Module):
    """``Linear+bias -> ReLU -> Dropout`` for every layer but the last, then ``Linear+bias`` (projection.py:36-61).

    ``projection_dim`` must be a list of widths, e.g. ``[768, 512]`` (configs/projection/2xLinear512.yaml:5); an int raises
    ``TypeError`` as in the reference (projection.py:45).  State-dict keys: ``layers.{i}.weight`` / ``layers.{i}.bias``.
    """

    def __init__(self, embedding_dim, projection_dim=[], dropout=0.5, precision=None):  # noqa: B006 (reference signature)
        super().__init__()
        self.embedding_dim = embedding_dim
        self.projection_dim = projection_dim
        self.layers = nn.ModuleList()
        self.layers.append(nn.Linear(embedding_dim, projection_dim[0]))
        for i in range(len(projection_dim) - 1):
            self.layers.append(nn.Linear(projection_dim[i], projection_dim[i + 1]))
        self.dropout = nn.Dropout(dropout)  # holds p; the mask is applied inside the fused operator
        self.relu = nn.ReLU()
        self.precision = precision

    def forward(self, x):
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            if i < last:
                p = self.dropout.p if self.training else 0.0
                # bias + ReLU live in the contraction's epilogue; the keep mask is drawn (Philox), applied and recorded by
                # one launch on its output (the reference: nn.Dropout, projection.py:51,59)
                x = ops.linear(x, layer.weight, layer.bias, relu=True, prec=self.precision, drop_p=p)
            else:
                x = ops.linear(x, layer.weight, layer.bias, prec=self.precision)
        return x


class MLPProjectionHead(nn.Module):
    """``p = Linear(x); LayerNorm(Dropout(Linear(GELU(p))) + p)`` (projection.py:85-101).

    State-dict keys: ``projection.*``, ``fc.*``, ``layer_norm.*``.  GELU is the exact-erf form and LayerNorm uses eps=1e-5
    (torch defaults, as in the reference).
    """

    def __init__(self, embedding_dim, projection_dim, dropout=0.5, precision=None):
        super().__init__()
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)
        self.precision = precision

    def forward(self, x):
        projected = ops.linear(x, self.projection.weight, self.projection.bias, prec=self.precision)
        h = ops.gelu(projected)
        p = self.dropout.p if self.training else 0.0
        h = ops.linear(h, self.fc.weight, self.fc.bias, prec=self.precision, drop_p=p)
        h = ops.residual_add(h, projected)
        return ops.layer_norm(h, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)

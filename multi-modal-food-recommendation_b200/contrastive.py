"""Contrastive terms of CLUSSL on gathered `[2B, d]` views.

`correlation_distance` is the live term (FoodRec/models/pricai_modelx.py:263,409-437), `info_nce`
the dormant `CL_loss` (:354-378).  Both are launch/latency-bound at 2B = 1024.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _centred_distance(X: torch.Tensor) -> torch.Tensor:
    r = (X * X).sum(1, keepdim=True)
    D = torch.sqrt(torch.clamp_min(r - 2 * (X @ X.t()) + r.t(), 0.0) + 1e-8)
    return D - D.mean(0, keepdim=True) - D.mean(1, keepdim=True) + D.mean()


def _dcov(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    n = float(A.shape[0])
    return torch.sqrt(torch.clamp_min((A * B).sum() / (n * n), 0.0) + 1e-8)


def correlation_distance(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    D1, D2 = _centred_distance(x), _centred_distance(y)
    d12, d11, d22 = _dcov(D1, D2), _dcov(D1, D1), _dcov(D2, D2)
    return (d12 / torch.sqrt(torch.clamp_min(d11 * d22, 0.0) + 1e-10)).reshape(1)


def info_nce(hidden: torch.Tensor, temperature: float = 0.5, hidden_norm: bool = True) -> torch.Tensor:
    b = hidden.shape[0] // 2
    if hidden_norm:
        hidden = F.normalize(hidden, p=2, dim=-1)
    h1, h2 = hidden[:b], hidden[b:2 * b]
    big = torch.eye(b, device=hidden.device, dtype=hidden.dtype) * 1e9
    aa = h1 @ h1.t() / temperature - big
    bb = h2 @ h2.t() / temperature - big
    ab = h1 @ h2.t() / temperature
    ba = h2 @ h1.t() / temperature
    lab = torch.arange(b, device=hidden.device)
    la = F.cross_entropy(torch.cat([ab, aa], 1), lab)
    lb = F.cross_entropy(torch.cat([ba, bb], 1), lab)
    return (la + lb) / b

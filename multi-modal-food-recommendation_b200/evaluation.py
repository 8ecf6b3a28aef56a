"""Full-sort evaluation: user x item scores, optional history mask, top-K, metrics.

Replaces `Trainer.evaluate`'s per-user `full_sort_predict` + `torch.topk`
(FoodRec/common/trainer.py:476-503) for the dot-product models.
"""
from __future__ import annotations

import numpy as np
import torch


def full_sort_scores(user_all: torch.Tensor, item_all: torch.Tensor, users: torch.Tensor) -> torch.Tensor:
    """Dense fp32 `[n_batch_users, n_items]` scores (the matrix the fused top-K path avoids)."""
    return user_all[users.reshape(-1).long()] @ item_all.t()

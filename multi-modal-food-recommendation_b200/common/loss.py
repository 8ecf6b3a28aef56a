"""Parameter-free loss modules kept so that `str(model)` and attribute names match the reference
(`self.mf_loss`, `self.reg_loss`; FoodRec/common/loss.py:8-50).  The arithmetic runs in the fused
kernel behind `ops.rank_loss`; calling these modules directly routes to small fused kernels too."""
import torch
import torch.nn as nn


class BPRLoss(nn.Module):
    def __init__(self, gamma=1e-10):
        super().__init__()
        self.gamma = gamma


class EmbLoss(nn.Module):
    def __init__(self, norm=2):
        super().__init__()
        self.norm = norm

"""`BPRLoss` / `EmbLoss` with the reference's names and call signatures (FoodRec/common/loss.py:8-50).

Inside the drop-in models these modules are hyper-parameter holders: the gathers, the BPR term and the regulariser of a
training batch run in ONE fused launch (`ops.rank_loss`, csrc/rank_loss.cu), which reads `gamma` / `norm` from here.
Called directly -- `model.mf_loss(pos_scores, neg_scores)`, `model.reg_loss(u_ego, pos_ego, neg_ego)`, as reference
code does -- they compute the same values on scores / rows the caller already holds, through the stand-alone kernels
`fr_bpr_scores_fwd` / `fr_l2_norm_f32` (CUDA fp32 only: there is no CPU path)."""
import torch
import torch.nn as nn

from .. import _lib


def _cuda_f32(t, what):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise _lib.FoodRecError(f"{what}: expected a CUDA float32 tensor, got {t.device} {t.dtype} (no CPU fallback)")
    return t.contiguous()


class _BprScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, neg, gamma):
        pos, neg = _cuda_f32(pos, "BPRLoss"), _cuda_f32(neg, "BPRLoss")
        if pos.shape != neg.shape:
            raise ValueError(f"BPRLoss: pos {tuple(pos.shape)} and neg {tuple(neg.shape)} differ")
        out = torch.empty(1, device=pos.device)
        coef = torch.empty_like(pos)
        _lib.check(_lib.lib.fr_bpr_scores_fwd(pos.data_ptr(), neg.data_ptr(), pos.numel(), gamma, out.data_ptr(),
                                              coef.data_ptr(), _lib.stream_ptr()), "fr_bpr_scores_fwd")
        ctx.save_for_backward(coef)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (coef,) = ctx.saved_tensors
        d = coef * g
        return d, -d, None


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _cuda_f32(x, "EmbLoss")
        out = torch.empty(1, device=x.device)
        _lib.check(_lib.lib.fr_l2_norm_f32(x.data_ptr(), x.numel(), out.data_ptr(), _lib.stream_ptr()), "fr_l2_norm_f32")
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, nrm = ctx.saved_tensors
        return x * (g / nrm.clamp_min(1e-30))      # torch.norm's subgradient at 0 is 0


class BPRLoss(nn.Module):
    """loss.py:8-34: `-mean(log(gamma + sigmoid(pos - neg)))`."""

    def __init__(self, gamma=1e-10):
        super().__init__()
        self.gamma = gamma

    def forward(self, pos_score, neg_score):
        return _BprScores.apply(pos_score, neg_score, float(self.gamma))


class EmbLoss(nn.Module):
    """loss.py:37-50: `sum_k ||E_k||_2 / E_last.shape[0]`, shape [1] (the un-squared norm, as the reference has it)."""

    def __init__(self, norm=2):
        super().__init__()
        self.norm = norm

    def forward(self, *embeddings):
        if self.norm != 2:
            raise NotImplementedError("EmbLoss: only norm=2 (the value every reference model uses)")
        total = None
        for e in embeddings:
            n = _L2Norm.apply(e)
            total = n if total is None else total + n
        return total / embeddings[-1].shape[0]

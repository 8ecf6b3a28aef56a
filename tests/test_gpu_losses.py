"""GPU parity of the fused loss kernels against reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_clussl import close

pytestmark = pytest.mark.gpu


def test_rank_loss_matches_reference_modules():
    from foodrec_b200 import ops
    from oracle import losses
    torch.manual_seed(3)
    U, I, d, B = 50, 70, 64, 97
    emb = (torch.randn(U + I, d) * 0.3).requires_grad_(True)
    uw, iw = (torch.randn(U, d) * 0.2).requires_grad_(True), (torch.randn(I, d) * 0.2).requires_grad_(True)
    gw = (torch.randn(11, d) * 0.2).requires_grad_(True)
    u, p, n = torch.randint(0, U, (B,)), torch.randint(0, I, (B,)), torch.randint(0, I, (B,))
    ing = torch.randint(0, 11, (B, 20))
    mf = losses.bpr_from_tables(emb[:U], emb[U:], u, p, n)
    gpad = torch.nn.functional.embedding(ing, gw, padding_idx=10)
    reg = losses.emb_loss(uw[u], iw[p], iw[n], gpad)
    (mf * 0.7 + reg.sum() * 1.3).backward()
    c = lambda t: t.detach().cuda().requires_grad_(True)
    emb_d, uw_d, iw_d, gw_d = c(emb), c(uw), c(iw), c(gw)
    mf_d, reg_d = ops.rank_loss(emb_d, U, u.cuda(), p.cuda(), n.cuda(),
                                [(uw_d, u.cuda()), (iw_d, p.cuda()), (iw_d, n.cuda()), (gw_d, ing.cuda(), 10)],
                                reg_den=float(B))
    (mf_d * 0.7 + reg_d * 1.3).backward()
    close(mf_d, mf.detach().numpy())
    close(reg_d, reg.detach().numpy()[0])
    for a, b in ((emb_d, emb), (uw_d, uw), (iw_d, iw), (gw_d, gw)):
        close(a.grad, b.grad.numpy(), rtol=2e-5)
    assert float(gw_d.grad[10].abs().max()) == 0.0  # padding row gets no gradient


def test_bpr_and_embloss_golden():
    from foodrec_b200 import ops
    g = load_golden("primitives.npz")
    # BPR on scores: build rows whose dot products reproduce the golden scores exactly
    pos, neg = torch.from_numpy(g["bpr/pos"]), torch.from_numpy(g["bpr/neg"])
    B = pos.numel()
    emb = torch.zeros(1 + 2 * B, 32)
    emb[0, 0] = 1.0
    emb[1:1 + B, 0] = pos
    emb[1 + B:, 0] = neg
    z = torch.zeros(B, dtype=torch.long)
    e1, e2, e3 = (torch.from_numpy(g[k]) for k in ("emb/e1", "emb/e2", "emb/e3"))
    embs = [torch.nn.functional.pad(e, (0, 0, 0, 0))[:, :].contiguous() for e in (e1, e2, e3)]
    tabs = [torch.cat([e[:, :32], e[:, 32:]], 0).cuda() for e in embs]  # [2n, 32]: same Frobenius norm
    mf, reg = ops.rank_loss(emb.cuda(), 1, z.cuda(), torch.arange(B).cuda(), (torch.arange(B) + B).cuda(),
                            [(t, torch.arange(t.shape[0]).cuda()) for t in tabs], reg_den=float(e3.shape[0]))
    close(mf, g["bpr/out"])
    close(reg, g["emb/out"][0])


def test_loss_modules_called_directly_match_reference_golden():
    """`model.mf_loss(pos, neg)` / `model.reg_loss(e1, e2, e3)` as reference code calls them (FoodRec/common/loss.py:31-34,
    44-50): values against the reference-run golden, gradients against the same formula in torch-CPU."""
    from foodrec_b200.common.loss import BPRLoss, EmbLoss
    g = load_golden("primitives.npz")
    pos, neg = torch.from_numpy(g["bpr/pos"]), torch.from_numpy(g["bpr/neg"])
    pd, nd = pos.cuda().requires_grad_(True), neg.cuda().requires_grad_(True)
    out = BPRLoss()(pd, nd)
    close(out, g["bpr/out"])
    out.backward()
    pc, nc = pos.clone().requires_grad_(True), neg.clone().requires_grad_(True)
    (-torch.log(1e-10 + torch.sigmoid(pc - nc)).mean()).backward()
    close(pd.grad, pc.grad.numpy(), rtol=2e-5)
    close(nd.grad, nc.grad.numpy(), rtol=2e-5)
    es = [torch.from_numpy(g[k]) for k in ("emb/e1", "emb/e2", "emb/e3")]
    ed = [e.cuda().requires_grad_(True) for e in es]
    reg = EmbLoss()(*ed)
    assert tuple(reg.shape) == (1,)
    close(reg, g["emb/out"])
    (reg.sum() * 1.7).backward()
    ec = [e.clone().requires_grad_(True) for e in es]
    (sum(torch.norm(e, p=2) for e in ec) / ec[-1].shape[0] * 1.7).backward()
    for a, b in zip(ed, ec):
        close(a.grad, b.grad.numpy(), rtol=2e-5)
    with pytest.raises(Exception):
        BPRLoss()(pos, neg)          # CPU tensors: no fallback


def test_distance_correlation_golden_and_grad():
    """Against the reference golden (fp32, CPU) and against the same formula evaluated in fp64.

    The reference's fp32 result carries noise of its own: its diagonal
    D_ii = sqrt(max(r_i - 2 x_i.x_i + r_i, 0) + 1e-8) is BLAS rounding (1e-4..3e-4 instead of 1e-4),
    the diagonal carries ~n/(n + 0.01 n^2) of dcov_xx, and autograd multiplies that noise by
    1/(2 D_ii) = 5000 before it cancels.  So: tight bound against fp64, loose against fp32."""
    from foodrec_b200 import ops
    from oracle import losses
    g = load_golden("primitives.npz")
    x = torch.from_numpy(g["dcor/x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(g["dcor/y"]).cuda().requires_grad_(True)
    d = ops.correlation_distance(x, y)
    d.sum().backward()
    close(d, g["dcor/out"].reshape(1), rtol=1e-4)
    close(x.grad, g["dcor/gx"], rtol=3e-3)
    close(y.grad, g["dcor/gy"], rtol=3e-3)
    x64 = torch.from_numpy(g["dcor/x"]).double().requires_grad_(True)
    y64 = torch.from_numpy(g["dcor/y"]).double().requires_grad_(True)
    d64 = losses.correlation_distance(x64, y64)
    d64.sum().backward()
    close(d, d64.detach().numpy().reshape(1), rtol=1e-5)
    close(x.grad, x64.grad.numpy(), rtol=1e-4)
    close(y.grad, y64.grad.numpy(), rtol=1e-4)


@pytest.mark.parametrize("n", [100, 1024])
def test_three_view_dcor_vs_oracle(n):
    from foodrec_b200 import ops
    from oracle import losses
    torch.manual_seed(n)
    rows = 3000
    tabs = [(torch.randn(rows + k * 10, 64) * 0.1) for k in range(3)]
    idx = torch.randint(0, rows, (n,))
    idx[5] = idx[17]  # duplicate rows in the batch (same item as pos and neg)
    w = torch.tensor([0.3, 1.0, -0.5])
    refs = {}
    for dt in (torch.float32, torch.float64):
        ts = [t.detach().clone().to(dt).requires_grad_(True) for t in tabs]
        a, b, c = (t[idx] for t in ts)
        ref = torch.stack([losses.correlation_distance(a, b), losses.correlation_distance(a, c),
                           losses.correlation_distance(c, b)]).reshape(-1)
        (ref * w.to(dt)).sum().backward()
        refs[dt] = (ref.detach().numpy(), [t.grad.numpy() for t in ts])
    tabs_d = [t.cuda().requires_grad_(True) for t in tabs]
    out = ops.dcor_terms(tabs_d, idx.cuda(), [(0, 1), (0, 2), (2, 1)])
    (out * w.cuda()).sum().backward()
    close(out, refs[torch.float64][0], rtol=1e-5)
    close(out, refs[torch.float32][0], rtol=1e-4)
    for td, g64, g32 in zip(tabs_d, refs[torch.float64][1], refs[torch.float32][1]):
        close(td.grad, g64, rtol=1e-4)
        close(td.grad, g32, rtol=1e-2)  # the fp32 autograd reference is itself ~5e-3 noisy at n=1024 (see above)


def test_info_nce_golden():
    from foodrec_b200 import ops
    g = load_golden("primitives.npz")
    h = torch.from_numpy(g["nce/h"]).cuda().requires_grad_(True)
    c = ops.info_nce(h)
    c.backward()
    close(c, g["nce/out"])
    close(h.grad, g["nce/gh"], rtol=1e-4)


@pytest.mark.parametrize("rows,temperature,norm", [(1024, 0.5, True), (130, 0.2, True), (257, 0.5, False)])
def test_info_nce_vs_oracle(rows, temperature, norm):
    """Fused NT-Xent (one Gram matrix) vs the reference's four-block formulation restated in the oracle,
    including an odd trailing row (ignored, zero gradient) and the un-normalised variant."""
    from foodrec_b200 import ops
    from oracle import losses
    torch.manual_seed(rows)
    h = (torch.randn(rows, 64) * (1.0 if norm else 0.3)).requires_grad_(True)
    ref = losses.info_nce(h, temperature=temperature, hidden_norm=norm)
    ref.backward()
    hd = h.detach().cuda().requires_grad_(True)
    out = ops.info_nce(hd, temperature=temperature, hidden_norm=norm)
    out.backward()
    close(out, ref.detach().numpy())
    close(hd.grad, h.grad.numpy(), rtol=2e-5)


def test_cosine_mean_vs_torch():
    """HealthRec's KD cosine term: fused gather + cosine + mean vs `F.cosine_similarity(A, T[idx]).mean()`."""
    from foodrec_b200 import ops
    torch.manual_seed(8)
    A = torch.randn(1024, 64).requires_grad_(True)
    T = torch.randn(5000, 64).requires_grad_(True)
    idx = torch.randint(0, 5000, (1024,))
    idx[3] = idx[9]
    ref = torch.nn.functional.cosine_similarity(A, T[idx], dim=-1).mean()
    (ref * 1.7).backward()
    Ad, Td = A.detach().cuda().requires_grad_(True), T.detach().cuda().requires_grad_(True)
    out = ops.cosine_mean(Ad, Td, idx.cuda())
    (out * 1.7).backward()
    close(out, ref.detach().numpy())
    close(Ad.grad, A.grad.numpy(), rtol=2e-5)
    close(Td.grad, T.grad.numpy(), rtol=2e-5)

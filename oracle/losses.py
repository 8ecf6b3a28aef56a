"""Ranking / contrastive losses on gathered rows (oracle).  Test infrastructure only."""
import torch
import torch.nn.functional as F


def bpr_loss(pos_score, neg_score, gamma=1e-10):
    """FoodRec/common/loss.py:28-34."""
    return -torch.log(gamma + torch.sigmoid(pos_score - neg_score)).mean()


def emb_loss(*embs):
    """Un-squared Frobenius norms summed, divided by the LAST argument's leading extent.
    FoodRec/common/loss.py:45-50."""
    tot = torch.zeros(1, device=embs[0].device)      # the reference: `torch.zeros(1).to(embeddings[-1].device)`, loss.py:46
    for e in embs:
        tot = tot + torch.norm(e, p=2)
    return tot / embs[-1].shape[0]


def bpr_from_tables(user_all, item_all, u, p, n):
    """Gathers + row dot products + BPR.  FoodRec/models/cikm_model.py:255-261
    (same at pricai_modelx.py:252-258, lightgcn.py:159-166)."""
    ue, pe, ne = user_all[u], item_all[p], item_all[n]
    return bpr_loss((ue * pe).sum(1), (ue * ne).sum(1))


def correlation_distance(x, y):
    """Distance correlation of two [n, d] views.  FoodRec/models/pricai_modelx.py:409-437."""
    zero = torch.zeros(1, device=x.device)           # pricai_modelx.py:410 `torch.zeros(1).to(self.device)`

    def centred(X):
        r = (X * X).sum(1, keepdim=True)
        D = torch.sqrt(torch.maximum(r - 2 * (X @ X.t()) + r.t(), zero) + 1e-8)
        return D - D.mean(0, keepdim=True) - D.mean(1, keepdim=True) + D.mean()

    def dcov(A, B):
        n = float(A.shape[0])
        return torch.sqrt(torch.maximum((A * B).sum() / (n * n), zero) + 1e-8)

    D1, D2 = centred(x), centred(y)
    d12, d11, d22 = dcov(D1, D2), dcov(D1, D1), dcov(D2, D2)
    return d12 / torch.sqrt(torch.maximum(d11 * d22, zero) + 1e-10)


def info_nce(hidden, temperature=0.5, hidden_norm=True):
    """SimCLR NT-Xent over the two halves of `hidden`.  FoodRec/models/pricai_modelx.py:354-378
    (`CL_loss`; dormant in the reference, call commented out at :259)."""
    b = hidden.shape[0] // 2
    if hidden_norm:
        hidden = F.normalize(hidden, p=2, dim=-1)
    h1, h2 = hidden[:b], hidden[b:2 * b]
    eye = torch.eye(b, device=hidden.device) * 1e9
    aa = h1 @ h1.t() / temperature - eye
    bb = h2 @ h2.t() / temperature - eye
    ab = h1 @ h2.t() / temperature
    ba = h2 @ h1.t() / temperature
    lab = torch.arange(b, device=hidden.device)
    la = F.cross_entropy(torch.cat([ab, aa], 1), lab)
    lb = F.cross_entropy(torch.cat([ba, bb], 1), lab)
    return (la + lb) / b


def kd_cosine_loss(item_know, pos_e, neg_e, threshold):
    """`max(0, 1 - mean(cos(item_know, [pos;neg])) - thr)`.  FoodRec/models/cikm_model.py:263-264,304-308."""
    kd = 1 - F.cosine_similarity(item_know, torch.cat([pos_e, neg_e], 0), dim=-1).mean()
    return torch.max(torch.tensor(0.0), kd - threshold)


def clussl_loss(fwd_out, user_w, item_w, u, p, n, reg_weight, loss_cl):
    """The three terms `PRICAI_ModelX.calculate_loss` returns.  FoodRec/models/pricai_modelx.py:234-276."""
    user_all, item_all, (img, txt, ing) = fwd_out
    allit = torch.cat([p, n], 0)
    a, b, c = img[allit], txt[allit], ing[allit]
    mf = bpr_from_tables(user_all, item_all, u, p, n)
    cl = correlation_distance(a, b) + correlation_distance(a, c) + correlation_distance(c, b)
    reg = reg_weight * emb_loss(user_w[u], item_w[p], item_w[n])
    return mf, loss_cl * cl, reg

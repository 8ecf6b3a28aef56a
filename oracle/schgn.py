"""CPU oracle for SCHGN (FoodRec/models/schgn.py).  TEST INFRASTRUCTURE ONLY.

Two things live here:

* `GCNConvRestated` / `install_pyg_stub()` -- a torch-only stand-in for
  `torch_geometric.nn.GCNConv` (PyG is not installable offline).  `tests/golden/make_golden.py`
  installs it so that the reference's own `SCHGN` class can be imported and EXECUTED; every other
  line of the model that produced `tests/golden/schgn_mini.npz` is the reference's.  The stand-in
  follows PyG's documented defaults (`add_self_loops=True, normalize=True, improved=False,
  cached=False, bias=True`, flow source->target, `lin` without bias, glorot weight / zero bias
  init) -- that one operator stays "parity unpinned".
* a functional restatement of the model's arithmetic over a `state_dict` (`P`), used by the GPU
  parity tests at sizes beyond the golden and on machines without the reference checkout.
  Dropout is the identity here (the goldens are produced with `F.dropout` patched to the identity
  and the module in `eval()` mode; the reference's masks depend on the RNG stream of the device).
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

from .adjacency import gcn_norm_edges, schgn_edge_index
from .propagation import gcn_conv_tanh


class GCNConvRestated(torch.nn.Module):
    """`torch_geometric.nn.GCNConv(in, out)` with default arguments (call site schgn.py:34)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = torch.nn.Module()                               # PyG's own Linear: `weight` only,
        self.lin.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        self.lin.weight.data.uniform_(-a, a)                       # initialised once, `glorot`
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))  # PyG `zeros`

    def forward(self, x, edge_index):
        src, dst, w = gcn_norm_edges(edge_index, x.shape[0])
        h = x @ self.lin.weight.t()
        out = torch.zeros_like(h).index_add_(0, dst, h[src] * w[:, None])
        return out + self.bias


def install_pyg_stub():
    mod, nn_mod = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn")
    nn_mod.GCNConv = GCNConvRestated
    mod.nn = nn_mod
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = mod, nn_mod


# ----------------------------------------------------------------------------- functional restatement
def _lin(P, name, x):
    b = P.get(name + ".bias")
    return F.linear(x, P[name + ".weight"], b)


def gcn_tables(P, edge_index, sizes):
    """schgn.py:241-250: x = cat(user, item, ingredient, health) -> tanh(GCNConv) -> split."""
    x = torch.cat([P["user_embed"], P["item_embed"], P["ingre_embed_first"], P["health_embed"]], 0)
    src, dst, w = gcn_norm_edges(edge_index, x.shape[0])
    g = gcn_conv_tanh(x, src, dst, w, P["new_gcn.conv1.lin.weight"], P["new_gcn.conv1.bias"])
    return torch.split(g, list(sizes), 0)


def attention_ingredient_level(P, ingre_emb, u_emb, img_emb, ingre_num):
    """schgn.py:159-184."""
    n = ingre_emb.shape[1]
    cat = torch.cat([ingre_emb, u_emb[:, None, :].expand(-1, n, -1), img_emb[:, None, :].expand(-1, n, -1)], 2)
    a = _lin(P, "h_att_ingre", torch.tanh(_lin(P, "W_att_ingre", cat))).squeeze(-1)
    valid = (torch.arange(n)[None, :] < ingre_num[:, None]).float()
    a = torch.softmax(a + (1.0 - valid) * -1e12, dim=1)
    return (a[:, :, None] * ingre_emb).sum(1)


def attention_id_ingre_image(P, u, i_emb, ingre_att, img_emb, hl_emb):
    """schgn.py:186-206.  The reference stacks the four (user, component) pairs along dim 0
    (`[4b, 2e]`, component-major) and then reads the `[4b]` logits back with `.view(b, -1)`: row r of
    the softmax holds flat entries `4r .. 4r+3`, i.e. logit (component (4r+k) // b, sample
    (4r+k) % b), not the four components of sample r.  Restated as written (results must match)."""
    b = u.shape[0]
    comps = torch.stack([i_emb, ingre_att, img_emb, hl_emb], 0)                       # [4, b, e] component-major
    cp = torch.cat([u[None].expand(4, -1, -1), comps], 2)
    logit = _lin(P, "h_att_comp", torch.tanh(_lin(P, "W_att_comp", cp))).reshape(4 * b)
    B = torch.softmax(logit.view(b, 4), 1)                                            # rows mix samples
    return (B[:, :, None] * comps.permute(1, 0, 2)).sum(1)


def compute_score(P, tables, user, item, ingre, ingre_num, img, hl, ingre_embedding):
    """schgn.py:233-268 without dropout.  `tables` = `gcn_tables(...)`."""
    ug, ig, gg, hg = tables
    ingre_g = torch.cat([gg, P["ingre_embed_second"], P["ingre_embed_mask"]], 0)
    u_emb, i_emb = P["user_embed"][user], P["item_embed"][item]
    ingre_emb, hl_emb = ingre_embedding[ingre], P["health_embed"][hl]
    img_emb = _lin(P, "img_trans", img.to(torch.float32))
    u_f, i_f = u_emb + ug[user], i_emb + ig[item]
    ingre_f, hl_f = ingre_emb + ingre_g[ingre], hl_emb + hg[hl]
    ingre_att = attention_ingredient_level(P, ingre_f, u_f, img_emb, ingre_num)
    item_att = attention_id_ingre_image(P, u_f, i_f, ingre_att, img_emb, hl_f)
    hidden = _lin(P, "W_concat", torch.cat([u_f, item_att, u_f * item_att], 1))
    score = _lin(P, "output_mlp", torch.relu(hidden)).squeeze(-1)
    return score, u_emb, i_emb, ingre_emb, hl_emb, ingre_g, item_att


def _encoder(P, x, mask, n_layers, n_heads):
    """FoodRec/common/module.py:48-190 in eval mode (no dropout)."""
    for l in range(n_layers):
        p = f"ingre_encoder.layer.{l}."
        B, L, H = x.shape
        dh = H // n_heads

        def heads(t):
            return t.view(B, L, n_heads, dh).permute(0, 2, 1, 3)

        q, k, v = (heads(_lin(P, p + "attention." + n, x)) for n in ("query", "key", "value"))
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + mask, -1)
        ctx = (att @ v).permute(0, 2, 1, 3).reshape(B, L, H)
        h = F.layer_norm(_lin(P, p + "attention.dense", ctx) + x, (H,), P[p + "attention.LayerNorm.weight"],
                         P[p + "attention.LayerNorm.bias"], 1e-12)
        t = _lin(P, p + "intermediate.dense_1", h)
        t = t * 0.5 * (1.0 + torch.erf(t / math.sqrt(2.0)))
        x = F.layer_norm(_lin(P, p + "intermediate.dense_2", t) + h, (H,), P[p + "intermediate.LayerNorm.weight"],
                         P[p + "intermediate.LayerNorm.bias"], 1e-12)
    return x


def ssl_loss(P, ingre_embedding, ingre_g, masked_seq, pos_seq, neg_seq, n_ingredients, n_layers, n_heads):
    """schgn.py:208-231 (masked-ingredient prediction)."""
    mask = ((masked_seq == n_ingredients).float() * -1e8)[:, None, None, :]
    new = _encoder(P, ingre_g[masked_seq], mask, n_layers, n_heads)
    e = _lin(P, "mip_norm", new.reshape(-1, new.shape[-1]))
    pos = torch.sigmoid((e * ingre_embedding[pos_seq].reshape(e.shape)).sum(-1))
    neg = torch.sigmoid((e * ingre_embedding[neg_seq].reshape(e.shape)).sum(-1))
    d = torch.sigmoid(pos - neg)
    bce = F.binary_cross_entropy(d, torch.ones_like(d), reduction="none")
    return (bce * (masked_seq == n_ingredients + 1).float().flatten()).sum()


def calculate_loss(P, batch, cfg, edge_index, sizes):
    """schgn.py:270-316 -> (bpr, reg, ssl)."""
    ingre_embedding = torch.cat([P["ingre_embed_first"], P["ingre_embed_second"], P["ingre_embed_mask"]], 0)
    tables = gcn_tables(P, edge_index, sizes)
    pos = compute_score(P, tables, batch["u_id"], batch["pos_i_id"], batch["pos_ingre_code"], batch["pos_ingre_num"],
                        batch["pos_img"], batch["pos_cl"].long(), ingre_embedding)
    neg = compute_score(P, tables, batch["u_id"], batch["neg_i_id"], batch["neg_ingre_code"], batch["neg_ingre_num"],
                        batch["neg_img"], batch["neg_cl"].long(), ingre_embedding)
    ssl = cfg["ssl"] * ssl_loss(P, ingre_embedding, pos[5], batch["masked_ingre_seq"], batch["pos_ingre_seq"],
                                batch["neg_ingre_seq"], sizes[2], cfg["num_hidden_layers"],
                                cfg["num_attention_heads"])
    bpr = -torch.log(torch.sigmoid(pos[0] - neg[0])).sum()
    sq = lambda t: (t ** 2).sum()  # noqa: E731
    reg = cfg["regs"] * (sq(pos[1]) + sq(pos[2]) + sq(neg[2]) + sq(pos[3]) + sq(neg[3]))
    reg = reg + cfg["reg_health"] * (sq(pos[4]) + sq(neg[4]))
    reg = reg + cfg["reg_image"] * sq(P["img_trans.weight"])
    reg = reg + cfg["reg_w"] * (sq(P["W_concat.weight"]) + sq(P["output_mlp.weight"]))
    reg = reg + cfg["reg_g"] * sq(P["new_gcn.conv1.lin.weight"])
    return bpr, reg, ssl


def full_sort_scores(P, ds, user, edge_index, sizes):
    """schgn.py:318-345: one user against every item (inference embedding table has no mask row)."""
    I = ds.n_items
    ingre_embedding = torch.cat([P["ingre_embed_first"], P["ingre_embed_second"]], 0)
    tables = gcn_tables(P, edge_index, sizes)
    out = compute_score(P, tables, torch.full((I,), int(user), dtype=torch.long), torch.arange(I),
                        torch.from_numpy(np.asarray(ds.ingredientCodeDict)).long(),
                        torch.from_numpy(np.asarray(ds.ingredientNum)).long(),
                        torch.from_numpy(np.asarray(ds.embImage, dtype=np.float32)),
                        torch.from_numpy(np.asarray(ds.cal_level)).long(), ingre_embedding)
    return out[0]


__all__ = ["GCNConvRestated", "install_pyg_stub", "gcn_tables", "compute_score", "ssl_loss", "calculate_loss",
           "full_sort_scores", "schgn_edge_index"]

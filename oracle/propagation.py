"""Graph propagation forwards (oracle).  Test infrastructure only."""
import torch


def layer_mean_propagate(S, ego: torch.Tensor, n_layers: int) -> torch.Tensor:
    """`mean_l(S^l ego)`, l = 0..n_layers, via `torch.sparse.mm` then `stack(dim=1).mean(dim=1)`.

    FoodRec/models/cikm_model.py:185-190 / 196-202; pricai_modelx.py:181-186 etc.; lightgcn.py:135-142.
    """
    outs = [ego]
    x = ego
    for _ in range(n_layers):
        x = torch.sparse.mm(S, x)
        outs.append(x)
    return torch.stack(outs, dim=1).mean(dim=1)


def healthrec_forward(S_ui, S_ri, user_w, item_w, ingre_w, n_users, n_items, n_ingredients,
                      n_layers, ui_layers):
    """HealthRec (`CIKM_Model.forward`), FoodRec/models/cikm_model.py:182-208.

    `ingre_w` has the padding row last ([G+1, d]); it is dropped before propagation (:184).
    Returns (user_all, item_all, ingre_ir).
    """
    ir = layer_mean_propagate(S_ri, torch.cat([item_w, ingre_w[:-1]], 0), n_layers)
    item_ir, ingre_ir = torch.split(ir, [n_items, n_ingredients])
    ui = layer_mean_propagate(S_ui, torch.cat([user_w, item_ir], 0), ui_layers)
    user_all, item_all = torch.split(ui, [n_users, n_items])
    return user_all, item_all, ingre_ir


def clussl_forward(S_ui, S_ingre, S_image, S_text, user_w, item_w, ingre_w, image_proto, text_proto,
                   n_users, n_items, n_ingredients, n_cluster, n_ri_layers, n_ui_layers):
    """CLUSSL (`PRICAI_ModelX.forward`), FoodRec/models/pricai_modelx.py:179-232.

    All three item-side propagations run `n_ri_layers` layers (`n_mm_layers` is read but unused,
    :34,196,210).  `image_proto` / `text_proto` are the [n_cluster, d] tables entering the graph
    (already projected by image_trs/text_trs when centre embeddings are in use, :190-192,204-206).
    Returns (user_all, item_all, (item_image, item_text, item_ingre)).
    """
    ing = layer_mean_propagate(S_ingre, torch.cat([item_w, ingre_w[:-1]], 0), n_ri_layers)
    item_ingre = ing[:n_items]
    img = layer_mean_propagate(S_image, torch.cat([item_w, image_proto], 0), n_ri_layers)
    item_image = img[:n_items]
    txt = layer_mean_propagate(S_text, torch.cat([item_w, text_proto], 0), n_ri_layers)
    item_text = txt[:n_items]
    item_emb = item_ingre + item_image + item_text
    ui = layer_mean_propagate(S_ui, torch.cat([user_w, item_emb], 0), n_ui_layers)
    user_all, item_all = torch.split(ui, [n_users, n_items])
    return user_all, item_all, (item_image, item_text, item_ingre)


def lightgcn_forward(S_ui, user_w, item_ego, n_users, n_items, n_layers):
    """FoodRec/models/lightgcn.py:134-147; `item_ego = image_trs(image_embedding.weight)` (:122-132)."""
    ui = layer_mean_propagate(S_ui, torch.cat([user_w, item_ego], 0), n_layers)
    return torch.split(ui, [n_users, n_items])


def gcn_conv_tanh(x, src, dst, w, lin_weight, bias):
    """`tanh(GCNConv(x, edge_index))`, FoodRec/models/schgn.py:29-41: `lin` (no bias) first,
    weighted scatter-add of source rows into targets, `+ bias`, tanh.  (src, dst, w) come from
    `adjacency.gcn_norm_edges`, which the reference recomputes per call (`cached=False`).
    """
    h = x @ lin_weight.t()
    out = torch.zeros_like(h).index_add_(0, dst, h[src] * w[:, None])
    return torch.tanh(out + bias)

"""The drop-in models driven through the reference trainer's exact call sequence.

Restated from FoodRec/common/trainer.py (the file cannot travel to the GPU box): `_train_epoch` :156-229 (train mode,
batch moved to the device, `copy.deepcopy` of the batch, `calculate_loss` -> tuple, `sum`, per-term `.item()`, NaN
check, `backward`, `optimizer.step`), `_valid_by_user_epoch` :231-282 (`forward()` unpacked into exactly three values,
`validRatings[user_idx]`, per-user `inference_fast` -> `.cpu().detach().numpy().copy()` -> `np.argsort(...)[::-1]`)
and `evaluate` :476-503 (`full_sort_predict(batch)` -> `torch.topk(scores, max(topk), dim=-1)`,
`validRatings[batch_idx]`).  Every step is checked against the CPU oracle."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Config(dict):
    """FoodRec/utils/configurator.py:121-125: missing keys read as None."""

    def __getitem__(self, k):
        return self.get(k)


def clussl(ds):
    from foodrec_b200.models.pricai_modelx import PRICAI_ModelX
    cfg = Config(device="cuda", embedding_size=64, train_batch_size=64, is_multimodal_model=True, end2end=False,
                 use_health_level_multi_hot=True, n_ri_layers=2, n_mm_layers=1, n_ui_layers=1, reg_weight=0.01,
                 loss_cl=0.1, n_cluster=ds.cfg.n_cluster, graph_inference_fast=True, neg_sample_num=50, topk=[10, 20, 50])
    torch.manual_seed(999)
    m = PRICAI_ModelX(cfg, ds)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.to(cfg["device"]), cfg, sd


def test_train_epoch_call_sequence(mini_ds):
    import bench
    from foodrec_b200.synth import sample_train_batches
    ds = mini_ds
    model, cfg, sd = clussl(ds)
    device = torch.device(cfg["device"])
    optimizer = torch.optim.Adam(model.parameters(), lr=0.002)               # trainer.py:142-143
    oracle = bench.OracleClussl(ds, sd, 0.002)
    train_data = [{k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}   # what default_collate yields (CPU tensors)
                  for b in sample_train_batches(ds, 64, 4, seed=17)]
    # ---- trainer.py:171-224
    model.train()
    loss_func = model.calculate_loss
    total_loss, loss_batches = None, []
    for batch_idx, interaction in enumerate(train_data):
        ref_terms = [float(t) for t in oracle.step({k: v.numpy() for k, v in interaction.items()})]
        interaction = {k: v.to(device, non_blocking=True) for k, v in interaction.items()}
        optimizer.zero_grad()
        second_inter = copy.deepcopy(interaction)                           # :181
        losses = loss_func(interaction)
        assert isinstance(losses, tuple) and all(torch.is_tensor(x) and x.numel() == 1 for x in losses)
        loss = sum(losses)
        loss_tuple = tuple(per_loss.item() for per_loss in losses)          # :186
        total_loss = loss_tuple if total_loss is None else tuple(map(sum, zip(total_loss, loss_tuple)))
        assert not torch.isnan(loss).any()                                  # _check_nan
        loss.backward()
        optimizer.step()
        loss_batches.append(loss.detach())
        for name, a, r in zip(("mf", "cl", "reg"), loss_tuple, ref_terms):
            tol = 2e-4 if name == "cl" else 1e-5
            assert abs(a - r) <= tol * max(abs(r), 1e-3), (batch_idx, name, a, r)
        assert all(torch.equal(second_inter[k], interaction[k]) for k in interaction)   # the batch is not modified
    assert len(total_loss) == 3 and len(loss_batches) == 4


def test_valid_by_user_epoch_call_sequence(mini_ds):
    from oracle import adjacency, propagation
    ds = mini_ds
    model, cfg, sd = clussl(ds)
    device = torch.device(cfg["device"])
    rng = np.random.default_rng(4)
    model.train()                                # the trainer leaves the mode to the caller: both must work
    user_emb, item_emb, ingre_emb = model.forward()                         # trainer.py:235-236: exactly three values
    assert len(ingre_emb) == 3
    S = [adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items),
         adjacency.norm_adj_item_side(ds.rIngre_triples, ds.n_items, ds.num_ingredients),
         adjacency.norm_adj_item_side(ds.image_cluster_triples, ds.n_items, ds.cfg.n_cluster),
         adjacency.norm_adj_item_side(ds.text_cluster_triples, ds.n_items, ds.cfg.n_cluster)]
    ref_u, ref_i, _ = propagation.clussl_forward(
        S[0], S[1], S[2], S[3], sd["user_embedding.weight"], sd["item_embedding.weight"], sd["ingre_embedding.weight"],
        sd["image_prototype_embedding.weight"], sd["text_prototype_embedding.weight"], ds.n_users, ds.n_items,
        ds.num_ingredients, ds.cfg.n_cluster, 2, 1)
    model.eval()
    for user_idx in range(0, 40):
        pos_items = model.dataset.validRatings[user_idx]                    # :239
        negs = rng.choice(ds.n_items, size=50, replace=False)
        cand = np.concatenate([np.asarray(pos_items, dtype=np.int64), negs])   # positives first (dataloader.py:228-302)
        user_batch = {"user_input": torch.full((cand.size,), user_idx, dtype=torch.int64), "item_input": torch.from_numpy(cand)}
        user_batch = {k: v.to(device, non_blocking=True) for k, v in user_batch.items()}
        predictions = model.inference_fast(user_batch, user_emb, item_emb)  # :243
        predictions = predictions.cpu().detach().numpy().copy()             # :247
        assert predictions.shape == (cand.size,)
        want = (ref_u[user_idx][None, :] * ref_i[cand]).sum(1).detach().numpy()
        assert np.allclose(predictions, want, rtol=1e-5, atol=1e-7)
        pred_idx = np.argsort(predictions)[::-1]                            # :253
        if np.abs(np.diff(np.sort(want))).min() > 1e-6:
            assert np.array_equal(pred_idx, np.argsort(want)[::-1])
        alt = model.inference_by_user(user_batch).cpu().numpy()             # :245 (graph_inference_fast off)
        assert np.allclose(alt, want, rtol=1e-5, atol=1e-7)


def test_evaluate_full_sort_call_sequence(mini_ds):
    from foodrec_b200 import evaluation as E, metrics as Mx
    ds = mini_ds
    model, cfg, sd = clussl(ds)
    device = torch.device(cfg["device"])
    model.eval()
    topk = cfg["topk"]
    batch_matrix_list, pos = [], []
    with torch.no_grad():
        ua, ia = model._tables()
        dense = (ua @ ia.t()).cpu()
    for batch_idx in range(60):                                              # trainer.py:489-503, one user per batch
        batched_data = {"u_id": torch.tensor([batch_idx])}
        batched_data = {k: v.to(device, non_blocking=True) for k, v in batched_data.items()}
        pos_items = model.dataset.validRatings[batch_idx]                   # :492
        with torch.no_grad():
            scores = model.full_sort_predict(batched_data)                  # :495
        assert scores.shape == (ds.n_items,)
        _, topk_index = torch.topk(scores, max(topk), dim=-1)               # :497
        batch_matrix_list.append(topk_index.cpu().tolist())
        pos.append(pos_items)
        assert torch.allclose(scores.cpu(), dense[batch_idx], rtol=1e-5, atol=1e-7)
    # the same users through the fused path (never materialises the score rows): identical lists away from ties,
    # identical metrics to 4 decimals
    _, fused = E.full_sort_topk(ua, ia, torch.arange(60, device=device), max(topk))
    fused = fused.cpu().numpy()
    ref = np.asarray(batch_matrix_list)
    mism = fused != ref
    if mism.any():
        rows = np.nonzero(mism.any(1))[0]
        for r in rows:
            assert np.abs(dense[r][fused[r]].numpy() - dense[r][ref[r]].numpy()).max() <= 2e-6
    assert Mx.topk_metrics(fused, pos, topk=tuple(topk)) == Mx.topk_metrics(ref, pos, topk=tuple(topk))

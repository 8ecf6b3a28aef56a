"""GPU parity of the SCHGN drop-in: the reference's own class executed (golden, GCNConv restated),
the oracle restatement at C1 scale, and the fused full-sort pair scorer (`fr_schgn_attend` /
`fr_schgn_score`) against both."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_clussl import Cfg, close, dev_batch

pytestmark = pytest.mark.gpu

CFG = dict(device="cuda", embedding_size=64, train_batch_size=64, is_multimodal_model=True, end2end=False,
           num_attention_heads=2, num_hidden_layers=2, hidden_act="gelu", inner_size=256, hidden_dropout_prob=0.5,
           attention_probs_dropout_prob=0.5, regs=0.01, reg_image=1, reg_w=0.05, reg_g=0.01, reg_health=0.01, ssl=0.008,
           SCHGN_ssl=True, neg_sample_num=4)


@pytest.fixture()
def no_dropout(monkeypatch):
    """The goldens / the oracle take dropout as the identity (masks depend on the device RNG)."""
    monkeypatch.setattr(torch.nn.functional, "dropout", lambda x, p=0.5, training=True, inplace=False: x)


def golden_model(mini_ds):
    from foodrec_b200.models.schgn import SCHGN
    g = load_golden("schgn_mini.npz")
    torch.manual_seed(999)
    m = SCHGN(Cfg(CFG), mini_ds)
    for k, v in m.state_dict().items():       # same seed => the reference's initial weights, bit for bit
        assert np.array_equal(v.numpy(), g["sd/" + k]), k
    return m.to("cuda").eval(), g


def test_gcn_and_losses_match_reference_run(mini_ds, no_dropout):
    from foodrec_b200.synth import sample_train_batches
    m, g = golden_model(mini_ds)
    close(torch.cat(m.gcn_tables(), 0), g["gcn/out"])
    for b, batch in enumerate(sample_train_batches(mini_ds, 64, 2, seed=11, schgn=True)):
        m.zero_grad()
        losses = m.calculate_loss(dev_batch(batch))
        close(torch.stack([x.reshape(()) for x in losses]), g[f"loss/{b}"])
        sum(losses).backward()
        for name, p in m.named_parameters():
            key = f"grad/{name}/{b}"
            if key in g and "key.bias" not in name:     # softmax is shift-invariant: that gradient is 0 + noise
                # table gradients at 2e-5; the attention / concat weights have gradients of ~1e-4 formed
                # from O(1) terms, so fp32 summation order (GPU vs CPU GEMMs) shows at ~1e-8 absolute
                close(p.grad, g[key], rtol=2e-5 if "embed" in name else 5e-4, atol=1e-9)


def test_full_sort_and_candidate_scores_match_reference_run(mini_ds):
    m, g = golden_model(mini_ds)
    for u in (0, 7, 200):
        s = m.full_sort_predict({"u_id": torch.tensor([u], device="cuda")})
        assert s.shape == (mini_ds.n_items,)
        assert np.abs(s.cpu().numpy() - g[f"full_sort/{u}"]).max() <= 2e-6
    cand = g["by_user/cand"]
    t = lambda a: torch.from_numpy(np.asarray(a)).cuda()  # noqa: E731
    with torch.no_grad():
        sc = m.inference_by_user({
        "user_input": torch.full((len(cand),), 3, device="cuda"), "item_input": t(cand),
        "img_input": t(mini_ds.embImage[cand]), "ingre_num_input": t(mini_ds.ingredientNum[cand]),
        "ingre_input": t(mini_ds.ingredientCodeDict[cand]), "cal_level_input": t(mini_ds.cal_level[cand])})
    assert np.abs(sc.cpu().numpy() - g["by_user/scores"]).max() <= 2e-6


def test_user_blocks_are_independent(mini_ds):
    """40 users (two full groups of 16 + a ragged one) against single-user calls; the per-user
    projections are cuBLAS GEMMs whose summation order depends on the batch size, hence not bitwise."""
    m, g = golden_model(mini_ds)
    users = torch.arange(3, 43, device="cuda")
    block = m.full_sort_scores(users)
    for r in (0, 15, 16, 33, 39):
        single = m.full_sort_scores(users[r:r + 1])[0]
        assert float((block[r] - single).abs().max()) <= 1e-6, r
    assert m.full_sort_scores(users[:0]).shape == (0, mini_ds.n_items)


@pytest.fixture(scope="module")
def c1_model():
    from foodrec_b200.synth import make_dataset
    from foodrec_b200.models.schgn import SCHGN
    ds = make_dataset("C1")
    torch.manual_seed(5)
    m = SCHGN(Cfg(CFG), ds)
    with torch.no_grad():   # move away from the near-symmetric initial point so that scores spread out
        for p in (m.user_embed, m.item_embed, m.ingre_embed_first, m.health_embed):
            p.mul_(20.0)
    return m.to("cuda").eval(), ds


def test_c1_full_sort_matches_oracle_and_topk_with_mask(c1_model):
    from foodrec_b200 import evaluation
    from oracle import schgn as O
    m, ds = c1_model
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ei = O.schgn_edge_index(ds)
    sizes = (ds.n_users, ds.n_items, ds.num_ingredients, ds.num_calories_level)
    users = torch.tensor([0, 11, 4999, 2500], device="cuda")
    got = m.full_sort_scores(users).cpu()
    ref = torch.stack([O.full_sort_scores(P, ds, int(u), ei, sizes) for u in users.cpu()])
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1e-5 * scale + 2e-6
    assert float(ref.std()) > 100 * 2e-6          # the comparison is not vacuous

    hist = evaluation.HistoryCSR(ds.train_coo_matrix, ds.n_users, "cuda")
    k = 20
    vals, idx = m.full_sort_topk(users, k, hist=hist)
    for r, u in enumerate(users.cpu().tolist()):
        seen = hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]]
        assert not np.isin(idx[r].cpu().numpy(), seen).any()
        masked = ref[r].clone()
        masked[torch.from_numpy(seen).long()] = float("-inf")
        rv, ri = torch.topk(masked, k)
        same = idx[r].cpu() == ri
        # positions may differ only where the fp32 scores tie within the comparison tolerance
        assert float((masked[idx[r].cpu()] - rv).abs().max()) <= 1e-5 * scale + 2e-6
        assert same.float().mean() >= 0.8
    v2, i2 = m.full_sort_topk(users, k)               # unmasked, as the reference evaluates
    assert torch.equal(i2, torch.topk(m.full_sort_scores(users), k).indices)


@pytest.mark.parametrize("k", [1, 20, 50, 64])
def test_fused_topk_equals_dense_scores_then_topk(c1_model, k):
    """`fr_schgn_score_topk` (selection fused into the scorer, `[users, items]` never written) against the dense
    scores of `fr_schgn_score` + mask + `torch.topk`, for users spanning several 16-user passes: identical values, and
    identical indices wherever the k-th score is not tied."""
    from foodrec_b200 import evaluation
    m, ds = c1_model
    users = torch.arange(3, 3 + 53 * 41, 41, device="cuda") % ds.n_users            # 53 users
    hist = evaluation.HistoryCSR(ds.train_coo_matrix, ds.n_users, "cuda")
    for h in (None, hist):
        dense = m.full_sort_scores(users)
        if h is not None:
            h.mask_scores_(dense, users)
        rv, ri = torch.topk(dense, k, dim=-1)
        v, i = m.full_sort_topk(users, k, hist=h)
        assert v.shape == (users.numel(), k) and i.dtype == torch.int64
        assert torch.equal(v, rv)                                   # same kernel arithmetic: bit-identical scores
        assert torch.equal(torch.gather(dense, 1, i), v)            # every index carries its score
        untied = (rv[:, :-1] != rv[:, 1:]).all(dim=1) if k > 1 else torch.ones(users.numel(), dtype=torch.bool, device="cuda")
        assert torch.equal(i[untied], ri[untied])
        # ties go to the lower item id
        assert bool(((v[:, :-1] > v[:, 1:]) | (i[:, :-1] < i[:, 1:])).all()) if k > 1 else True
    v70, i70 = m.full_sort_topk(users[:5], 70)                      # beyond the kernel's 64: dense fallback
    assert torch.equal(i70, torch.topk(m.full_sort_scores(users[:5]), 70).indices)


def test_c1_evaluate_full_sort_metrics_match_oracle_ranking(c1_model):
    """`evaluation.evaluate_full_sort` on SCHGN (the reference's `Trainer.evaluate` loop, no mask): Recall /
    NDCG / Precision / MAP @5/10/20/50 over 64 users equal those of the CPU restatement's ranking to 4 d.p."""
    from foodrec_b200 import evaluation, metrics as M
    from oracle import schgn as O
    m, ds = c1_model
    users = list(range(0, 5000, 79))[:64]
    pos = [ds.testRatings[u] for u in users]
    got, top = evaluation.evaluate_full_sort(m, users, pos)
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ei = O.schgn_edge_index(ds)
    sizes = (ds.n_users, ds.n_items, ds.num_ingredients, ds.num_calories_level)
    ref_top = np.stack([torch.topk(O.full_sort_scores(P, ds, u, ei, sizes), 50).indices.numpy() for u in users[:16]])
    want = M.topk_metrics(ref_top, pos[:16])
    have = M.topk_metrics(top[:16], pos[:16])
    assert have == want, (have, want)
    assert set(got) == set(want)


def test_c1_training_step_matches_oracle(c1_model, no_dropout):
    """One `calculate_loss` + backward at C1 scale against the CPU restatement (dropout = identity)."""
    from foodrec_b200.synth import sample_train_batches
    from oracle import schgn as O
    m, ds = c1_model
    batch = sample_train_batches(ds, 256, 1, seed=2, schgn=True)[0]
    P = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and k != "ingre_embed_second")
         for k, v in m.state_dict().items()}
    sizes = (ds.n_users, ds.n_items, ds.num_ingredients, ds.num_calories_level)
    ref = O.calculate_loss(P, {k: torch.from_numpy(np.asarray(v)) for k, v in batch.items()}, Cfg(CFG),
                           O.schgn_edge_index(ds), sizes)
    sum(ref).backward()
    m.zero_grad()
    got = m.calculate_loss(dev_batch(batch))
    close(torch.stack([x.reshape(()) for x in got]), np.array([float(x.detach()) for x in ref]))
    sum(got).backward()
    for name, p in m.named_parameters():
        if p.grad is None or "key.bias" in name:
            continue
        close(p.grad, P[name].grad.numpy(), rtol=5e-5, atol=1e-9)
    m.zero_grad()


def test_sample_sort_predict_matches_oracle(c1_model):
    from oracle import schgn as O
    m, ds = c1_model
    rng = np.random.default_rng(0)
    n, neg = 6, CFG["neg_sample_num"]
    u = rng.integers(0, ds.n_users, n)
    pos = rng.integers(0, ds.n_items, n)
    negs = rng.integers(0, ds.n_items, (n, neg))
    t = lambda a: torch.from_numpy(np.asarray(a)).cuda()  # noqa: E731
    batch = {"u_id": t(u), "pos_i_id": t(pos), "neg_i_id": t(negs),
             "pos_ingre_code": t(ds.ingredientCodeDict[pos]), "neg_ingre_code": t(ds.ingredientCodeDict[negs]),
             "pos_ingre_num": t(ds.ingredientNum[pos]), "neg_ingre_num": t(ds.ingredientNum[negs]),
             "pos_img": t(ds.embImage[pos]), "neg_img": t(ds.embImage[negs]),
             "pos_cl": t(ds.cal_level[pos]), "neg_cl": t(ds.cal_level[negs])}
    with torch.no_grad():
        got = m.sample_sort_predict(batch).cpu()
    assert got.shape == (n, neg + 1)
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sizes = (ds.n_users, ds.n_items, ds.num_ingredients, ds.num_calories_level)
    tables = O.gcn_tables(P, O.schgn_edge_index(ds), sizes)
    items = np.concatenate([negs, pos[:, None]], 1).reshape(-1)
    users = np.repeat(u, neg + 1)
    f = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731
    ref = O.compute_score(P, tables, f(users), f(items), f(ds.ingredientCodeDict[items]), f(ds.ingredientNum[items]),
                          f(ds.embImage[items]), f(ds.cal_level[items]),
                          torch.cat([P["ingre_embed_first"], P["ingre_embed_second"]], 0))[0].view(n, neg + 1)
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 2e-6

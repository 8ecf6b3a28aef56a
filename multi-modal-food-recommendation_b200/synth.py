"""Seeded synthetic food-recommendation datasets shaped like the reference's `FoodData`.

The reference reads Allrecipes / Foodcom from disk (FoodRec/utils/dataset.py:351-370); neither
dataset ships with it and there is no network, so every test and benchmark here runs on a
duck-typed stand-in carrying exactly the attributes the reference models consume
(FoodRec/models/cikm_model.py:20-25,49,94; pricai_modelx.py:22-29,57-63; schgn.py:50-55,141-148;
common/trainer.py:239-241,492-494).  Generation is pure numpy under `default_rng(seed)` so the same
bytes come out in this container (where goldens are made from the reference) and on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

MAX_INGRE = 20  # FoodRec/utils/dataloader.py:17 (max_len)

# Named scales from SURVEY.md section 8 (C1..C4); C5 is built directly as CSR on the device by bench.py.
SCALES = {
    "mini": dict(n_users=257, n_items=193, n_inter=2600, n_ingredients=61, n_cluster=16, dv=48, dt=24),
    "C1": dict(n_users=5000, n_items=3000, n_inter=50000, n_ingredients=800, n_cluster=200, dv=128, dt=64),
    "C2": dict(n_users=70000, n_items=45000, n_inter=1000000, n_ingredients=20000, n_cluster=2000, dv=4096, dt=384),
    "C3": dict(n_users=7600, n_items=30000, n_inter=192000, n_ingredients=5000, n_cluster=2000, dv=2048, dt=512),
}


@dataclass
class SynthConfig:
    n_users: int
    n_items: int
    n_inter: int
    n_ingredients: int
    n_cluster: int = 2000
    dv: int = 2048
    dt: int = 512
    n_cal_level: int = 80
    n_health: int = 7
    cluster_k: int = 6
    seed: int = 2024
    features: bool = True
    clusters: bool = True


class SynthFoodData:
    """Attribute bag; see module docstring for which reference code reads which field."""


def _popularity_cdf(n_items: int, alpha: float = 0.7) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
    c = np.cumsum(w)
    return c / c[-1]


def _draw_items(rng, cdf, perm, n):
    return perm[np.minimum(np.searchsorted(cdf, rng.random(n)), len(cdf) - 1)]


def topk_nearest_centres(x: np.ndarray, centres: np.ndarray, k: int, block: int = 4096) -> np.ndarray:
    """Exact k nearest (Euclidean) centres per row, nearest first, ties to the lower centre id.

    Same ranking as the notebook's per-item `argsort([norm(x - c) for c in centres])[:k]`
    (dataset_process/allrecipes_kmeans.ipynb, code cells 0-3) but computed blockwise in fp64.
    """
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(centres, dtype=np.float64)
    c2 = (c * c).sum(1)
    out = np.empty((x.shape[0], k), dtype=np.int64)
    for s in range(0, x.shape[0], block):
        xb = x[s:s + block]
        d2 = (xb * xb).sum(1)[:, None] - 2.0 * (xb @ c.T) + c2[None, :]
        out[s:s + block] = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return out


def make_dataset(cfg: SynthConfig | str, **overrides) -> SynthFoodData:
    if isinstance(cfg, str):
        cfg = SynthConfig(**{**SCALES[cfg], **overrides})
    rng = np.random.default_rng(cfg.seed)
    U, I = cfg.n_users, cfg.n_items
    ds = SynthFoodData()
    ds.cfg = cfg
    ds.n_users, ds.n_items, ds.num_items = U, I, I
    ds.num_ingredients = cfg.n_ingredients
    ds.num_calories_level = cfg.n_cal_level
    ds.num_health_level = cfg.n_health

    # ---- interactions: log-normal user degrees (min 1), power-law item popularity, de-duplicated
    mean_deg = cfg.n_inter / U
    deg = rng.lognormal(mean=0.0, sigma=1.0, size=U)
    deg = np.maximum(1, np.round(deg * (mean_deg / deg.mean()))).astype(np.int64)
    deg = np.minimum(deg, max(1, I // 2))
    users = np.repeat(np.arange(U, dtype=np.int64), deg)
    cdf = _popularity_cdf(I)
    perm = rng.permutation(I).astype(np.int64)
    items = _draw_items(rng, cdf, perm, users.shape[0])
    keys = np.unique(users * I + items)
    tr_u, tr_i = keys // I, keys % I
    # every user keeps at least one train interaction (train users are contiguous from 0,
    # FoodRec/utils/dataset.py:137-176)
    missing = np.setdiff1d(np.arange(U), tr_u)
    if missing.size:
        extra = missing * I + _draw_items(rng, cdf, perm, missing.size)
        keys = np.unique(np.concatenate([keys, extra]))
        tr_u, tr_i = keys // I, keys % I
    ds.train_coo_matrix = sp.coo_matrix(
        (np.ones(keys.shape[0], dtype=np.float32), (tr_u, tr_i)), shape=(U, I))
    ds.uRecipe_triples = np.stack([tr_u, tr_i], axis=1)
    ds.n_train = int(keys.shape[0])

    # ---- held-out positives per user (ground truth for full-sort evaluation)
    cand = _draw_items(rng, cdf, perm, U * 12).reshape(U, 12)
    ckeys = np.arange(U, dtype=np.int64)[:, None] * I + cand
    fresh = ~np.isin(ckeys, keys)
    n_valid = 1 + rng.poisson(0.8, size=U)
    n_test = 1 + rng.poisson(2.0, size=U)
    validRatings, testRatings = [], []
    for u in range(U):
        c = cand[u][fresh[u]]
        _, first = np.unique(c, return_index=True)
        c = c[np.sort(first)]
        if c.size < 2:  # pathological: fall back to any two non-train items
            row = set(tr_i[tr_u == u].tolist())
            c = np.array([j for j in range(I) if j not in row][:2], dtype=np.int64)
        nv = min(int(n_valid[u]), max(1, c.size // 2))
        validRatings.append(c[:nv].tolist())
        testRatings.append(c[nv:nv + int(n_test[u])].tolist() or c[:1].tolist())
    ds.validRatings, ds.testRatings = validRatings, testRatings

    # ---- recipe -> ingredient codes (20 wide, padded with n_ingredients), FoodRec/utils/dataset.py:51-53
    G = cfg.n_ingredients
    cnt = np.clip(np.round(rng.normal(8.7, 3.5, size=I)), 1, min(MAX_INGRE, G)).astype(np.int64)
    g_cdf = _popularity_cdf(G, 0.9)
    g_perm = rng.permutation(G).astype(np.int64)
    codes = np.full((I, MAX_INGRE), G, dtype=np.int64)
    raw = _draw_items(rng, g_cdf, g_perm, I * MAX_INGRE * 2).reshape(I, MAX_INGRE * 2)
    for i in range(I):
        _, first = np.unique(raw[i], return_index=True)
        uniq = raw[i][np.sort(first)][:cnt[i]]
        cnt[i] = uniq.size
        codes[i, :uniq.size] = uniq
    ds.ingredientCodeDict = codes
    ds.ingredientNum = cnt
    rows = np.repeat(np.arange(I, dtype=np.int64), cnt)
    ds.rIngre_triples = np.stack([rows, codes[codes != G]], axis=1)

    ds.cal_level = rng.integers(0, cfg.n_cal_level, size=I)
    ds.rCalories_triples = np.stack([np.arange(I, dtype=np.int64), ds.cal_level], axis=1)
    ds.health_level_multi_hot = (rng.random((I, cfg.n_health)) < 0.35).astype(np.float32)
    ds.health_level = rng.integers(0, 6, size=I)
    in_train = np.zeros(I, dtype=bool)
    in_train[tr_i] = True
    ds.cold_num = int((~in_train).sum())

    # ---- item modality features mixed from latent centres; centre tables double as k-means centres
    ds.embImage = ds.embText = None
    ds.image_center = ds.text_center = None
    ds.image_cluster_triples = ds.text_cluster_triples = None
    if cfg.features:
        C = cfg.n_cluster
        assign = rng.integers(0, C, size=I)
        ds.image_center = rng.standard_normal((C, cfg.dv), dtype=np.float32)
        ds.text_center = rng.standard_normal((C, cfg.dt), dtype=np.float32)
        ds.embImage = (0.7 * ds.image_center[assign]
                       + 0.7 * rng.standard_normal((I, cfg.dv), dtype=np.float32)).astype(np.float32)
        assign_t = np.where(rng.random(I) < 0.8, assign, rng.integers(0, C, size=I))
        ds.embText = (0.7 * ds.text_center[assign_t]
                      + 0.7 * rng.standard_normal((I, cfg.dt), dtype=np.float32)).astype(np.float32)
        ds.image_size = cfg.dv
        if cfg.clusters:
            k = min(cfg.cluster_k, C)
            it = np.repeat(np.arange(I, dtype=np.int64), k)
            ds.image_cluster_triples = np.stack(
                [it, topk_nearest_centres(ds.embImage, ds.image_center, k).reshape(-1)], axis=1)
            ds.text_cluster_triples = np.stack(
                [it, topk_nearest_centres(ds.embText, ds.text_center, k).reshape(-1)], axis=1)
    return ds


def masked_ingredient_task(rng, codes: np.ndarray, counts: np.ndarray, n_ingredients: int, masked_p: float = 0.2):
    """Vectorised `TrainDataLoader.ssl_task` (FoodRec/utils/dataloader.py:117-143): each real
    ingredient slot is replaced by the mask token `n_ingredients + 1` with probability `masked_p`
    and gets a random negative ingredient that is not in the recipe; other slots pass through."""
    B, L = codes.shape
    real = np.arange(L)[None, :] < counts[:, None]
    hide = real & (rng.random((B, L)) < masked_p)
    neg = rng.integers(0, n_ingredients, size=(B, L))
    for _ in range(50):
        clash = hide & (neg[:, :, None] == np.where(real, codes, -1)[:, None, :]).any(-1)
        if not clash.any():
            break
        neg[clash] = rng.integers(0, n_ingredients, size=int(clash.sum()))
    masked = np.where(hide, n_ingredients + 1, codes)
    return masked.astype(np.int64), codes.astype(np.int64), np.where(hide, neg, codes).astype(np.int64)


def sample_train_batches(ds: SynthFoodData, batch_size: int, n_batches: int, seed: int = 7, schgn: bool = False):
    """Host-side (u, pos, neg) batches with rejection-sampled negatives.

    Mirrors what `TrainDataLoader.__getitem__` + default collation hands the model
    (FoodRec/utils/dataloader.py:50-115,145-151): int64 `u_id/pos_i_id/neg_i_id` of shape [B], the
    two 20-wide ingredient code rows and counts, and the multi-hot health rows; `schgn=True` adds the
    image rows (float64, as the loader yields them), calorie levels and the masked-ingredient task.
    """
    rng = np.random.default_rng(seed)
    coo = ds.train_coo_matrix
    I = ds.n_items
    keys = np.sort(coo.row.astype(np.int64) * I + coo.col.astype(np.int64))
    out = []
    for _ in range(n_batches):
        idx = rng.integers(0, coo.nnz, size=batch_size)
        u = coo.row[idx].astype(np.int64)
        p = coo.col[idx].astype(np.int64)
        n = rng.integers(0, I, size=batch_size)
        for _try in range(50):
            bad = np.isin(u * I + n, keys)
            if not bad.any():
                break
            n[bad] = rng.integers(0, I, size=int(bad.sum()))
        out.append({
            "u_id": u, "pos_i_id": p, "neg_i_id": n,
            "pos_ingre_code": ds.ingredientCodeDict[p], "neg_ingre_code": ds.ingredientCodeDict[n],
            "pos_ingre_num": ds.ingredientNum[p], "neg_ingre_num": ds.ingredientNum[n],
            "pos_hl_mh": ds.health_level_multi_hot[p], "neg_hl_mh": ds.health_level_multi_hot[n],
        })
        if schgn:  # the extra fields SCHGN.calculate_loss reads (dataloader.py:62-66,75-80)
            m, ps, ns = masked_ingredient_task(rng, ds.ingredientCodeDict[p], ds.ingredientNum[p], ds.num_ingredients)
            out[-1].update({
                "pos_img": ds.embImage[p].astype(np.float64), "neg_img": ds.embImage[n].astype(np.float64),
                "pos_cl": ds.cal_level[p], "neg_cl": ds.cal_level[n],
                "masked_ingre_seq": m, "pos_ingre_seq": ps, "neg_ingre_seq": ns,
            })
    return out

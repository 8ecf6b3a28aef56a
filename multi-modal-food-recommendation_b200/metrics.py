"""Top-K ranking metrics with the reference's definitions, vectorised.

`topk_metrics` follows `TopKEvaluator.evaluate` (FoodRec/utils/topk_evaluator.py:68-115) and
FoodRec/common/matrics.py:9-102 (Recall, Recall2, NDCG, Precision, MAP; values rounded to 4 d.p.);
the hit matrix replaces the reference's python double loop (`i in m` over lists) by a sorted-key
membership test.  Host numpy, like the reference: it runs once per evaluation on `[n_users, 50]`.
"""
from __future__ import annotations

import numpy as np


def hit_matrix(topk_index: np.ndarray, pos_items) -> np.ndarray:
    topk_index = np.asarray(topk_index, dtype=np.int64)
    n, K = topk_index.shape
    lens = np.fromiter((len(p) for p in pos_items), dtype=np.int64, count=n)
    flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in pos_items]) if lens.sum() else np.zeros(0, np.int64)
    stride = int(max(topk_index.max(initial=0), flat.max(initial=0))) + 2
    keys = np.unique(np.repeat(np.arange(n, dtype=np.int64), lens) * stride + flat)
    q = np.arange(n, dtype=np.int64)[:, None] * stride + topk_index
    pos = np.searchsorted(keys, q)
    pos[pos >= keys.size] = max(keys.size - 1, 0)
    return (keys[pos] == q) & (topk_index >= 0) if keys.size else np.zeros_like(q, dtype=bool)


def topk_metrics(topk_index, pos_items, metrics=("recall", "ndcg", "precision", "map"), topk=(5, 10, 20, 50)):
    hits = hit_matrix(topk_index, pos_items)
    pos_len = np.fromiter((len(p) for p in pos_items), dtype=np.int64, count=hits.shape[0])
    n, K = hits.shape
    ranks = np.arange(1, K + 1)
    cum = np.cumsum(hits, axis=1)
    cut = np.minimum(pos_len, K)
    out = {}
    for m in metrics:
        m = m.lower()
        if m == "recall":
            v = (cum / pos_len.reshape(-1, 1)).mean(axis=0)
        elif m == "recall2":
            v = cum.sum(axis=0) / pos_len.sum()
        elif m == "precision":
            v = (cum / ranks).mean(axis=0)
        elif m == "ndcg":
            disc = (1.0 / np.log2(ranks.astype(np.float32) + 1)).astype(np.float32)
            ideal = np.cumsum(disc)
            idcg = ideal[np.minimum(np.arange(K)[None, :], (cut - 1)[:, None])]
            dcg = np.cumsum(np.where(hits, disc, 0), axis=1)
            v = (dcg / idcg).mean(axis=0)
        elif m == "map":
            pre = cum / ranks
            sum_pre = np.cumsum(pre * hits.astype(np.float32), axis=1)
            den = np.minimum(ranks[None, :], cut[:, None])
            v = (sum_pre / den).astype(np.float32).mean(axis=0)
        else:
            raise ValueError(f"There is no user grouped topk metric named {m}!")
        for k in topk:
            out[f"{m}@{k}"] = round(float(v[k - 1]), 4)
    return out


def metrics_by_user(doc_list, rel_list):
    """Recall and NDCG of one ranked list (FoodRec/common/trainer.py:55-69)."""
    rel = set(rel_list)
    gains = [1.0 / np.log2(i + 2) for i, d in enumerate(doc_list) if d in rel]
    idcg = sum(1.0 / np.log2(i + 2) for i in range(min(len(doc_list), len(rel_list))))
    return len(gains) / len(rel_list), float(sum(gains)) / idcg


def by_user_metrics(scores: np.ndarray, ptr: np.ndarray, n_pos: np.ndarray, neg_num: int = 500):
    """The reference's by-user evaluation (FoodRec/common/trainer.py:49-69,231-282) on a flat score
    array: user `r` owns `scores[ptr[r]:ptr[r+1]]`, its first `n_pos[r]` entries are the positives.
    Returns {'AUC', 'Recall@10', 'Recall@20', 'NDCG@10', 'NDCG@20'} averaged over users."""
    disc = 1.0 / np.log2(np.arange(2, 22))
    rows = np.zeros((len(n_pos), 5))
    for r in range(len(n_pos)):
        pred = scores[ptr[r]:ptr[r + 1]]
        p = int(n_pos[r])
        order = np.argsort(pred)[::-1]
        neg = pred[p:]
        rows[r, 0] = sum(float(np.sum(neg < pred[i])) for i in range(p)) / (p * neg_num)
        hit = order[:20] < p
        for j, k in enumerate((10, 20)):
            h = hit[:k]
            rows[r, 1 + j] = h.sum() / p
            rows[r, 3 + j] = (disc[:len(h)] * h).sum() / disc[:min(len(h), p)].sum()
    m = rows.mean(axis=0)
    return {"AUC": m[0], "Recall@10": m[1], "Recall@20": m[2], "NDCG@10": m[3], "NDCG@20": m[4]}

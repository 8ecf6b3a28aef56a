"""Full-sort scoring, top-K and ranking metrics (oracle).  Test infrastructure only."""
import math

import numpy as np
import torch


def full_sort_scores(user_all, item_all, users):
    """Dense fp32 `user_all[u] @ item_all^T` -- what `inference_fast` computes restricted to
    candidates (FoodRec/models/cikm_model.py:294-302), taken over all items."""
    return user_all[users] @ item_all.t()


def full_sort_topk(user_all, item_all, users, k, hist_ptr=None, hist_idx=None):
    """`torch.topk(scores, k, dim=-1)` as `Trainer.evaluate` does (FoodRec/common/trainer.py:495-497).
    With a history CSR (`hist_ptr`/`hist_idx` over ALL users) the user's training items are set to
    -inf first -- the MMRec behaviour the north star asks for; the reference itself applies no mask
    (SURVEY.md D1), which is `hist_ptr=None`."""
    s = full_sort_scores(user_all, item_all, users).clone()
    if hist_ptr is not None:
        for r, u in enumerate(users.tolist()):
            s[r, hist_idx[hist_ptr[u]:hist_ptr[u + 1]]] = -float("inf")
    return torch.topk(s, k, dim=-1)


def inference_scores(user_all, item_all, user_input, item_input):
    """FoodRec/models/cikm_model.py:294-302 / pricai_modelx.py:278-286."""
    return (user_all[user_input] * item_all[item_input]).sum(1)


def hit_matrix(topk_index, pos_items):
    """FoodRec/utils/topk_evaluator.py:103-106."""
    return np.asarray([[int(i) in set(m) for i in row] for m, row in zip(pos_items, topk_index)])


def topk_metrics(topk_index, pos_items, metrics=("recall", "ndcg", "precision", "map"),
                 topk=(5, 10, 20, 50)):
    """`TopKEvaluator.evaluate` restated (FoodRec/utils/topk_evaluator.py:68-115) on top of the
    metric definitions in FoodRec/common/matrics.py:9-102.  Values rounded to 4 d.p. as there."""
    hits = hit_matrix(topk_index, pos_items)
    pos_len = np.asarray([len(m) for m in pos_items])
    n, K = hits.shape
    ranks = np.arange(1, K + 1)
    cum = np.cumsum(hits, axis=1)
    res = {}
    for m in metrics:
        if m == "recall":
            v = (cum / pos_len.reshape(-1, 1)).mean(0)
        elif m == "recall2":
            v = cum.sum(0) / pos_len.sum()
        elif m == "precision":
            v = (cum / ranks).mean(0)
        elif m == "ndcg":
            disc = (1.0 / np.log2(ranks.astype(np.float32) + 1)).astype(np.float32)
            idcg = np.tile(np.cumsum(disc), (n, 1))
            cut = np.minimum(pos_len, K)
            for r, c in enumerate(cut):
                idcg[r, c:] = idcg[r, c - 1]
            dcg = np.cumsum(np.where(hits, disc, 0), axis=1)
            v = (dcg / idcg).mean(0)
        elif m == "map":
            pre = cum / ranks
            sum_pre = np.cumsum(pre * hits.astype(np.float32), axis=1)
            cut = np.minimum(pos_len, K)
            out = np.zeros((n, K), dtype=np.float32)
            for r, c in enumerate(cut):
                rg = ranks.copy()
                rg[c:] = rg[c - 1]
                out[r] = sum_pre[r] / rg
            v = out.mean(0)
        else:
            raise ValueError(m)
        for k in topk:
            res[f"{m}@{k}"] = round(float(v[k - 1]), 4)
    return res


def metrics_by_user(doc_list, rel_list):
    """FoodRec/common/trainer.py:55-69."""
    dcg = hit = 0.0
    rel = set(rel_list)
    for i, d in enumerate(doc_list):
        if d in rel:
            dcg += 1 / (math.log(i + 2) / math.log(2))
            hit += 1
    idcg = sum(1 / (math.log(i + 2) / math.log(2)) for i in range(min(len(doc_list), len(rel_list))))
    return hit / len(rel_list), dcg / idcg


def auc_fast(n_pos, predictions, neg_num):
    """FoodRec/common/trainer.py:49-52 (positives occupy the first `n_pos` slots)."""
    neg = predictions[n_pos:]
    return float(sum(np.sum(neg < predictions[i]) for i in range(n_pos))) / (n_pos * neg_num)


def by_user_eval(scores_per_user, n_pos_per_user, neg_num=500):
    """`_valid_by_user_epoch` restated (FoodRec/common/trainer.py:231-282): argsort descending,
    AUC, recall/ndcg at 10 and 20, mean over users."""
    rows = []
    for pred, n_pos in zip(scores_per_user, n_pos_per_user):
        order = np.argsort(pred)[::-1]
        auc = auc_fast(n_pos, pred, neg_num)
        rec, nd = [], []
        for k in (10, 20):
            r, g = metrics_by_user(order[:k].tolist(), list(range(n_pos)))
            rec.append(r)
            nd.append(g)
        rows.append((rec, nd, [auc, auc]))
    rec, nd, auc = np.array(rows).mean(0).tolist()
    return {"AUC": auc[0], "Recall@10": rec[0], "Recall@20": rec[1], "NDCG@10": nd[0], "NDCG@20": nd[1]}

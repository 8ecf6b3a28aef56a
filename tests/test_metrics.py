"""CPU: the vectorised product metrics equal the reference's TopKEvaluator (goldens) and the oracle."""
import numpy as np

from conftest import load_golden

import foodrec_b200  # noqa: F401
from foodrec_b200 import metrics
from oracle import ranking


def test_metrics_match_reference_evaluator():
    g = load_golden("primitives.npz")
    ptr, idx = g["rank/pos_ptr"], g["rank/pos_idx"]
    pos = [idx[ptr[u]:ptr[u + 1]].tolist() for u in range(len(ptr) - 1)]
    res = metrics.topk_metrics(g["rank/topi"], pos)
    for k, v in zip(g["rank/metric_keys"], g["rank/metric_vals"]):
        assert res[str(k)] == v, (k, res[str(k)], v)
    toy = metrics.topk_metrics(np.array([[4, 1, 7], [0, 2, 9], [5, 6, 3], [8, 8, 1]]),
                               [[1], [9, 0], [2], [1, 8, 4]], topk=(1, 3))
    for k, v in zip(g["rank/toy_keys"], g["rank/toy_vals"]):
        assert toy[str(k)] == v, (k, toy[str(k)], v)
    r, n = metrics.metrics_by_user([3, 0, 9, 1, 7], [0, 1])
    assert abs(r - g["rank/by_user"][0]) < 1e-12 and abs(n - g["rank/by_user"][1]) < 1e-12


def test_metrics_random_vs_oracle():
    rng = np.random.default_rng(0)
    for _ in range(5):
        n, I = 200, 500
        top = np.stack([rng.permutation(I)[:50] for _ in range(n)])
        pos = [rng.choice(I, size=int(rng.integers(1, 70)), replace=False).tolist() for _ in range(n)]
        a = metrics.topk_metrics(top, pos, metrics=("recall", "recall2", "ndcg", "precision", "map"))
        b = ranking.topk_metrics(top, pos, metrics=("recall", "recall2", "ndcg", "precision", "map"))
        assert a == b
        assert np.array_equal(metrics.hit_matrix(top, pos), ranking.hit_matrix(top, pos))


def test_by_user_metrics_match_oracle():
    rng = np.random.default_rng(1)
    n_pos = rng.integers(1, 6, size=30)
    lens = n_pos + 50
    ptr = np.concatenate([[0], np.cumsum(lens)])
    scores = rng.standard_normal(int(ptr[-1])).astype(np.float32)
    got = metrics.by_user_metrics(scores, ptr, n_pos, neg_num=50)
    ref = ranking.by_user_eval([scores[ptr[r]:ptr[r + 1]] for r in range(30)], n_pos.tolist(), neg_num=50)
    for k in ref:
        assert abs(got[k] - ref[k]) < 1e-12, (k, got[k], ref[k])

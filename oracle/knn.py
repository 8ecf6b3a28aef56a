"""Cosine kNN graph and centroid-assignment top-k (oracle).  Test infrastructure only."""
import numpy as np
import torch


def build_sim(x):
    """FoodRec/utils/utils.py:132-135 (no eps in the norm)."""
    xn = x / torch.norm(x, p=2, dim=-1, keepdim=True)
    return xn @ xn.t()


def knn_neighbourhood(adj, k):
    """FoodRec/utils/utils.py:118-121: row top-k (self included), scattered into a dense matrix."""
    val, ind = torch.topk(adj, k, dim=-1)
    return torch.zeros_like(adj).scatter_(-1, ind, val), val, ind


def normalized_laplacian(adj):
    """FoodRec/utils/utils.py:124-130 (= `get_dense_laplacian(.., 'sym')`, :153-159)."""
    d = torch.pow(adj.sum(-1), -0.5)
    d[torch.isinf(d)] = 0.0
    return d[:, None] * adj * d[None, :]


def sparse_laplacian(row, col, w, n, normalization="sym"):
    """FoodRec/utils/utils.py:138-151 with `torch_scatter.scatter_add` restated as `index_add_`."""
    deg = torch.zeros(n, dtype=w.dtype).index_add_(0, row, w)
    if normalization == "sym":
        dis = deg.pow(-0.5)
        dis[dis == float("inf")] = 0
        return dis[row] * w * dis[col]
    if normalization == "rw":
        di = 1.0 / deg
        di[di == float("inf")] = 0
        return di[row] * w
    return w


def knn_normalized_graph(adj, k, norm_type="sym"):
    """Sparse branch of `build_knn_normalized_graph`, FoodRec/utils/utils.py:170-180: returns
    (row, col, weight) with row-major order, `k` entries per row in top-k order."""
    val, ind = torch.topk(adj, k, dim=-1)
    n = adj.shape[0]
    row = torch.arange(n).repeat_interleave(k)
    col = ind.reshape(-1)
    return row, col, sparse_laplacian(row, col, val.reshape(-1), n, norm_type)


def centroid_topk(x, centres, k):
    """Per item `argsort([norm(x - c) for c in centres])[:10][:6]` in fp64 numpy.
    dataset_process/allrecipes_kmeans.ipynb code cells 0-3 (the python loop itself for small inputs)."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(centres, dtype=np.float64)
    out = np.empty((x.shape[0], k), dtype=np.int64)
    for i in range(x.shape[0]):
        d = np.linalg.norm(x[i][None, :] - c, axis=1)
        out[i] = np.argsort(d)[:10][:k]
    return out

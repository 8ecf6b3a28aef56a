"""Normalised adjacency construction (oracle).  Test infrastructure only."""
import numpy as np
import scipy.sparse as sp
import torch


def _sym_normalise(A: sp.spmatrix):
    """`D^-1/2 A D^-1/2` with `D = rowcount(A>0) + 1e-7`, returned as a torch sparse COO fp32.

    Follows FoodRec/models/cikm_model.py:166-180 (identical in pricai_modelx.py:163-177,
    lightgcn.py:106-120): integer degree from `(A > 0).sum(axis=1)`, `+1e-7` promotes to fp64,
    `np.power(., -0.5)`, `D * A * D` in scipy (fp64), `coo_matrix(L)` row-major, values cast to fp32
    by `torch.FloatTensor`.  The reference fills a `dok_matrix` through a dict first; the entries
    are 0/1 with duplicates collapsed, which is what the CSR built here holds.
    """
    A = sp.csr_matrix(A, dtype=np.float32)
    A.sum_duplicates()
    A.data[:] = 1.0
    deg = np.asarray((A > 0).sum(axis=1)).reshape(-1) + 1e-7
    d = np.power(deg, -0.5)
    D = sp.diags(d)
    L = sp.coo_matrix(D * A * D)
    idx = torch.from_numpy(np.stack([L.row, L.col]).astype(np.int64))
    val = torch.from_numpy(L.data.astype(np.float32))
    return torch.sparse_coo_tensor(idx, val, L.shape)


def norm_adj_user_item(train_coo: sp.coo_matrix, n_users: int, n_items: int):
    """`[[0,R],[R^T,0]]` normalised.  FoodRec/models/cikm_model.py:136-180."""
    R = sp.csr_matrix(train_coo, dtype=np.float32)
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    assert A.shape == (n_users + n_items, n_users + n_items)
    return _sym_normalise(A)


def norm_adj_item_side(triples: np.ndarray, n_items: int, n_side: int):
    """Item/side-node graph: edge `(side + n_items, item)` symmetrised, then normalised.

    FoodRec/models/cikm_model.py:91-134 and pricai_modelx.py:88-131 (`load_graph` +
    `get_norm_adj_recipe_*`); `triples[:, 0]` is the item id, `triples[:, 1]` the side-node id.
    """
    t = np.asarray(triples, dtype=np.int64)
    n = n_items + n_side
    r, c = t[:, 1] + n_items, t[:, 0]
    M = sp.coo_matrix((np.ones(len(t), dtype=np.float32), (r, c)), shape=(n, n))
    return _sym_normalise(M + M.T)


def gcn_norm_edges(edge_index: torch.Tensor, n_nodes: int):
    """PyG `GCNConv` default normalisation (add_self_loops=True, improved=False, flow
    source->target): one unit self-loop per node, `deg[t] = sum of weights into t`,
    `w_e = deg[s]^-1/2 * deg[t]^-1/2` with inf -> 0.  Call site FoodRec/models/schgn.py:34,39,247;
    semantics from torch_geometric.nn.conv.gcn_conv.gcn_norm (not installed; parity unpinned).
    Returns (src, dst, w) including the self-loops appended last.
    """
    src, dst = edge_index[0].long(), edge_index[1].long()
    loops = torch.arange(n_nodes, dtype=torch.long)
    src = torch.cat([src, loops])
    dst = torch.cat([dst, loops])
    w = torch.ones(src.numel(), dtype=torch.float32)
    deg = torch.zeros(n_nodes, dtype=torch.float32).index_add_(0, dst, w)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0.0
    return src, dst, dis[src] * w * dis[dst]


def schgn_edge_index(ds):
    """Directed heterogeneous edge list of SCHGN, row 0 = source, row 1 = target.

    FoodRec/models/schgn.py:139-151 builds rows `[item+U, user]`, `[ingre+U+I, item+U]`,
    `[cal+U+I+G, item+U]`; :241,247 concatenates them and transposes, so column 0 is the source.
    """
    U, I, G = ds.n_users, ds.n_items, ds.num_ingredients
    ur = np.asarray(ds.uRecipe_triples, dtype=np.int64)
    ri = np.asarray(ds.rIngre_triples, dtype=np.int64)
    rc = np.asarray(ds.rCalories_triples, dtype=np.int64)
    src = np.concatenate([ur[:, 1] + U, ri[:, 1] + U + I, rc[:, 1] + U + I + G])
    dst = np.concatenate([ur[:, 0], ri[:, 0] + U, rc[:, 0] + U])
    return torch.from_numpy(np.stack([src, dst]))

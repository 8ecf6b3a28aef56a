"""Normalised adjacencies as device-resident CSR + segment plans (built once per model).

Host-side construction is vectorised numpy with the reference's exact arithmetic
(FoodRec/models/cikm_model.py:108-180; pricai_modelx.py:105-177; lightgcn.py:76-120):
degree = count of stored entries per row, `d = (deg + 1e-7) ** -0.5` in fp64, value
`fp32((d[r] * 1.0) * d[c])`.  The reference fills a `dok_matrix` through a python dict and
multiplies scipy matrices on every model construction; the values produced here are bit-identical
(tests/test_host_graph.py checks that against the committed goldens).

`PropGraph` holds what `fr_spmm_csr_f32` consumes: `col`/`val` (CSR payload), the segment plan
(`seg`, `long_rows`) and the long-row workspace.  For the symmetric graphs `graph.T is graph`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

SEG = int(os.environ.get("FR_SPMM_SEG", "64"))  # nonzeros per segment (<= FR_SPMM_SEG = 128)


class PropGraph:
    def __init__(self, row_ptr: np.ndarray, col: np.ndarray, val: np.ndarray, n_cols: int, device,
                 transpose: "PropGraph | None | str" = None):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.n_rows = int(row_ptr.shape[0] - 1)
        self.n_cols = int(n_cols)
        self.nnz = int(row_ptr[-1])
        self.device = torch.device(device)
        self.row_ptr_host = row_ptr
        n_seg, n_long, n_part = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib.fr_spmm_plan_sizes(row_ptr.ctypes.data, self.n_rows, SEG, C.byref(n_seg), C.byref(n_long),
                                               C.byref(n_part)), "fr_spmm_plan_sizes")
        self.n_seg, self.n_long, self.n_part = n_seg.value, n_long.value, n_part.value
        seg = np.empty((max(self.n_seg, 1), 4), dtype=np.int32)
        lrows = np.empty((max(self.n_long, 1), 4), dtype=np.int32)
        _lib.check(_lib.lib.fr_spmm_plan_fill(row_ptr.ctypes.data, self.n_rows, SEG, seg.ctypes.data, lrows.ctypes.data),
                   "fr_spmm_plan_fill")
        self.seg_host, self.long_rows_host = seg, lrows
        dev = self.device
        # payload: host arrays are uploaded once; device tensors (a graph assembled on the GPU) are adopted as they are
        self.col = (col.to(dev, torch.int32).contiguous() if torch.is_tensor(col)
                    else torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)).to(dev))
        self.val = (val.to(dev, torch.float32).contiguous() if torch.is_tensor(val)
                    else torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(dev))
        self.seg = torch.from_numpy(seg).to(dev)
        self.long_rows = torch.from_numpy(lrows).to(dev)
        self.counters = torch.zeros(max(self.n_long, 1), dtype=torch.int32, device=dev)
        self._partial = {}
        self.T = self if transpose == "self" else transpose

    @classmethod
    def from_plan(cls, row_ptr, col, val, n_cols: int, device, seg, long_rows, counts, symmetric: bool = False):
        """Adopt a stored segment plan (`cache.load_graph`) instead of rebuilding it from the row pointers."""
        self = cls.__new__(cls)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.n_rows, self.n_cols, self.nnz = int(row_ptr.shape[0] - 1), int(n_cols), int(row_ptr[-1])
        self.device = dev = torch.device(device)
        self.row_ptr_host = row_ptr
        self.n_seg, self.n_long, self.n_part = (int(c) for c in counts)
        self.seg_host = np.ascontiguousarray(seg, dtype=np.int32).reshape(-1, 4)
        self.long_rows_host = np.ascontiguousarray(long_rows, dtype=np.int32).reshape(-1, 4)
        if self.seg_host.shape[0] < max(self.n_seg, 1) or self.long_rows_host.shape[0] < max(self.n_long, 1):
            raise _lib.FoodRecError("stored segment plan is shorter than its counts")
        def up(a, dt):           # (memory-mapped inputs are read-only: copy before handing them to torch)
            return torch.from_numpy(np.array(a, dtype=dt, copy=True)).to(dev)
        self.col, self.val = up(col, np.int32), up(val, np.float32)
        self.seg, self.long_rows = up(self.seg_host, np.int32), up(self.long_rows_host, np.int32)
        self.counters = torch.zeros(max(self.n_long, 1), dtype=torch.int32, device=dev)
        self._partial = {}
        self.T = self if symmetric else None
        return self

    def partial(self, d: int) -> torch.Tensor:
        buf = self._partial.get(d)
        if buf is None:
            buf = torch.empty(max(self.n_part, 1) * d, dtype=torch.float32, device=self.device)
            self._partial[d] = buf
        return buf

    def spmm_bytes(self, d: int, with_z: bool = False) -> int:
        """Algorithmic bytes of one launch: CSR payload + X read once + Y written once."""
        b = 8 * self.nnz + 4 * (self.n_rows + 1) + 4 * d * self.n_cols + 4 * d * self.n_rows
        return b + (4 * d * self.n_rows if with_z else 0)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.row_ptr_host),
                             shape=(self.n_rows, self.n_cols))


def _csr_from_pairs(rows: np.ndarray, cols: np.ndarray, n: int):
    """Sorted, de-duplicated CSR structure of a 0/1 matrix given (row, col) pairs."""
    key = np.unique(rows.astype(np.int64) * n + cols.astype(np.int64))
    r = (key // n).astype(np.int64)
    c = (key % n).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=row_ptr[1:])
    return row_ptr, r, c


def _sym_norm_values(row_ptr, r, c):
    deg = np.diff(row_ptr).astype(np.int64) + 1e-7          # (A > 0).sum(axis=1) + 1e-7  -> fp64
    d = np.power(deg, -0.5)
    return ((d[r] * np.float64(1.0)) * d[c]).astype(np.float32)  # D * A * D, then FloatTensor cast


def symmetric_normalised(rows, cols, n, device) -> PropGraph:
    """`D^-1/2 (M + M^T) D^-1/2` for the 0/1 pattern M given by (rows, cols)."""
    rr = np.concatenate([rows, cols])
    cc = np.concatenate([cols, rows])
    row_ptr, r, c = _csr_from_pairs(rr, cc, n)
    if row_ptr[-1] >= 2 ** 31:
        raise _lib.FoodRecError("graph has >= 2^31 stored entries; int32 CSR offsets would overflow")
    return PropGraph(row_ptr, c, _sym_norm_values(row_ptr, r, c), n, device, transpose="self")


def symmetric_normalised_device(rows: torch.Tensor, cols: torch.Tensor, n: int) -> PropGraph:
    """`symmetric_normalised` with the edge list already on the GPU (int64 tensors): de-duplication, degrees and
    the `(deg + 1e-7)^-1/2` products (fp64, cast to fp32 -- the reference's arithmetic, cikm_model.py:136-180) run
    as device ops and only the row pointers (4 (n + 1) bytes) visit the host for the segment plan.  Used for
    graphs whose host-side construction would take minutes (C5, 400 M stored entries)."""
    dev = rows.device
    key = torch.unique(torch.cat([rows * n + cols, cols * n + rows]))           # sorted, duplicate-free
    if key.numel() >= 2 ** 31:
        raise _lib.FoodRecError("graph has >= 2^31 stored entries; int32 CSR offsets would overflow")
    r = torch.div(key, n, rounding_mode="floor")
    c = key - r * n
    del key
    deg = torch.bincount(r, minlength=n)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=row_ptr[1:])
    d = (deg.to(torch.float64) + 1e-7).pow(-0.5)
    val = ((d[r] * 1.0) * d[c]).to(torch.float32)
    return PropGraph(row_ptr.cpu().numpy(), c.to(torch.int32), val, n, dev, transpose="self")


def norm_adj_user_item(train_coo, n_users: int, n_items: int, device) -> PropGraph:
    """User-item bipartite graph `[[0,R],[R^T,0]]`, normalised.  cikm_model.py:136-180."""
    u = np.asarray(train_coo.row, dtype=np.int64)
    i = np.asarray(train_coo.col, dtype=np.int64) + n_users
    return symmetric_normalised(u, i, n_users + n_items, device)


def norm_adj_item_side(triples, n_items: int, n_side: int, device) -> PropGraph:
    """Item / side-node (ingredient or cluster) graph; `triples[:,0]` item, `triples[:,1]` side id.
    cikm_model.py:91-134, pricai_modelx.py:88-131."""
    t = np.asarray(triples, dtype=np.int64)
    return symmetric_normalised(t[:, 1] + n_items, t[:, 0], n_items + n_side, device)


def gcn_normalised(src, dst, n_nodes: int, device) -> PropGraph:
    """PyG `GCNConv` normalisation of a directed edge list (source -> target) with one unit
    self-loop per node; duplicates are kept (they add).  Rows of the returned CSR are TARGETS:
    `out[t] = sum_e w_e x[s_e]`; `.T` is the CSR of the transpose for the backward.
    Call site FoodRec/models/schgn.py:34,39,139-151,241-250."""
    src = np.concatenate([np.asarray(src, dtype=np.int64), np.arange(n_nodes, dtype=np.int64)])
    dst = np.concatenate([np.asarray(dst, dtype=np.int64), np.arange(n_nodes, dtype=np.int64)])
    deg = np.bincount(dst, minlength=n_nodes).astype(np.float32)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, np.float32(-0.5), dtype=np.float32)
    dis[np.isinf(dis)] = 0.0
    w = (dis[src] * np.float32(1.0) * dis[dst]).astype(np.float32)

    def build(rows, cols):
        order = np.lexsort((cols, rows))
        rp = np.zeros(n_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n_nodes), out=rp[1:])
        return rp, cols[order].astype(np.int32), w[order]

    fwd = PropGraph(*build(dst, src), n_nodes, device)
    bwd = PropGraph(*build(src, dst), n_nodes, device)
    fwd.T, bwd.T = bwd, fwd
    return fwd


def from_torch_sparse(S: torch.Tensor, symmetric: bool = True) -> PropGraph:
    """PropGraph from a reference-built adjacency: a row-major-sorted, duplicate-free
    `torch.sparse` COO on the device (what `get_norm_adj_mat` returns after `.to(device)`,
    FoodRec/models/cikm_model.py:74,174-180).  Row pointers are computed on the device
    (`fr_csr_from_coo`); values and column order are taken as they are."""
    idx, val = S._indices(), S._values()
    if idx.device.type != "cuda":
        raise _lib.FoodRecError("from_torch_sparse needs a CUDA sparse tensor")
    n = int(S.shape[0])
    rows = idx[0].contiguous()
    if rows.numel() > 1 and bool((rows[1:] < rows[:-1]).any()):
        raise _lib.FoodRecError("COO rows are not sorted; coalesce() first")
    row_ptr = torch.empty(n + 1, dtype=torch.int32, device=idx.device)
    scratch = torch.empty(max(n, 1), dtype=torch.int32, device=idx.device)
    _lib.check(_lib.lib.fr_csr_from_coo(rows.data_ptr(), rows.numel(), n, row_ptr.data_ptr(), scratch.data_ptr(),
                                        _lib.stream_ptr()), "fr_csr_from_coo")
    return PropGraph(row_ptr.cpu().numpy(), idx[1].to(torch.int32).cpu().numpy(), val.float().cpu().numpy(), int(S.shape[1]),
                     idx.device, transpose="self" if symmetric else None)

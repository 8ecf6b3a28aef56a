"""Binary on-disk formats for the graphs of the path.

The reference re-parses its graphs from text on every start: `np.loadtxt` over the edge files
(FoodRec/utils/dataset.py:325-343: `ri_graph.txt`, `image_cluster_edge.txt`, `text_cluster_edge.txt`, ...; a python
tokeniser, ~1 us per character) and un-pickles a scipy COO for the interactions (dataset.py:88-91), then every model
rebuilds its normalised adjacency through python dicts and a `dok_matrix` (cikm_model.py:91-180).  Here:

* `save_arrays` / `load_arrays`: one little-endian container file -- magic, JSON directory (name, dtype, shape, offset),
  64-byte-aligned raw arrays -- read back with `np.memmap`, so loading is a page-in, not a parse.
* `load_edge_triples(txt)`: the drop-in for `np.loadtxt(path)` on an edge file: parses the text once with the C
  tokeniser of pandas, writes `<txt>.frcache` next to it and afterwards serves the cache while the text file's size
  and mtime are unchanged.  Same values and the same dtype as the reference call (`np.loadtxt` returns float64 unless
  the caller passes `dtype=np.int_`, dataset.py:342).
* `save_interactions` / `load_interactions`: the pickled COO of dataset.py:88-91 as row / col / shape arrays; the
  loader returns the same `scipy.sparse.coo_matrix` (float32 ones).
* `save_graph` / `load_graph`: a built `PropGraph` (normalised CSR + segment plan + its transpose link) -- a model
  start-up uploads five arrays instead of re-deriving degrees, values and the plan.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

MAGIC = b"FRB200\x01\x00"
_ALIGN = 64


def save_arrays(path: str, meta: dict | None = None, **arrays) -> None:
    """Write `arrays` (name -> ndarray) and a small JSON `meta` dict to `path` atomically."""
    entries, off = [], 0
    arrs = {}
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        arrs[name] = a
        entries.append({"name": name, "dtype": a.dtype.str, "shape": list(a.shape), "offset": off, "nbytes": int(a.nbytes)})
        off += -(-a.nbytes // _ALIGN) * _ALIGN
    head = json.dumps({"meta": meta or {}, "arrays": entries}).encode()
    data0 = -(-(len(MAGIC) + 8 + len(head)) // _ALIGN) * _ALIGN
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", len(head), data0))
        f.write(head)
        for e in entries:
            f.seek(data0 + e["offset"])
            f.write(arrs[e["name"]].tobytes())
        f.truncate(data0 + off)
    os.replace(tmp, path)


def load_arrays(path: str, mmap: bool = True):
    """-> (arrays: dict name -> ndarray (memory-mapped, read-only, when `mmap`), meta: dict)."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a foodrec_b200 array container")
        n_head, data0 = struct.unpack("<II", f.read(8))
        head = json.loads(f.read(n_head).decode())
    out = {}
    for e in head["arrays"]:
        shape, dt = tuple(e["shape"]), np.dtype(e["dtype"])
        if e["nbytes"] == 0:
            out[e["name"]] = np.empty(shape, dtype=dt)
        elif mmap:
            out[e["name"]] = np.memmap(path, dtype=dt, mode="r", offset=data0 + e["offset"], shape=shape)
        else:
            out[e["name"]] = np.fromfile(path, dtype=dt, count=int(np.prod(shape)), offset=data0 + e["offset"]).reshape(shape)
    return out, head["meta"]


# ------------------------------------------------------------------------------------------ edge files
def _parse_edge_text(path: str, dtype):
    import pandas as pd
    df = pd.read_csv(path, sep=r"\s+", header=None, comment="#", dtype=np.float64 if np.dtype(dtype).kind == "f" else None,
                     engine="c")
    a = df.to_numpy()
    if a.shape[1] == 1:          # np.loadtxt squeezes a single column
        a = a[:, 0]
    return np.ascontiguousarray(a.astype(dtype, copy=False))


def load_edge_triples(path: str, dtype=np.float64, cache: bool = True, refresh: bool = False) -> np.ndarray:
    """`np.loadtxt(path, dtype=dtype)` for a whitespace-separated numeric edge file, served from `<path>.frcache` when
    that exists and was written for the file as it is now (same size and mtime)."""
    st = os.stat(path)
    stamp = {"size": int(st.st_size), "mtime_ns": int(st.st_mtime_ns), "dtype": np.dtype(dtype).str}
    cpath = path + ".frcache"
    if cache and not refresh and os.path.exists(cpath):
        try:
            arrs, meta = load_arrays(cpath)
            if meta.get("source") == stamp:
                return arrs["triples"]
        except (ValueError, KeyError, OSError):
            pass
    a = _parse_edge_text(path, dtype)
    if cache:
        try:
            save_arrays(cpath, meta={"source": stamp}, triples=a)
        except OSError:
            pass                 # read-only data directory: parse every time
    return a


# -------------------------------------------------------------------------------------- interactions
def save_interactions(path: str, coo) -> None:
    coo = coo.tocoo()
    save_arrays(path, meta={"shape": [int(coo.shape[0]), int(coo.shape[1])]}, row=np.asarray(coo.row, dtype=np.int32),
                col=np.asarray(coo.col, dtype=np.int32), data=np.asarray(coo.data, dtype=np.float32))


def load_interactions(path: str):
    """The `train_coo_matrix` of FoodRec/utils/dataset.py:88-91 (`pickle.load(f).astype(np.float32)`)."""
    import scipy.sparse as sp
    a, meta = load_arrays(path)
    return sp.coo_matrix((np.asarray(a["data"]), (np.asarray(a["row"]), np.asarray(a["col"]))), shape=tuple(meta["shape"]),
                         dtype=np.float32)


def interactions_from_pickle(pickle_path: str, cache_path: str | None = None):
    """Load the reference's pickled COO once and leave the binary form beside it."""
    import pickle
    cache_path = cache_path or pickle_path + ".frcache"
    if os.path.exists(cache_path) and os.stat(cache_path).st_mtime_ns >= os.stat(pickle_path).st_mtime_ns:
        return load_interactions(cache_path)
    with open(pickle_path, "rb") as f:
        coo = pickle.load(f).astype(np.float32)
    try:
        save_interactions(cache_path, coo)
    except OSError:
        pass
    return coo


# --------------------------------------------------------------------------------------------- graphs
# Layout version of the stored segment plan (`fr_spmm_plan_fill`): 3 = long-row segments carry their fold slot, two-level
# fold entries.  A file written with another layout is rejected (the caller rebuilds the graph and saves it again).
PLAN_VERSION = 3


def save_graph(path: str, g) -> None:
    """A built `graph.PropGraph`: CSR payload, segment plan and whether it is its own transpose."""
    if g.T is not None and g.T is not g:
        raise ValueError("save_graph stores symmetric graphs (g.T is g) or graphs without a transpose; save g.T separately")
    save_arrays(path, meta={"n_rows": g.n_rows, "n_cols": g.n_cols, "nnz": g.nnz, "n_seg": g.n_seg, "n_long": g.n_long,
                            "n_part": g.n_part, "symmetric": g.T is g, "plan_version": PLAN_VERSION},
                row_ptr=g.row_ptr_host, col=g.col.cpu().numpy(), val=g.val.cpu().numpy(), seg=g.seg_host,
                long_rows=g.long_rows_host)


def load_graph(path: str, device):
    """-> `graph.PropGraph` with the stored plan adopted as is (no degree pass, no plan build)."""
    from .graph import PropGraph
    a, meta = load_arrays(path)
    if meta.get("plan_version") != PLAN_VERSION:
        raise ValueError(f"{path}: segment plan layout {meta.get('plan_version')} != {PLAN_VERSION}; rebuild the graph and save it again")
    return PropGraph.from_plan(a["row_ptr"], a["col"], a["val"], meta["n_cols"], device, a["seg"], a["long_rows"],
                               (meta["n_seg"], meta["n_long"], meta["n_part"]), symmetric=bool(meta["symmetric"]))

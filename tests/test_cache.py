"""CPU tests of the binary graph / edge / interaction caches (FoodRec/utils/dataset.py:88-91, 325-343)."""
import os
import pickle
import time

import numpy as np
import pytest

import foodrec_b200  # noqa: F401
from foodrec_b200 import cache


def test_array_container_round_trip(tmp_path):
    p = str(tmp_path / "a.frb")
    arrs = dict(i32=np.arange(7, dtype=np.int32), f32=np.random.default_rng(0).standard_normal((5, 3)).astype(np.float32),
                empty=np.empty((0, 4), dtype=np.int64), f64=np.array([1.5, -2.25]))
    cache.save_arrays(p, meta={"k": 3}, **arrs)
    for mmap in (True, False):
        got, meta = cache.load_arrays(p, mmap=mmap)
        assert meta == {"k": 3} and set(got) == set(arrs)
        for k, v in arrs.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(np.asarray(got[k]), v)
    with open(p, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        cache.load_arrays(p)


@pytest.mark.parametrize("dtype", [np.float64, np.int_])
def test_edge_file_equals_loadtxt_and_is_served_from_cache(tmp_path, dtype):
    rng = np.random.default_rng(1)
    t = np.stack([rng.integers(0, 3000, 500), rng.integers(0, 2000, 500), rng.integers(0, 5, 500)], 1)
    p = str(tmp_path / "ri_graph.txt")
    np.savetxt(p, t, fmt="%d", delimiter=" ")
    ref = np.loadtxt(p, dtype=dtype)                       # the reference call (dataset.py:325-343)
    a = cache.load_edge_triples(p, dtype=dtype)
    assert a.dtype == ref.dtype and np.array_equal(a, ref)
    assert os.path.exists(p + ".frcache")
    b = cache.load_edge_triples(p, dtype=dtype)            # second load: the memory-mapped cache
    assert isinstance(b, np.memmap) and np.array_equal(np.asarray(b), ref)
    # the text file changes -> the stale cache is ignored and rewritten
    time.sleep(0.01)
    np.savetxt(p, t[:100] + 1, fmt="%d", delimiter=" ")
    c = cache.load_edge_triples(p, dtype=dtype)
    assert np.array_equal(np.asarray(c), np.loadtxt(p, dtype=dtype))


def test_interactions_cache_equals_pickled_coo(tmp_path):
    from foodrec_b200.synth import make_dataset
    ds = make_dataset("mini")
    pk = str(tmp_path / "train_coo_matrix.pkl")
    with open(pk, "wb") as f:
        pickle.dump(ds.train_coo_matrix.astype(np.float64), f)
    with open(pk, "rb") as f:
        ref = pickle.load(f).astype(np.float32)            # dataset.py:88-91
    for _ in range(2):                                     # first call converts, second reads the binary form
        got = cache.interactions_from_pickle(pk)
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.array_equal(got.row, ref.row) and np.array_equal(got.col, ref.col) and np.array_equal(got.data, ref.data)
    assert os.path.exists(pk + ".frcache")


def test_graph_cache_round_trip(tmp_path):
    from foodrec_b200 import graph as G
    from foodrec_b200.synth import make_dataset
    ds = make_dataset("mini")
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
    p = str(tmp_path / "ui.frb")
    cache.save_graph(p, g)
    h = cache.load_graph(p, "cpu")
    assert (h.n_rows, h.n_cols, h.nnz, h.n_seg, h.n_long, h.n_part) == (g.n_rows, g.n_cols, g.nnz, g.n_seg, g.n_long, g.n_part)
    assert h.T is h
    for name in ("col", "val", "seg", "long_rows"):
        assert np.array_equal(getattr(h, name).numpy(), getattr(g, name).numpy()), name
    assert np.array_equal(h.row_ptr_host, g.row_ptr_host)

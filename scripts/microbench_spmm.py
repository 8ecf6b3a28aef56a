"""SpMM micro-benchmark on the C2 user-item graph: fused kernel vs stock torch.sparse.mm on the same GPU."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import graph as G, ops
from foodrec_b200.synth import make_dataset


def timeit(fn, iters=50, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    scale = sys.argv[1] if len(sys.argv) > 1 else "C2"
    t0 = time.time()
    ds = make_dataset(scale, features=False)
    print("dataset", time.time() - t0, "s", flush=True)
    t0 = time.time()
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    print("graph", time.time() - t0, "s  N", g.n_rows, "nnz", g.nnz, "n_seg", g.n_seg, "n_long", g.n_long, flush=True)
    N, d = g.n_rows, 64
    X = torch.randn(N, d, device="cuda")
    Z = torch.randn(N, d, device="cuda")
    Y = torch.empty(N, d, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for name, fl in (("warm_l2", None), ("flushed", flush)):
        t = timeit(lambda: ops.spmm(g, X, out=Y), flush=fl)
        tz = timeit(lambda: ops.spmm(g, X, Z=Z, alpha=0.5, beta=0.5, out=Y), flush=fl)
        res[name] = {"spmm_ms": t, "spmm_GBs": g.spmm_bytes(d) / t / 1e6, "spmm_z_ms": tz,
                     "spmm_z_GBs": g.spmm_bytes(d, True) / tz / 1e6}
    import numpy as np
    rows = np.repeat(np.arange(N), np.diff(g.row_ptr_host))
    S = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, g.col.cpu().numpy().astype(np.int64)])).cuda(),
                                g.val, (N, N))
    res["torch_sparse_mm_coo_ms"] = timeit(lambda: torch.sparse.mm(S, X))
    Sc = S.coalesce().to_sparse_csr()
    res["torch_sparse_mm_csr_ms"] = timeit(lambda: Sc @ X)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

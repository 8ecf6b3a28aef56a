"""Propagation on a graph whose embedding table does not fit the 126 MB L2 (HBM-bound regime, a scaled-down C5):
U users, I items, E interactions, power-law item popularity, log-normal user degrees; one SpMM layer, d = 64."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import foodrec_b200  # noqa
from foodrec_b200 import graph as G, ops


def main():
    U = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    I = int(float(sys.argv[2])) if len(sys.argv) > 2 else 400_000
    E = int(float(sys.argv[3])) if len(sys.argv) > 3 else 40_000_000
    rng = np.random.default_rng(0)
    t0 = time.time()
    deg = rng.lognormal(0.0, 1.0, size=U)
    deg = np.maximum(1, np.round(deg * (E / U / deg.mean()))).astype(np.int64)
    users = np.repeat(np.arange(U, dtype=np.int64), deg)
    w = 1.0 / np.power(np.arange(1, I + 1, dtype=np.float64), 0.7)
    cdf = np.cumsum(w); cdf /= cdf[-1]
    items = rng.permutation(I)[np.minimum(np.searchsorted(cdf, rng.random(users.shape[0])), I - 1)]
    g = G.symmetric_normalised(users, items + U, U + I, "cuda")
    print(f"graph: N={g.n_rows} nnz={g.nnz} n_seg={g.n_seg} n_long={g.n_long} built in {time.time() - t0:.1f}s", flush=True)
    N, d = g.n_rows, 64
    X = torch.randn(N, d, device="cuda"); Z = torch.randn(N, d, device="cuda"); Y = torch.empty(N, d, device="cuda")
    for _ in range(3):
        ops.spmm(g, X, Z=Z, alpha=0.5, beta=0.5, out=Y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.spmm(g, X, Z=Z, alpha=0.5, beta=0.5, out=Y)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    alg = g.spmm_bytes(d)
    print(json.dumps({"N": N, "nnz": g.nnz, "table_MB": N * d * 4 / 1e6, "ms": ms, "algorithmic_GB": alg / 1e9,
                      "algorithmic_GBs": alg / ms / 1e6, "gather_bound_GB": (g.nnz * (8 + 4 * d) + 8 * d * N) / 1e9,
                      "gather_GBs": (g.nnz * (8 + 4 * d) + 8 * d * N) / ms / 1e6}))


if __name__ == "__main__":
    main()

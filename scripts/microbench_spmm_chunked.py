"""HBM-regime experiment: the C5-shaped graph of bench.py (2 M users, 0.4 M items, 79 M stored entries, 614 MB table)
as ONE launch vs split launches -- user rows (their gathers hit the 102 MB item block, L2-resident) and item rows
by blocks of user columns that fit the L2 (Y = S_c X + Y, block after block)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import foodrec_b200  # noqa
from foodrec_b200 import graph as G, ops

dev = torch.device("cuda")
nu, ni = 2_000_000, 400_000
g = bench._c5_shaped_graph(dev, nu, ni, 40_000_000)
n = g.n_rows
X = torch.randn(n, 64, device=dev) * 0.1
Z = torch.randn(n, 64, device=dev)
Y = torch.empty(n, 64, device=dev)
t = lambda fn, it=5: bench.timed_ms(fn, it, warm=2)
out = {"full_ms": t(lambda: ops.spmm(g, X, Z=Z, alpha=0.5, beta=0.5, out=Y))}
ref = Y.clone()
rp = torch.from_numpy(g.row_ptr_host.astype(np.int64)).to(dev)
rows = torch.repeat_interleave(torch.arange(n, device=dev), rp[1:] - rp[:-1])


def sub_graph(mask, r0, r1):
    """rows [r0, r1) restricted to the entries selected by `mask`; output rows are numbered from r0."""
    sel = mask & (rows >= r0) & (rows < r1)
    r = rows[sel] - r0
    deg = torch.bincount(r, minlength=r1 - r0)
    p = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=p[1:])
    return G.PropGraph(p.cpu().numpy(), g.col[sel], g.val[sel], n, dev)


all_ = torch.ones_like(rows, dtype=torch.bool)
gu = sub_graph(all_, 0, nu)
gi = sub_graph(all_, nu, n)
Yu, Yi = Y[:nu], Y[nu:]
Zu, Zi = Z[:nu].contiguous(), Z[nu:].contiguous()
out["user_rows_ms"] = t(lambda: ops.spmm(gu, X, Z=Zu, alpha=0.5, beta=0.5, out=Yu))
out["item_rows_ms"] = t(lambda: ops.spmm(gi, X, Z=Zi, alpha=0.5, beta=0.5, out=Yi))
for K in (4, 8, 16, 32):
    edges = np.linspace(0, nu, K + 1).astype(np.int64)
    subs = [sub_graph((g.col >= int(edges[k])) & (g.col < int(edges[k + 1])), nu, n) for k in range(K)]
    Ya, Yb = torch.empty_like(Zi), torch.empty_like(Zi)

    def run():
        src, a, b = Zi, Ya, Yb
        for k, sg in enumerate(subs):
            ops.spmm(sg, X, Z=src, alpha=0.5, beta=0.5 if k == 0 else 1.0, out=a)
            src, a, b = a, b, a
        return src
    out[f"item_rows_{K}_chunks_ms"] = t(run)
    res = run()
    out[f"item_rows_{K}_chunks_max_rel_err"] = float((res - ref[nu:]).abs().max() / ref[nu:].abs().max())
    del subs
print(json.dumps(out))

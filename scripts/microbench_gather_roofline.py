"""Gather roofline vs the propagation kernel on the C2 user-item graph (one B200).

Times (a) `fr_probe_gather`: nothing but the 256-byte row gathers of one propagation launch, with the graph's
own column indices and with uniform random indices, for an L2-resident table (C2: 29 MB) and an HBM-sized one
(2.4 M rows = 614 MB); (b) the propagation kernel itself on the same graph.  Prints one JSON line.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import foodrec_b200  # noqa: F401
from foodrec_b200 import _lib, graph as G, ops
from foodrec_b200.synth import make_dataset


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def probe(tab, idx, inflight, blocks):
    out = torch.empty(blocks * 32, device="cuda")

    def run():
        _lib.check(_lib.lib.fr_probe_gather(tab.data_ptr(), 64, idx.data_ptr(), idx.numel(), inflight, blocks,
                                            out.data_ptr(), _lib.stream_ptr()), "fr_probe_gather")
    return timed(run)


def main():
    ds = make_dataset("C2", features=False)
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    X = torch.randn(g.n_rows, 64, device="cuda")
    res = {"graph": {"N": g.n_rows, "nnz": g.nnz}}
    ms = timed(lambda: ops.spmm(g, X))
    gather_bytes = g.nnz * 256
    res["spmm"] = {"us": ms * 1e3, "algorithmic_GBs": g.spmm_bytes(64) / ms / 1e6, "gathered_GBs": gather_bytes / ms / 1e6}
    best = {}
    for name, idx, tab in (("graph_cols_L2_table", g.col, X),
                           ("random_L2_table", torch.randint(0, g.n_rows, (g.nnz,), device="cuda", dtype=torch.int32), X)):
        for inflight in (4, 8):
            for blocks in (148 * 4, 148 * 8, 148 * 16, 148 * 32):
                t = probe(tab, idx, inflight, blocks)
                cur = best.get(name)
                if cur is None or t < cur["us"] / 1e3:
                    best[name] = {"us": t * 1e3, "GBs": idx.numel() * 256 / t / 1e6, "inflight": inflight, "blocks": blocks}
    big = torch.randn(2_400_000, 64, device="cuda")
    idx = torch.randint(0, big.shape[0], (20_000_000,), device="cuda", dtype=torch.int32)
    for inflight in (4, 8):
        for blocks in (148 * 8, 148 * 16, 148 * 32):
            t = probe(big, idx, inflight, blocks)
            cur = best.get("random_HBM_table")
            if cur is None or t < cur["us"] / 1e3:
                best["random_HBM_table"] = {"us": t * 1e3, "GBs": idx.numel() * 256 / t / 1e6, "inflight": inflight,
                                            "blocks": blocks}
    res["gather_only"] = best
    res["spmm_fraction_of_gather_roofline"] = res["spmm"]["gathered_GBs"] / best["graph_cols_L2_table"]["GBs"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""Per-graph propagation launch time on C2 (four graphs of CLUSSL), warm L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import graph as G, ops
from foodrec_b200.synth import make_dataset
from microbench_spmm import timeit

ds = make_dataset("C2")
gs = {"ui": G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda"),
      "ingre": G.norm_adj_item_side(ds.rIngre_triples, ds.n_items, ds.num_ingredients, "cuda"),
      "image": G.norm_adj_item_side(ds.image_cluster_triples, ds.n_items, ds.cfg.n_cluster, "cuda"),
      "text": G.norm_adj_item_side(ds.text_cluster_triples, ds.n_items, ds.cfg.n_cluster, "cuda")}
out = {}
for name, g in gs.items():
    X = torch.randn(g.n_cols, 64, device="cuda"); Z = torch.randn(g.n_rows, 64, device="cuda"); Y = torch.empty_like(Z)
    t = timeit(lambda: ops.spmm(g, X, Z=Z, alpha=0.5, beta=0.5, out=Y))
    out[name] = {"rows": g.n_rows, "nnz": g.nnz, "n_seg": g.n_seg, "n_long": g.n_long, "us": t * 1e3,
                 "GBs": g.spmm_bytes(64) / t / 1e6}
print(json.dumps(out, indent=1))

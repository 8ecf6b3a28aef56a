"""CPU oracle for the graph-propagation + full-ranking hot path.  TEST INFRASTRUCTURE ONLY.

A restatement, in torch-CPU / numpy, of what the reference computes on this path.  Every function
cites the reference lines it follows (paths relative to the reference checkout, `FoodRec/...`).
The arithmetic the reference delegates to third-party ops (`torch.sparse.mm`, `torch.topk`,
`torch.mm`; versions unpinned by the reference, torch 2.11 here) is delegated to the same ops.

Pinning: the reference has no tests and no golden vectors, so the oracle is
pinned against outputs of the reference itself, executed in the build container by
`tests/golden/make_golden.py` and committed under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays them.  SCHGN's GCNConv (torch_geometric, not installable
offline) is restated from PyG's documented semantics and is "parity unpinned".

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this package.  Nothing under `multi-modal-food-recommendation_b200/` does.
"""

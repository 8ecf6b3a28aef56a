"""C2 full-sort evaluation (70 000 users x 45 000 items, k = 20, history mask) of the CLUSSL drop-in after a few
hundred training steps, for different candidate slacks: total time, kernel time and how many rows take each path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E
from foodrec_b200.models.pricai_modelx import PRICAI_ModelX
from foodrec_b200.synth import make_dataset, sample_train_batches
from foodrec_b200.train import FusedAdam, GraphedTrainStep

dev = torch.device("cuda")
ds = make_dataset("C2")
cfg = bench.model_cfg(ds, "cuda")
torch.manual_seed(999)
m = PRICAI_ModelX(cfg, ds).to(dev).train()
opt = FusedAdam(m.parameters(), lr=cfg["learning_rate"])
keys = ("u_id", "pos_i_id", "neg_i_id")
bt = [{k: torch.from_numpy(b[k]).to(dev) for k in keys} for b in sample_train_batches(ds, 512, 16, seed=7)]
step = GraphedTrainStep(m, opt, bt[0], keys=keys)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1300):
    step(bt[i % 16])
m.eval()
with torch.no_grad():
    ua, ia = m._tables()
    ua, ia = ua.contiguous(), ia.contiguous()
hist = E.HistoryCSR(ds.train_coo_matrix, ds.n_users, dev)
users = torch.arange(ds.n_users, device=dev)
ib, bmax = E.to_bf16(ia), E.max_row_norm(ia)
out = {}
for slack in (12, 20, 28, 36, 44):
    st = {}
    fn = lambda: E.full_sort_topk(ua, ia, users, 20, hist=hist, B_bf16=ib, b_max_norm=bmax, index_dtype=torch.int32, stats=st, slack=slack)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        E.PROFILE = prof = []
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        E.PROFILE = None
        ts.append((a.elapsed_time(b), prof[0][0].elapsed_time(prof[0][1])))
    t = min(ts)
    out[f"slack_{slack}"] = {"total_ms": round(t[0], 3), "main_kernel_ms": round(t[1], 3), "stats": dict(st)}
print(json.dumps(out))

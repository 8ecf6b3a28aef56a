"""Per-kernel time of one eager CLUSSL train step at C2 (library kernels, CUDA events), plus the graph-replay step time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import foodrec_b200  # noqa
from foodrec_b200 import _lib
from foodrec_b200.models.pricai_modelx import PRICAI_ModelX
from foodrec_b200.synth import make_dataset, sample_train_batches
from foodrec_b200.train import GraphedTrainStep, eager_step

scale = sys.argv[1] if len(sys.argv) > 1 else "C2"
ds = make_dataset(scale, features=(scale != "C2") or True)
dev = torch.device("cuda")
cfg = bench.model_cfg(ds, "cuda")
torch.manual_seed(999)
model = PRICAI_ModelX(cfg, ds).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.002, capturable=True)
bs = sample_train_batches(ds, 512, 4, seed=7)
res = [{k: torch.from_numpy(b[k]).to(dev) for k in ("u_id", "pos_i_id", "neg_i_id")} for b in bs]
for i in range(3):
    eager_step(model, opt, res[i % 4])
torch.cuda.synchronize()
with _lib.kernel_profile() as prof:
    for i in range(5):
        eager_step(model, opt, res[i % 4])
tot = 0
for k, (n, us) in sorted(prof.result.items(), key=lambda x: -x[1][1]):
    print(f"{k:28s} {n/5:6.1f} launches/step {us/5:9.1f} us/step")
    tot += us / 5
print("library kernels total us/step", round(tot, 1))
g = GraphedTrainStep(model, opt, res[0])
for i in range(10):
    g(res[i % 4])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(100):
    g(res[i % 4])
b.record()
torch.cuda.synchronize()
print("graph replay ms/step", a.elapsed_time(b) / 100)

"""Kernel timeline of ONE graph-replayed CLUSSL C2 train step (torch profiler / CUPTI): start offset, duration, stream."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import foodrec_b200  # noqa
from foodrec_b200.models.pricai_modelx import PRICAI_ModelX
from foodrec_b200.synth import make_dataset, sample_train_batches
from foodrec_b200.train import FusedAdam, GraphedTrainStep

ds = make_dataset("C2")
dev = torch.device("cuda")
cfg = bench.model_cfg(ds, "cuda")
torch.manual_seed(999)
model = PRICAI_ModelX(cfg, ds).to(dev).train()
opt = FusedAdam(model.parameters(), lr=0.002)
bs = sample_train_batches(ds, 512, 4, seed=7)
res = [{k: torch.from_numpy(b[k]).to(dev) for k in ("u_id", "pos_i_id", "neg_i_id")} for b in bs]
g = GraphedTrainStep(model, opt, res[0], keys=("u_id", "pos_i_id", "neg_i_id"))
for i in range(20):
    g(res[i % 4])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        g(res[i % 4])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last replay: events after the last memcpy burst
starts = [e.time_range.start for e in ev]
t_end = max(e.time_range.end for e in ev)
# split into replays by gaps > 100 us
groups, cur = [], [ev[0]]
for a, b in zip(ev, ev[1:]):
    if b.time_range.start - a.time_range.end > 100:
        groups.append(cur); cur = []
    cur.append(b)
groups.append(cur)
last = groups[-1]
t0 = last[0].time_range.start
print(f"step span {last[-1].time_range.end - t0:.1f} us, {len(last)} kernels")
for e in last:
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:90]}")

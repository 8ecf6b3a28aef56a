"""Kernel-time breakdown of one eager HealthRec (CIKM_Model) train step on the C2 data (torch profiler, CUDA time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import foodrec_b200  # noqa
from foodrec_b200.models.cikm_model import CIKM_Model
from foodrec_b200.synth import make_dataset, sample_train_batches
from foodrec_b200.train import FusedAdam, eager_step

ds = make_dataset("C2")
dev = torch.device("cuda")
cfg = bench.Cfg(device="cuda", embedding_size=64, train_batch_size=512, is_multimodal_model=True, end2end=False,
                use_health_level_multi_hot=True, num_attention_heads=2, num_hidden_layers=2, attention_probs_dropout_prob=0.5,
                hidden_act="gelu", n_layers=2, ui_layers=1, reg_weight=0.5, loss_kd=0.05, loss_health=0.1, kd_threshold=0.4,
                learning_rate=0.001)
torch.manual_seed(999)
m = CIKM_Model(cfg, ds).to(dev).train()
opt = FusedAdam(m.parameters(), lr=1e-3)
bs = sample_train_batches(ds, 512, 2, seed=21)
res = [{k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in b.items()} for b in bs]
for i in range(3):
    eager_step(m, opt, res[i % 2])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        eager_step(m, opt, res[i % 2])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))

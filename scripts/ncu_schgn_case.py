"""One attend + score launch pair of the SCHGN full-sort scorer at C2 (256 users) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200.models.schgn import SCHGN
from foodrec_b200.synth import make_dataset


class Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


ds = make_dataset("C2", clusters=False)
cfg = Cfg(device="cuda", embedding_size=64, train_batch_size=512, is_multimodal_model=True, end2end=False,
          num_attention_heads=2, num_hidden_layers=2, hidden_act="gelu", inner_size=256, hidden_dropout_prob=0.5,
          attention_probs_dropout_prob=0.5, regs=0.01, reg_image=1, reg_w=0.05, reg_g=0.01, reg_health=0.01, ssl=0.008)
torch.manual_seed(999)
m = SCHGN(cfg, ds).to("cuda").eval()
with torch.no_grad():
    for _ in range(2):
        s = m.full_sort_scores(torch.arange(256, device="cuda"))
torch.cuda.synchronize()
print("ok", float(s[0, 0]))

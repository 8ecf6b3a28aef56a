"""SCHGN at C2 scale on one B200: train step (eager) and fused full-sort pair scorer.

python scripts/microbench_schgn.py [scale] [n_users]  ->  one JSON line.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import foodrec_b200  # noqa: F401
from foodrec_b200 import _lib
from foodrec_b200.models.schgn import SCHGN
from foodrec_b200.synth import make_dataset, sample_train_batches


class Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    scale = sys.argv[1] if len(sys.argv) > 1 else "C2"
    nu = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    t0 = time.time()
    ds = make_dataset(scale, clusters=False)
    cfg = Cfg(device="cuda", embedding_size=64, train_batch_size=512, is_multimodal_model=True, end2end=False,
              num_attention_heads=2, num_hidden_layers=2, hidden_act="gelu", inner_size=256, hidden_dropout_prob=0.5,
              attention_probs_dropout_prob=0.5, regs=0.01, reg_image=1, reg_w=0.05, reg_g=0.01, reg_health=0.01,
              ssl=0.008, SCHGN_ssl=True, neg_sample_num=4)
    torch.manual_seed(999)
    m = SCHGN(cfg, ds).to("cuda")
    out = {"scale": scale, "users": ds.n_users, "items": ds.n_items, "setup_s": round(time.time() - t0, 1)}

    # ---- train step, eager (calculate_loss -> backward -> Adam), B = 512
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, fused=True)
    batches = [{k: torch.from_numpy(np.asarray(v)).cuda() for k, v in b.items()}
               for b in sample_train_batches(ds, 512, 4, seed=1, schgn=True)]
    state = {"i": 0}

    def step():
        opt.zero_grad(set_to_none=True)
        loss = sum(m.calculate_loss(batches[state["i"] % len(batches)]))
        loss.backward()
        opt.step()
        state["i"] += 1
    for _ in range(5):
        step()
    out["train_step_ms"] = round(timed(step, 30), 3)
    if os.environ.get("FR_TORCH_PROFILE"):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60), file=sys.stderr)
    try:  # the same step as one CUDA-graph replay (static batch buffers, capturable Adam)
        from foodrec_b200.train import GraphedTrainStep
        opt_g = torch.optim.Adam(m.parameters(), lr=5e-4, fused=True, capturable=True)
        gstep = GraphedTrainStep(m, opt_g, batches[0], keys=tuple(batches[0].keys()))
        for b in batches:
            gstep(b)
        out["train_step_graph_ms"] = round(timed(lambda: gstep(batches[state["i"] % len(batches)]), 30), 3)
        out["graph_losses"] = [round(float(x), 4) for x in gstep(batches[1])]
    except Exception as exc:  # noqa: BLE001
        out["train_step_graph_error"] = repr(exc)[:300]
    x = torch.cat([m.user_embed, m.item_embed, m.ingre_embed_first, m.health_embed], 0).detach()
    ei = m._edges(m.g2i_edges, m.i2u_edges)
    g = m.new_gcn.plan(ei, x.shape[0])
    h = m.new_gcn.conv1.lin(x).detach()
    from foodrec_b200 import ops
    out["gcn_propagate_ms"] = round(timed(lambda: ops.gcn_propagate_tanh(g, h, m.new_gcn.conv1.bias.detach()), 50), 4)
    out["gcn_nodes"], out["gcn_nnz"] = g.n_rows, g.nnz

    # ---- full sort
    m.eval()
    users = torch.arange(nu, device="cuda")
    with torch.no_grad():
        t1 = time.time()
        m.full_sort_scores(users[:16])
        torch.cuda.synchronize()
        out["item_side_first_call_s"] = round(time.time() - t1, 2)
        m.full_sort_scores(users)
        ms = timed(lambda: m.full_sort_scores(users), 3)
        with _lib.kernel_profile() as prof:
            m.full_sort_scores(users)
        out["full_sort_ms"] = round(ms, 2)
        out["full_sort_users_per_s"] = round(nu / ms * 1e3, 1)
        out["kernels_ms"] = {k: round(us / 1e3, 3) for k, (n, us) in prof.result.items()}
        a = m.full_sort_scores(users[:64])
        out["score_std"] = float(a.std())
        t1 = time.time()
        v, i = m.full_sort_topk(users, 20)
        torch.cuda.synchronize()
        out["full_sort_topk_ms"] = round((time.time() - t1) * 1e3, 2)
    pairs = nu * ds.n_items
    out["pairs"] = pairs
    print(json.dumps(out))


if __name__ == "__main__":
    main()

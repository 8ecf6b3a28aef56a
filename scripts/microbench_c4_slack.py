"""C4-shaped full sort on one rank's shard (default 125 000 users x 500 000 items, k = 20, 20-item history mask): total
time of `evaluation.gemm_topk` (tensor-core pass + fp32 re-score + certificate + fallbacks) for different candidate
slacks, with the per-path row counts.  usage: python scripts/microbench_c4_slack.py [M]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E

dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
N, K, k = 500_000, 64, 20
g = torch.Generator(device=dev).manual_seed(4)
I = torch.randn(N, K, device=dev, generator=g) * 0.1
U = torch.randn(M, K, device=dev, generator=g) * 0.1
hist_idx = torch.sort(torch.randint(0, N, (M, 20), device=dev, generator=g), dim=1)[0].to(torch.int32).reshape(-1)


class H:
    ptr = torch.arange(0, 20 * M + 1, 20, device=dev, dtype=torch.int64)
    idx = hist_idx


Ib, bmax = E.to_bf16(I), E.max_row_norm(I)
rid = torch.arange(M, device=dev)
out = {"M": M}
for slack in (None, 12, 20):
    st, prof = {}, []
    fn = lambda: E.gemm_topk(U, I, k, row_ids=rid, hist=H, B_bf16=Ib, b_max_norm=bmax, index_dtype=torch.int32, stats=st, slack=slack)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        E.PROFILE = prof = []
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        E.PROFILE = None
        ts.append((a.elapsed_time(b), prof[0][0].elapsed_time(prof[0][1])))
    t = min(ts)
    out[f"slack_{slack}"] = {"total_ms": t[0], "main_kernel_ms": t[1], "stats": dict(st)}
print(json.dumps(out))

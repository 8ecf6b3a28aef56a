"""Cosine kNN item-graph build at C2 scale (45 000 items, 384-d text / 4096-d image features, k = 10):
fused tcgen05 GEMM + top-k vs torch (dense similarity + topk) on the same GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E, _lib
from microbench_eval import timeit

N = int(sys.argv[1]) if len(sys.argv) > 1 else 45000
out = {}
for D in (384, 4096):
    torch.manual_seed(D)
    x = torch.randn(N, D, device="cuda")
    xn = (x / x.norm(dim=-1, keepdim=True)).contiguous()
    xb = E.to_bf16(xn)
    prof = []
    E.PROFILE = prof
    t_all = timeit(lambda: E.gemm_topk(xn, xn, 10, A_bf16=xb, B_bf16=xb), iters=3, warm=1)
    E.PROFILE = None
    kms = sorted(p[0].elapsed_time(p[1]) for p in prof)[len(prof) // 2]
    res = {"N": N, "D": D, "kernel_ms": kms, "pipeline_ms": t_all, "tflops": 2.0 * N * N * D / kms / 1e9}
    if N * N * 4 <= 10e9:
        def torch_path():
            for s in range(0, N, 4096):
                torch.topk(xn[s:s + 4096] @ xn.t(), 10, dim=-1)
        res["torch_fp32_ms"] = timeit(torch_path, iters=2, warm=1)
        def torch_bf16():
            for s in range(0, N, 4096):
                torch.topk(xb[s:s + 4096] @ xb.t(), 10, dim=-1)
        res["torch_bf16_ms"] = timeit(torch_bf16, iters=2, warm=1)
    out[D] = res
    del x, xn, xb
print(json.dumps(out))

"""One fused GEMM+top-k launch at kNN shape (N x N x 4096) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E
N = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
torch.manual_seed(0)
x = torch.randn(N, 4096, device="cuda")
xn = (x / x.norm(dim=-1, keepdim=True)).contiguous()
xb = E.to_bf16(xn)
for _ in range(3):
    v, i = E.gemm_topk(xn, xn, 10, A_bf16=xb, B_bf16=xb)
torch.cuda.synchronize()
print("ok", int(i[0, 0]))

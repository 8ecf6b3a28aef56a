"""Full-sort eval micro-benchmark: fused tensor-core score+top-K vs torch matmul+topk on the same GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E, _lib


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M, N = int(sys.argv[1]) if len(sys.argv) > 1 else 70000, int(sys.argv[2]) if len(sys.argv) > 2 else 45000
    K, k = 64, 20
    torch.manual_seed(0)
    U = torch.randn(M, K, device="cuda") * 0.1
    I = torch.randn(N, K, device="cuda") * 0.1
    Ub, Ib = E.to_bf16(U), E.to_bf16(I)
    cv = torch.empty(M, 32, device="cuda"); ci = torch.empty(M, 32, dtype=torch.int32, device="cuda")

    ws = E._workspace(U.device, int(_lib.lib.fr_gemm_topk_ws_bytes(M)))

    def raw():
        _lib.check(_lib.lib.fr_gemm_topk_bf16(Ub.data_ptr(), M, Ib.data_ptr(), N, K, 1.0, None, None, None, None, 32,
                                              cv.data_ptr(), ci.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    t_raw = timeit(raw)
    t_full = timeit(lambda: E.gemm_topk(U, I, k))
    res = {"M": M, "N": N, "K": K, "gemm_topk_ms": t_raw, "tflops": 2.0 * M * N * K / t_raw / 1e9,
           "users_per_s_kernel": M / t_raw * 1e3, "full_pipeline_ms": t_full, "users_per_s_pipeline": M / t_full * 1e3}
    if M * N <= 4e9:
        def torch_path():
            for s in range(0, M, 8192):
                torch.topk(U[s:s + 8192] @ I.t(), k, dim=-1)
        res["torch_fp32_matmul_topk_ms"] = timeit(torch_path, iters=3, warm=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

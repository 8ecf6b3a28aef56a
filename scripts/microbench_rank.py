"""Ranking kernel microbenchmark (`rank_topk_pair_kernel`) on the judged shapes.
usage: python scripts/microbench_rank.py [c4|knn|text|all]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda"
    out = {"impl": "rank_topk_pair_kernel"}
    peak = 1645.6
    if which in ("c4", "all"):
        M, N, K, k = 148 * 128 * 4, 500_000, 64, 32
        g = torch.Generator(device=dev).manual_seed(4)
        U = torch.randn(M, K, device=dev, generator=g) * 0.1
        I = torch.randn(N, K, device=dev, generator=g) * 0.1
        Ub, Ib = E.to_bf16(U), E.to_bf16(I)
        hidx = torch.sort(torch.randint(0, N, (M, 20), device=dev, generator=g), dim=1)[0].to(torch.int32).reshape(-1)

        class H:
            ptr = torch.arange(0, 20 * M + 1, 20, device=dev, dtype=torch.int64)
            idx = hidx
        rid = torch.arange(M, device=dev)
        for name, kw in (("c4_slice_mask_kc32", dict(row_ids=rid, hist=H)), ("c4_slice_nomask_kc32", {})):
            t = timeit(lambda: E.gemm_topk(U, I, k, exact=False, A_bf16=Ub, B_bf16=Ib, **kw))
            out[name] = {"ms": t, "tflops": 2.0 * M * N * K / t / 1e9, "frac": 2.0 * M * N * K / t / 1e9 / peak}
        # correctness spot check vs fp32 (exact path)
        st = {}
        v, i = E.gemm_topk(U, I, 20, row_ids=rid, hist=H, A_bf16=Ub, B_bf16=Ib, stats=st)
        sub = torch.arange(0, M, M // 128, device=dev)[:128]
        S = (U[sub].double() @ I.double().t()).float()
        S.scatter_(1, hidx.view(M, 20)[sub].long(), float("-inf"))
        ref = torch.topk(S, 20, dim=-1)
        out["c4_check"] = {"idx_mismatch": int((i[sub] != ref[1]).sum()), "stats": st,
                           "max_score_diff": float((torch.gather(S, 1, i[sub]) - ref[0]).abs().max())}
        # C2 eval shape
        M2, N2 = 70_000, 45_000
        U2 = torch.randn(M2, K, device=dev, generator=g) * 0.1
        I2 = torch.randn(N2, K, device=dev, generator=g) * 0.1
        U2b, I2b = E.to_bf16(U2), E.to_bf16(I2)
        t = timeit(lambda: E.gemm_topk(U2, I2, 32, exact=False, A_bf16=U2b, B_bf16=I2b))
        out["c2_eval_nomask_kc32"] = {"ms": t, "tflops": 2.0 * M2 * N2 * K / t / 1e9, "frac": 2.0 * M2 * N2 * K / t / 1e9 / peak}
    for name, D in (("knn", 4096), ("text", 384)):
        if which in (name, "all"):
            n = 45_000
            g = torch.Generator(device=dev).manual_seed(7)
            x = torch.randn(n, D, device=dev, generator=g)
            xn = (x / x.norm(dim=-1, keepdim=True)).contiguous()
            xb = E.to_bf16(xn)
            t = timeit(lambda: E.gemm_topk(xn, xn, 22, exact=False, A_bf16=xb, B_bf16=xb), iters=2)
            out[f"knn_D{D}_kc22"] = {"ms": t, "tflops": 2.0 * n * n * D / t / 1e9, "frac": 2.0 * n * n * D / t / 1e9 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""One timing of the fused score + top-K kernel on the C4 slice shape under the current FR_TOPK_* environment
(the environment is read once per process, so variants are separate processes).
usage: python scripts/microbench_rank_probe.py [mask|nomask] [k] [M]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foodrec_b200  # noqa
from foodrec_b200 import evaluation as E


def main():
    mask = (sys.argv[1] if len(sys.argv) > 1 else "nomask") == "mask"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 128 * 4
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    N, K = 500_000, 64
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(4)
    U = torch.randn(M, K, device=dev, generator=g) * 0.1
    I = torch.randn(N, K, device=dev, generator=g) * 0.1
    Ub, Ib = E.to_bf16(U), E.to_bf16(I)
    kw = {}
    if mask:
        hidx = torch.sort(torch.randint(0, N, (M, max(H, 1)), device=dev, generator=g), dim=1)[0].to(torch.int32).reshape(-1)

        class Hist:
            ptr = torch.arange(0, H * M + 1, max(H, 1), device=dev, dtype=torch.int64) if H > 0 else torch.zeros(M + 1, device=dev, dtype=torch.int64)
            idx = hidx
        kw = dict(row_ids=torch.arange(M, device=dev), hist=Hist)
    fn = lambda: E.gemm_topk(U, I, k, exact=False, A_bf16=Ub, B_bf16=Ib, **kw)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    env = {k_: v for k_, v in os.environ.items() if k_.startswith("FR_TOPK")}
    print(json.dumps({"env": env, "mask": mask, "hist_len": H if mask else None, "k": k, "M": M, "ms": round(t, 3), "tflops": round(2.0 * M * N * K / t / 1e9, 1),
                      "frac_burst": round(2.0 * M * N * K / t / 1e9 / 1645.6, 4)}))


if __name__ == "__main__":
    main()

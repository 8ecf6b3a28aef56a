"""Mini-batch k-means fit on the device.

The reference builds CLUSSL's item-cluster graphs in notebooks (dataset_process/allrecipes_kmeans.ipynb,
foodcom_kmeans.ipynb, cells 0 and 2): `MiniBatchKMeans(n_clusters=2000, init_size=512, batch_size=1024,
random_state=2024).fit(features)` on the host, then a python loop over items for the six nearest centres
(`evaluation.centroid_topk` replaces that loop).  This module is the fit, following scikit-learn's
algorithm (`sklearn/cluster/_kmeans.py`, MiniBatchKMeans: k-means++ seeding on `init_size` samples,
uniformly sampled mini-batches, per-centre running means `c <- (c * n_c + sum of assigned rows) / (n_c + m_c)`,
stop after `max_no_improvement` steps without a better smoothed inertia or `max_iter` epochs):

* the assignment step -- nearest centre of every row of the batch -- is the fused bf16 tensor-core GEMM +
  top-1 kernel with the `-|c|^2 / 2` column bias and an exact fp32 re-score (`fr_gemm_topk_bf16`,
  `fr_rescore_topk_f32`), the per-centre sums are one `fr_scatter_add_rows` launch;
* scikit-learn's random re-assignment of starved centres is not reproduced, and the random stream is torch's,
  not numpy's: the result is a k-means solution of the same quality, not the same centres (parity is
  statistical: tests compare the inertia with scikit-learn's on the same data).
"""
from __future__ import annotations

import torch

from . import _lib, evaluation

_L = _lib.lib


def _sq_dist_to(x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    return ((x - c) ** 2).sum(1)


def kmeans_plus_plus(x: torch.Tensor, n_clusters: int, gen: torch.Generator) -> torch.Tensor:
    """D^2 seeding (Arthur & Vassilvitskii) with scikit-learn's `2 + log(k)` local trials per step."""
    n = x.shape[0]
    trials = 2 + int(torch.log(torch.tensor(float(n_clusters))).item())
    first = int(torch.randint(n, (1,), generator=gen, device=x.device))
    centres = [x[first]]
    closest = _sq_dist_to(x, x[first])
    for _ in range(1, n_clusters):
        cand = torch.multinomial(closest.clamp_min(0) + 1e-30, trials, replacement=True, generator=gen)
        d = torch.cdist(x[cand], x).pow(2)                       # [trials, n]
        pot = torch.minimum(d, closest[None, :]).sum(1)
        best = int(pot.argmin())
        centres.append(x[cand[best]])
        closest = torch.minimum(closest, d[best])
    return torch.stack(centres)


def assign(x: torch.Tensor, centres: torch.Tensor):
    """Nearest centre (index, squared distance) of every row of `x` through the fused GEMM + top-1 kernel."""
    idx = evaluation.centroid_topk(x, centres, 1)[:, 0]
    return idx, _sq_dist_to(x, centres[idx])


def minibatch_kmeans(features: torch.Tensor, n_clusters: int, batch_size: int = 1024, max_iter: int = 100,
                     init_size: int | None = None, max_no_improvement: int = 10, seed: int = 2024):
    """Returns (`centres [n_clusters, D]` fp32, `inertia` of the full data set under them)."""
    x = features.detach().float().contiguous()
    if x.device.type != "cuda":
        raise _lib.FoodRecError("minibatch_kmeans needs a CUDA tensor (no CPU path)")
    n, d = x.shape
    if d % 8:
        raise _lib.FoodRecError(f"feature width {d} must be a multiple of 8")
    gen = torch.Generator(device=x.device)
    gen.manual_seed(seed)
    init_size = min(n, max(init_size or 3 * batch_size, 3 * n_clusters))    # scikit-learn's adjustment
    batch_size = min(batch_size, n)
    sub = x[torch.randperm(n, device=x.device, generator=gen)[:init_size]]
    centres = kmeans_plus_plus(sub, n_clusters, gen).contiguous()
    counts = torch.zeros(n_clusters, device=x.device)
    steps = max(1, (max_iter * n) // batch_size)
    ewa, best, stall = None, float("inf"), 0
    alpha = min(1.0, batch_size * 2.0 / (n + 1))
    for step in range(steps):
        rows = torch.randint(n, (batch_size,), device=x.device, generator=gen)
        xb = x[rows]
        idx, dist = assign(xb, centres)
        m = torch.bincount(idx, minlength=n_clusters).float()
        sums = torch.zeros_like(centres)
        _lib.check(_L.fr_scatter_add_rows(xb.data_ptr(), d, idx.contiguous().data_ptr(), idx.numel(), sums.data_ptr(),
                                          _lib.stream_ptr()), "fr_scatter_add_rows")
        new_counts = counts + m
        touched = m > 0
        centres = torch.where(touched[:, None], (centres * counts[:, None] + sums) / new_counts.clamp_min(1)[:, None],
                              centres).contiguous()
        counts = new_counts
        if step % 8 == 7:                                          # smoothed batch inertia, checked every 8 steps
            inertia = float(dist.sum())
            ewa = inertia if ewa is None else ewa * (1 - alpha) + inertia * alpha
            if ewa < best:
                best, stall = ewa, 0
            else:
                stall += 1
                if max_no_improvement is not None and stall >= max_no_improvement:
                    break
    total = 0.0
    for s in range(0, n, 65536):
        total += float(assign(x[s:s + 65536], centres)[1].sum())
    return centres, total

"""GPU parity of the fused tensor-core score + top-K kernel against the CPU oracle (`torch.topk`)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def check_topk(val, idx, scores_cpu, k, atol=2e-6, excluded=None):
    """Indices identical to fp32 `topk` except where the fp32 scores tie (within fp32 summation noise)."""
    ref_v, ref_i = torch.topk(scores_cpu, k, dim=-1)
    idx, val = idx.cpu(), val.cpu()
    assert idx.shape == ref_i.shape
    got_scores = torch.gather(scores_cpu, 1, idx.clamp_min(0))
    assert torch.allclose(val, got_scores, atol=atol, rtol=1e-5)           # reported values are the exact scores
    assert torch.allclose(got_scores, ref_v, atol=atol, rtol=1e-5)         # rank-by-rank the same score
    mism = idx != ref_i
    if mism.any():                                                         # allowed only at (near-)ties
        assert float((got_scores - ref_v).abs()[mism].max()) <= atol
    for r in range(idx.shape[0]):
        assert len(set(idx[r].tolist())) == k
    return int(mism.sum())


@pytest.mark.parametrize("M,N,K,k", [(1, 300, 64, 5), (128, 256, 64, 20), (300, 1000, 64, 20), (257, 2049, 64, 50),
                                     (513, 5000, 128, 10), (200, 777, 384, 7), (130, 520, 72, 52)])
def test_gemm_topk_matches_fp32_topk(M, N, K, k):
    from foodrec_b200 import evaluation as E
    torch.manual_seed(M * 7 + N)
    A, B = torch.randn(M, K) * 0.1, torch.randn(N, K) * 0.1
    val, idx = E.gemm_topk(A.cuda(), B.cuda(), k)
    check_topk(val, idx, A @ B.t(), k)


def test_raw_bf16_scores_are_the_tensor_core_product():
    """Without the fp32 re-score the values must equal the bf16-rounded inputs' product (validates the
    UMMA descriptors / swizzle end to end, every K step and every column block)."""
    from foodrec_b200 import evaluation as E
    torch.manual_seed(5)
    M, N, K, k = 256, 1024, 192, 64
    A, B = torch.randn(M, K), torch.randn(N, K)
    val, idx = E.gemm_topk(A.cuda(), B.cuda(), k, exact=False)
    S = A.bfloat16().float() @ B.bfloat16().float().t()
    ref_v, _ = torch.topk(S, k, dim=-1)
    assert torch.allclose(val.cpu(), ref_v, atol=1e-3, rtol=1e-4)
    assert torch.allclose(torch.gather(S, 1, idx.cpu()), val.cpu(), atol=1e-3, rtol=1e-4)


def test_history_mask_and_ragged_rows():
    from foodrec_b200 import evaluation as E
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    U, I, K, k = 500, 3000, 64, 20
    deg = rng.integers(0, 60, size=U)
    deg[7] = 0
    deg[11] = I - 5  # a user who has seen almost everything: fewer than k items remain
    rows = np.repeat(np.arange(U), deg)
    cols = np.concatenate([rng.choice(I, size=d, replace=False) for d in deg])
    coo = sp.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(U, I))
    hist = E.HistoryCSR(coo, U, "cuda")
    torch.manual_seed(1)
    ue, ie = torch.randn(U, K) * 0.1, torch.randn(I, K) * 0.1
    users = torch.from_numpy(rng.permutation(U)[:333].copy())
    val, idx = E.full_sort_topk(ue.cuda(), ie.cuda(), users.cuda(), k, hist=hist)
    S = ue[users] @ ie.t()
    for r, u in enumerate(users.tolist()):
        S[r, hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].astype(np.int64)] = -float("inf")
    ref_v, ref_i = torch.topk(S, k, dim=-1)
    idx, val = idx.cpu(), val.cpu()
    finite = torch.isfinite(ref_v)
    assert torch.equal(idx[~finite], torch.full_like(idx[~finite], -1))       # padding where nothing is left
    assert torch.allclose(val[finite], ref_v[finite], atol=2e-6)
    mism = (idx != ref_i) & finite
    if mism.any():
        assert float((val - ref_v).abs()[mism].max()) <= 2e-6
    # no masked item is ever returned
    for r, u in enumerate(users.tolist()):
        seen = set(hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].tolist())
        assert not (set(idx[r][idx[r] >= 0].tolist()) & seen)


def test_full_sort_golden_and_metrics():
    """The reference's `torch.topk(scores, 50)` indices (golden, CPU) and the 4-d.p. metrics."""
    from foodrec_b200 import evaluation as E, metrics
    g = load_golden("primitives.npz")
    ue, ie = torch.from_numpy(g["rank/ue"]), torch.from_numpy(g["rank/ie"])
    val, idx = E.full_sort_topk(ue.cuda(), ie.cuda(), None, 50)
    n_diff = check_topk(val, idx, ue @ ie.t(), 50)
    ptr, pidx = g["rank/pos_ptr"], g["rank/pos_idx"]
    pos = [pidx[ptr[u]:ptr[u + 1]].tolist() for u in range(len(ptr) - 1)]
    res = metrics.topk_metrics(idx.cpu().numpy(), pos)
    for kname, v in zip(g["rank/metric_keys"], g["rank/metric_vals"]):
        assert res[str(kname)] == v, (kname, res[str(kname)], v, n_diff)


def test_cosine_knn_and_laplacian_golden():
    from foodrec_b200 import evaluation as E
    from oracle import knn
    g = load_golden("primitives.npz")
    feat = torch.from_numpy(g["knn/feat"])
    val, ind = E.knn_topk(feat.cuda(), 7)
    sim = torch.from_numpy(g["knn/sim"])
    check_topk(val, ind, sim, 7)
    assert torch.equal(ind.cpu()[:, 0], torch.arange(feat.shape[0]))  # self is the nearest neighbour and is kept
    ei, w = E.knn_normalized_graph(feat.cuda(), 7, "sym")
    row, col, w_ref = knn.knn_normalized_graph(sim, 7, "sym")
    assert torch.equal(ei[0].cpu(), row)
    same = ei[1].cpu() == col
    assert float(same.float().mean()) > 0.99
    assert torch.allclose(w.cpu()[same], w_ref[same], rtol=1e-5, atol=1e-7)


def test_knn_larger_feature_width():
    from foodrec_b200 import evaluation as E
    torch.manual_seed(2)
    feat = torch.randn(1500, 384)
    val, ind = E.knn_topk(feat.cuda(), 10)
    xn = feat / feat.norm(dim=-1, keepdim=True)
    check_topk(val, ind, xn @ xn.t(), 10, atol=5e-6)


def test_centroid_assignment_golden(mini_ds):
    from foodrec_b200 import evaluation as E
    g = load_golden("primitives.npz")
    got = E.centroid_topk(torch.from_numpy(mini_ds.embImage[:64]).cuda(), torch.from_numpy(mini_ds.image_center).cuda(), 6)
    assert np.array_equal(got.cpu().numpy(), g["centroid/top6"])


def test_centroid_assignment_c1_scale():
    from foodrec_b200 import evaluation as E
    from foodrec_b200.synth import make_dataset
    ds = make_dataset("C1")
    got = E.centroid_topk(torch.from_numpy(ds.embImage).cuda(), torch.from_numpy(ds.image_center).cuda(), 6).cpu().numpy()
    ref = ds.image_cluster_triples[:, 1].reshape(ds.n_items, 6)   # exact fp64 assignment from the generator
    assert (got == ref).mean() > 0.9999
    assert (np.sort(got, 1) == np.sort(ref, 1)).all(axis=1).mean() > 0.999


def test_rejects_bad_shapes():
    from foodrec_b200 import _lib, evaluation as E
    A, B = torch.randn(4, 60).cuda(), torch.randn(9, 60).cuda()
    with pytest.raises(_lib.FoodRecError):
        E.gemm_topk(A, B, 3)            # K not a multiple of 8
    with pytest.raises(_lib.FoodRecError):
        E.gemm_topk(torch.randn(4, 64).cuda(), torch.randn(9, 64).cuda(), 65)
    with pytest.raises(_lib.FoodRecError):
        E.gemm_topk(torch.randn(4, 64).cuda(), torch.randn(9, 64).cuda(), 3, index_dtype=torch.int16)


def test_k_near_the_candidate_capacity_is_still_exact():
    """k = 60 leaves only 4 spare candidates: the certificate (not a slack heuristic) guarantees the result, rows it
    cannot certify go through the exact fp32 kernel."""
    from foodrec_b200 import evaluation as E
    torch.manual_seed(8)
    A, B = torch.randn(70, 64) * 0.1, torch.randn(900, 64) * 0.1
    st = {}
    val, idx = E.gemm_topk(A.cuda(), B.cuda(), 60, stats=st)
    check_topk(val, idx, A @ B.t(), 60)
    assert st["rows"] == 70 and st["kc"] == 64


def test_wide_inner_dimension_single_sweep():
    """K = 4096 (image features): many k-blocks per tile, single-sweep path, partial last column block."""
    from foodrec_b200 import evaluation as E
    torch.manual_seed(11)
    feat = torch.randn(700, 4096)
    val, ind = E.knn_topk(feat.cuda(), 10)
    xn = feat / feat.norm(dim=-1, keepdim=True)
    check_topk(val, ind, xn @ xn.t(), 10, atol=5e-6)


def test_minibatch_kmeans_matches_sklearn_quality():
    """`kmeans.minibatch_kmeans` (dataset_process/*_kmeans.ipynb cell 0) against scikit-learn's MiniBatchKMeans
    on the same blobs: the inertia is within 10 % (different random stream: statistical parity), every point's
    assignment is its exact nearest centre, and a second fit with the same seed lands on the same quality (the
    per-centre sums are fp32 atomics, so the centres themselves may differ in the last bits)."""
    from sklearn.cluster import MiniBatchKMeans
    from foodrec_b200 import kmeans
    rng = np.random.default_rng(0)
    true = rng.standard_normal((24, 32)).astype(np.float32) * 3
    x = (true[rng.integers(0, 24, 6000)] + rng.standard_normal((6000, 32)).astype(np.float32) * 0.6).astype(np.float32)
    ref = MiniBatchKMeans(n_clusters=24, init_size=512, batch_size=256, random_state=2024, n_init=1).fit(x)
    xt = torch.from_numpy(x).cuda()
    centres, inertia = kmeans.minibatch_kmeans(xt, 24, batch_size=256, init_size=512, seed=2024)
    assert centres.shape == (24, 32)
    assert inertia <= 1.10 * ref.inertia_, (inertia, ref.inertia_)
    idx, dist = kmeans.assign(xt, centres)
    exact = torch.cdist(xt, centres).pow(2)
    assert float((dist - exact.min(1).values).abs().max()) <= 1e-3
    assert float((exact.gather(1, idx[:, None])[:, 0] - exact.min(1).values).abs().max()) <= 1e-4 * float(exact.max())
    c2, i2 = kmeans.minibatch_kmeans(xt, 24, batch_size=256, init_size=512, seed=2024)
    assert abs(i2 - inertia) <= 0.02 * inertia


# ------------------------------------------------------------------------------------------------------------------
# The paths that only large shapes reach (they ran only inside bench.py's spot checks): the bounding sweep
# on every second column tile (N >= 262 144), the persistent loop over more row blocks than CTA pairs
# (M > 148 * 128) and users whose history is longer than the NG = 128 group maxima of the bounding pass.
def _sampled_reference(U, I, hist, rows, k):
    S = (U[rows].double() @ I.double().t()).float().cpu()           # fp64 accumulate -> the fp32 ranking without summation noise
    S32 = (U[rows] @ I.t()).cpu()
    for r, u in enumerate(rows.tolist()):
        if hist is not None:
            cols = hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].astype(np.int64)
            S[r, cols] = -float("inf")
            S32[r, cols] = -float("inf")
    return S, S32


@pytest.mark.parametrize("k", [20, 50])
def test_large_shape_two_sweep_persistent_long_history(k):
    from foodrec_b200 import evaluation as E
    import scipy.sparse as sp
    dev = "cuda"
    M, N, K = 20_000, 300_000, 64
    g = torch.Generator(device=dev).manual_seed(1234 + k)
    U = torch.randn(M, K, device=dev, generator=g) * 0.1
    I = torch.randn(N, K, device=dev, generator=g) * 0.1
    rng = np.random.default_rng(k)
    deg = rng.integers(0, 40, size=M)
    long_users = rng.choice(M, size=64, replace=False)
    deg[long_users] = rng.integers(129, 700, size=64)                # history > NG: the bound falls back to -inf
    rows = np.repeat(np.arange(M), deg)
    cols = rng.integers(0, N, size=rows.size)
    hist = E.HistoryCSR(sp.coo_matrix((np.ones(rows.size, np.float32), (rows, cols)), shape=(M, N)), M, dev)
    st = {}
    val, idx = E.gemm_topk(U, I, k, hist=hist, stats=st, index_dtype=torch.int32)
    assert idx.dtype == torch.int32 and idx.shape == (M, k)
    sample = torch.from_numpy(np.unique(np.concatenate([long_users, rng.choice(M, size=512, replace=False),
                                                        np.arange(M - 130, M)]))).to(dev)
    S, _ = _sampled_reference(U, I, hist, sample.cpu(), k)
    check_topk(val[sample], idx[sample].long(), S, k, atol=3e-6)
    for r, u in enumerate(sample.tolist()):                          # no masked item is ever returned
        seen = set(hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].tolist())
        assert not (set(idx[u].tolist()) & seen)
    # the certificate passes for almost every row at k = 20; at k = 50 the widest candidate set (64) leaves only 14
    # ranks of slack against a 2^-8 |u| max|i| rounding bound, so more rows take the wide / exact fp32 paths
    assert st["exact_rows"] <= st["uncertified"] <= (M // 50 if k == 20 else M // 4)


def test_persistent_loop_more_row_blocks_than_ctas():
    """M > 148 * 128 rows: every CTA pair walks several 256-row blocks (TMEM / barrier phases carry over)."""
    from foodrec_b200 import evaluation as E
    dev = "cuda"
    M, N, K, k = 148 * 128 + 777, 4096, 64, 20
    g = torch.Generator(device=dev).manual_seed(99)
    U = torch.randn(M, K, device=dev, generator=g) * 0.1
    I = torch.randn(N, K, device=dev, generator=g) * 0.1
    val, idx = E.gemm_topk(U, I, k)
    sample = torch.cat([torch.arange(0, M, 37, device=dev), torch.arange(M - 300, M, device=dev)])
    check_topk(val[sample], idx[sample], (U[sample].double() @ I.double().t()).float().cpu(), k, atol=3e-6)


def test_certificate_catches_adversarial_near_ties():
    """Columns engineered so that bf16 rounding reorders them across the candidate cut: many near-duplicates of the
    best items differ by less than bf16 resolution.  The bf16 pass alone cannot rank them; the certificate must flag
    those rows and the fallback must still return the fp32 top-k."""
    from foodrec_b200 import evaluation as E
    torch.manual_seed(3)
    M, N, K, k = 96, 6000, 64, 20
    A = torch.randn(M, K) * 0.1
    B = torch.randn(N, K) * 0.02
    base = torch.randn(1, K) * 0.3
    B[:200] = base + torch.randn(200, K) * 2e-4          # 200 columns within ~1e-4 of each other: far below 2^-8 relative
    A[:48] = base * 0.5 + torch.randn(48, K) * 0.01      # rows that score that cluster highest
    st = {}
    val, idx = E.gemm_topk(A.cuda(), B.cuda(), k, stats=st)
    S = (A.double() @ B.double().t()).float()
    check_topk(val, idx, S, k, atol=4e-6)
    assert st["uncertified"] >= 40 and st["exact_rows"] >= 40        # the adversarial rows were caught ...
    ref_i = torch.topk(S, k, dim=-1)[1]
    gap = (torch.topk(S, k + 1, dim=-1)[0][:, k - 1] - torch.topk(S, k + 1, dim=-1)[0][:, k]).abs()
    clear = gap > 8e-6
    assert torch.equal(torch.sort(idx.cpu()[clear], 1)[0], torch.sort(ref_i[clear], 1)[0])   # ... and answered exactly
    raw_v, raw_i = E.gemm_topk(A.cuda(), B.cuda(), k, exact=False)                            # the bf16 pass alone is wrong here
    assert not torch.equal(torch.sort(raw_i.cpu()[:48], 1)[0], torch.sort(ref_i[:48], 1)[0])


@pytest.mark.parametrize("metric", [0, 1])
def test_exact_fp32_row_kernel(metric):
    """`fr_exact_topk_f32` on its own: arbitrary row subset, history mask, ties to the lower column, padding."""
    from foodrec_b200 import evaluation as E
    import scipy.sparse as sp
    torch.manual_seed(21)
    M, N, K, k = 50, 3001, 72, 33
    A, B = torch.randn(M, K), torch.randn(N, K)
    B[100] = B[7]                                         # exact duplicates: the lower column must come first
    B[2000] = B[7]
    rows = torch.tensor([3, 49, 0, 17, 17, 8])
    rng = np.random.default_rng(5)
    hr = np.repeat(np.arange(M), 30)
    hc = rng.integers(0, N, size=hr.size)
    hist = E.HistoryCSR(sp.coo_matrix((np.ones(hr.size, np.float32), (hr, hc)), shape=(M, N)), M, "cuda")
    bias = torch.randn(N) if metric == 0 else None
    v, i = E.exact_topk_rows(A.cuda(), rows.cuda(), B.cuda(), k, scale=0.5 if metric == 0 else 1.0,
                             bias=None if bias is None else bias.cuda(), metric=metric, hist=hist)
    if metric == 0:
        S = 0.5 * (A[rows].double() @ B.double().t()).float() + bias
    else:
        S = -(torch.cdist(A[rows].double(), B.double()) ** 2).float()
    for r, u in enumerate(rows.tolist()):
        S[r, hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].astype(np.int64)] = -float("inf")
    check_topk(v, i, S, k, atol=2e-4 if metric else 2e-5)
    i = i.cpu()
    for r in range(rows.numel()):
        pos = {c: int((i[r] == c).nonzero()[0]) for c in (7, 100, 2000) if (i[r] == c).any()}
        assert list(pos) == sorted(pos, key=pos.get)     # duplicates appear in ascending column order
    # fewer eligible columns than k: padded with -1 / -inf
    tiny_v, tiny_i = E.exact_topk_rows(A.cuda(), rows[:2].cuda(), B[:10].cuda(), 33, metric=metric)
    assert tiny_i.shape == (2, 33) and int((tiny_i >= 0).sum()) == 20 and bool(torch.isinf(tiny_v[:, 10:]).all())
    # thresholded variant (`fr_exact_topk_thr_f32`): a lower bound of the k-th best score per row -> same result, no
    # dense score block; a useless bound (-inf: every column qualifies, the lists overflow) falls back to the dense path
    kth = torch.topk(S, k, dim=1)[0][:, -1]
    for thr in (kth - 1e-3 * kth.abs() - 1e-3, torch.full_like(kth, -float("inf"))):
        v2, i2 = E.exact_topk_rows(A.cuda(), rows.cuda(), B.cuda(), k, scale=0.5 if metric == 0 else 1.0,
                                   bias=None if bias is None else bias.cuda(), metric=metric, hist=hist, thr=thr.cuda())
        assert torch.equal(i2.cpu(), i) and torch.equal(v2, v)

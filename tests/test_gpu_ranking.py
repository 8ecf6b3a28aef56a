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
        E.gemm_topk(torch.randn(4, 64).cuda(), torch.randn(900, 64).cuda(), 60)  # no room for bf16 slack


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

"""GPU parity: propagation kernel (through the C ABI) vs the CPU oracle's `torch.sparse.mm`."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: embeddings within 1e-5 relative in fp32


def close(a, b, rtol=RTOL):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a - b).abs().max())
    assert err <= rtol * scale, (err, scale)


def random_csr(n_rows, n_cols, degs, seed):
    rng = np.random.default_rng(seed)
    rp = np.concatenate([[0], np.cumsum(degs)]).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(n_cols, size=d, replace=d > n_cols)) for d in degs] + [np.zeros(0, int)])
    # ~1/sqrt(deg) magnitudes like a normalised adjacency, so row sums stay O(1)
    val = (rng.standard_normal(int(rp[-1])) / np.sqrt(np.repeat(np.maximum(degs, 1), degs))).astype(np.float32)
    return rp, col.astype(np.int32), val


def to_torch_sparse(rp, col, val, n_rows, n_cols):
    rows = np.repeat(np.arange(n_rows), np.diff(rp))
    return torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, col.astype(np.int64)])), torch.from_numpy(val),
                                   (n_rows, n_cols))


@pytest.mark.parametrize("d", [32, 64, 128])
@pytest.mark.parametrize("case", ["ragged", "long", "huge", "many_rows", "empty", "single"])
def test_spmm_matches_sparse_mm(d, case):
    from foodrec_b200 import graph as G, ops
    rng = np.random.default_rng(1)
    n_cols = 3000
    if case == "ragged":
        degs = rng.integers(0, 40, size=2000)
    elif case == "long":  # rows far beyond one segment, incl. exact multiples of the segment length
        degs = np.array([5000, 128, 129, 256, 0, 1, 2999, 127, 640] + list(rng.integers(0, 300, size=300)))
    elif case == "huge":  # thousands of segments per row: the two-level fold (children of ~sqrt(k) segments + a parent)
        degs = np.array([300000, 3, 70000, 64 * 16, 64 * 16 + 1, 64 * 17] + list(rng.integers(0, 90, size=200)))
    elif case == "many_rows":  # hundreds of thousands of short rows (grid of > 8 000 blocks) plus a few long ones
        degs = np.concatenate([rng.integers(0, 5, size=280000), [700, 64, 65]])
    elif case == "empty":
        degs = np.zeros(257, dtype=np.int64)
    else:
        degs = np.array([1])
    rp, col, val = random_csr(len(degs), n_cols, degs, 2)
    g = G.PropGraph(rp, col, val, n_cols, "cuda")
    X = torch.randn(n_cols, d)
    Z = torch.randn(len(degs), d)
    bias = torch.randn(d)
    S = to_torch_sparse(rp, col, val, len(degs), n_cols)
    ref = torch.sparse.mm(S, X)
    if case == "huge":
        # a sequential fp32 sum of 300 000 terms (torch-CPU's row order) is itself 4e-6 .. 1.2e-5 away from the exact
        # product, so the yardstick here is the fp64 product; the segmented fixed-order fold must sit well inside 1e-5
        ref = torch.sparse.mm(S.double(), X.double()).float()
    close(ops.spmm(g, X.cuda()), ref)
    close(ops.spmm(g, X.cuda(), Z=Z.cuda(), alpha=0.25, beta=0.5), 0.25 * ref + 0.5 * Z)
    close(ops.spmm(g, X.cuda(), bias=bias.cuda(), act=1), torch.tanh(ref + bias))
    # bit-reproducible (long rows are reduced in fixed order, not by arrival)
    a, b = ops.spmm(g, X.cuda()), ops.spmm(g, X.cuda())
    assert torch.equal(a, b)
    assert int(g.counters.abs().sum()) == 0


def test_rejects_bad_arguments():
    from foodrec_b200 import _lib, graph as G, ops
    rp, col, val = random_csr(4, 10, [1, 2, 0, 3], 0)
    g = G.PropGraph(rp, col, val, 10, "cuda")
    with pytest.raises(_lib.FoodRecError):
        ops.spmm(g, torch.randn(10, 48).cuda())           # unsupported width
    with pytest.raises(_lib.FoodRecError):
        ops.spmm(g, torch.randn(9, 64).cuda())            # wrong row count
    with pytest.raises(_lib.FoodRecError):
        ops.spmm(g, torch.randn(10, 64))                  # CPU tensor: no fallback
    with pytest.raises(_lib.FoodRecError):
        ops.spmm(g, torch.randn(10, 64).cuda().double())  # dtype


@pytest.mark.parametrize("layers", [0, 1, 2, 3])
def test_layer_mean_propagation_fwd_bwd(mini_ds, layers):
    from foodrec_b200 import graph as G, ops
    from oracle import adjacency, propagation
    ds = mini_ds
    S = adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items)
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    torch.manual_seed(0)
    ego = (torch.randn(ds.n_users + ds.n_items, 64) * 0.1).requires_grad_(True)
    w = torch.randn(ds.n_users + ds.n_items, 64)
    ref = propagation.layer_mean_propagate(S, ego, layers)
    (ref * w).sum().backward()
    ego_d = ego.detach().cuda().requires_grad_(True)
    out = ops.propagate_mean(g, ego_d, layers)
    (out * w.cuda()).sum().backward()
    close(out, ref)
    close(ego_d.grad, ego.grad)


def test_directed_gcn_propagation_fwd_bwd(mini_ds):
    from foodrec_b200 import graph as G, ops
    from oracle import adjacency, propagation
    ds = mini_ds
    n = ds.n_users + ds.n_items + ds.num_ingredients + ds.num_calories_level
    ei = adjacency.schgn_edge_index(ds)
    src, dst, w = adjacency.gcn_norm_edges(ei, n)
    g = G.gcn_normalised(ei[0].numpy(), ei[1].numpy(), n, "cuda")
    torch.manual_seed(1)
    x = (torch.randn(n, 64) * 0.3).requires_grad_(True)
    W = (torch.randn(64, 64) * 0.2).requires_grad_(True)
    b = (torch.randn(64) * 0.1).requires_grad_(True)
    t = torch.randn(n, 64)
    ref = propagation.gcn_conv_tanh(x, src, dst, w, W, b)
    (ref * t).sum().backward()
    xd, Wd, bd = (v.detach().cuda().requires_grad_(True) for v in (x, W, b))
    out = ops.gcn_propagate_tanh(g, xd @ Wd.t(), bd)
    (out * t.cuda()).sum().backward()
    close(out, ref)
    close(xd.grad, x.grad, 2e-5)
    close(Wd.grad, W.grad, 2e-5)
    close(bd.grad, b.grad, 2e-5)


def test_c1_scale_forward_vs_oracle():
    from foodrec_b200 import graph as G, ops
    from foodrec_b200.synth import make_dataset
    from oracle import adjacency, propagation
    ds = make_dataset("C1", features=False)
    S = adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items)
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    assert g.n_long > 0  # popular items exceed one segment at this scale
    torch.manual_seed(0)
    ego = torch.randn(ds.n_users + ds.n_items, 64) * 0.1
    close(ops.propagate_mean(g, ego.cuda(), 2), propagation.layer_mean_propagate(S, ego, 2))


def test_propgraph_from_reference_style_sparse_tensor(mini_ds):
    """A reference-built adjacency (row-sorted torch COO on the device) becomes a plan through the
    device-side COO -> CSR kernel and propagates like `torch.sparse.mm` on it."""
    from foodrec_b200 import graph as G, ops
    from oracle import adjacency
    ds = mini_ds
    S = adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items)
    g = G.from_torch_sparse(S.cuda())
    g0 = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    assert np.array_equal(g.row_ptr_host, g0.row_ptr_host)
    assert torch.equal(g.col, g0.col) and torch.equal(g.val, g0.val)
    X = torch.randn(S.shape[0], 64)
    close(ops.spmm(g, X.cuda()), torch.sparse.mm(S, X))


@pytest.mark.parametrize("layers", [1, 2, 3])
def test_masked_backward_propagation_is_bit_identical(layers):
    """Skipping the gathers of all-zero gradient rows (row-activity masks) must not change a single bit, and the
    mask the kernel emits for its output must cover every nonzero row."""
    from foodrec_b200 import graph as G, ops
    from foodrec_b200.synth import make_dataset
    ds = make_dataset("C1", features=False)
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cuda")
    N = g.n_rows
    torch.manual_seed(layers)
    grad = torch.zeros(N, 64, device="cuda")
    rows = torch.randint(0, N, (1500,), device="cuda")
    grad[rows] = torch.randn(1500, 64, device="cuda")
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
    mask[rows] = 1
    dense = ops.propagate_mean_raw(g, grad, layers)
    sparse, out_mask = ops.propagate_mean_masked(g, grad, layers, mask)
    assert torch.equal(dense, sparse)
    nz = (dense != 0).any(dim=1)
    assert bool((out_mask.bool() | ~nz).all())          # every nonzero row is marked active
    # a conservative (all-ones) mask is also exact
    sparse2, _ = ops.propagate_mean_masked(g, grad, layers, torch.ones(N, dtype=torch.uint8, device="cuda"))
    assert torch.equal(dense, sparse2)


def test_row_masks_do_not_change_model_gradients(mini_ds, mini_batches):
    """CLUSSL loss + backward with and without the row-mask fast path: same losses, gradients equal up to the
    order of fp32 atomics in the loss kernels."""
    from foodrec_b200 import ops
    from foodrec_b200.models.pricai_modelx import PRICAI_ModelX
    from test_gpu_clussl import cfg_for, dev_batch
    torch.manual_seed(999)
    m = PRICAI_ModelX(cfg_for(mini_ds), mini_ds).to("cuda")
    grads = {}
    for flag in (True, False):
        ops.USE_ROW_MASKS = flag
        try:
            m.zero_grad()
            losses = m.calculate_loss(dev_batch(mini_batches[0]))
            sum(losses).backward()
            grads[flag] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        finally:
            ops.USE_ROW_MASKS = True
    for n in grads[True]:
        a, b = grads[True][n], grads[False][n]
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()) + 1e-12, n


def test_push_epilogue_single_rank_matches_plain_propagation(mini_ds):
    """World-size-1 run of the peer-memory path (`fr_peer_alloc`, `fr_push_rows`, `fr_spmm_csr_f32_push`): the
    pushed copies feed the next layer, so the result must be bit-identical to `propagate_mean`, forward and
    backward, across repeated calls (alternating table pairs).  Multi-rank runs: scripts/dist_propagation_check.py."""
    import os
    import torch.distributed as dist
    from foodrec_b200 import dist as D, graph as G, ops
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        g = G.norm_adj_user_item(mini_ds.train_coo_matrix, mini_ds.n_users, mini_ds.n_items, "cuda")
        pg = D.RowPartitionedGraph.from_graph(g, 0, 1, "cuda")
        tables = D.PeerTables(pg.n_padded, 64, "cuda")
        torch.manual_seed(3)
        ego = torch.randn(g.n_rows, 64, device="cuda") * 0.1
        w = torch.randn(g.n_rows, 64, device="cuda")
        for layers in (1, 2, 3):
            a = ego.clone().requires_grad_(True)
            b = ego.clone().requires_grad_(True)
            ref = ops.propagate_mean(g, a, layers)
            (ref * w).sum().backward()
            got = D.propagate_mean_pushed(pg, b, layers, tables)
            (got * w).sum().backward()
            assert torch.equal(got, ref), layers
            assert torch.equal(b.grad, a.grad), layers
        tables.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_row_collectives_through_the_c_abi_single_rank(mini_ds):
    """`fr_comm_*` / `fr_allgather_rows` / `fr_reduce_scatter_rows` (NCCL bound at run time) with a one-rank
    communicator: the partitioned propagation that exchanges through `dist.RowComm` equals `propagate_mean` bit for bit,
    forward and backward, and the two collectives are the identity.  Multi-rank: scripts/dist_propagation_check.py."""
    from foodrec_b200 import _lib, dist as D, graph as G, ops
    assert _lib.lib.fr_comm_version() >= 21000
    comm = D.RowComm(device="cuda")
    assert (comm.rank, comm.world) == (0, 1)
    try:
        x = torch.randn(37, 64, device="cuda")
        assert torch.equal(comm.all_gather_rows(x), x)
        assert torch.equal(comm.reduce_scatter_rows(x), x)
        g = G.norm_adj_user_item(mini_ds.train_coo_matrix, mini_ds.n_users, mini_ds.n_items, "cuda")
        pg = D.RowPartitionedGraph.from_graph(g, 0, 1, "cuda")
        torch.manual_seed(5)
        ego = torch.randn(g.n_rows, 64, device="cuda") * 0.1
        w = torch.randn(g.n_rows, 64, device="cuda")
        a, b = ego.clone().requires_grad_(True), ego.clone().requires_grad_(True)
        ref = ops.propagate_mean(g, a, 2)
        (ref * w).sum().backward()
        got = D.propagate_mean_partitioned(pg, b, 2, group=comm)
        (got * w).sum().backward()
        assert torch.equal(got, ref) and torch.equal(b.grad, a.grad)
        full = D.gather_rows_autograd(b, comm)          # all-gather whose backward is the reduce-scatter
        assert torch.equal(full, b)
    finally:
        comm.close()


def test_grouped_launch_is_bit_identical_to_separate_launches():
    """`fr_spmm_csr_f32_grouped`: three graphs of different sizes (ragged, long rows, an empty one among them) in one
    grid, two-segment operands, fused `+ beta Z`; forward and backward through the layer-mean propagation."""
    from foodrec_b200 import graph as G, ops
    rng = np.random.default_rng(7)
    specs = [(900, list(rng.integers(0, 30, size=900))),
             (400, [3000, 129, 700] + list(rng.integers(0, 200, size=397))),
             (64, [0] * 64)]
    graphs, tops, bottoms = [], [], []
    for k, (n, degs) in enumerate(specs):
        degs = np.asarray(degs)
        rp, col, val = random_csr(n, n, degs, 10 + k)
        gt = G.PropGraph(*_transpose_csr(rp, col, val, n), n, "cuda")
        g = G.PropGraph(rp, col, val, n, "cuda", transpose=gt)
        gt.T = g
        graphs.append(g)
        n_top = n // 3
        tops.append(torch.randn(n_top, 64, device="cuda", requires_grad=True))
        bottoms.append(torch.randn(n - n_top, 64, device="cuda", requires_grad=True))
    grp = ops.PropGroup(graphs)
    assert grp.n_blocks == sum(-(-g.n_seg // 32) for g in graphs)
    outs = ops.propagate_mean_grouped(grp, tops, bottoms, 2)
    w = [torch.randn_like(o) for o in outs]
    sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
    got = [(t.grad.clone(), b.grad.clone()) for t, b in zip(tops, bottoms)]
    for t, b in zip(tops, bottoms):
        t.grad = b.grad = None
    prev = ops.USE_ROW_MASKS
    ops.USE_ROW_MASKS = False
    try:
        ref = [ops.propagate_mean(g, t, 2, bottom=b) for g, t, b in zip(graphs, tops, bottoms)]
        sum((o * wi).sum() for o, wi in zip(ref, w)).backward()
    finally:
        ops.USE_ROW_MASKS = prev
    for o, r in zip(outs, ref):
        assert torch.equal(o, r)
    for (gt_, gb_), t, b in zip(got, tops, bottoms):
        assert torch.equal(gt_, t.grad) and torch.equal(gb_, b.grad)
    assert all(int(g.counters.abs().sum()) == 0 for g in graphs)


def _transpose_csr(rp, col, val, n):
    import scipy.sparse as sp
    m = sp.csr_matrix((val, col, rp), shape=(n, n)).T.tocsr()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float32)

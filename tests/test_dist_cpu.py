"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row partitioning + per-layer all-gather
reproduce the single-rank propagation (forward and backward), user-sharded ranking reproduces the
single-rank top-K.  The device kernels are replaced by torch-CPU stand-ins injected through the
`spmm=` / `topk_fn=` hooks -- the collective plumbing and the partition arithmetic are what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import foodrec_b200  # noqa: F401


class HostGraph:
    """CPU stand-in for PropGraph (same constructor contract), used only by these tests."""

    def __init__(self, row_ptr, col, val, n_cols, device, transpose=None):
        self.n_rows = len(row_ptr) - 1
        rows = np.repeat(np.arange(self.n_rows), np.diff(row_ptr))
        self.S = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, np.asarray(col, dtype=np.int64)])),
                                         torch.from_numpy(np.asarray(val, dtype=np.float32)), (self.n_rows, n_cols))
        self.T = self


def host_spmm(graph, x_full, z, alpha, beta):
    return alpha * torch.sparse.mm(graph.S, x_full) + beta * z


def host_topk(user_all, item_all, users, k, hist=None):
    return torch.topk(user_all[users] @ item_all.t(), k, dim=-1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_layers, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from foodrec_b200 import dist as D, graph as G
        from foodrec_b200.synth import make_dataset
        ds = make_dataset("mini")
        g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
        N = g.n_rows
        pg = D.RowPartitionedGraph(g.row_ptr_host, g.col.numpy(), g.val.numpy(), N, rank, world, "cpu", graph_cls=HostGraph)
        torch.manual_seed(0)
        ego = torch.randn(N, 64) * 0.1
        w = torch.randn(N, 64)
        ego_l = pg.local_rows(ego).requires_grad_(True)
        res = D.propagate_mean_partitioned(pg, ego_l, n_layers, spmm=host_spmm)
        (res * pg.local_rows(w)).sum().backward()
        users = torch.arange(ds.n_users)
        full = torch.randn(N, 64, generator=torch.Generator().manual_seed(1))
        top = D.full_sort_topk_sharded(full[:ds.n_users], full[ds.n_users:], users, 10, topk_fn=host_topk)
        out[rank] = (pg.lo, pg.hi, res.detach()[:pg.hi - pg.lo].clone(), ego_l.grad[:pg.hi - pg.lo].clone(), top.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_layers", [1, 3])
def test_row_partitioned_propagation_and_sharded_eval_world2(n_layers):
    from foodrec_b200 import graph as G
    from foodrec_b200.synth import make_dataset
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_layers, out), nprocs=world, join=True)
    ds = make_dataset("mini")
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
    N = g.n_rows
    S = HostGraph(g.row_ptr_host, g.col.numpy(), g.val.numpy(), N, "cpu").S
    torch.manual_seed(0)
    ego = (torch.randn(N, 64) * 0.1).requires_grad_(True)
    w = torch.randn(N, 64)
    from oracle import propagation
    ref = propagation.layer_mean_propagate(S, ego, n_layers)
    (ref * w).sum().backward()
    covered = 0
    for r in range(world):
        lo, hi, res, grad, top = out[r]
        assert torch.allclose(res, ref.detach()[lo:hi], rtol=1e-5, atol=1e-7)
        assert torch.allclose(grad, ego.grad[lo:hi], rtol=1e-5, atol=1e-7)
        covered += hi - lo
    assert covered == N
    full = torch.randn(N, 64, generator=torch.Generator().manual_seed(1))
    ref_top = torch.topk(full[:ds.n_users] @ full[ds.n_users:].t(), 10, dim=-1)[1]
    assert torch.equal(out[0][4], ref_top) and torch.equal(out[1][4], ref_top)


def test_shard_arithmetic():
    from foodrec_b200 import dist as D
    assert D.shard_rows(10, 4) == (3, 12)
    assert D.shard_rows(8, 4) == (2, 8)
    u = torch.arange(10)
    parts = [D.shard_users(u, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), u) and [p.numel() for p in parts] == [3, 3, 3, 1]


def _host_losses(full, ego_full, u, p, n, n_users, reg_weight):
    from oracle import losses
    mf = losses.bpr_from_tables(full[:n_users], full[n_users:], u, p, n)
    reg = losses.emb_loss(ego_full[u], ego_full[n_users + p], ego_full[n_users + n])
    return mf, reg_weight * reg


def _train_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from foodrec_b200 import dist as D, graph as G
        from foodrec_b200.synth import make_dataset, sample_train_batches
        ds = make_dataset("mini")
        g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
        pg = D.RowPartitionedGraph(g.row_ptr_host, g.col.numpy(), g.val.numpy(), g.n_rows, rank, world, "cpu",
                                   graph_cls=HostGraph)
        torch.manual_seed(0)
        ego = torch.randn(g.n_rows, 64) * 0.1
        ego_l = pg.local_rows(ego).requires_grad_(True)
        b = sample_train_batches(ds, 48, world, seed=4)[rank]                  # every rank its own mini-batch
        batch = {k: torch.from_numpy(b[k]) for k in ("u_id", "pos_i_id", "neg_i_id")}
        losses = D.partitioned_bpr_losses(
            pg, ego_l, ds.n_users, 2, batch, 0.1, spmm=host_spmm,
            loss_fn=lambda full, ego_full, u, p, n: _host_losses(full, ego_full, u, p, n, ds.n_users, 0.1))
        (sum(losses) / world).backward()
        out[rank] = (pg.lo, pg.hi, [float(x) for x in losses], ego_l.grad[:pg.hi - pg.lo].clone())
    finally:
        dist.destroy_process_group()


def test_row_partitioned_training_step_world2():
    """Losses and the owner's parameter gradient of the row-partitioned data-parallel step (all-gather forward,
    reduce-scatter backward, per-rank mini-batches) equal the single-process mean over the two batches."""
    from foodrec_b200 import graph as G
    from foodrec_b200.synth import make_dataset, sample_train_batches
    from oracle import propagation
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ds = make_dataset("mini")
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
    S = HostGraph(g.row_ptr_host, g.col.numpy(), g.val.numpy(), g.n_rows, "cpu").S
    torch.manual_seed(0)
    ego = (torch.randn(g.n_rows, 64) * 0.1).requires_grad_(True)
    full = propagation.layer_mean_propagate(S, ego, 2)
    total, per_rank = 0.0, []
    for b in sample_train_batches(ds, 48, world, seed=4):
        u, p, n = (torch.from_numpy(b[k]) for k in ("u_id", "pos_i_id", "neg_i_id"))
        mf, reg = _host_losses(full, ego, u, p, n, ds.n_users, 0.1)
        per_rank.append([float(mf), float(reg)])
        total = total + (mf + reg) / world
    total.backward()
    covered = 0
    for r in range(world):
        lo, hi, losses, grad = out[r]
        np.testing.assert_allclose(losses, per_rank[r], rtol=1e-5)
        assert torch.allclose(grad, ego.grad[lo:hi], rtol=1e-4, atol=1e-8)
        covered += hi - lo
    assert covered == g.n_rows


def _worker_interleaved(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from foodrec_b200 import dist as D, graph as G
        from foodrec_b200.synth import make_dataset
        ds = make_dataset("mini")
        g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, "cpu")
        N = g.n_rows
        pg = D.RowPartitionedGraph(g.row_ptr_host, g.col.numpy(), g.val.numpy(), N, rank, world, "cpu", graph_cls=HostGraph,
                                   balance="interleave")
        torch.manual_seed(0)
        ego = torch.randn(N, 64) * 0.1
        w = torch.randn(N, 64)
        ego_l = pg.local_rows(ego).requires_grad_(True)
        res = D.propagate_mean_partitioned(pg, ego_l, 2, spmm=host_spmm)
        (res * pg.local_rows(w)).sum().backward()
        full_res = pg.to_original(D._all_gather_rows(res.detach()))
        full_grad = pg.to_original(D._all_gather_rows(ego_l.grad))
        out[rank] = (int(pg.local.S._nnz()), full_res.clone(), full_grad.clone())
    finally:
        dist.destroy_process_group()


def test_interleaved_row_partition_balances_entries_and_reproduces_single_rank():
    """`balance="interleave"`: rank p owns rows p, p + P, ...; both ranks hold about half of the stored entries (the
    contiguous split of a users-then-items numbering does not) and the gathered result equals the single-rank one."""
    from foodrec_b200 import graph as G
    from foodrec_b200.synth import make_dataset
    from oracle import adjacency, propagation
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_interleaved, args=(world, _free_port(), out), nprocs=world, join=True)
    ds = make_dataset("mini")
    S = adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items)
    N = ds.n_users + ds.n_items
    torch.manual_seed(0)
    ego = (torch.randn(N, 64) * 0.1).requires_grad_(True)
    w = torch.randn(N, 64)
    ref = propagation.layer_mean_propagate(S, ego, 2)
    (ref * w).sum().backward()
    nnz = [out[r][0] for r in range(world)]
    assert abs(nnz[0] - nnz[1]) <= 0.1 * sum(nnz), nnz
    for r in range(world):
        assert torch.allclose(out[r][1], ref.detach(), rtol=1e-5, atol=1e-7)
        assert torch.allclose(out[r][2], ego.grad, rtol=1e-5, atol=1e-7)

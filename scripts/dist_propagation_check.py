"""torchrun --nproc-per-node N scripts/dist_propagation_check.py [scale]
Row-partitioned propagation (one NCCL all-gather per layer, fwd + bwd) on real GPUs: every rank checks its
row block against the single-GPU kernel result, then the step is timed (max over ranks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import foodrec_b200  # noqa
from foodrec_b200 import dist as D, graph as G, ops
from foodrec_b200.synth import make_dataset


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    scale = sys.argv[1] if len(sys.argv) > 1 else "C2"
    layers = 3
    ds = make_dataset(scale, features=False)
    g = G.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items, dev)
    pg = D.RowPartitionedGraph.from_graph(g, rank, world, dev)
    torch.manual_seed(0)
    ego = torch.randn(g.n_rows, 64, device=dev) * 0.1
    w = torch.randn(g.n_rows, 64, device=dev)
    # single-GPU result (every rank computes it: the graph is small enough)
    e1 = ego.clone().requires_grad_(True)
    ref = ops.propagate_mean(g, e1, layers)
    (ref * w).sum().backward()
    el = pg.local_rows(ego).requires_grad_(True)
    out = D.propagate_mean_partitioned(pg, el, layers)
    (out * pg.local_rows(w)).sum().backward()
    n = pg.hi - pg.lo
    err_f = float((out[:n] - ref[pg.lo:pg.hi]).abs().max() / ref.abs().max())
    err_b = float((el.grad[:n] - e1.grad[pg.lo:pg.hi]).abs().max() / e1.grad.abs().max())
    ok = torch.tensor([float(err_f < 1e-6 and err_b < 1e-6)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)

    # the same exchange through this library's C ABI (fr_allgather_rows / fr_reduce_scatter_rows on its own communicator)
    comm = D.RowComm(device=dev)
    ec = pg.local_rows(ego).requires_grad_(True)
    outc = D.propagate_mean_partitioned(pg, ec, layers, group=comm)
    (outc * pg.local_rows(w)).sum().backward()
    abi_same = torch.tensor([float(torch.equal(outc, out) and torch.equal(ec.grad, el.grad))], device=dev)
    dist.all_reduce(abi_same, op=dist.ReduceOp.MIN)

    # the same propagation with the exchange fused into the kernel's epilogue (peer memory, no all-gather)
    tables = D.PeerTables(pg.n_padded, 64, dev)
    ep = pg.local_rows(ego).requires_grad_(True)
    outp = D.propagate_mean_pushed(pg, ep, layers, tables)
    (outp * pg.local_rows(w)).sum().backward()
    push_same = torch.tensor([float(torch.equal(outp, out) and torch.equal(ep.grad, el.grad))], device=dev)
    dist.all_reduce(push_same, op=dist.ReduceOp.MIN)

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def step():
        el.grad = None
        o = D.propagate_mean_partitioned(pg, el, layers)
        (o * o).sum().backward()

    def step_push():
        ep.grad = None
        o = D.propagate_mean_pushed(pg, ep, layers, tables)
        (o * o).sum().backward()

    def step_abi():
        ec.grad = None
        o = D.propagate_mean_partitioned(pg, ec, layers, group=comm)
        (o * o).sum().backward()
    t_gather, t_push, t_abi = timed(step), timed(step_push), timed(step_abi)

    # row-partitioned data-parallel training step (per-rank mini-batches, all-gather forward / reduce-scatter
    # backward around the fused BPR + EmbLoss kernel) against the same step computed on one GPU
    from foodrec_b200.synth import sample_train_batches
    hb = sample_train_batches(ds, 512, world, seed=4)
    batches = [{k: torch.from_numpy(b[k]).to(dev) for k in ("u_id", "pos_i_id", "neg_i_id")} for b in hb]
    pe = pg.local_rows(ego).requires_grad_(True)
    losses = D.partitioned_bpr_losses(pg, pe, ds.n_users, layers, batches[rank], 0.1, group=comm)   # C-ABI collectives
    (sum(losses) / world).backward()
    e2 = ego.clone().requires_grad_(True)
    full = ops.propagate_mean(g, e2, layers)
    tot, mine = 0.0, None
    for r, b in enumerate(batches):
        mf, reg = ops.rank_loss(full, ds.n_users, b["u_id"], b["pos_i_id"], b["neg_i_id"],
                                [(e2, b["u_id"], None), (e2, b["pos_i_id"] + ds.n_users, None),
                                 (e2, b["neg_i_id"] + ds.n_users, None)], 512.0)
        if r == rank:
            mine = (float(mf), 0.1 * float(reg))
        tot = tot + (mf + 0.1 * reg) / world
    tot.backward()
    gref = e2.grad[pg.lo:pg.hi]
    err_l = max(abs(float(losses[0]) - mine[0]) / abs(mine[0]), abs(float(losses[1]) - mine[1]) / abs(mine[1]))
    err_g = float((pe.grad[:n] - gref).abs().max() / gref.abs().max())
    train_ok = torch.tensor([float(err_l < 1e-5 and err_g < 1e-5)], device=dev)
    dist.all_reduce(train_ok, op=dist.ReduceOp.MIN)

    def train_step():
        pe.grad = None
        ls = D.partitioned_bpr_losses(pg, pe, ds.n_users, layers, batches[rank], 0.1)
        (sum(ls) / world).backward()
    t_train = timed(train_step)
    if rank == 0:
        print(json.dumps({"world": world, "scale": scale, "layers": layers, "N": g.n_rows, "nnz": g.nnz,
                          "rel_err_fwd": err_f, "rel_err_bwd": err_b, "all_ranks_ok": bool(ok.item()),
                          "fwd_bwd_ms_max_over_ranks": t_gather,
                          "push_bit_identical_to_all_gather_path": bool(push_same.item()),
                          "push_fwd_bwd_ms_max_over_ranks": t_push,
                          "c_abi_collectives_bit_identical_to_torch_distributed": bool(abi_same.item()),
                          "c_abi_fwd_bwd_ms_max_over_ranks": t_abi,
                          "partitioned_train_step": {"loss_rel_err": err_l, "grad_rel_err": err_g,
                                                     "all_ranks_ok": bool(train_ok.item()), "ms_max_over_ranks": t_train}}))
    tables.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

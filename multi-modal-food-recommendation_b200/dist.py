"""Multi-GPU: row-partitioned propagation (one all-gather per layer) and user-sharded evaluation.

One process per GPU, `torch.distributed` (NCCL over NVLink) for the plumbing; the per-layer exchange itself is also
available behind the C ABI (`RowComm`: `fr_allgather_rows` / `fr_reduce_scatter_rows` on this library's own communicator).  The reference is a
single-process, single-GPU program, so there is no reference behaviour to match
here: correctness is "N ranks reproduce the 1-rank result" (tests/test_dist_cpu.py, gloo).

Propagation.  Rows of the normalised adjacency S -- and of every layer's embeddings -- are split into
`world` equal blocks (the last one zero-padded).  Rank p owns rows [p*R, (p+1)*R) of S as its own CSR
whose columns index the FULL node set; a layer is `all_gather(X_local) -> local SpMM`.  For the
symmetric graphs the backward is the same thing on the gathered upstream gradient
(`dX_p = S_p . all_gather(dY)`), i.e. one all-gather per layer in each direction and no reduce-scatter.
The layer-mean epilogue (Horner form, see ops.py) is row-local.

`propagate_mean_pushed` is the same computation with the exchange fused into the kernel: the SpMM's
epilogue stores each finished row into a full-size table on EVERY rank (CUDA-IPC peer memory over
NVLink / NVSwitch), so the next layer's input is complete when the kernels end and the transfer
overlaps the gathers; a 4-byte all-reduce orders the ranks.  Results are bit-identical to the
all-gather path (same local kernel, same inputs).

Evaluation.  Users are sharded across ranks, the item table is replicated; each rank runs the fused
score + top-K kernel on its users and the `[U/P, k]` results are gathered once at the end.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .graph import PropGraph


def shard_rows(n_rows: int, world: int):
    """Equal row blocks (`rows_per_rank`, padded total)."""
    r = -(-n_rows // world)
    return r, r * world


class RowPartitionedGraph:
    """This rank's row block of a square normalised adjacency given as CSR arrays (host numpy or torch tensors).

    `balance="block"`: rank p owns the contiguous rows [p R, (p + 1) R).  With users numbered before items the blocks
    hold very different numbers of stored entries (item rows are ~5x longer than user rows on the C5-shaped graph: the
    slowest rank had 70 % of them at N = 2).  `balance="interleave"` relabels the nodes `new = (old % P) R + old // P`
    -- rank p owns the rows p, p + P, p + 2P, ... -- so users and items, short rows and long rows, are dealt round-robin
    and every rank gets ~1/P of the entries; the all-gather still moves equal contiguous blocks.  `local_rows` /
    `to_original` convert between the original numbering and a rank's block / the gathered table."""

    def __init__(self, row_ptr, col, val, n_nodes: int, rank: int, world: int, device, symmetric: bool = True,
                 graph_cls=PropGraph, balance: str = "block"):
        if balance not in ("block", "interleave"):
            raise ValueError("balance must be 'block' or 'interleave'")
        self.n_nodes, self.rank, self.world, self.balance = n_nodes, rank, world, balance
        self.rows_per_rank, self.n_padded = shard_rows(n_nodes, world)
        R = self.rows_per_rank
        rp = np.asarray(row_ptr, dtype=np.int64)
        if balance == "block":
            lo = min(rank * R, n_nodes)
            hi = min(lo + R, n_nodes)
            self.lo, self.hi = lo, hi
            local_rp = np.full(R + 1, rp[hi] - rp[lo], dtype=np.int64)
            local_rp[:hi - lo + 1] = rp[lo:hi + 1] - rp[lo]
            lcol, lval = col[rp[lo]:rp[hi]], val[rp[lo]:rp[hi]]
        else:
            rows = np.arange(rank, n_nodes, world, dtype=np.int64)          # original ids of this rank's rows
            self.lo, self.hi = 0, rows.size
            deg = rp[rows + 1] - rp[rows]
            local_rp = np.full(R + 1, int(deg.sum()), dtype=np.int64)
            local_rp[0] = 0
            np.cumsum(deg, out=local_rp[1:rows.size + 1])
            tcol = col if torch.is_tensor(col) else torch.from_numpy(np.ascontiguousarray(col))
            tval = val if torch.is_tensor(val) else torch.from_numpy(np.ascontiguousarray(val))
            dv = tcol.device
            t_deg = torch.from_numpy(deg).to(dv)
            shift = torch.from_numpy(rp[rows] - local_rp[:rows.size]).to(dv)       # global offset - local offset, per row
            elem = torch.repeat_interleave(shift, t_deg) + torch.arange(int(local_rp[-1]), device=dv)
            c = tcol[elem].to(torch.int64)
            lcol = ((c % world) * R + c // world).to(torch.int32)                    # position in the gathered table
            lval = tval[elem]
            if not torch.is_tensor(col):
                lcol, lval = lcol.numpy(), lval.numpy()
        # columns index the all-gathered [n_padded, d] table; padding rows are empty
        self.local = graph_cls(local_rp, lcol, lval, self.n_padded, device, transpose="self" if symmetric else None)
        self.symmetric = symmetric

    @classmethod
    def from_graph(cls, g: PropGraph, rank: int, world: int, device, **kw):
        return cls(g.row_ptr_host, g.col.cpu().numpy(), g.val.cpu().numpy(), g.n_rows, rank, world, device, **kw)

    def local_rows(self, full: torch.Tensor) -> torch.Tensor:
        """This rank's (padded) row block of a full `[n_nodes, d]` table in the ORIGINAL numbering."""
        out = torch.zeros((self.rows_per_rank, full.shape[1]), dtype=full.dtype, device=full.device)
        if self.balance == "block":
            out[:self.hi - self.lo] = full[self.lo:self.hi]
        else:
            out[:self.hi] = full[self.rank::self.world]
        return out

    def to_original(self, gathered: torch.Tensor) -> torch.Tensor:
        """`[n_nodes, d]` table in the original numbering from the all-gathered `[n_padded, d]` blocks."""
        if self.balance == "block":
            return gathered[:self.n_nodes]
        R, P = self.rows_per_rank, self.world
        return gathered.view(P, R, -1).transpose(0, 1).reshape(P * R, -1)[:self.n_nodes]


class RowComm:
    """This library's own NCCL communicator for the row-block collectives of the partitioned propagation
    (`fr_comm_init`, `fr_allgather_rows`, `fr_reduce_scatter_rows`: the C-ABI form of the one exchange per layer, SURVEY.md
    8b / 8e).  Built from the ranks of a `torch.distributed` group -- rank 0's 128-byte NCCL unique id travels through
    that group once -- or stand-alone for a single process (`world == 1`).  Pass it wherever this module takes `group`."""

    def __init__(self, group=None, device=None):
        import ctypes as C
        from . import _lib
        self._lib = _lib
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.FoodRecError("RowComm needs a CUDA device (NCCL); CPU groups use torch.distributed (gloo) directly")
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(_lib.lib.fr_comm_unique_id(uid), "fr_comm_unique_id")
        if self.world > 1:
            box = [uid.raw]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = C.create_string_buffer(box[0], 128)
        comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.fr_comm_init(uid, self.rank, self.world, C.byref(comm)), "fr_comm_init")
        self._comm = comm

    def all_gather_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        x_local = x_local.contiguous()
        if not (x_local.is_cuda and x_local.dtype == torch.float32 and x_local.dim() == 2):
            raise self._lib.FoodRecError("RowComm.all_gather_rows: expected a CUDA float32 [rows, d] block")
        full = torch.empty((x_local.shape[0] * self.world, x_local.shape[1]), dtype=torch.float32, device=x_local.device)
        self._lib.check(self._lib.lib.fr_allgather_rows(self._comm, x_local.data_ptr(), x_local.shape[0], x_local.shape[1],
                                                        full.data_ptr(), self._lib.stream_ptr()), "fr_allgather_rows")
        return full

    def reduce_scatter_rows(self, g_full: torch.Tensor) -> torch.Tensor:
        g_full = g_full.contiguous()
        rows = g_full.shape[0] // self.world
        if not (g_full.is_cuda and g_full.dtype == torch.float32 and g_full.dim() == 2 and rows * self.world == g_full.shape[0]):
            raise self._lib.FoodRecError("RowComm.reduce_scatter_rows: expected a CUDA float32 [world * rows, d] table")
        out = torch.empty((rows, g_full.shape[1]), dtype=torch.float32, device=g_full.device)
        self._lib.check(self._lib.lib.fr_reduce_scatter_rows(self._comm, g_full.data_ptr(), rows, g_full.shape[1],
                                                             out.data_ptr(), self._lib.stream_ptr()), "fr_reduce_scatter_rows")
        return out

    def close(self):
        if self._comm is not None and self._comm.value:
            torch.cuda.synchronize(self.device)
            self._lib.check(self._lib.lib.fr_comm_destroy(self._comm), "fr_comm_destroy")
        self._comm = None


def _all_gather_rows(x_local: torch.Tensor, group=None) -> torch.Tensor:
    if isinstance(group, RowComm):
        return group.all_gather_rows(x_local)
    world = dist.get_world_size(group)
    full = torch.empty((x_local.shape[0] * world, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(full, x_local.contiguous(), group=group)
    return full


def propagate_mean_partitioned_raw(pg: RowPartitionedGraph, ego_local: torch.Tensor, n_layers: int, spmm, group=None):
    """`mean_l S^l ego` for this rank's rows; `spmm(graph, X_full, Z, alpha, beta)` is the local kernel."""
    if n_layers == 0:
        return ego_local.clone()
    inv = 1.0 / (n_layers + 1)
    t = ego_local
    for layer in range(n_layers):
        last = layer == n_layers - 1
        x_full = _all_gather_rows(t, group)          # the one exchange step of the layer
        t = spmm(pg.local, x_full, ego_local, inv if last else 1.0, inv if last else 1.0)
    return t


class _PartitionedPropagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego_local, pg, n_layers, spmm, group):
        ctx.pg, ctx.n_layers, ctx.spmm, ctx.group = pg, n_layers, spmm, group
        return propagate_mean_partitioned_raw(pg, ego_local.contiguous(), n_layers, spmm, group)

    @staticmethod
    def backward(ctx, g):
        if not ctx.pg.symmetric:
            raise RuntimeError("row-partitioned backward needs a symmetric graph (S^T = S)")
        return propagate_mean_partitioned_raw(ctx.pg, g.contiguous(), ctx.n_layers, ctx.spmm, ctx.group), None, None, None, None


def _device_spmm(graph, x_full, z, alpha, beta):
    from . import ops
    return ops.spmm(graph, x_full, Z=z, alpha=alpha, beta=beta)


def propagate_mean_partitioned(pg: RowPartitionedGraph, ego_local: torch.Tensor, n_layers: int, group=None, spmm=None):
    """Differentiable row-partitioned layer-mean propagation (this rank's rows in, this rank's rows out)."""
    return _PartitionedPropagate.apply(ego_local, pg, n_layers, spmm or _device_spmm, group)


# ----------------------------------------------------------------------- row-partitioned training step
class _GatherRowsAutograd(torch.autograd.Function):
    """`all_gather` of row blocks whose backward is the matching reduce-scatter (sum over ranks): every rank
    reads rows of the full table for its own mini-batch, the row's owner receives the summed gradient."""

    @staticmethod
    def forward(ctx, x_local, group):
        ctx.group, ctx.rows = group, x_local.shape[0]
        return _all_gather_rows(x_local, group)

    @staticmethod
    def backward(ctx, g_full):
        g_full = g_full.contiguous()
        if isinstance(ctx.group, RowComm):
            return ctx.group.reduce_scatter_rows(g_full), None
        out = torch.empty((ctx.rows, g_full.shape[1]), dtype=g_full.dtype, device=g_full.device)
        if dist.get_backend(ctx.group) == "gloo":          # gloo has no reduce-scatter: all-reduce and slice
            dist.all_reduce(g_full, group=ctx.group)
            r = dist.get_rank(ctx.group)
            out.copy_(g_full[r * ctx.rows:(r + 1) * ctx.rows])
        else:
            dist.reduce_scatter_tensor(out, g_full, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None


def gather_rows_autograd(x_local: torch.Tensor, group=None) -> torch.Tensor:
    return _GatherRowsAutograd.apply(x_local, group)


def partitioned_bpr_losses(pg: RowPartitionedGraph, ego_local: torch.Tensor, n_users: int, n_layers: int, batch: dict,
                           reg_weight: float, group=None, spmm=None, loss_fn=None, propagate=None):
    """One data-parallel training step's losses on a ROW-PARTITIONED embedding table (the
    LightGCN / CLUSSL user-item tower at graphs too large for one GPU): this rank owns `ego_local`
    (`[rows_per_rank, d]`, users first then items in the global numbering) and a mini-batch of global ids.

      propagate (one exchange per layer, row-local layer mean)            -> out_local
      all-gather out_local and ego_local once (backward: reduce-scatter)   -> full tables for the batch gathers
      BPR on the propagated rows + EmbLoss on the ego rows of THIS rank's batch (`fr_rank_loss_*`)

    Returns `(mf_loss, reg_loss)` of this rank's batch; the caller backpropagates `sum(losses) / world`, which
    makes the parameter gradient the mean over ranks -- the same semantics as the replicated data-parallel
    step of `bench.py` (`train.OverlappedGradAllReduce`).  `loss_fn(full, ego_full, u, p, n)` and `spmm` are
    injectable for the CPU (gloo) tests."""
    propagate = propagate or (lambda e: propagate_mean_partitioned(pg, e, n_layers, group, spmm))
    out_local = propagate(ego_local)
    full = gather_rows_autograd(out_local, group)[:pg.n_nodes]
    ego_full = gather_rows_autograd(ego_local, group)[:pg.n_nodes]
    u, p, n = batch["u_id"], batch["pos_i_id"], batch["neg_i_id"]
    if loss_fn is not None:
        return loss_fn(full, ego_full, u, p, n)
    from . import ops
    mf, reg = ops.rank_loss(full.contiguous(), n_users, u, p, n,
                            [(ego_full.contiguous(), u, None), (ego_full.contiguous(), p + n_users, None),
                             (ego_full.contiguous(), n + n_users, None)], float(u.numel()))
    return mf, reg_weight * reg


# ------------------------------------------------------- propagation with the exchange fused into the kernel
class _DevicePtr:
    """Raw device allocation exposed through `__cuda_array_interface__` so torch can view it."""

    def __init__(self, ptr, rows, d):
        self.__cuda_array_interface__ = {"shape": (rows, d), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerTables:
    """Four full-size `[n_padded, d]` fp32 tables on every rank of the node (two per call, calls
    alternate between the two pairs), each mapped into every other rank with CUDA IPC (`fr_peer_alloc` /
    `fr_peer_open`).  `peers[b]` is the ctypes array of table b's address on ranks 0..world-1 (this
    rank's own copy included), the destination list of the push epilogue; `table(b)` views this rank's
    copy.  Alternating pairs make an end-of-call barrier unnecessary: a rank can only start writing
    pair A again after every rank has passed the barriers of the call that used pair B, i.e. after every
    rank has finished reading pair A."""

    def __init__(self, n_padded: int, d: int, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self.group = _lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise _lib.FoodRecError("peer tables span one node: at most 8 ranks")
        self.n_padded, self.d, self.device = n_padded, d, torch.device(device)
        self.own, handles = [], []
        for _ in range(4):
            ptr, h = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(_lib.lib.fr_peer_alloc(n_padded * d * 4, C.byref(ptr), h), "fr_peer_alloc")
            self.own.append(ptr.value)
            handles.append(h.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        self.opened, self.peers = [], []
        for b in range(4):
            addrs = []
            for q in range(self.world):
                if q == self.rank:
                    addrs.append(self.own[b])
                    continue
                ptr = C.c_void_p()
                _lib.check(_lib.lib.fr_peer_open(everyone[q][b], C.byref(ptr)), "fr_peer_open")
                self.opened.append(ptr.value)
                addrs.append(ptr.value)
            self.peers.append((C.c_void_p * self.world)(*addrs))
        self._views = [torch.as_tensor(_DevicePtr(p, n_padded, d), device=self.device) for p in self.own]
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._calls = 0
        self.barrier()

    def next_pair(self) -> int:
        """First table index of the pair the next propagation call uses (ranks call in lockstep)."""
        self._calls += 1
        return 2 * (self._calls & 1)

    def table(self, b: int) -> torch.Tensor:
        return self._views[b]

    def barrier(self):
        """Stream-ordered barrier between ranks: a 4-byte all-reduce (the collective library does the
        waiting; no kernel of this package spins on another rank)."""
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        torch.cuda.synchronize(self.device)
        self.barrier()
        torch.cuda.synchronize(self.device)
        self._views = []
        for p in self.opened:
            self._lib.check(self._lib.lib.fr_peer_close(p), "fr_peer_close")
        self.opened = []
        dist.barrier(group=self.group)
        for p in self.own:
            self._lib.check(self._lib.lib.fr_peer_free(p), "fr_peer_free")
        self.own = []


def propagate_mean_pushed_raw(pg: RowPartitionedGraph, ego_local: torch.Tensor, n_layers: int, tables: PeerTables):
    """`mean_l S^l ego` for this rank's rows with the per-layer exchange fused into the producing kernel:
    layer l's epilogue stores its output rows into table (l+1) % 2 of every rank (`fr_spmm_csr_f32_push`),
    so there is no all-gather; the layer-0 input is exchanged by `fr_push_rows`.  A 4-byte all-reduce orders
    the ranks after every exchange (L per call)."""
    from . import _lib, ops
    L = _lib.lib
    if n_layers == 0:
        return ego_local.clone()
    g, d = pg.local, ego_local.shape[1]
    if (tables.n_padded, tables.d) != (pg.n_padded, d):
        raise _lib.FoodRecError("peer tables do not match the partitioned graph")
    ego_local = ego_local.contiguous()
    row_off = pg.rank * pg.rows_per_rank
    st = _lib.stream_ptr()
    base = tables.next_pair()
    _lib.check(L.fr_push_rows(ego_local.data_ptr(), pg.rows_per_rank, d, tables.peers[base], tables.world, row_off, st),
               "fr_push_rows")
    tables.barrier()
    inv = 1.0 / (n_layers + 1)
    for layer in range(n_layers):
        x_full = tables.table(base + layer % 2)
        if layer == n_layers - 1:
            return ops.spmm(g, x_full, Z=ego_local, alpha=inv, beta=inv)
        else:
            _lib.check(L.fr_spmm_csr_f32_push(
                g.seg.data_ptr(), g.n_seg, g.long_rows.data_ptr(), g.n_long, g.col.data_ptr(), g.val.data_ptr(), d,
                x_full.data_ptr(), ego_local.data_ptr(), 1.0, 1.0, None, g.partial(d).data_ptr(), g.counters.data_ptr(),
                tables.peers[base + (layer + 1) % 2], tables.world, row_off, st), "fr_spmm_csr_f32_push")
            tables.barrier()


class _PushedPropagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego_local, pg, n_layers, tables):
        ctx.pg, ctx.n_layers, ctx.tables = pg, n_layers, tables
        return propagate_mean_pushed_raw(pg, ego_local, n_layers, tables)

    @staticmethod
    def backward(ctx, g):
        if not ctx.pg.symmetric:
            raise RuntimeError("row-partitioned backward needs a symmetric graph (S^T = S)")
        return propagate_mean_pushed_raw(ctx.pg, g.contiguous(), ctx.n_layers, ctx.tables), None, None, None


def propagate_mean_pushed(pg: RowPartitionedGraph, ego_local: torch.Tensor, n_layers: int, tables: PeerTables):
    """Differentiable row-partitioned layer-mean propagation over peer memory (no all-gather)."""
    return _PushedPropagate.apply(ego_local, pg, n_layers, tables)


# ------------------------------------------------------------------------------------ evaluation
def shard_users(users: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    n = users.numel()
    per = -(-n // world)
    return users[rank * per:min((rank + 1) * per, n)]


def full_sort_topk_sharded(user_all, item_all, users, k, hist=None, group=None, topk_fn=None):
    """Each rank ranks its slice of `users`; the `[n_users, k]` indices are gathered on every rank."""
    from . import evaluation
    topk_fn = topk_fn or evaluation.full_sort_topk
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = shard_users(users, rank, world)
    per = -(-users.numel() // world)
    idx = torch.full((per, k), -1, dtype=torch.int64, device=user_all.device)
    if mine.numel():
        idx[:mine.numel()] = topk_fn(user_all, item_all, mine, k, hist=hist)[1]
    out = torch.empty((per * world, k), dtype=torch.int64, device=user_all.device)
    dist.all_gather_into_tensor(out, idx, group=group)       # the only communication of the evaluation
    return out[:users.numel()]

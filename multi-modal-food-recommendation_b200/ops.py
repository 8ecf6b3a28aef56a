"""Differentiable operators over the C ABI (raw device pointers in, no torch types cross the boundary).

Each op launches the hand-written sm_100a kernels on torch's current CUDA stream, so they compose
with autograd, CUDA graphs and `torch.distributed` like any other op.  Nothing here computes on
the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .graph import PropGraph

_L = _lib.lib
# bench.py sets this to a list to collect (start_event, end_event, algorithmic_bytes, graph, has_z) per propagation launch
PROFILE = None


def _chk_f32(t: torch.Tensor, name: str):
    if t.device.type != "cuda":
        raise _lib.FoodRecError(f"{name} must be a CUDA tensor (foodrec_b200 has no CPU path), got {t.device}")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.FoodRecError(f"{name} must be contiguous float32, got {t.dtype} contiguous={t.is_contiguous()}")


def _idx(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.device.type != "cuda":
        raise _lib.FoodRecError(f"{name} must be a CUDA tensor, got {t.device}")
    return t.contiguous() if t.dtype == torch.int64 else t.to(torch.int64).contiguous()


# ------------------------------------------------------------------------------ row-activity masks
# The gradient of a mini-batch loss w.r.t. a propagated table is exactly zero outside the rows the batch
# touched, and stays sparse through the first backward propagation layers.  Producers of such gradients
# (`rank_loss`, `item_views`, the masked propagation itself) register a uint8 row mask for the gradient tensor
# they hand to autograd; the propagation backward looks its upstream gradient up here and, on a hit, skips
# the gathers of all-zero rows (results are bit-identical).  An entry keeps a reference to its tensor, so a
# data pointer found here cannot have been recycled for another tensor; entries are popped when consumed and
# the registry is cleared at the start of every backward pass of `rank_loss`.
_ROW_MASKS = {}
USE_ROW_MASKS = True


def _register_mask(t: torch.Tensor, mask: torch.Tensor):
    if USE_ROW_MASKS:
        _ROW_MASKS[(t.data_ptr(), tuple(t.shape))] = (t, mask)


def _take_mask(t: torch.Tensor):
    """Pop the row mask registered for gradient tensor `t`.  Autograd synchronises streams for the gradient
    tensors it routes, but it never sees the masks: the consumer may run on another stream than the producer
    (CLUSSL's item-side graphs back-propagate on side streams), so the mask is recorded on the consuming
    stream -- the caching allocator then cannot hand its block back to the producer's stream while the
    consumer's kernel is still queued."""
    hit = _ROW_MASKS.pop((t.data_ptr(), tuple(t.shape)), None)
    if hit is None:
        return None
    mask = hit[1]
    base = mask._base if mask._base is not None else mask
    base.record_stream(torch.cuda.current_stream())
    return mask


def spmm_masked(graph: PropGraph, X, Z, alpha, beta, x_mask, want_mask: bool):
    """Backward-pass launch: `alpha * S @ X + beta * Z` skipping rows of X whose mask byte is 0; optionally
    returns the row mask of the result."""
    d = X.shape[1]
    out = torch.empty((graph.n_rows, d), dtype=torch.float32, device=X.device)
    y_mask = torch.empty(graph.n_rows, dtype=torch.uint8, device=X.device) if want_mask else None
    prof = PROFILE
    if prof is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(_L.fr_spmm_csr_f32_masked(
        graph.seg.data_ptr(), graph.n_seg, graph.long_rows.data_ptr(), graph.n_long, graph.col.data_ptr(),
        graph.val.data_ptr(), d, X.data_ptr(), _lib.ptr(Z), float(alpha), float(beta), out.data_ptr(),
        graph.partial(d).data_ptr(), graph.counters.data_ptr(), x_mask.data_ptr(), _lib.ptr(y_mask),
        _lib.stream_ptr()), "fr_spmm_csr_f32_masked")
    if prof is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        prof.append((ev0, ev1, graph.spmm_bytes(d), graph, Z is not None))
    return out, y_mask


def propagate_mean_masked(graph: PropGraph, g: torch.Tensor, n_layers: int, mask: torch.Tensor):
    """`mean_l S^l g` for an upstream gradient `g` with row mask `mask` (Horner form; every layer's input is
    `previous + g`, whose mask is the previous layer's output mask).  Returns (result, its row mask)."""
    inv = 1.0 / (n_layers + 1)
    t, m = g, mask
    for layer in range(n_layers):
        last = layer == n_layers - 1
        t, m = spmm_masked(graph, t, g, inv if last else 1.0, inv if last else 1.0, m, want_mask=True)
    return t, m


# ------------------------------------------------------------------------------------ propagation
def spmm(graph: PropGraph, X: torch.Tensor, Z: torch.Tensor | None = None, alpha: float = 1.0, beta: float = 0.0,
         bias: torch.Tensor | None = None, act: int = 0, out: torch.Tensor | None = None,
         X1: torch.Tensor | None = None, Z1: torch.Tensor | None = None) -> torch.Tensor:
    """`out = act(alpha * S @ [X; X1] + beta * [Z; Z1] + bias)`; no autograd.  `X1` / `Z1` are optional
    second segments (rows stacked under `X` / `Z`), read in place instead of concatenating."""
    _chk_f32(X, "X")
    rows = X.shape[0] + (X1.shape[0] if X1 is not None else 0)
    if rows != graph.n_cols:
        raise _lib.FoodRecError(f"X has {rows} rows, graph has {graph.n_cols} columns")
    d = X.shape[1]
    if out is None:
        out = torch.empty((graph.n_rows, d), dtype=torch.float32, device=X.device)
    for t, name in ((X1, "X1"), (Z, "Z"), (Z1, "Z1"), (bias, "bias")):
        if t is not None:
            _chk_f32(t, name)
    if Z is not None and Z.shape[0] + (Z1.shape[0] if Z1 is not None else 0) != graph.n_rows:
        raise _lib.FoodRecError("Z does not have one row per graph row")
    prof = PROFILE
    if prof is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(_L.fr_spmm_csr_f32_split(
        graph.seg.data_ptr(), graph.n_seg, graph.long_rows.data_ptr(), graph.n_long, graph.col.data_ptr(),
        graph.val.data_ptr(), d, X.data_ptr(), _lib.ptr(X1), X.shape[0] if X1 is not None else 0, _lib.ptr(Z), _lib.ptr(Z1),
        Z.shape[0] if (Z is not None and Z1 is not None) else 0, float(alpha), float(beta), _lib.ptr(bias), int(act), out.data_ptr(),
        graph.partial(d).data_ptr(), graph.counters.data_ptr(), _lib.stream_ptr()), "fr_spmm_csr_f32")
    if prof is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        prof.append((ev0, ev1, graph.spmm_bytes(d), graph, Z is not None))
    return out


def propagate_mean_raw(graph: PropGraph, ego: torch.Tensor, n_layers: int, ego1: torch.Tensor | None = None) -> torch.Tensor:
    """`mean_{l=0..L} S^l [ego; ego1]` in Horner form, L fused launches, no layer stack (and no
    concatenated ego table) in memory."""
    if n_layers == 0:
        return ego.clone() if ego1 is None else torch.cat((ego, ego1), 0)
    inv = 1.0 / (n_layers + 1)
    t, t1 = ego, ego1
    for layer in range(n_layers):
        last = layer == n_layers - 1
        t = spmm(graph, t, Z=ego, alpha=inv if last else 1.0, beta=inv if last else 1.0, X1=t1, Z1=ego1)
        t1 = None
    return t


class _PropagateMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego, graph, n_layers):
        ctx.graph, ctx.n_layers = graph, n_layers
        return propagate_mean_raw(graph, ego.contiguous(), n_layers)

    @staticmethod
    def backward(ctx, g):
        # d/d ego of mean_l S^l ego is mean_l (S^T)^l g: the same Horner recurrence on S^T
        gt = ctx.graph.T
        if gt is None:
            raise _lib.FoodRecError("graph has no transpose plan; build it with a `.T`")
        g = g.contiguous()
        mask = _take_mask(g) if ctx.n_layers > 0 else None
        if mask is not None:
            r, rm = propagate_mean_masked(gt, g, ctx.n_layers, mask)
            _register_mask(r, rm)
            return r, None, None
        return propagate_mean_raw(gt, g, ctx.n_layers), None, None


class _PropagateMean2(torch.autograd.Function):
    """Layer-mean propagation of the stacked table [top; bottom] without building it; the gradient is one
    buffer whose two row blocks are handed back as views."""

    @staticmethod
    def forward(ctx, top, bottom, graph, n_layers):
        ctx.graph, ctx.n_layers, ctx.n_top = graph, n_layers, top.shape[0]
        return propagate_mean_raw(graph, top.contiguous(), n_layers, bottom.contiguous())

    @staticmethod
    def backward(ctx, g):
        gt = ctx.graph.T
        if gt is None:
            raise _lib.FoodRecError("graph has no transpose plan; build it with a `.T`")
        g = g.contiguous()
        mask = _take_mask(g) if ctx.n_layers > 0 else None
        if mask is not None:
            r, rm = propagate_mean_masked(gt, g, ctx.n_layers, mask)
            top, bottom = r[:ctx.n_top], r[ctx.n_top:]
            _register_mask(top, rm[:ctx.n_top])
            _register_mask(bottom, rm[ctx.n_top:])
            return top, bottom, None, None
        r = propagate_mean_raw(gt, g, ctx.n_layers)
        return r[:ctx.n_top], r[ctx.n_top:], None, None


def propagate_mean(graph: PropGraph, ego: torch.Tensor, n_layers: int, bottom: torch.Tensor | None = None) -> torch.Tensor:
    """Differentiable layer-mean propagation (replaces the `torch.sparse.mm` loop + stack/mean).  With
    `bottom`, the propagated table is `[ego; bottom]` (the reference's `torch.cat` of two tables) read in
    place."""
    if bottom is None:
        return _PropagateMean.apply(ego, graph, n_layers)
    return _PropagateMean2.apply(ego, bottom, graph, n_layers)


# ----------------------------------------------------------------------------- grouped propagation
class PropGroup:
    """Several independent graphs propagated by ONE launch per layer (`fr_spmm_csr_f32_grouped`).  The block map
    interleaves the graphs' blocks in proportion to their sizes, each graph in its own plan order (long-row segments
    first), and is uploaded once."""

    def __init__(self, graphs):
        import numpy as np
        self.graphs = list(graphs)
        if not 1 <= len(self.graphs) <= 4:
            raise _lib.FoodRecError("a PropGroup holds 1..4 graphs")
        nb = [int(_L.fr_spmm_task_blocks(g.n_seg)) for g in self.graphs]
        task = np.concatenate([np.full(n, t, dtype=np.int32) for t, n in enumerate(nb)])
        blk = np.concatenate([np.arange(n, dtype=np.int32) for n in nb])
        key = np.concatenate([(np.arange(n) + 0.5) / max(n, 1) for n in nb])
        order = np.argsort(key, kind="stable")
        self.n_blocks = int(task.size)
        self.blk_map = torch.from_numpy(np.stack([task[order], blk[order]], 1).copy()).to(self.graphs[0].device)
        self._T = None

    @property
    def T(self):
        if self._T is None:
            ts = [g.T for g in self.graphs]
            if any(t is None for t in ts):
                raise _lib.FoodRecError("a graph of the group has no transpose plan")
            self._T = self if all(t is g for t, g in zip(ts, self.graphs)) else PropGroup(ts)
        return self._T


def spmm_grouped(group: PropGroup, Xs, Zs, alpha: float, beta: float, X1s=None, Z1s=None, outs=None):
    """One launch: `out[t] = alpha * S_t @ [Xs[t]; X1s[t]] + beta * [Zs[t]; Z1s[t]]` for every graph of the group;
    no autograd.  Bit-identical to the separate `spmm` calls."""
    n = len(group.graphs)
    X1s = X1s or [None] * n
    Z1s = Z1s or [None] * n
    d = Xs[0].shape[1]
    given, outs, tasks = outs, [], (_lib.SpmmTask * n)()
    prof = PROFILE
    for t, g in enumerate(group.graphs):
        X, X1, Z, Z1 = Xs[t], X1s[t], Zs[t], Z1s[t]
        for a, name in ((X, "X"), (X1, "X1"), (Z, "Z"), (Z1, "Z1")):
            if a is not None:
                _chk_f32(a, name)
        if X.shape[0] + (X1.shape[0] if X1 is not None else 0) != g.n_cols or X.shape[1] != d:
            raise _lib.FoodRecError(f"task {t}: operand rows do not match the graph")
        out = given[t] if given is not None else torch.empty((g.n_rows, d), dtype=torch.float32, device=X.device)
        outs.append(out)
        tk = tasks[t]
        tk.seg, tk.n_seg, tk.long_rows, tk.n_long = g.seg.data_ptr(), g.n_seg, g.long_rows.data_ptr(), g.n_long
        tk.col_idx, tk.val = g.col.data_ptr(), g.val.data_ptr()
        tk.X0, tk.X1, tk.x_split = X.data_ptr(), _lib.ptr(X1), X.shape[0] if X1 is not None else 0
        tk.Z0, tk.Z1 = _lib.ptr(Z), _lib.ptr(Z1)
        tk.z_split = Z.shape[0] if (Z is not None and Z1 is not None) else 0
        tk.alpha, tk.beta = float(alpha), float(beta)
        tk.Y, tk.partial, tk.counters = out.data_ptr(), g.partial(d).data_ptr(), g.counters.data_ptr()
    if prof is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(_L.fr_spmm_csr_f32_grouped(tasks, n, d, group.blk_map.data_ptr(), group.n_blocks, _lib.stream_ptr()),
               "fr_spmm_csr_f32_grouped")
    if prof is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        prof.append((ev0, ev1, sum(g.spmm_bytes(d) for g in group.graphs), group, Zs[0] is not None))
    return outs


def _propagate_mean_grouped_raw(group, tops, bottoms, n_layers):
    inv = 1.0 / (n_layers + 1)
    t, t1 = list(tops), list(bottoms)
    for layer in range(n_layers):
        last = layer == n_layers - 1
        t = spmm_grouped(group, t, list(tops), inv if last else 1.0, inv if last else 1.0, X1s=t1, Z1s=list(bottoms))
        t1 = [None] * len(t)
    return t


class _PropagateMeanGrouped(torch.autograd.Function):
    """Layer-mean propagation of `[top_t; bottom_t]` over graph t, all graphs of the group in one launch per layer
    (forward and backward)."""

    @staticmethod
    def forward(ctx, group, n_layers, *tabs):
        n = len(group.graphs)
        tops, bottoms = [t.contiguous() for t in tabs[:n]], [t.contiguous() for t in tabs[n:]]
        ctx.group, ctx.n_layers, ctx.n_top = group, n_layers, [t.shape[0] for t in tops]
        # CLUSSL feeds the SAME item table to every graph of the group (pricai_modelx.py:181,195,209): its gradient is
        # then the sum of the graphs' top blocks, produced by one `fr_sum_rows` launch in the backward instead of
        # autograd's chain of elementwise adds
        ctx.shared_top = n > 1 and all(t.data_ptr() == tops[0].data_ptr() and t.shape == tops[0].shape for t in tops)
        return tuple(_propagate_mean_grouped_raw(group, tops, bottoms, n_layers))

    @staticmethod
    def backward(ctx, *gs):
        gt = ctx.group.T
        gs = list(gs)
        d = next(g.shape[1] for g in gs if g is not None)
        for k, g in enumerate(gs):
            if g is None:      # an output nobody differentiated: its cotangent is zero
                gr = ctx.group.graphs[k]
                gs[k] = torch.zeros((gr.n_rows, d), dtype=torch.float32, device=gr.device)
            else:
                gs[k] = g.contiguous()
                _take_mask(gs[k])          # the grouped launch does not use row masks; drop the registration
        r = _propagate_mean_grouped_raw(gt, gs, [None] * len(gs), ctx.n_layers)
        tops = [x[:nt] for x, nt in zip(r, ctx.n_top)]
        bottoms = [x[nt:] for x, nt in zip(r, ctx.n_top)]
        if ctx.shared_top:
            total = torch.empty((ctx.n_top[0], d), dtype=torch.float32, device=r[0].device)
            _lib.check(_L.fr_sum_rows(_ptr_array(r), len(r), d, ctx.n_top[0], total.data_ptr(), _lib.stream_ptr()), "fr_sum_rows")
            tops = [total] + [None] * (len(r) - 1)
        return (None, None, *tops, *bottoms)


def propagate_mean_grouped(group: PropGroup, tops, bottoms, n_layers: int):
    """Differentiable `mean_l S_t^l [tops[t]; bottoms[t]]` for every graph t of `group`, one launch per layer for the
    whole group (CLUSSL's three item-side graphs, pricai_modelx.py:179-218)."""
    if n_layers == 0:
        return [torch.cat((a, b), 0) for a, b in zip(tops, bottoms)]
    return list(_PropagateMeanGrouped.apply(group, n_layers, *tops, *bottoms))


class _SpmmBiasTanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, graph):
        y = spmm(graph, x.contiguous(), bias=bias.contiguous(), act=1)
        ctx.save_for_backward(y)
        ctx.graph = graph
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        dpre = (g * (1.0 - y * y)).contiguous()
        return spmm(ctx.graph.T, dpre), dpre.sum(0), None


def gcn_propagate_tanh(graph: PropGraph, h: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """`tanh(S @ h + bias)` -- the propagate/bias/tanh part of SCHGN's GraphConv (schgn.py:38-41)."""
    return _SpmmBiasTanh.apply(h, bias, graph)


# ---------------------------------------------------------------------------------- ranking loss
_WS = {}


def _workspace(device) -> torch.Tensor:
    ws = _WS.get(device)
    if ws is None:
        ws = torch.zeros(int(_L.fr_rank_loss_ws_floats()), dtype=torch.float32, device=device)
        _WS[device] = ws
    return ws


def _ptr_array(tensors):
    return (C.c_void_p * max(len(tensors), 1))(*[None if t is None else t.data_ptr() for t in tensors])


class _RankLoss(torch.autograd.Function):
    """(mf_loss, reg_loss) = BPR on propagated rows + EmbLoss on ego rows, one launch each way."""

    @staticmethod
    def forward(ctx, emb, item_off, u, p, n, gamma, reg_den, reg_idx, reg_pad, *reg_tabs):
        _chk_f32(emb, "emb")
        B, d = u.numel(), emb.shape[1]
        for t in reg_tabs:
            _chk_f32(t, "regulariser table")
            if t.shape[1] != d:
                raise _lib.FoodRecError("regulariser tables must share the embedding width")
        dev = emb.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        coef = torch.empty(B, dtype=torch.float32, device=dev)
        gnorm = torch.empty(max(len(reg_tabs), 1), dtype=torch.float32, device=dev)
        cnt = (C.c_int64 * max(len(reg_tabs), 1))(*[int(i.numel()) for i in reg_idx])
        _lib.check(_L.fr_rank_loss_fwd(
            emb.data_ptr(), d, int(item_off), u.data_ptr(), p.data_ptr(), n.data_ptr(), B, float(gamma),
            len(reg_tabs), _ptr_array(reg_tabs), _ptr_array(reg_idx), cnt, float(reg_den), out.data_ptr(),
            coef.data_ptr(), gnorm.data_ptr(), _workspace(dev).data_ptr(), _lib.stream_ptr()), "fr_rank_loss_fwd")
        ctx.save_for_backward(emb, u, p, n, coef, gnorm, *reg_idx, *reg_tabs)
        ctx.meta = (int(item_off), float(reg_den), len(reg_tabs), tuple(int(x) for x in reg_pad))
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_mf, g_reg):
        item_off, reg_den, ng, pads = ctx.meta
        emb, u, p, n, coef, gnorm = ctx.saved_tensors[:6]
        reg_idx = ctx.saved_tensors[6:6 + ng]
        reg_tabs = ctx.saved_tensors[6 + ng:6 + 2 * ng]
        B, d = u.numel(), emb.shape[1]
        g_out = torch.stack([g_mf.reshape(()), g_reg.reshape(())]).to(torch.float32).contiguous()
        # one zero-filled allocation (one fill launch) holds d_emb and one dense gradient per distinct
        # regulariser table (the same table may back several groups)
        shapes = [emb.shape] if ctx.needs_input_grad[0] else []
        keys = []
        for k, t in enumerate(reg_tabs):
            if ctx.needs_input_grad[9 + k] and t.data_ptr() not in keys:
                keys.append(t.data_ptr())
                shapes.append(t.shape)
        _ROW_MASKS.clear()   # a new backward pass starts here: nothing registered earlier may be used again
        n_flat = sum(sh[0] * sh[1] for sh in shapes)
        want_mask = USE_ROW_MASKS and ctx.needs_input_grad[0]
        n_mask_words = (emb.shape[0] + 3) // 4 if want_mask else 0     # the row mask shares the zero fill
        flat = torch.zeros(n_flat + n_mask_words, dtype=torch.float32, device=emb.device)
        emb_mask = flat[n_flat:].view(torch.uint8)[:emb.shape[0]] if want_mask else None
        views, o = [], 0
        for sh in shapes:
            views.append(flat[o:o + sh[0] * sh[1]].view(sh[0], sh[1]))
            o += sh[0] * sh[1]
        d_emb = views.pop(0) if ctx.needs_input_grad[0] else None
        uniq = dict(zip(keys, views))
        d_tabs = [uniq[t.data_ptr()] if ctx.needs_input_grad[9 + k] else None for k, t in enumerate(reg_tabs)]
        cnt = (C.c_int64 * max(ng, 1))(*[int(i.numel()) for i in reg_idx])
        pad = (C.c_int64 * max(ng, 1))(*pads)
        _lib.check(_L.fr_rank_loss_bwd(
            emb.data_ptr(), d, item_off, u.data_ptr(), p.data_ptr(), n.data_ptr(), B, coef.data_ptr(),
            g_out.data_ptr(), _lib.ptr(d_emb), ng, _ptr_array(reg_tabs), _ptr_array(reg_idx), cnt, pad, reg_den,
            gnorm.data_ptr(), _ptr_array(d_tabs), _lib.ptr(emb_mask), _lib.stream_ptr()), "fr_rank_loss_bwd")
        if emb_mask is not None:
            _register_mask(d_emb, emb_mask)
        # a table shared by several groups gets its (already summed) gradient once
        seen, grads = set(), []
        for k, t in enumerate(reg_tabs):
            if d_tabs[k] is None or t.data_ptr() in seen:
                grads.append(None)
            else:
                seen.add(t.data_ptr())
                grads.append(d_tabs[k])
        return (d_emb, None, None, None, None, None, None, None, None, *grads)


def rank_loss(emb: torch.Tensor, item_off: int, u, p, n, reg_groups, reg_den: float, gamma: float = 1e-10):
    """BPR over rows of `emb` (users at `u`, items at `item_off + p/n`) and the un-weighted EmbLoss
    `sum_g ||T_g[idx_g]||_F / reg_den` over `reg_groups = [(table, idx, pad_idx|-1), ...]`.
    Returns two 0-dim tensors `(mf_loss, reg_loss)`."""
    u, p, n = _idx(u, "u"), _idx(p, "p"), _idx(n, "n")
    tabs = [g[0] for g in reg_groups]
    idxs = [_idx(g[1].reshape(-1), "reg idx") for g in reg_groups]
    pads = [(-1 if (len(g) < 3 or g[2] is None) else int(g[2])) for g in reg_groups]
    return _RankLoss.apply(emb, item_off, u, p, n, gamma, reg_den, idxs, pads, *tabs)


# ------------------------------------------------------------------------------- gathers / scores
class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tab, idx):
        _chk_f32(tab, "table")
        out = torch.empty((idx.numel(), tab.shape[1]), dtype=torch.float32, device=tab.device)
        _lib.check(_L.fr_gather_rows(tab.data_ptr(), tab.shape[1], idx.data_ptr(), idx.numel(), out.data_ptr(),
                                     _lib.stream_ptr()), "fr_gather_rows")
        ctx.save_for_backward(idx)
        ctx.shape = tab.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        d_tab = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        g = g.contiguous()
        _lib.check(_L.fr_scatter_add_rows(g.data_ptr(), g.shape[1], idx.data_ptr(), idx.numel(), d_tab.data_ptr(),
                                          _lib.stream_ptr()), "fr_scatter_add_rows")
        return d_tab, None


def gather_rows(tab: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """`tab[idx]` with a scatter-add backward."""
    return _GatherRows.apply(tab.contiguous(), _idx(idx, "idx"))


class _CosineMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, T, idx):
        _chk_f32(A, "A")
        _chk_f32(T, "table")
        n, d, dev = A.shape[0], A.shape[1], A.device
        out = torch.empty(1, dtype=torch.float32, device=dev)
        state = torch.empty((3, n), dtype=torch.float32, device=dev)
        _lib.check(_L.fr_cosine_mean_fwd(A.data_ptr(), T.data_ptr(), idx.data_ptr(), n, d, out.data_ptr(),
                                         state[0].data_ptr(), state[1].data_ptr(), state[2].data_ptr(), _lib.stream_ptr()),
                   "fr_cosine_mean_fwd")
        ctx.save_for_backward(A, T, idx, state)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        A, T, idx, state = ctx.saved_tensors
        n, d = A.shape
        g = g.to(torch.float32).reshape(1).contiguous()
        dA = torch.empty_like(A) if ctx.needs_input_grad[0] else None
        dT = torch.zeros_like(T) if ctx.needs_input_grad[1] else None
        _lib.check(_L.fr_cosine_mean_bwd(A.data_ptr(), T.data_ptr(), idx.data_ptr(), n, d, state[0].data_ptr(),
                                         state[1].data_ptr(), state[2].data_ptr(), g.data_ptr(), _lib.ptr(dA), _lib.ptr(dT),
                                         _lib.stream_ptr()), "fr_cosine_mean_bwd")
        return dA, dT, None


def cosine_mean(A: torch.Tensor, table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """`cosine_similarity(A, table[idx], dim=-1).mean()` without materialising the gathered rows (HealthRec's
    KD term, cikm_model.py:263); scatter-add backward into the dense table gradient."""
    return _CosineMean.apply(A.contiguous(), table.contiguous(), _idx(idx.reshape(-1), "idx"))


def pair_scores(user_tab: torch.Tensor, item_tab: torch.Tensor, user: torch.Tensor, item: torch.Tensor):
    """`(user_tab[user] * item_tab[item]).sum(1)` (inference_fast / inference_by_user); no autograd."""
    user_tab, item_tab = user_tab.detach(), item_tab.detach()
    if not user_tab.is_contiguous():
        user_tab = user_tab.contiguous()
    if not item_tab.is_contiguous():
        item_tab = item_tab.contiguous()
    _chk_f32(user_tab, "user table")
    _chk_f32(item_tab, "item table")
    user, item = _idx(user, "user"), _idx(item, "item")
    out = torch.empty(user.numel(), dtype=torch.float32, device=user_tab.device)
    _lib.check(_L.fr_pair_scores(user_tab.data_ptr(), item_tab.data_ptr(), user_tab.shape[1], user.data_ptr(),
                                 item.data_ptr(), user.numel(), out.data_ptr(), _lib.stream_ptr()), "fr_pair_scores")
    return out


# --------------------------------------------------------------------------- contrastive terms
_DCOR_WS = {}


def _dcor_ws(device, n):
    need = int(_L.fr_dcor_ws_floats(n))
    ws = _DCOR_WS.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.float32, device=device)
        _DCOR_WS[device] = ws
    return ws


def _dcor_forward(tabs, idx, pairs, scale):
    V, P, n, d = len(tabs), len(pairs), idx.numel(), tabs[0].shape[1]
    for t in tabs:
        _chk_f32(t, "view table")
    dev = tabs[0].device
    Dm = torch.empty((V, n, n), dtype=torch.float32, device=dev)
    rowmean = torch.empty((V, n), dtype=torch.float32, device=dev)
    out = torch.empty(P + 1, dtype=torch.float32, device=dev)
    dfds = torch.empty(3 * P, dtype=torch.float32, device=dev)
    gm = torch.empty(V, dtype=torch.float32, device=dev)
    pr = (C.c_int32 * (2 * P))(*[int(x) for ab in pairs for x in ab])
    _lib.check(_L.fr_dcor_fwd(_ptr_array(tabs), V, d, idx.data_ptr(), n, pr, P, float(scale), Dm.data_ptr(),
                              rowmean.data_ptr(), out.data_ptr(), dfds.data_ptr(), gm.data_ptr(),
                              _dcor_ws(dev, n).data_ptr(), _lib.stream_ptr()), "fr_dcor_fwd")
    return out, (Dm, rowmean, dfds, gm)


def _dcor_backward(tabs, idx, pairs, state, g_terms, d_tabs, masks=None):
    """Accumulates sum_p g_terms[p] * d term_p / d tab_v into the dense `d_tabs[v]` (None = skip); `masks[v]`
    (uint8 per row, optional) get the touched rows marked."""
    Dm, rowmean, dfds, gm = state
    V, P, n, d = len(tabs), len(pairs), idx.numel(), tabs[0].shape[1]
    pr = (C.c_int32 * (2 * P))(*[x for ab in pairs for x in ab])
    ws = torch.empty(int(_L.fr_dcor_bwd_ws_floats(n)), dtype=torch.float32, device=idx.device)   # W matrices + row sums
    _lib.check(_L.fr_dcor_bwd(_ptr_array(tabs), V, d, idx.data_ptr(), n, pr, P, Dm.data_ptr(), rowmean.data_ptr(),
                              dfds.data_ptr(), gm.data_ptr(), g_terms.data_ptr(), _ptr_array(d_tabs),
                              _ptr_array(masks) if masks is not None else None, ws.data_ptr(), _lib.stream_ptr()), "fr_dcor_bwd")


def _term_grads(g_terms, g_total, P):
    """Upstream gradient per term: explicit per-term gradients plus the gradient of their sum."""
    if g_terms is None:
        return g_total.to(torch.float32).reshape(1).expand(P).contiguous()
    g = g_terms.to(torch.float32)
    if g_total is not None:
        g = g + g_total.reshape(1)
    return g.contiguous()


class _DcorTerms(torch.autograd.Function):
    """(terms [P], their sum [1]) -- both scaled by `scale` inside the kernel."""

    @staticmethod
    def forward(ctx, idx, pairs, scale, *tabs):
        pairs = [tuple(int(x) for x in ab) for ab in pairs]
        out, state = _dcor_forward(tabs, idx, pairs, scale)
        ctx.save_for_backward(idx, *state, *tabs)
        ctx.pairs = pairs
        return out[:len(pairs)], out[len(pairs):]

    @staticmethod
    def backward(ctx, g_terms, g_total):
        idx = ctx.saved_tensors[0]
        state, tabs = ctx.saved_tensors[1:5], ctx.saved_tensors[5:]
        g = _term_grads(g_terms, g_total, len(ctx.pairs))
        d_tabs = [torch.zeros_like(t) if ctx.needs_input_grad[3 + k] else None for k, t in enumerate(tabs)]
        _dcor_backward(tabs, idx, ctx.pairs, state, g, d_tabs)
        return (None, None, None, *d_tabs)


def dcor_terms(tabs, idx: torch.Tensor, pairs, scale: float = 1.0, with_total: bool = False):
    """Distance correlations `[P]` (times `scale`) between the views `tabs[v][idx]` for the view pairs
    `pairs` (FoodRec/models/pricai_modelx.py:245-247,263,409-437): gathers, distance matrices, centring,
    covariances and their backward in four launches.  `with_total=True` also returns their sum `[1]`."""
    terms, total = _DcorTerms.apply(_idx(idx.reshape(-1), "idx"), list(pairs), float(scale), *[t.contiguous() for t in tabs])
    return (terms, total) if with_total else terms


class _GradHolder:
    """Hand-over between the two halves of `item_views` in the backward: the contrastive half runs first
    (its upstream gradient is known at the start of the backward) on its own stream and parks the dense
    table gradients here; the item_emb half adds its rows to them and returns them.  Either order of the
    two backward calls gives the same gradients."""

    def __init__(self):
        self.d_tabs = None
        self.masks = None
        self.stream = None
        self.done = False
        self.consumed = False
        self._armed = False

    def arm(self):
        """Queue the end-of-backward check once per backward pass.  The hand-over is only correct when BOTH
        halves are differentiated in the same pass (`loss = mf + cl + reg` as the trainer does).  If only the
        contrastive total was differentiated (`cl_loss.backward(retain_graph=True)`, `autograd.grad(cl_loss, ..)`)
        its table gradients are still parked here when the pass ends: that is reported loudly instead of
        silently returning zero gradients, and the state is reset so nothing leaks into a later pass."""
        if self._armed:
            return
        self._armed = True

        def _end_of_pass():
            lost = self.done and not self.consumed
            self.d_tabs = self.masks = self.stream = None
            self.done = self.consumed = self._armed = False
            if lost:
                raise _lib.FoodRecError(
                    "ops.item_views: the contrastive total was back-propagated without item_emb in the same backward "
                    "pass, so its table gradients were not delivered. Differentiate the summed loss in one pass, or "
                    "use ops.dcor_terms(...) for a stand-alone contrastive term.")
        torch.autograd.Variable._execution_engine.queue_callback(_end_of_pass)


class _DcorTotalInto(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, pairs, scale, holder, *tabs):
        pairs = [tuple(int(x) for x in ab) for ab in pairs]
        out, state = _dcor_forward(tabs, idx, pairs, scale)
        ctx.save_for_backward(idx, *state, *tabs)
        ctx.pairs, ctx.holder = pairs, holder
        return out[len(pairs):]

    @staticmethod
    def backward(ctx, g_total):
        idx = ctx.saved_tensors[0]
        state, tabs = ctx.saved_tensors[1:5], ctx.saved_tensors[5:]
        h = ctx.holder
        h.arm()
        d_tabs = [torch.zeros_like(t) for t in tabs]
        masks = None
        if USE_ROW_MASKS and not h.consumed:
            masks = [torch.zeros(t.shape[0], dtype=torch.uint8, device=t.device) for t in tabs]
        _dcor_backward(tabs, idx, ctx.pairs, state, _term_grads(None, g_total, len(ctx.pairs)), d_tabs, masks)
        if h.consumed:                       # the other half already ran: ordinary gradients, autograd adds them
            return (None, None, None, None, *d_tabs)
        h.d_tabs, h.masks, h.stream, h.done = d_tabs, masks, torch.cuda.current_stream(), True
        return (None, None, None, None) + (None,) * len(tabs)


class _SumRowsInto(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_items, holder, *tabs):
        d = tabs[0].shape[1]
        for t in tabs:
            _chk_f32(t, "table")
        out = torch.empty((n_items, d), dtype=torch.float32, device=tabs[0].device)
        _lib.check(_L.fr_sum_rows(_ptr_array(tabs), len(tabs), d, n_items, out.data_ptr(), _lib.stream_ptr()), "fr_sum_rows")
        ctx.n_items, ctx.holder = n_items, holder
        ctx.shapes = [tuple(t.shape) for t in tabs]
        ctx.dev = tabs[0].device
        return out

    @staticmethod
    def backward(ctx, g_item):
        h, d = ctx.holder, ctx.shapes[0][1]
        h.arm()
        g_item = g_item.contiguous()
        rows = (C.c_int64 * len(ctx.shapes))(*[int(sh[0]) for sh in ctx.shapes])
        src_mask = _take_mask(g_item)
        if h.done and not h.consumed:        # add to the parked contrastive gradients (other stream: join first)
            cur = torch.cuda.current_stream()
            if h.stream is not None and h.stream != cur:
                cur.wait_stream(h.stream)
                for t in h.d_tabs + (h.masks or []):
                    t.record_stream(cur)
            d_tabs, masks, acc = h.d_tabs, h.masks, 1
        else:
            d_tabs, acc = [torch.empty(sh, dtype=torch.float32, device=ctx.dev) for sh in ctx.shapes], 0
            masks = ([torch.empty(sh[0], dtype=torch.uint8, device=ctx.dev) for sh in ctx.shapes]
                     if (USE_ROW_MASKS and src_mask is not None) else None)
        h.consumed = True
        _lib.check(_L.fr_spread_rows(g_item.data_ptr(), d, ctx.n_items, _ptr_array(d_tabs), rows, len(d_tabs), acc,
                                     _lib.ptr(src_mask), _ptr_array(masks) if masks is not None else None,
                                     _lib.stream_ptr()), "fr_spread_rows")
        if masks is not None:
            for t, m in zip(d_tabs, masks):
                _register_mask(t, m)
        return (None, None, *d_tabs)


def item_views(tabs, idx: torch.Tensor, pairs, scale: float, n_items: int, side_stream=None):
    """CLUSSL's consumer of the three item-side tables: `(sum_v tabs[v][:n_items], scale * sum_p dcor_p)`
    (pricai_modelx.py:219,245-263).  The contrastive half is latency-bound, so it is launched on
    `side_stream` (forward and, through autograd, backward) beside the user-item propagation; in the
    backward its dense table gradients are written once and the item_emb gradient is added to their first
    `n_items` rows in place -- no slice copies, no extra zero fills, no gradient-sum kernels."""
    tabs = [t.contiguous() for t in tabs]
    idx = _idx(idx.reshape(-1), "idx")
    holder = _GradHolder()
    item_emb = _SumRowsInto.apply(int(n_items), holder, *tabs)
    if side_stream is None:
        total = _DcorTotalInto.apply(idx, list(pairs), float(scale), holder, *tabs)
    else:
        cur = torch.cuda.current_stream()
        side_stream.wait_stream(cur)
        with torch.cuda.stream(side_stream):
            total = _DcorTotalInto.apply(idx, list(pairs), float(scale), holder, *tabs)
    return item_emb, total


def correlation_distance(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """`PRICAI_ModelX.correlation_distance(x, y)` for two `[n, d]` views; returns shape `[1]`."""
    idx = torch.arange(x.shape[0], device=x.device)
    return dcor_terms([x, y], idx, [(0, 1)])


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, temperature, hidden_norm):
        _chk_f32(hidden, "hidden")
        b, d = hidden.shape[0] // 2, hidden.shape[1]
        n, dev = 2 * b, hidden.device
        hn = torch.empty((n, d), dtype=torch.float32, device=dev)
        norm = torch.empty(n, dtype=torch.float32, device=dev)
        G = torch.empty((n, n), dtype=torch.float32, device=dev)
        lse = torch.empty(n, dtype=torch.float32, device=dev)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = torch.empty(int(_L.fr_infonce_ws_floats(n)), dtype=torch.float32, device=dev)
        _lib.check(_L.fr_infonce_fwd(hidden.data_ptr(), b, d, float(temperature), int(bool(hidden_norm)), hn.data_ptr(),
                                     norm.data_ptr(), G.data_ptr(), lse.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                     _lib.stream_ptr()), "fr_infonce_fwd")
        ctx.save_for_backward(hn, norm, G, lse)
        ctx.meta = (b, d, float(temperature), int(bool(hidden_norm)), hidden.shape[0])
        return out[0]

    @staticmethod
    def backward(ctx, g):
        hn, norm, G, lse = ctx.saved_tensors
        b, d, tau, nrm, rows = ctx.meta
        g = g.to(torch.float32).reshape(1).contiguous()
        dh = torch.empty((rows, d), dtype=torch.float32, device=hn.device)
        if rows > 2 * b:
            dh[2 * b:].zero_()          # an odd trailing row takes no part (hidden.shape[0] // 2 in the reference)
        _lib.check(_L.fr_infonce_bwd(hn.data_ptr(), norm.data_ptr(), G.data_ptr(), lse.data_ptr(), b, d, tau, nrm,
                                     g.data_ptr(), dh.data_ptr(), _lib.stream_ptr()), "fr_infonce_bwd")
        return dh, None, None


def info_nce(hidden: torch.Tensor, temperature: float = 0.5, hidden_norm: bool = True) -> torch.Tensor:
    """SimCLR NT-Xent over the two halves of `hidden` (`CL_loss`, pricai_modelx.py:354-378): three fused
    launches forward, one backward (one symmetric Gram matrix instead of four logit blocks)."""
    return _InfoNCE.apply(hidden.contiguous(), float(temperature), bool(hidden_norm))

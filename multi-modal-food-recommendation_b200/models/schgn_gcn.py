"""SCHGN's graph-convolution block on the B200 propagation kernel.

`GraphConv` is a drop-in for the class of the same name in FoodRec/models/schgn.py:29-41
(`tanh(GCNConv(x, edge_index))`, torch_geometric defaults: one self-loop per node, symmetric
in-degree normalisation, `lin` without bias before propagation, `+ bias` after).  Parameter names
match PyG's (`conv1.lin.weight` `[out, in]`, `conv1.bias`) so SCHGN checkpoints load.  The reference
recomputes the normalisation on every call (`cached=False`) and calls the block twice per batch and
once per user in `full_sort_predict` (schgn.py:247,284-300,339); here the normalised CSR and its
transpose are built once per edge list (`graph.gcn_normalised`) and a call is one dense `lin` GEMM plus
one fused propagate + bias + tanh launch (backward: one launch on the transposed plan).
The rest of SCHGN (per-pair attention scorer, SSL encoder) is outside the hot path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import graph as G
from .. import ops


def truncated_normal_(tensor, mean=0.0, std=0.01):
    """schgn.py:18-26: resample-free truncated normal (first of four draws inside +-2 sigma)."""
    with torch.no_grad():
        tmp = tensor.new_empty(tensor.shape + (4,)).normal_()
        valid = (tmp < 2) & (tmp > -2)
        ind = valid.max(-1, keepdim=True)[1]
        tensor.data.copy_(tmp.gather(-1, ind).squeeze(-1))
        tensor.data.mul_(std).add_(mean)
    return tensor


class _GCNConvParams(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.lin = nn.Linear(in_channel, out_channel, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channel))


class GraphConv(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.in_channel, self.out_channel = in_channel, out_channel
        self.conv1 = _GCNConvParams(in_channel, out_channel)
        std = float(np.sqrt(2.0 / (in_channel + out_channel)))
        truncated_normal_(self.conv1.lin.weight, std=std)
        truncated_normal_(self.conv1.bias, std=std)
        self._plans = {}

    def plan(self, edge_index: torch.Tensor, n_nodes: int):
        """Normalised CSR (+ transpose) for `edge_index` `[2, E]` (row 0 = source, row 1 = target);
        cached per edge tensor."""
        key = (edge_index.data_ptr(), int(edge_index.shape[1]), n_nodes, edge_index._version)
        hit = self._plans.get(key)
        if hit is not None:
            return hit[1]
        ei = edge_index.detach().cpu().numpy()
        g = G.gcn_normalised(ei[0], ei[1], n_nodes, self.conv1.bias.device)
        # the entry holds the edge tensor itself: while it is cached its address cannot be handed to another
        # edge list of the same size (a temporary built per call would otherwise alias a stale plan)
        if len(self._plans) >= 4:
            self._plans.pop(next(iter(self._plans)))
        self._plans[key] = (edge_index, g)
        return g

    def forward(self, x, edge_index):
        g = self.plan(edge_index, x.shape[0])
        return ops.gcn_propagate_tanh(g, self.conv1.lin(x), self.conv1.bias)

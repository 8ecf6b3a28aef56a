"""BERT-style ingredient-sequence encoder used by SCHGN's masked-ingredient task.

Batch-sized dense torch (B x 20 x 64): outside the B200 hot path, kept as a torch module whose
parameter names, shapes and construction order follow FoodRec/common/module.py:48-190 so that
`state_dict`s interchange and a seed reproduces the reference's initial weights (one block is built,
then deep-copied per layer, as the reference does -- the copies share their initial values).
"""
import copy
import math

import torch
from torch import nn
import torch.nn.functional as F


def _erf_gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


_ACTIVATIONS = {"gelu": _erf_gelu, "relu": F.relu, "swish": lambda x: x * torch.sigmoid(x)}


class SelfAttention(nn.Module):
    def __init__(self, n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps):
        super().__init__()
        if hidden_size % n_heads:
            raise ValueError(f"hidden size {hidden_size} is not a multiple of the number of heads {n_heads}")
        self.n_heads, self.head_dim = n_heads, hidden_size // n_heads
        for name in ("query", "key", "value"):                      # creation order = RNG order
            setattr(self, name, nn.Linear(hidden_size, hidden_size))
        self.attn_dropout = nn.Dropout(attn_dropout_prob)
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.out_dropout = nn.Dropout(hidden_dropout_prob)

    def forward(self, x, mask):
        B, L, H = x.shape
        q, k, v = (proj(x).view(B, L, self.n_heads, self.head_dim).transpose(1, 2)
                   for proj in (self.query, self.key, self.value))
        prob = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(self.head_dim) + mask, dim=-1)
        ctx = (self.attn_dropout(prob) @ v).transpose(1, 2).reshape(B, L, H)
        return self.LayerNorm(self.out_dropout(self.dense(ctx)) + x)


class Intermediate(nn.Module):
    def __init__(self, hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps):
        super().__init__()
        self.dense_1 = nn.Linear(hidden_size, inner_size)
        self.act = _ACTIVATIONS[hidden_act] if isinstance(hidden_act, str) else hidden_act
        self.dense_2 = nn.Linear(inner_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.dropout = nn.Dropout(hidden_dropout_prob)

    def forward(self, x):
        return self.LayerNorm(self.dropout(self.dense_2(self.act(self.dense_1(x)))) + x)


class Layer(nn.Module):
    def __init__(self, n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                 layer_norm_eps):
        super().__init__()
        self.attention = SelfAttention(n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps)
        self.intermediate = Intermediate(hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps)

    def forward(self, x, mask):
        return self.intermediate(self.attention(x, mask))


class Encoder(nn.Module):
    def __init__(self, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
                 attn_dropout_prob=0.5, hidden_act="gelu", layer_norm_eps=1e-12):
        super().__init__()
        block = Layer(n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                      layer_norm_eps)
        self.layer = nn.ModuleList(copy.deepcopy(block) for _ in range(n_layers))

    def forward(self, hidden_states, attention_mask, output_all_encoded_layers=True):
        outputs = []
        for block in self.layer:
            hidden_states = block(hidden_states, attention_mask)
            outputs.append(hidden_states)
        return outputs if output_all_encoded_layers else outputs[-1:]

"""Parameter initialisers matching FoodRec/common/init.py:7-42 (same RNG consumption order, so the
same seed yields the same initial `state_dict` as the reference model)."""
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_


def xavier_uniform_initialization(module):
    if isinstance(module, nn.Embedding):
        xavier_uniform_(module.weight.data)
    elif isinstance(module, nn.Linear):
        xavier_uniform_(module.weight.data)
        if module.bias is not None:
            constant_(module.bias.data, 0)


def xavier_normal_initialization(module):
    if isinstance(module, nn.Embedding):
        xavier_normal_(module.weight.data)
    elif isinstance(module, nn.Linear):
        xavier_normal_(module.weight.data)
        if module.bias is not None:
            constant_(module.bias.data, 0)

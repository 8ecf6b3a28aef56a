"""Parameter initialisers with the reference's semantics (FoodRec/common/init.py:7-42): the weight
of every `nn.Embedding` / `nn.Linear` is re-drawn with a Xavier scheme and `nn.Linear` biases are
zeroed, module by module in `Module.apply` order -- so one seed yields the reference's initial
`state_dict` (checked bit-for-bit in tests/test_host.py)."""
import torch.nn as nn
from torch.nn import init as _init


def _make_initializer(draw):
    def initializer(module):
        is_linear = isinstance(module, nn.Linear)
        if not (is_linear or isinstance(module, nn.Embedding)):
            return
        draw(module.weight.data)
        if is_linear and module.bias is not None:
            _init.constant_(module.bias.data, 0)
    return initializer


xavier_uniform_initialization = _make_initializer(_init.xavier_uniform_)
xavier_normal_initialization = _make_initializer(_init.xavier_normal_)

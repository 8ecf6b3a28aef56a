"""Contrastive terms of CLUSSL on gathered `[2B, d]` views -- thin names over the fused kernels.

`correlation_distance` is the live term (FoodRec/models/pricai_modelx.py:263,409-437), `info_nce` the dormant
`CL_loss` (:354-378).  Both are launch/latency-bound at 2B = 1024; see `csrc/dcor.cu` and
`csrc/infonce.cu`.
"""
from .ops import correlation_distance, info_nce  # noqa: F401

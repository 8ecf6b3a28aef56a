"""foodrec_b200: B200-native graph-propagation + full-ranking hot path for the MMRec-derived food
recommenders (HealthRec / CLUSSL / SCHGN / LightGCN).

Host side (this package) mirrors the reference's model API; all device arithmetic is in
`csrc/*.cu`, reached through the C ABI declared in `include/foodrec_b200.h`
(`libfoodrec_b200.so`, loaded with ctypes by `_lib.py`).  There is no CPU fallback: importing
`foodrec_b200.ops` without the built library raises.
"""
__version__ = "0.1.0"

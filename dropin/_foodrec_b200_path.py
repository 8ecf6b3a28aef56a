"""Makes `foodrec_b200` importable from wherever the reference is run."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
import foodrec_b200  # noqa: E402,F401

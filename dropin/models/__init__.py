"""Overlay of the reference's `models` package: the four hot-path models resolve to the B200 drop-ins, every other
model module (`models.bm3`, `models.fgcn`, ...) falls through to the reference's own file.

The reference discovers models with `importlib.import_module('models.' + name.lower())`
(FoodRec/utils/utils.py:27-40) from the working directory `FoodRec/`.  Putting THIS directory's parent ahead of
`FoodRec/` on `sys.path` is all it takes -- no file of the reference is edited:

    cd /path/to/reference/FoodRec
    python /path/to/this/repo/dropin/run.py --model PRICAI_ModelX --dataset allrecipes      # runner.py's own arguments

(`PYTHONPATH=/path/to/this/repo/dropin` is enough for `python -c` / `python -m` entry points; `python runner.py` puts
the script's directory first, which is what `dropin/run.py` works around.)
"""
import os

_here = os.path.dirname(os.path.abspath(__file__))
for _cand in (os.environ.get("FOODREC_DIR"), os.getcwd()):
    if _cand:
        _ref = os.path.join(os.path.abspath(_cand), "models")
        if os.path.isdir(_ref) and os.path.abspath(_ref) != _here and _ref not in __path__:
            __path__.append(_ref)       # modules this overlay does not provide come from the reference

"""`get_model('CIKM_Model')` (FoodRec/utils/utils.py:27-40) resolves here when `dropin/` precedes `FoodRec/` on sys.path."""
from _foodrec_b200_path import foodrec_b200  # noqa: F401  (puts the repo root on sys.path)
from foodrec_b200.models.cikm_model import CIKM_Model  # noqa: E402,F401

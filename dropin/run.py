#!/usr/bin/env python
"""Run the UNMODIFIED reference (`FoodRec/runner.py`) on the B200 drop-in models.

    cd /path/to/reference/FoodRec && python /path/to/repo/dropin/run.py [runner.py's arguments]

`python runner.py` would put `FoodRec/` first on `sys.path`, so `models.pricai_modelx` would be the reference's own
file; this launcher puts `dropin/` first and then executes `runner.py` exactly as `python runner.py` does
(`runpy.run_path` leaves `sys.path` alone for a plain script).  Nothing in the reference tree is edited.
"""
import os
import runpy
import sys

here = os.path.dirname(os.path.abspath(__file__))
foodrec = os.path.abspath(os.environ.get("FOODREC_DIR", os.getcwd()))
if not os.path.isfile(os.path.join(foodrec, "runner.py")):
    sys.exit(f"dropin/run.py: no runner.py in {foodrec} (cd into the reference's FoodRec/ or set FOODREC_DIR)")
os.chdir(foodrec)
os.environ["FOODREC_DIR"] = foodrec
sys.path[:0] = [here, foodrec]
sys.argv[0] = os.path.join(foodrec, "runner.py")
runpy.run_path(sys.argv[0], run_name="__main__")

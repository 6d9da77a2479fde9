"""TEST / BENCH INFRASTRUCTURE: materialise the reference's own EM for Python 3.

The reference's `vp_localisation.py`, `probability_functions.py`, `coordinate_conversion.py` and
`calc_horizon.py` are Python 2.  They are read from /root/reference, given the five mechanical
patches of SURVEY.md section 8(c) (no arithmetic changes) and written to `dst`:

  1. print statements -> print()            2. `/` used for slicing -> `//`
  3. np.linalg.linalg.LinAlgError -> np.linalg.LinAlgError
  4. sklearn kwarg affinity= -> metric=     5. np.array(to_be_removed) -> dtype=int

`dst` is a temp directory (golden generation) or the git-ignored `baseline/_ref/` (BASELINE.md
section 3: the CPU baseline `bench.py` prints beside the port; it travels to the GPU box with the
work tree but never enters the repository's history).  Nothing here is imported by the product.
"""
import importlib
import os
import re
import sys
import warnings

REF = "/root/reference"
MODULES = ("vp_localisation", "probability_functions", "coordinate_conversion", "calc_horizon")


def available():
    return all(os.path.exists(os.path.join(REF, m + ".py")) for m in MODULES)


def materialise(dst):
    os.makedirs(dst, exist_ok=True)
    for name in MODULES:
        src = open(os.path.join(REF, name + ".py")).read()
        src = re.sub(r'^(\s*)print (.+)$', r'\1print(\2)', src, flags=re.M)
        src = src.replace("np.linalg.linalg.LinAlgError", "np.linalg.LinAlgError")
        src = src.replace("affinity='precomputed'", "metric='precomputed'")
        src = src.replace("np.array(to_be_removed)", "np.array(to_be_removed, dtype=int)")
        if name == "vp_localisation":
            src = src.replace("ra*sA/rA:(ra+1)*sA/rA, rb*sB/rB:(rb+1)*sB/rB",
                              "ra*sA//rA:(ra+1)*sA//rA, rb*sB//rB:(rb+1)*sB//rB")
            src = src.replace("max_response[0] + ra*sA/rA", "max_response[0] + ra*sA//rA")
            src = src.replace("max_response[1] + rb*sB/rB", "max_response[1] + rb*sB//rB")
        with open(os.path.join(dst, name + ".py"), "w") as fh:
            fh.write(src)
    return dst


def load(dst):
    """Import the materialised modules from `dst` (joblib workers re-import them: PYTHONPATH is set)."""
    if dst not in sys.path:
        sys.path.insert(0, dst)
    os.environ["PYTHONPATH"] = dst + os.pathsep + os.environ.get("PYTHONPATH", "")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vp = importlib.import_module("vp_localisation")
        prob = importlib.import_module("probability_functions")
    return vp, prob

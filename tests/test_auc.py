"""Row N4 (evaluation): auc.calc_auc mirror against golden vectors of the reference's own auc.py
(oracle/make_golden_auc.py)."""
import os

import numpy as np

from vanishing_points_2017_b200 import auc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "auc_cases.npz")


def test_calc_auc_matches_the_reference():
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        a, pts = auc.calc_auc(g["err_%d" % i].reshape(-1, 1), cutoff=float(g["cutoff_%d" % i]))
        np.testing.assert_allclose(a, float(g["auc_%d" % i]), rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(pts, g["pts_%d" % i], rtol=0, atol=0)

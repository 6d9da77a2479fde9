"""Row N1: the CPU restatement of calc_horizon.py against the golden
vectors produced by the reference's own calculate_horizon_and_ortho_vp
(oracle/make_golden_horizon.py -> tests/golden/horizon_cases.npz)."""
import os

import numpy as np

from oracle import horizon_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden", "horizon_cases.npz")


def cases():
    g = np.load(GOLD)
    for i in range(g["n_vp"].shape[0]):
        m = int(g["n_vp"][i])
        yield i, {"vp": g["vp"][i, :m].copy(), "counts": g["counts"][i, :m].copy()}, int(g["maxbest"][i]), \
            float(g["theta_vmin"][i]), float(g["theta_z"][i]), g["points"][i], g["combo"][i], bool(g["from_reference"][i])


def test_golden_file_is_meaningful():
    g = np.load(GOLD)
    assert g["from_reference"].sum() >= 120
    # a good share of the triplet cases select something other than the first triplet
    three = g["combo"][:, 2] >= 0
    assert three.sum() >= 90


def test_oracle_matches_reference_outputs():
    n_checked = 0
    for i, em, maxbest, tv, tz, points, combo, from_ref in cases():
        out = ho.calculate_horizon_and_ortho_vp(em, maxbest=maxbest, theta_vmin=tv, theta_z=tz)
        k = int((combo >= 0).sum())
        np.testing.assert_array_equal(np.asarray(out[5]).reshape(-1), combo[:k], err_msg="case %d" % i)
        for q in range(5):
            np.testing.assert_allclose(out[q], points[q], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg="case %d output %d" % (i, q))
        n_checked += 1
    assert n_checked == 160

"""Row N2: the CPU restatement of the LSD-output normalisation against golden
vectors produced by the reference's own detect_lsd_lines source (oracle/make_golden_lsd.py)."""
import os

import numpy as np

from oracle import lsd_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lsd_norm_cases.npz")


def test_oracle_is_bit_identical_to_the_reference():
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        out = lsd_oracle.segments_from_lsd(g["lsd_%d" % i], tuple(g["shape_%d" % i]))
        np.testing.assert_array_equal(out["segments"], g["segments_%d" % i])
        np.testing.assert_array_equal(np.signbit(out["segments"]), np.signbit(g["segments_%d" % i]))
        np.testing.assert_array_equal(out["nfa"], g["nfa_%d" % i])

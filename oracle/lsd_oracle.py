"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the LSD-output normalisation of
evaluation.detect_lsd_lines (reference evaluation.py:227-251, the part after the detector call) and of
the line construction of evaluation.py:158-168.

Pinned: oracle/make_golden_lsd.py executes the reference's own `detect_lsd_lines` source (extracted from
/root/reference/evaluation.py, with the detector call stubbed to return a prepared array) and stores
inputs and outputs in tests/golden/lsd_norm_cases.npz."""
import numpy as np


def segments_from_lsd(lsd_lines, image_shape):
    rows = np.array(lsd_lines, dtype=np.float64, copy=True)
    height, width = image_shape[0], image_shape[1]
    scale = np.maximum(width, height)                    # :233-234
    rows[:, 0] -= width / 2.0                            # :240-243
    rows[:, 1] -= height / 2.0
    rows[:, 2] -= width / 2.0
    rows[:, 3] -= height / 2.0
    rows[:, 0:4] /= (scale / 2.0)                        # :244-247
    rows[:, 1] *= -1                                     # :248-249
    rows[:, 3] *= -1
    return {"segments": rows[:, 0:4], "nfa": rows[:, 6] if rows.shape[1] > 6 else None}


def lines_from_segments(seg):
    """evaluation.py:161-168: [x1,y1,1] x [x2,y2,1] per segment."""
    p1 = np.concatenate([seg[:, 0:2], np.ones((seg.shape[0], 1))], axis=1)
    p2 = np.concatenate([seg[:, 2:4], np.ones((seg.shape[0], 1))], axis=1)
    return np.cross(p1, p2)

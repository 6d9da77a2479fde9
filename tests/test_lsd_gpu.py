"""Row N2 parity: segments_from_lsd_kernel (through the C ABI) vs the golden vectors of the reference's
detect_lsd_lines, bit for bit, and the pipeline fed with raw LSD rows vs the pipeline fed with segments."""
import os

import numpy as np
import pytest

from oracle import lsd_oracle
from vanishing_points_2017_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "lsd_norm_cases.npz")


def test_batch_is_bit_identical_to_the_reference():
    from vanishing_points_2017_b200 import evaluation
    g = np.load(GOLD)
    n = int(g["n_cases"])
    rows = [g["lsd_%d" % i] for i in range(n)]
    shapes = [tuple(g["shape_%d" % i]) for i in range(n)]
    out = evaluation.segments_from_lsd_batch(rows, shapes)
    off = out["offsets"]
    for i in range(n):
        seg = out["segments"][off[i]:off[i + 1]]
        np.testing.assert_array_equal(seg, g["segments_%d" % i])
        np.testing.assert_array_equal(np.signbit(seg), np.signbit(g["segments_%d" % i]))
        np.testing.assert_array_equal(out["nfa"][off[i]:off[i + 1]], g["nfa_%d" % i])
        np.testing.assert_array_equal(out["lines"][off[i]:off[i + 1]], lsd_oracle.lines_from_segments(g["segments_%d" % i]))
    single = evaluation.segments_from_lsd(rows[0], shapes[0])
    np.testing.assert_array_equal(single["segments"], g["segments_0"])
    np.testing.assert_array_equal(single["nfa"], g["nfa_0"])


def test_pipeline_from_raw_lsd_rows_equals_pipeline_from_segments():
    from vanishing_points_2017_b200 import cnn as vcnn, pipeline
    ws, bs = vcnn.random_weights(0, scale=3.0)
    pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
    batch = synth.make_batch(2, n_images=5)
    seg, off = batch["segments"], batch["offsets"]
    # raw pixel rows that normalise to the batch's segments: invert evaluation.py:240-249 for 640x480 images
    w, h = 640, 480
    raw = np.zeros((seg.shape[0], 7))
    raw[:, 0] = seg[:, 0] * (w / 2.0) + w / 2.0
    raw[:, 2] = seg[:, 2] * (w / 2.0) + w / 2.0
    raw[:, 1] = -seg[:, 1] * (w / 2.0) + h / 2.0
    raw[:, 3] = -seg[:, 3] * (w / 2.0) + h / 2.0
    norm = np.concatenate([lsd_oracle.segments_from_lsd(raw[off[b]:off[b + 1]], (h, w))["segments"] for b in range(5)])
    ref = pipe(norm, off)
    pipe.upload_lsd(raw, off, [w] * 5, [h] * 5)
    pipe.run()
    got = pipe.fetch()
    for a, b in zip(ref, got):
        assert (a["vp"] is None) == (b["vp"] is None)
        if a["vp"] is not None:
            np.testing.assert_array_equal(a["vp"], b["vp"])
            np.testing.assert_array_equal(a["vp_assoc"], b["vp_assoc"])

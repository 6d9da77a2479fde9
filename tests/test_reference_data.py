"""Row N3: the pickle-compatible export of the stage hand-off has the layout the
reference's consumers read (evaluation.py:152-183, 283-289, 344-350; example.py:62-65; benchmark.py:228-241)."""
import pickle

import numpy as np

from oracle import lsd_oracle
from vanishing_points_2017_b200 import evaluation, synth


def fake_results(B, ns):
    rs = np.random.RandomState(3)
    out = []
    for b in range(B):
        if b == 1:
            out.append({"vp_assoc": None, "vp": None, "counts": None, "count_id": None, "decision_metric": None,
                        "iterations": 0, "status": 2})
            continue
        m = 3
        out.append({"vp_assoc": rs.randint(-1, m, ns[b]).astype(np.int64), "vp": np.eye(3), "counts": np.array([5., 4., 3.]),
                    "counts_weighted": np.ones(3), "count_id": None, "decision_metric": None, "iterations": 7,
                    "sigma": np.full(3, 1e-6), "status": 0})
    return out


def test_layout_matches_the_reference_pickles(tmp_path):
    ns = [40, 25, 60]
    scenes = [synth.make_scene(70 + i, n) for i, n in enumerate(ns)]
    seg = np.concatenate([s["segments"] for s in scenes])
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    sph = np.zeros((3, 500, 500), np.uint8)
    resp = np.zeros((3, 20, 20), np.float32)
    data = evaluation.reference_data(fake_results(3, ns), seg, off, sph, resp, image_shapes=[(480, 640)] * 3,
                                     extra=[{"dataset": "yud", "image_file": "img%d.jpg" % b} for b in range(3)])
    assert len(data) == 3
    for b, d in enumerate(data):
        assert set(d) == {"lines", "sphere_image", "cnn_prediction", "EM_result"}                 # evaluation.py:177, 285, 349
        assert {"image_shape", "image", "line_segments", "lines", "dataset", "image_file"} <= set(d["lines"])   # :152, :170-171
        assert d["lines"]["line_segments"].shape == (ns[b], 4) and d["lines"]["lines"].shape == (ns[b], 3)
        assert d["sphere_image"].dtype == np.uint8 and d["cnn_prediction"].shape == (20, 20)
        raw = lsd_oracle.lines_from_segments(d["lines"]["line_segments"])
        # the EM normalises the rows in place and the array is stored back (vp_localisation.py:186, evaluation.py:350)
        np.testing.assert_allclose(np.linalg.norm(d["lines"]["lines"], axis=1), 1.0, rtol=1e-14)
        np.testing.assert_allclose(np.cross(d["lines"]["lines"], raw), 0.0, atol=1e-12)
        em = d["EM_result"]
        assert "status" not in em
        for k in ("vp", "counts", "vp_assoc", "decision_metric", "iterations", "count_id"):       # vp_localisation.py:441-442
            assert k in em
    assert data[1]["EM_result"]["vp"] is None                                                     # vp_localisation.py:205-206
    # what example.py:62-65 does with a pickle
    paths = [str(tmp_path / ("img%d.pkl" % b)) for b in range(3)]
    evaluation.dump_reference_pickles(data, paths)
    with open(paths[0], "rb") as fp:
        datum = pickle.load(fp)
    assert datum["EM_result"]["vp"].shape == (3, 3) and datum["lines"]["image_shape"] == (480, 640)

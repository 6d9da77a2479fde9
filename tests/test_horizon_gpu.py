"""Row N1 parity: horizon_kernel (through the C ABI) vs the golden vectors of the reference's
calc_horizon.calculate_horizon_and_ortho_vp and vs the CPU oracle on EM results of the pipeline."""
import os

import numpy as np
import pytest

from oracle import horizon_oracle as ho
from vanishing_points_2017_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "horizon_cases.npz")
TOL = dict(rtol=1e-9, atol=1e-12, equal_nan=True)       # float64, same operation order: rounding-level differences only


def test_against_reference_golden():
    from vanishing_points_2017_b200 import calc_horizon
    g = np.load(GOLD)
    n = g["n_vp"].shape[0]
    for i in range(n):
        m = int(g["n_vp"][i])
        em = {"vp": g["vp"][i, :m].copy(), "counts": g["counts"][i, :m].copy()}
        kw = dict(maxbest=int(g["maxbest"][i]), theta_vmin=float(g["theta_vmin"][i]), theta_z=float(g["theta_z"][i]))
        out = calc_horizon.calculate_horizon_and_ortho_vp_batch([em], **kw)[0]
        k = int((g["combo"][i] >= 0).sum())
        np.testing.assert_array_equal(out[5], g["combo"][i, :k], err_msg="case %d" % i)      # index work: exact
        for q in range(5):
            np.testing.assert_allclose(out[q], g["points"][i, q], err_msg="case %d output %d" % (i, q), **TOL)


def test_batch_equals_single_and_drop_in_signature():
    from vanishing_points_2017_b200 import calc_horizon
    g = np.load(GOLD)
    sel = [i for i in range(g["n_vp"].shape[0]) if g["maxbest"][i] == 10 and abs(g["theta_vmin"][i] - np.pi / 10) < 1e-12
           and abs(g["theta_z"][i] - np.pi / 4) < 1e-12 and g["n_vp"][i] > 0]
    ems = [{"vp": g["vp"][i, :int(g["n_vp"][i])].copy(), "counts": g["counts"][i, :int(g["n_vp"][i])].copy()} for i in sel]
    batch = calc_horizon.calculate_horizon_and_ortho_vp_batch(ems)
    for em, b in zip(ems, batch):
        single = calc_horizon.calculate_horizon_and_ortho_vp(em)              # the reference's defaults
        for q in range(6):
            np.testing.assert_array_equal(single[q], b[q])
    with pytest.raises(AttributeError):
        calc_horizon.calculate_horizon_and_ortho_vp({"vp": None, "counts": None})
    # a failed image in a batch gets the reference's default horizon (calc_horizon.py:207-212)
    out = calc_horizon.calculate_horizon_and_ortho_vp_batch([{"vp": None, "counts": None}])[0]
    np.testing.assert_array_equal(out[0], [-1.0, 0.0, 1.0])
    np.testing.assert_array_equal(out[1], [1.0, 0.0, 1.0])


def test_pipeline_horizons_match_oracle_on_device_resident_em_result():
    from vanishing_points_2017_b200 import cnn as vcnn, pipeline
    ws, bs = vcnn.random_weights(0, scale=3.0)
    pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
    batch = synth.make_batch(2, n_images=6)
    res = pipe(batch["segments"], batch["offsets"])
    hz = pipe.horizons(maxbest=20)                                            # example.py:65 uses maxbest=20
    assert len(hz) == len(res)
    # with ground truth: same tuples plus the errors of benchmark.py:247-253
    truth = np.tile(np.array([0.02, 1.0, 0.05]), (len(res), 1))
    hz2, err = pipe.horizons(maxbest=20, true_horizons=truth, scales=640.0, image_heights=480.0)
    for a, b, e in zip(hz, hz2, err):
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_allclose(e, ho.horizon_error(a[0], a[1], truth[0], 640.0, 480.0), rtol=1e-12, equal_nan=True)
    for r, h in zip(res, hz):
        if r["vp"] is None:
            np.testing.assert_array_equal(h[0], [-1.0, 0.0, 1.0])
            continue
        ref = ho.calculate_horizon_and_ortho_vp(r, maxbest=20)
        np.testing.assert_array_equal(h[5], np.asarray(ref[5]).reshape(-1))
        for q in range(5):
            np.testing.assert_allclose(h[q], ref[q], **TOL)


def test_horizon_errors_and_auc_like_benchmark_py():
    """benchmark.py:233-262: horizon -> error against the ground truth -> AUC."""
    from vanishing_points_2017_b200 import auc, calc_horizon
    g = np.load(GOLD)
    sel = [i for i in range(g["n_vp"].shape[0]) if g["maxbest"][i] == 10 and g["n_vp"][i] > 2 and abs(g["theta_z"][i] - np.pi / 4) < 1e-12
           and abs(g["theta_vmin"][i] - np.pi / 10) < 1e-12]
    assert len(sel) >= 10
    ems = [{"vp": g["vp"][i, :int(g["n_vp"][i])].copy(), "counts": g["counts"][i, :int(g["n_vp"][i])].copy()} for i in sel]
    rs = np.random.RandomState(2)
    truth = np.stack([np.array([np.sin(a), np.cos(a), d]) for a, d in zip(rs.normal(0, 0.1, len(sel)), rs.normal(0, 0.2, len(sel)))])
    scales = rs.choice([640.0, 800.0], len(sel))
    heights = rs.choice([480.0, 533.0, 600.0], len(sel))
    out, err = calc_horizon.calculate_horizon_and_ortho_vp_batch(ems, true_horizons=truth, scales=scales, image_heights=heights)
    ref = np.array([ho.horizon_error(o[0], o[1], truth[k], scales[k], heights[k]) for k, o in enumerate(out)])
    np.testing.assert_allclose(err, ref, rtol=1e-12, atol=1e-15, equal_nan=True)
    a, pts = auc.calc_auc(err[np.isfinite(err)])
    assert 0.0 <= a <= 1.0 and pts.shape[1] == 2

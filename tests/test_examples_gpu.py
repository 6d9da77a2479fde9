"""BASELINE.json configs[0]: the reference's four bundled example images (assets/examples) through the
whole path -- raw LSD rows (tests/golden/examples_lsd.npz, produced with the reference's own lsd.c by
oracle/make_golden_examples.py) -> normalised segments -> lines -> sphere image -> CNN -> EM -> horizon --
every stage compared with its oracle on the pipeline's own intermediates."""
import os

import numpy as np
import pytest

from oracle import cnn_oracle, horizon_oracle, lsd_oracle, sphere_oracle, vp_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "examples_lsd.npz")


def same_vps(a, r):
    if (a["vp"] is None) != (r["vp"] is None):
        return False
    if r["vp"] is None:
        return True
    if a["vp"].shape != r["vp"].shape:
        return False
    return np.arccos(np.minimum(np.abs(np.sum(a["vp"] * r["vp"], axis=1)), 1.0)).max() < 1e-4      # north_star: 1e-4 rad


def test_example_images_stage_by_stage():
    from vanishing_points_2017_b200 import pipeline
    g = np.load(GOLD)
    n = int(g["n_images"])
    rows = [g["lsd_%d" % i] for i in range(n)]
    shapes = [tuple(int(v) for v in g["shape_%d" % i]) for i in range(n)]
    assert [r.shape[0] for r in rows] == [370, 628, 347, 1191]
    off = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows])]).astype(np.int32)
    ws, bs = cnn_oracle.random_weights(0, scale=3.0)
    pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
    pipe.upload_lsd(np.concatenate(rows), off, [s[1] for s in shapes], [s[0] for s in shapes])
    pipe.run()
    res, sig, sph = pipe.fetch(want_response=True, want_sphere=True)
    segs = [lsd_oracle.segments_from_lsd(rows[i], shapes[i])["segments"] for i in range(n)]
    # S0 + S1: bit-exact bins, hence identical images
    for i in range(n):
        lines = lsd_oracle.lines_from_segments(segs[i])
        ref_img = sphere_oracle.votes_to_image(sphere_oracle.sphere_votes(lines, 500))
        assert np.array_equal(sph[i], ref_img)
    # C1: bf16 tensor-core path vs the fp32 oracle
    rsig, rlog = cnn_oracle.forward(sph, ws, bs)
    assert np.max(np.abs(sig - rsig)) <= 0.25 * 1e-2 * np.max(np.abs(rlog)) + 1e-6
    # E0-E12 on the pipeline's own response and sphere image; N1 on the EM result
    decided, boundary = 0, []
    rs = np.random.RandomState(0)
    hz = pipe.horizons(maxbest=20)
    for i in range(n):
        s = segs[i]

        def oracle_em(sp):
            try:
                return vp_oracle.expectation_maximisation(lsd_oracle.lines_from_segments(sp), sp.copy(), sig[i].astype(np.float64),
                                                          sphere_image=sph[i])
            except ValueError:
                return {"vp": None}
        ref = oracle_em(s)
        if same_vps(res[i], ref):
            decided += 1
            if ref["vp"] is not None:
                np.testing.assert_array_equal(res[i]["counts"], ref["counts"])
                np.testing.assert_array_equal(res[i]["vp_assoc"], ref["vp_assoc"])
        else:
            # an image on a decision boundary of the reference algorithm itself decides nothing
            flips = sum(0 if same_vps(oracle_em(s * (1.0 + 1e-14 * rs.standard_normal(s.shape))), ref) else 1 for _ in range(3))
            assert flips > 0, "image %d: VPs differ from an oracle that is stable under perturbation" % i
            boundary.append(i)
        if res[i]["vp"] is not None:
            h = horizon_oracle.calculate_horizon_and_ortho_vp(res[i], maxbest=20)
            np.testing.assert_array_equal(hz[i][5], np.asarray(h[5]).reshape(-1))
            for q in range(5):
                np.testing.assert_allclose(hz[i][q], h[q], rtol=1e-9, atol=1e-12, equal_nan=True)
    # the set of decision-boundary images is pinned: a regression that turns a stable image into a skipped one fails here
    assert boundary == EXPECTED_BOUNDARY_IMAGES, "images skipped as decision-boundary cases: %r (pinned: %r)" % (
        boundary, EXPECTED_BOUNDARY_IMAGES)
    assert decided == n - len(EXPECTED_BOUNDARY_IMAGES)


# Images of assets/examples whose EM result flips in the ORACLE itself under a 1e-14 relative perturbation of the
# segments with this CNN response; measured on the B200 box, `profiles/r2_parity_report.txt`.
# Image 3 (N = 1191): the oracle's own result changes in 6 of 6 perturbed runs there.
EXPECTED_BOUNDARY_IMAGES = [3]

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
            boundary.append(i)
        if res[i]["vp"] is not None:
            h = horizon_oracle.calculate_horizon_and_ortho_vp(res[i], maxbest=20)
            np.testing.assert_array_equal(hz[i][5], np.asarray(h[5]).reshape(-1))
            for q in range(5):
                np.testing.assert_allclose(hz[i][q], h[q], rtol=1e-9, atol=1e-12, equal_nan=True)
    # Only the images KNOWN to be sensitive may be skipped: a regression that turns a stable image into a skipped one fails.
    assert set(boundary) <= set(SENSITIVE_IMAGES), "images skipped as decision-boundary cases: %r (allowed: %r)" % (
        boundary, SENSITIVE_IMAGES)
    assert decided >= n - len(SENSITIVE_IMAGES)
    # ... and EVERY image, the sensitive ones included, must agree with the oracle strictly when both are stopped after
    # 16 iterations: identical VP sets, counts and line association, VPs within 1e-4 rad (they agree to 1e-7 rad there,
    # profiles/r2_em_divergence.txt; the long runs of images 2 and 3 part ways later, at iterations 19 and > 21, when
    # rounding-level differences in (1 - |cos|)^2 of near-perfect lines have been amplified by ill-conditioned refits).
    from vanishing_points_2017_b200 import vp_localisation as em
    for i in range(n):
        lines = lsd_oracle.lines_from_segments(segs[i])
        ref = vp_oracle.expectation_maximisation(lines.copy(), segs[i].copy(), sig[i].astype(np.float64), sphere_image=sph[i], num_iter=16)
        got = em.expectation_maximisation(lines.copy(), segs[i].copy(), sig[i].astype(np.float64), sphere_image=sph[i], num_iter=16)
        assert same_vps(got, ref), "image %d: truncated runs differ" % i
        assert got["iterations"] == ref["iterations"]
        np.testing.assert_array_equal(got["counts"], ref["counts"])
        np.testing.assert_array_equal(got["vp_assoc"], ref["vp_assoc"])


# Images of assets/examples whose full EM run sits on a decision boundary of the reference algorithm with the random-init
# CNN response.  Image 3 (N = 1191): the oracle's own result changes in 6 of 6 runs on inputs
# perturbed by 1e-14 (profiles/r2_parity_report.txt).  Image 2 (N = 347): the convergence test of iteration 19
# (max_err = 4.2e-3 against the 5e-3 threshold) falls either way depending on last-bit differences in the E-step
# (profiles/r2_em_divergence.txt).  Which side the ORACLE falls on depends on the host's BLAS, so the set is an
# allow-list, not an expectation.
SENSITIVE_IMAGES = [2, 3]

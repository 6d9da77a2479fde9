"""Stage S0/S1 parity: CUDA kernels (through the C ABI) vs oracle/sphere_oracle.py."""
import numpy as np
import pytest

from oracle import sphere_oracle as so
from vanishing_points_2017_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from vanishing_points_2017_b200 import sphere_mapping
    return sphere_mapping


def test_lines_from_segments_bit_exact(sm):
    sc = synth.make_scene(5, 777)
    np.testing.assert_array_equal(sm.lines_from_segments(sc["segments"]), sc["lines"])


@pytest.mark.parametrize("N,S", [(2, 64), (37, 250), (300, 500), (700, 500)])
def test_votes_bins_bit_exact(sm, N, S):
    sc = synth.make_scene(100 + N, N)
    hist, img = sm.sphere_votes(sc["lines"], S)
    ref = so.sphere_votes(sc["lines"], S)
    assert hist.dtype == np.uint32 and hist.sum() == N * (N - 1) // 2
    np.testing.assert_array_equal(hist.astype(np.int64), ref)          # bit-exact bin indices
    np.testing.assert_array_equal(img, so.votes_to_image(ref))


def test_votes_weighted(sm):
    sc = synth.make_scene(9, 200)
    w = np.random.RandomState(0).uniform(0.1, 2.0, 200)
    wh, img = sm.sphere_votes(sc["lines"], 250, weights=w)
    ref = so.sphere_votes(sc["lines"], 250, weights=w)
    # tolerance stated by BASELINE.json north_star: accumulated histograms within 1e-5 relative
    np.testing.assert_allclose(wh, ref / 65536.0, rtol=1e-5, atol=0)
    np.testing.assert_array_equal(img, so.votes_to_image(ref))


def test_votes_ragged_batch_and_empty(sm):
    ns = [0, 1, 5, 130, 257, 0, 64]
    scs = [synth.make_scene(300 + i, n) if n else None for i, n in enumerate(ns)]
    lines = np.concatenate([s["lines"] for s in scs if s is not None] + [np.zeros((0, 3))])
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    r = sm.sphere_map_batch(lines, off, 128, "votes")
    for b, n in enumerate(ns):
        ref = so.sphere_votes(lines[off[b]:off[b + 1]], 128)
        np.testing.assert_array_equal(r["hist"][b].astype(np.int64), ref)
        np.testing.assert_array_equal(r["image"][b], so.votes_to_image(ref))


def test_votes_degenerate_pairs(sm):
    # identical / parallel lines: zero cross product must be skipped, not binned
    lines = np.array([[1.0, 2.0, 0.5], [1.0, 2.0, 0.5], [2.0, 4.0, 1.0], [0.3, -0.1, 0.2]])
    hist, _ = sm.sphere_votes(lines, 100)
    ref = so.sphere_votes(lines, 100)
    np.testing.assert_array_equal(hist.astype(np.int64), ref)
    assert hist.sum() == 3


@pytest.mark.parametrize("N,S", [(1, 100), (60, 250), (400, 500)])
def test_curves_counts_bit_exact(sm, N, S):
    sc = synth.make_scene(400 + N, N)
    r = sm.sphere_map_batch(sc["lines"], [0, N], S, "curves", alpha=0.1)
    ref = so.sphere_curve_counts(sc["lines"], S)
    np.testing.assert_array_equal(r["hist"][0].astype(np.int64), ref)
    np.testing.assert_array_equal(r["image"][0], so.curve_image(ref, 0.1))


def test_curves_kernel_variants_are_bit_identical(sm, monkeypatch):
    """The band kernel with float32 pre-binned rows (default), the same kernel in float64 only, and the
    one-CTA-per-line kernel with its difference array in HBM (grids beyond 1536 cells) give the same counts,
    on a ragged batch with an empty image and on lines of wild scales / zero components."""
    rs = np.random.RandomState(3)
    ns = [0, 700, 33, 1, 260]
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    lines = np.concatenate([synth.make_scene(900 + i, max(n, 1))["lines"][:n] for i, n in enumerate(ns)])
    wild = rs.standard_normal((260, 3)) * np.exp(rs.uniform(-8, 8, (260, 3)))
    wild[:20, 1] = 0.0
    wild[20:40, 0] = 0.0
    wild[40:60, 2] = 0.0
    lines[-260:] = wild
    for S in (500, 77):
        base = sm.sphere_map_batch(lines, off, S, "curves")
        for var in ("VPK_CURVES_F64", "VPK_CURVES_GLOBAL_DIFF"):
            monkeypatch.setenv(var, "1")
            other = sm.sphere_map_batch(lines, off, S, "curves")
            monkeypatch.delenv(var)
            np.testing.assert_array_equal(base["hist"], other["hist"])
            np.testing.assert_array_equal(base["image"], other["image"])
        assert base["hist"][0].sum() == 0
        ref = so.sphere_curve_counts(lines[off[1]:off[2]], S)
        np.testing.assert_array_equal(base["hist"][1].astype(np.int64), ref)
    big = sm.sphere_map_batch(lines[off[1]:off[2]], [0, 700], 1600, "curves")          # > kBandMaxS: the HBM form
    np.testing.assert_array_equal(big["hist"][0].astype(np.int64), so.sphere_curve_counts(lines[off[1]:off[2]], 1600))


def test_sphere_line_plot_drop_in(sm):
    sc = synth.make_scene(77, 150)
    lines = sc["lines"].copy()
    ref_lines = sc["lines"].copy()
    img = sm.sphere_line_plot(lines, 250, alpha=0.1, f=2.0)
    ref = so.sphere_line_plot(ref_lines, 250, alpha=0.1, f=2.0)
    assert img.dtype == np.uint8 and img.shape == (250, 250)
    np.testing.assert_array_equal(lines, ref_lines)       # same in-place mutation as the reference
    np.testing.assert_array_equal(img, ref)
    with pytest.raises(NotImplementedError):
        sm.sphere_line_plot(lines, 250, alternative=True)


def test_full_size_properties(sm):
    """HLW-shaped image count is checked through size-independent properties:
    every valid pair votes exactly once."""
    batch = synth.make_batch(4, n_images=24)
    off = batch["offsets"]
    r = sm.sphere_map_batch(batch["lines"], off, 500, "votes", want_image=False)
    n = np.diff(off).astype(np.int64)
    np.testing.assert_array_equal(r["hist"].reshape(len(n), -1).sum(axis=1), n * (n - 1) // 2)


def test_votes_against_the_reference_projection_golden(sm, golden_dir):
    """Directly against vectors of the reference's own coordinate_conversion.py (oracle/make_golden_sphere.py):
    every one of the 101 025 pairwise intersections of a 450-line scene must land in the cell that
    round(angle_to_index(point_to_angle(p))) gives (coordinate_conversion.py:23-35, 53-61)."""
    import os
    g = np.load(os.path.join(golden_dir, "sphere_cases.npz"))
    want = np.bincount(g["votes_rows"].astype(np.int64) * 500 + g["votes_cols"], minlength=250000).reshape(500, 500)
    hist, _ = sm.sphere_votes(g["votes_lines"], 500)
    np.testing.assert_array_equal(hist.astype(np.int64), want)
    # and from the segments (S0 + S1 on the device)
    lines = sm.lines_from_segments(g["votes_segments"])
    np.testing.assert_array_equal(lines, g["votes_lines"])


def test_curves_cover_every_sample_of_the_reference_expression(sm, golden_dir):
    """curves mode: every (row, column) sample of beta(alpha) as the reference's own expression
    (sphere_mapping.py:40, 61-63) produces it must be a covered pixel of that line's raster."""
    import os
    g = np.load(os.path.join(golden_dir, "sphere_cases.npz"))
    lines, rows = g["curves_lines"], g["curves_rows"].astype(np.int64)
    a = g["curves_alpha"]
    cols = np.clip(np.floor((a / np.pi + 0.5 - 0.5 / 500) * 500 + 0.5), 0, 499).astype(np.int64)
    for i in range(0, lines.shape[0], 6):
        r = sm.sphere_map_batch(lines[i:i + 1], [0, 1], 500, "curves")
        cover = r["hist"][0]
        ok = rows[i] >= 0
        assert cover.max() == 1
        assert np.all(cover[rows[i][ok], cols[ok]] == 1)

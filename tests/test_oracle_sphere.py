"""Rows T0 / S1: the sphere-mapping restatement (oracle/sphere_oracle.py) pinned against vectors computed
by the reference's OWN code (oracle/make_golden_sphere.py: coordinate_conversion.py imported unmodified;
the great-circle expression cut out of sphere_mapping.py:40, 61-63)."""
import os

import numpy as np
import pytest

from oracle import sphere_oracle as so

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sphere_cases.npz"))


@pytest.mark.parametrize("S", [500, 250, 20])
def test_index_angle_maps_bit_exact(S):
    got = np.stack([so.index_to_angle(i, (S, S)) for i in G["t0_idx_%d" % S]])
    np.testing.assert_array_equal(got, G["t0_idx2ang_%d" % S])
    got = np.stack([so.angle_to_index(a, (S, S)) for a in G["t0_ang_%d" % S]])
    np.testing.assert_array_equal(got, G["t0_ang2idx_%d" % S])


def test_angle_point_maps_bit_exact():
    got = np.stack([so.angle_to_point(a) for a in G["t0_a2p_in"]])
    np.testing.assert_array_equal(got, G["t0_a2p_out"])            # sign(0) == 0 quirk included
    got = np.stack([so.point_to_angle(p) for p in G["t0_p2a_in"]])
    np.testing.assert_array_equal(got, G["t0_p2a_out"])


def test_votes_bins_of_all_pairs_match_the_reference_projection():
    """101 025 intersections of a 450-line scene: cell = round(angle_to_index(point_to_angle(p)))."""
    lines = G["votes_lines"]
    ii, jj = np.triu_indices(lines.shape[0], 1)
    row, col, valid = so.pair_bins(lines[ii], lines[jj], 500)
    assert valid.all()
    np.testing.assert_array_equal(row, G["votes_rows"].astype(np.int64))
    np.testing.assert_array_equal(col, G["votes_cols"].astype(np.int64))
    hist = np.bincount(G["votes_rows"].astype(np.int64) * 500 + G["votes_cols"], minlength=250000).reshape(500, 500)
    np.testing.assert_array_equal(so.sphere_votes(lines, 500), hist)


def test_curve_rows_match_the_reference_expression():
    rows, cols = so.curve_rows(G["curves_lines"], 500)
    np.testing.assert_array_equal(rows, G["curves_rows"].astype(np.int64))
    a = G["curves_alpha"]
    assert a.shape == (10000,) and a[0] == -np.pi / 2 and a[-1] == np.pi / 2
    np.testing.assert_array_equal(cols, np.clip(np.floor((a / np.pi + 0.5 - 0.5 / 500) * 500 + 0.5), 0, 499).astype(np.int64))

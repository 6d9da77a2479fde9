"""The numpy restatement (oracle/vp_oracle.py) against vectors produced by the
reference's own EM (oracle/make_golden.py -> tests/golden/)."""
import glob
import os

import numpy as np
import pytest

from oracle import vp_oracle as vo

TOL = dict(rtol=1e-9, atol=1e-12)


@pytest.fixture(scope="module")
def fn(golden_dir):
    return np.load(os.path.join(golden_dir, "em_functions_n72.npz"))


def test_lsim(fn):
    np.testing.assert_allclose(vo.calc_lsim(fn["lp"], sigma=1), fn["lsim"], rtol=1e-9, atol=1e-25)


def test_line_rating(fn):
    np.testing.assert_allclose(vo.line_rating_knn(fn["lp"], k2=4), fn["lscore_k4"], **TOL)


def test_lines_angles(fn):
    np.testing.assert_allclose(vo.lines_angles(fn["lp"]), fn["langles"], **TOL)


def test_find_maxima(fn):
    np.testing.assert_array_equal(vo.find_maxima(fn["resp"]), fn["maxima"])


def test_find_initial_vps(fn):
    np.testing.assert_allclose(vo.find_initial_vps(fn["sphere_image"], fn["resp"], 25), fn["v0"], **TOL)


def test_pdf_params(fn):
    pp = vo.pdf_params(fn["resp"])
    np.testing.assert_allclose(pp.means, fn["pdf_means"], **TOL)
    np.testing.assert_allclose(pp.weights, fn["pdf_weights"], **TOL)
    assert pp.sigma == float(fn["pdf_sigma"])


def _estep(fn):
    pp = vo.pdf_params(fn["resp"])
    s = np.ones(fn["v0"].shape[0]) * pp.sigma * 1e-6
    return pp, s, vo.calc_probabilities(pp, fn["v0"], fn["lp"], s)


def test_estep(fn):
    _, _, p = _estep(fn)
    np.testing.assert_allclose(p.angles, fn["angles"], **TOL)
    np.testing.assert_allclose(p.v, fn["p_v"], rtol=1e-9, atol=1e-300)
    # lvsq = (1-|cos|)^2 is ill-conditioned near 0: compare 1-|cos| absolutely
    np.testing.assert_allclose(np.sqrt(p.lvsq), np.sqrt(fn["lvsq"]), rtol=0, atol=2e-15)
    np.testing.assert_allclose(p.lv, fn["p_lv"], rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(p.l, fn["p_l"], rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(p.vl, fn["p_vl"], rtol=1e-6, atol=1e-300)


def test_weight_matrix_counts_newvp(fn):
    _, s, p = _estep(fn)
    w = vo.weight_matrix(p.vl, fn["lweight"], fn["lsim"], bias=1)
    np.testing.assert_allclose(w, fn["w"], rtol=1e-6, atol=1e-300)
    c, cw, assoc = vo.calc_vp_line_counts(fn["v0"], fn["lp"], s, w, fn["lweight"], 1.96 ** 2)
    np.testing.assert_array_equal(c, fn["counts"])
    np.testing.assert_array_equal(assoc, fn["assoc"])
    np.testing.assert_allclose(cw, fn["counts_weighted"], **TOL)
    for m in range(w.shape[0]):
        nv = vo.calc_new_vanishing_point(fn["l"], w[m])
        if nv is None:
            assert not fn["newvp"][m].any()
        else:
            assert np.arccos(min(1.0, abs(nv @ fn["newvp"][m]))) < 1e-7


def test_merge(fn):
    pp = vo.pdf_params(fn["resp"])
    lsim = fn["lsim"]
    v = fn["merge_in_v"].copy()
    _, v_out, s_out = vo.merge_vps(v.copy(), v, fn["merge_in_s"].copy(), fn["l"], 1e-2, fn["lweight"], lsim,
                                   lsim.sum(axis=0), 1, pp, fn["lp"])
    assert v_out.shape == fn["merge_out_v"].shape
    np.testing.assert_allclose(v_out, fn["merge_out_v"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(s_out, fn["merge_out_s"], rtol=1e-6)


def test_split(golden_dir):
    g = np.load(os.path.join(golden_dir, "em_split_n160.npz"))
    v, s, added = vo.split_best_vp(g["v_in"].copy(), g["s_in"].copy(), g["lp"], g["l"], g["w"], g["lweight"],
                                   g["langles"], 1e-3)
    assert v.shape == g["v_out"].shape
    np.testing.assert_allclose(v, g["v_out"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(s, g["s_out"], rtol=1e-12)


def match_vps(a, b):
    """Greedy one-to-one angular matching; returns max angle over matched rows."""
    ang = np.arccos(np.minimum(np.abs(a @ b.T), 1.0))
    return ang


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "em_full_*.npz"))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_full_em(path):
    g = np.load(path)
    res = vo.expectation_maximisation(g["lines"].copy(), g["segments"].copy(), g["resp"].copy(),
                                      sphere_image=g["sphere_image"].copy())
    assert res["iterations"] == int(g["iterations"])
    assert res["vp"].shape == g["vp"].shape
    ang = np.arccos(np.minimum(np.abs(np.sum(res["vp"] * g["vp"], axis=1)), 1.0))
    assert ang.max() < 1e-7, ang
    np.testing.assert_array_equal(res["counts"], g["counts"])
    np.testing.assert_array_equal(res["vp_assoc"], g["vp_assoc"])
    np.testing.assert_allclose(res["sigma"], g["sigma"], rtol=1e-6)
    np.testing.assert_allclose(res["counts_weighted"], g["counts_weighted"], rtol=1e-9)


def test_split_real(golden_dir):
    """A split that really happened inside the reference's run of the n600_long scene."""
    g = np.load(os.path.join(golden_dir, "em_split_real_n600.npz"))
    v, s, added = vo.split_best_vp(g["v_in"].copy(), g["s_in"].copy(), g["lp"], g["l"], g["w"], g["lweight"],
                                   g["langles"], 1e-3)
    assert added == 1 and v.shape == g["v_out"].shape
    ang = np.arccos(np.minimum(np.abs(np.sum(v * g["v_out"], axis=1)), 1.0))
    assert ang.max() < 1e-7
    np.testing.assert_allclose(s, g["s_out"], rtol=1e-12)

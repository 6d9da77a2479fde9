"""The EM state machine the CUDA kernels execute (csrc/em_core.cuh), compiled for
the host with a one-thread team (tests/hostsim/, TEST INFRASTRUCTURE: not part of
libvpk.so), against the reference's golden vectors and the numpy oracle.  This is
the CPU-side check of the control flow; the GPU tests (test_em_gpu.py) check the
kernels themselves through the C ABI."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import sphere_oracle as so
from oracle import vp_oracle as vo
from vanishing_points_2017_b200 import _lib, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "em_hostsim.cpp")
CORE = os.path.join(HERE, "..", "vanishing_points_2017_b200", "csrc", "em_core.cuh")
SO = os.path.join(HERE, "hostsim", "_em_hostsim.so")
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "em_full_*.npz")))
VP_TOL_RAD = 1e-4
M = 64


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.hostsim_em.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.POINTER(_lib.EmConfig)] + [C.c_void_p] * 10

    def run(lines, segs, resp, sphere, init_vp=None, **kw):
        N = lines.shape[0]
        cfg = _lib.EmConfig(100, 25, 10, 3, 1, 1, 1, 1, 1.0, 1e-3, 1.96 ** 2, 5e-3, 1e-200)
        for k, v in kw.items():
            setattr(cfg, k, int(v) if isinstance(v, bool) else v)
        lines = np.ascontiguousarray(lines, np.float64)
        segs = np.ascontiguousarray(segs, np.float64)
        resp = np.ascontiguousarray(resp, np.float64)
        sphere = np.ascontiguousarray(sphere, np.uint8)
        iv = None if init_vp is None else np.ascontiguousarray(init_vp, np.float64)
        st, nv, it, steps = (np.zeros(1, np.int32) for _ in range(4))
        vp, sg, cw = np.zeros((M, 3)), np.zeros(M), np.zeros(M)
        cnt, assoc, dm = np.zeros(M, np.int32), np.zeros(max(N, 1), np.int32), np.zeros(max(M * N, 1))
        rc = lib.hostsim_em(_lib.ptr(lines), _lib.ptr(segs), N, _lib.ptr(resp), _lib.ptr(sphere), sphere.shape[0],
                            _lib.ptr(iv), 0 if iv is None else iv.shape[0], C.byref(cfg), *[_lib.ptr(a) for a in
                            (st, nv, it, vp, sg, cnt, cw, assoc, dm, steps)])
        assert rc == 0, "state machine did not terminate"
        if st[0] != 0:
            return {"vp": None, "status": int(st[0])}
        m = int(nv[0])
        return {"vp": vp[:m], "sigma": sg[:m], "counts": cnt[:m].astype(float), "counts_weighted": cw[:m],
                "vp_assoc": assoc[:N].astype(np.int64), "decision_metric": dm[:m * N].reshape(m, N),
                "iterations": int(it[0]), "supersteps": int(steps[0])}
    return run


def compare(res, ref):
    if ref["vp"] is None:
        assert res["vp"] is None
        return
    assert res["vp"] is not None
    assert res["iterations"] == int(ref["iterations"])
    assert res["vp"].shape == ref["vp"].shape
    ang = np.arccos(np.minimum(np.abs(np.sum(res["vp"] * ref["vp"], axis=1)), 1.0))
    assert ang.max() < VP_TOL_RAD, ang
    np.testing.assert_array_equal(res["counts"], ref["counts"])
    np.testing.assert_array_equal(res["vp_assoc"], ref["vp_assoc"])
    np.testing.assert_allclose(res["sigma"], ref["sigma"], rtol=1e-5)


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[8:-4])
def test_state_machine_against_reference_golden(sim, path):
    g = np.load(path)
    res = sim(g["lines"], g["segments"], g["resp"], g["sphere_image"])
    compare(res, {k: g[k] for k in ("vp", "counts", "vp_assoc", "sigma", "iterations")})
    np.testing.assert_allclose(res["decision_metric"], g["decision_metric"], rtol=1e-5, atol=1e-300)


@pytest.mark.parametrize("seed,N,noise", [(1, 3, 0.5), (2, 12, 0.5), (3, 64, 0.5), (4, 333, 1.0), (6, 520, 2.5)])
def test_state_machine_against_oracle(sim, seed, N, noise):
    sc = synth.make_scene(7000 + seed, N, 800, 600, noise_deg=noise)
    img = so.votes_to_image(so.sphere_votes(sc["lines"], 500))
    resp = synth.ideal_response(sc["vps"], seed=seed)
    try:
        ref = vo.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp.copy(), sphere_image=img)
    except ValueError:
        ref = {"vp": None}
    compare(sim(sc["lines"], sc["segments"], resp, img), ref)


def test_state_machine_kwargs(sim):
    sc = synth.make_scene(31, 200)
    img = so.votes_to_image(so.sphere_votes(sc["lines"], 250))
    resp = synth.ideal_response(sc["vps"], seed=31)
    for kw in (dict(do_merge=False), dict(do_split=False), dict(use_weights=False), dict(do_iterations=False),
               dict(num_init_vp=5), dict(num_iter=3)):
        ref = vo.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp.copy(), sphere_image=img, **kw)
        compare(sim(sc["lines"], sc["segments"], resp, img, **kw), ref)
    iv = sc["vps"] * 3.0
    ref = vo.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp.copy(), sphere_image=img, init_vp=iv)
    compare(sim(sc["lines"], sc["segments"], resp, img, init_vp=iv), ref)


def test_refit_matches_svd_also_when_one_line_dominates(sim):
    """E7: the smallest right-singular vector of diag(w / max w) l (vp_localisation.py:453-479, LAPACK SVD
    in the reference) from the 3x3 scatter matrix.  When one line dominates, the scatter matrix in the
    original basis loses the small singular values; the refinement sweep must recover the SVD's answer."""
    lib = C.CDLL(SO)
    lib.hostsim_refit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rs = np.random.RandomState(5)
    n_refined = 0
    for case in range(60):
        N = int(rs.randint(3, 40))
        l = rs.standard_normal((N, 3))
        l /= np.linalg.norm(l, axis=1, keepdims=True)
        w = rs.uniform(0.05, 1.0, N)
        if case % 3 == 1:
            w *= 10.0 ** -rs.uniform(4, 7)            # one line dominates: sigma_2 / sigma_1 ~ 1e-4 .. 1e-7
            w[rs.randint(N)] = 1.0
        elif case % 3 == 2:
            vp_true = rs.standard_normal(3)
            vp_true /= np.linalg.norm(vp_true)        # lines through one point + one dominant line
            l = np.cross(vp_true, rs.standard_normal((N, 3)))
            l += 1e-9 * rs.standard_normal((N, 3))
            l /= np.linalg.norm(l, axis=1, keepdims=True)
            w *= 1e-5
            w[0] = 1.0
        A = (w / w.max())[:, None] * l
        ref = np.linalg.svd(A, full_matrices=False)[2][2]
        ref = ref * np.sign(ref[2])
        sv = np.linalg.svd(A, compute_uv=False)
        vp, refined = np.zeros(3), np.zeros(1, np.int32)
        ok = lib.hostsim_refit(_lib.ptr(np.ascontiguousarray(l)), _lib.ptr(np.ascontiguousarray(w)), N, _lib.ptr(vp), _lib.ptr(refined))
        assert ok == 1
        n_refined += int(refined[0])
        ang = np.arccos(min(abs(float(vp @ ref)), 1.0))
        # the SVD itself is only defined to eps * sigma_1 / (sigma_2 - sigma_3)
        bound = max(1e-6, 1e3 * 2.2e-16 * sv[0] / max(sv[1] - sv[2], 1e-300))
        assert ang < bound, (case, ang, bound, sv, int(refined[0]))
    assert n_refined >= 10

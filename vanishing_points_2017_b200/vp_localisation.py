"""Drop-in for the reference's vp_localisation.expectation_maximisation
(stage 3) on the B200.

Same positional order, keyword names, defaults and result-dict keys as
reference vp_localisation.py:168-172 / :441-442; a failed image returns the
reference's skeleton dict of None (:205-206).  The whole EM (pair similarity,
kNN line rating, E/M steps, split, merge, final refit) runs inside one
persistent CUDA kernel behind libvpk.so's C ABI (`vpk_em`).
"""
import collections
import ctypes as C

import numpy as np

from . import _lib

# the reference's probability_functions.PDF (probability_functions.py:5): type of result['distribution']
PDF = collections.namedtuple("PDF", "v lv vl l lvsq angles")


def em_distribution(image, n_vp, n_lines, ctx=None):
    """result['distribution'] of image `image` of the LAST EM call on `ctx` (vp_localisation.py:441-442): evaluated on
    the device from the planes of the last E-step (`vpk_em_distribution`); ask before the next EM call."""
    ctx = ctx or _lib.default_context()
    M, N = int(n_vp), int(n_lines)
    a = {"v": np.empty(M), "lv": np.empty((N, M)), "vl": np.empty((M, N)), "l": np.empty(N), "lvsq": np.empty((N, M)),
         "angles": np.empty((M, 2))}
    _lib.check(ctx.lib.vpk_em_distribution(ctx.h, int(image), M, N, _lib.ptr(a["v"]), _lib.ptr(a["lv"]), _lib.ptr(a["vl"]),
                                           _lib.ptr(a["l"]), _lib.ptr(a["lvsq"]), _lib.ptr(a["angles"])), "vpk_em_distribution")
    return PDF(**a)

_SKELETON = {"vp_assoc": None, "vp": None, "counts": None, "count_id": None, "decision_metric": None,
             "iterations": 0}


def _alloc_result(B, sumN, want_dm):
    M = _lib.VPK_MAX_VP
    arrs = {
        # every entry the caller reads is written by the library (rows beyond n_vp are never read)
        "status": np.empty(B, np.int32), "n_vp": np.empty(B, np.int32), "iterations": np.empty(B, np.int32),
        "vp": np.empty((B, M, 3), np.float64), "sigma": np.empty((B, M), np.float64),
        "counts": np.empty((B, M), np.int32), "counts_weighted": np.empty((B, M), np.float64),
        "vp_assoc": np.empty(max(sumN, 1), np.int32),
        "decision_metric": np.zeros(max(M * sumN, 1), np.float64) if want_dm else None,
    }
    res = _lib.EmResult()
    for k, a in arrs.items():
        setattr(res, k, None if a is None else a.ctypes.data)
    return arrs, res


def unpack_results(arrs, offsets):
    """Per-image reference-style dicts from the flat result arrays."""
    out = []
    M = _lib.VPK_MAX_VP
    for b in range(len(offsets) - 1):
        if arrs["status"][b] != _lib.EM_OK:
            d = dict(_SKELETON)
            d["status"] = int(arrs["status"][b])
            out.append(d)
            continue
        m = int(arrs["n_vp"][b])
        o0, o1 = int(offsets[b]), int(offsets[b + 1])
        dm = None
        if arrs.get("decision_metric") is not None:
            dm = arrs["decision_metric"][M * o0:M * o0 + m * (o1 - o0)].reshape(m, o1 - o0).copy()
        out.append({"vp_assoc": arrs["vp_assoc"][o0:o1].astype(np.int64), "vp": arrs["vp"][b, :m].copy(),
                    "counts": arrs["counts"][b, :m].astype(np.float64),
                    "counts_weighted": arrs["counts_weighted"][b, :m].copy(), "count_id": None,
                    "decision_metric": dm, "iterations": int(arrs["iterations"][b]),
                    "sigma": arrs["sigma"][b, :m].copy(), "status": 0})
    return out


def expectation_maximisation_batch(lines, segments, offsets, responses, sphere_images, init_vp=None,
                                   init_vp_offsets=None, want_decision_metric=False, ctx=None, **kwargs):
    """Ragged batch form: lines (sumN,3), segments (sumN,4), offsets (B+1),
    responses (B,20,20), sphere_images (B,S,S) uint8 -> list of result dicts."""
    ctx = ctx or _lib.default_context()
    dm = kwargs.pop("distance_measure", "angle")
    if dm != "angle":
        if dm in ("dotprod", "area"):
            raise NotImplementedError("only distance_measure='angle' (the value every reference caller passes: "
                                      "example.py:28, benchmark.py:51) is built")
        assert False        # vp_localisation.py:203
    cfg = _lib.em_config(**{k: (int(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in kwargs.items()})
    lines = _lib.as_f64(lines, 3)
    segments = _lib.as_f64(segments, 4)
    off = _lib.as_offsets(offsets)
    B = off.size - 1
    if off[-1] != lines.shape[0] or lines.shape[0] != segments.shape[0]:
        raise ValueError("offsets / lines / segments disagree")
    resp = np.ascontiguousarray(responses, dtype=np.float64).reshape(B, _lib.VPK_GRID, _lib.VPK_GRID)
    sph = None
    S = 1
    if sphere_images is not None:
        sph = np.ascontiguousarray(sphere_images, dtype=np.uint8)
        sph = sph.reshape(B, sph.shape[-2], sph.shape[-1])
        if sph.shape[1] != sph.shape[2]:
            raise ValueError("sphere images must be square")
        S = sph.shape[1]
    iv = ioff = None
    if init_vp is not None:
        iv = _lib.as_f64(init_vp, 3)
        if init_vp_offsets is None and B != 1:
            raise ValueError("init_vp for a batch of %d images needs init_vp_offsets (B + 1 entries)" % B)
        ioff = _lib.as_offsets(init_vp_offsets if init_vp_offsets is not None else [0, iv.shape[0]])
        if ioff.size != B + 1 or ioff[-1] != iv.shape[0]:
            raise ValueError("init_vp_offsets must have B + 1 = %d entries ending at %d, got %r" % (B + 1, iv.shape[0], ioff))
    arrs, res = _alloc_result(B, int(off[-1]), want_decision_metric)
    _lib.check(ctx.lib.vpk_em(ctx.h, _lib.ptr(lines), _lib.ptr(segments), _lib.ptr(off), B, _lib.ptr(resp),
                              _lib.ptr(sph), S, _lib.ptr(iv), _lib.ptr(ioff), C.byref(cfg), C.byref(res)),
               "vpk_em")
    return unpack_results(arrs, off)


def expectation_maximisation(l, lp, cnn_response, num_iter=100, sphere_image=None,
                             init_vp=None, do_merge=True, do_split=True, do_iterations=True,
                             distance_measure="angle", use_weights=True, wbias=1, num_init_vp=25, split_merge_freq=10,
                             merge_thresh=1e-3, outlier_thresh=1.96 ** 2, final_convergence=5e-3,
                             s_thresh=1e-200, num_min_lines=3, ctx=None):
    """vp_localisation.expectation_maximisation (reference vp_localisation.py:168-450).

    Like the reference, `l` is row-normalised in place (:186, :226) so callers
    that re-pickle it (evaluation.py:350) see the same array."""
    N = l.shape[0]
    res = expectation_maximisation_batch(
        l, lp, [0, N], np.asarray(cnn_response)[None], None if sphere_image is None else sphere_image[None],
        init_vp=init_vp, want_decision_metric=True, ctx=ctx, num_iter=num_iter, do_merge=do_merge,
        do_split=do_split, do_iterations=do_iterations, distance_measure=distance_measure, use_weights=use_weights,
        wbias=float(wbias), num_init_vp=num_init_vp, split_merge_freq=split_merge_freq, merge_thresh=merge_thresh,
        outlier_thresh=outlier_thresh, final_convergence=final_convergence, s_thresh=s_thresh,
        num_min_lines=num_min_lines)[0]
    if N:
        nr = np.sqrt(np.sum(l * l, axis=1))
        l /= nr[:, None]
        l /= np.sqrt(np.sum(l * l, axis=1))[:, None]
    if res["vp"] is None:
        return {k: res[k] for k in _SKELETON}
    res.pop("status", None)
    res["distribution"] = em_distribution(0, res["vp"].shape[0], N, ctx)       # the PDF tuple of the last E-step (:441-442)
    return res

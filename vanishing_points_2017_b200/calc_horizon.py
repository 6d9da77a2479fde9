"""Drop-in for the reference's calc_horizon.py.

`calculate_horizon_and_ortho_vp(em_result, maxbest=10, theta_vmin=pi/10, theta_z=pi/4)`
keeps the reference's name, argument order, defaults and return tuple
(calc_horizon.py:19, :225; callers example.py:65, benchmark.py:233); the work is
done by `horizon_kernel` through the C ABI (`vpk_horizon`).  There is no CPU
fallback: without libvpk.so / a GPU the call raises VpkError.
"""
import numpy as np

from . import _lib


def _pack(em_results):
    B = len(em_results)
    vp = np.zeros((B, _lib.VPK_MAX_VP, 3), np.float64)
    counts = np.zeros((B, _lib.VPK_MAX_VP), np.int32)
    n_vp = np.zeros(B, np.int32)
    for b, r in enumerate(em_results):
        v = r.get("vp") if r is not None else None
        if v is None:
            continue                      # no VPs: the reference's default horizon (calc_horizon.py:207-212)
        v = np.asarray(v, np.float64).reshape(-1, 3)
        m = v.shape[0]
        if m > _lib.VPK_MAX_VP:
            raise ValueError("at most %d vanishing points per image" % _lib.VPK_MAX_VP)
        vp[b, :m] = v
        counts[b, :m] = np.asarray(r["counts"]).reshape(-1)[:m]
        n_vp[b] = m
    return vp, counts, n_vp


def _unpack(points, combo, n_vp, maxbest):
    out = []
    for b in range(points.shape[0]):
        k = 3 if min(int(maxbest), int(n_vp[b])) > 2 else 2
        hP1, hP2, zVP, hVP1, hVP2 = (points[b, q].copy() for q in range(5))
        out.append((hP1, hP2, zVP, hVP1, hVP2, combo[b, :k].astype(int)))
    return out


def _truth(true_horizons, scales, image_heights, B):
    """(B,3) ground-truth horizons, benchmark.py's `scale` and `imageHeight` per image -> contiguous float64."""
    if true_horizons is None:
        return None, None, None, None
    th = np.ascontiguousarray(true_horizons, np.float64).reshape(B, 3)
    sc = np.ascontiguousarray(np.broadcast_to(np.asarray(scales, np.float64), (B,)))
    hh = np.ascontiguousarray(np.broadcast_to(np.asarray(image_heights, np.float64), (B,)))
    return th, sc, hh, np.empty(B, np.float64)


def calculate_horizon_and_ortho_vp_batch(em_results, maxbest=10, theta_vmin=np.pi / 10., theta_z=np.pi / 4., ctx=None,
                                         true_horizons=None, scales=None, image_heights=None):
    """One (hP1, hP2, zVP, hVP1, hVP2, best_combo) tuple per EM result dict.  With ground-truth horizons
    (homogeneous lines), `scales` and `image_heights` as in benchmark.py:247-253, returns (tuples, errors)."""
    ctx = ctx or _lib.default_context()
    vp, counts, n_vp = _pack(em_results)
    B = len(em_results)
    points = np.empty((B, 5, 3), np.float64)
    combo = np.empty((B, 3), np.int32)
    th, sc, hh, err = _truth(true_horizons, scales, image_heights, B)
    _lib.check(ctx.lib.vpk_horizon(ctx.h, _lib.ptr(vp), _lib.ptr(counts), _lib.ptr(n_vp), B, int(maxbest), float(theta_vmin),
                                   float(theta_z), _lib.ptr(th), _lib.ptr(sc), _lib.ptr(hh), _lib.ptr(points), _lib.ptr(combo),
                                   _lib.ptr(err)), "vpk_horizon")
    out = _unpack(points, combo, n_vp, maxbest)
    return out if err is None else (out, err)


def calculate_horizon_and_ortho_vp(em_result, maxbest=10, theta_vmin=np.pi / 10., theta_z=np.pi / 4.):
    """calc_horizon.py:19-225.  Like the reference, an EM result without VPs
    (`em_result['vp'] is None`) is an error (the reference fails on `None.copy()`, :22)."""
    if em_result["vp"] is None:
        raise AttributeError("'NoneType' object has no attribute 'copy'")
    return calculate_horizon_and_ortho_vp_batch([em_result], maxbest, theta_vmin, theta_z)[0]

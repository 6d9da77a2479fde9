"""Drop-in for the reference's sphere_mapping module (stage S1) on the B200.

`sphere_line_plot(lines, size, alpha, f, alternative)` keeps the signature,
return type (uint8[size,size]) and the in-place `lines[:,0:2] *= f` side effect
of reference sphere_mapping.py:36-72; `get_sphere_image` mirrors
evaluation.py:12-14.  The batched forms take a ragged (flat, offsets) batch.
All work runs in hand-written CUDA kernels behind libvpk.so's C ABI.
"""
import numpy as np

from . import _lib


def lines_from_segments(segments, ctx=None):
    """evaluation.py:158-168 on the device: (N,4) segments -> (N,3) lines."""
    ctx = ctx or _lib.default_context()
    seg = _lib.as_f64(segments, 4)
    out = np.empty((seg.shape[0], 3), dtype=np.float64)
    _lib.check(ctx.lib.vpk_lines_from_segments(ctx.h, _lib.ptr(seg), seg.shape[0], _lib.ptr(out)),
               "vpk_lines_from_segments")
    return out


def sphere_map_batch(lines, offsets, size, mode="votes", alpha=0.1, weights=None, want_hist=True,
                     want_image=True, ctx=None):
    """Ragged batch -> dict(hist (B,S,S) uint32 | whist float32, image (B,S,S) uint8)."""
    ctx = ctx or _lib.default_context()
    lines = _lib.as_f64(lines, 3)
    off = _lib.as_offsets(offsets)
    if off[-1] != lines.shape[0]:
        raise ValueError("offsets[-1] must equal the number of lines")
    B = off.size - 1
    m = {"votes": _lib.SPHERE_VOTES, "curves": _lib.SPHERE_CURVES}[mode]
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    hist = np.empty((B, size, size), dtype=np.uint32) if (want_hist and w is None) else None
    whist = np.empty((B, size, size), dtype=np.float32) if (want_hist and w is not None) else None
    img = np.empty((B, size, size), dtype=np.uint8) if want_image else None
    _lib.check(ctx.lib.vpk_sphere_map(ctx.h, _lib.ptr(lines), _lib.ptr(off), B, int(size), m, float(alpha),
                                      _lib.ptr(w), _lib.ptr(hist), _lib.ptr(whist), _lib.ptr(img)),
               "vpk_sphere_map")
    return {"hist": hist, "whist": whist, "image": img}


def sphere_votes(lines, size, weights=None, ctx=None):
    """north_star formulation for one image: (S,S) histogram and uint8 image."""
    lines = _lib.as_f64(lines, 3)
    r = sphere_map_batch(lines, [0, lines.shape[0]], size, "votes", weights=weights, ctx=ctx)
    h = r["hist"] if weights is None else r["whist"]
    return h[0], r["image"][0]


def sphere_line_plot(lines, size, alpha=0.1, f=1.0, alternative=False, mode="curves", ctx=None):
    """sphere_mapping.sphere_line_plot (reference sphere_mapping.py:36-72).

    mode="curves" (default) rasterises the reference's great-circle geometry;
    mode="votes" returns the pairwise-intersection image of the north_star.
    """
    if alternative:
        raise NotImplementedError("alternative=True is never used by the reference's callers "
                                  "(sphere_mapping.py:58-59)")
    lines[:, 0] *= f            # the reference mutates its argument (:55-56)
    lines[:, 1] *= f
    r = sphere_map_batch(lines, [0, lines.shape[0]], size, mode, alpha=alpha, want_hist=False, ctx=ctx)
    return r["image"][0]


def get_sphere_image(lines, size=250, alpha=0.1, f=1.0, mode="curves", ctx=None):
    """evaluation.get_sphere_image (reference evaluation.py:12-14)."""
    return sphere_line_plot(lines, size, alpha=alpha, f=f, alternative=False, mode=mode, ctx=ctx)

"""Host mirror of the parts of the reference's evaluation.py that sit between the LSD detector and the
hot path: the normalisation of the raw LSD rows
(`detect_lsd_lines`, evaluation.py:227-251, everything after the `lsd.detect_line_segments` call) and the
line construction (evaluation.py:158-168), on the device through the C ABI (`vpk_segments_from_lsd`).
The LSD detector itself stays upstream (row N4).  No CPU fallback."""
import numpy as np

from . import _lib


def segments_from_lsd_batch(lsd_rows, image_shapes, want_lines=True, ctx=None):
    """lsd_rows: list of (N_b, >=4) arrays as lsd.detect_line_segments returns them (pixels);
    image_shapes: list of (height, width[, ...]) like `image.shape`.
    Returns dict(segments (sum N,4), lines (sum N,3) | None, nfa (sum N,) | None, offsets (B+1,) int32)."""
    ctx = ctx or _lib.default_context()
    B = len(lsd_rows)
    rows = [np.asarray(r, np.float64).reshape(-1, np.asarray(r).shape[-1] if np.asarray(r).ndim == 2 else 7) for r in lsd_rows]
    ncols = rows[0].shape[1] if B else 7
    if any(r.shape[1] != ncols for r in rows) or ncols < 4:
        raise ValueError("all LSD arrays must have the same number of columns (>= 4)")
    off = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows])]).astype(np.int32)
    flat = np.ascontiguousarray(np.concatenate(rows) if B else np.zeros((0, ncols)))
    widths = np.array([s[1] for s in image_shapes], np.int32)
    heights = np.array([s[0] for s in image_shapes], np.int32)
    n = int(off[-1])
    seg = np.empty((n, 4), np.float64)
    lines = np.empty((n, 3), np.float64) if want_lines else None
    nfa = np.empty(n, np.float64) if ncols >= 7 else None
    _lib.check(ctx.lib.vpk_segments_from_lsd(ctx.h, _lib.ptr(flat), ncols, _lib.ptr(off), _lib.ptr(widths), _lib.ptr(heights), B,
                                             _lib.ptr(seg), _lib.ptr(lines), _lib.ptr(nfa)), "vpk_segments_from_lsd")
    return {"segments": seg, "lines": lines, "nfa": nfa, "offsets": off}


def segments_from_lsd(lsd_lines, image_shape):
    """What detect_lsd_lines returns for one image (evaluation.py:251): {'segments': (N,4), 'nfa': (N,)}."""
    out = segments_from_lsd_batch([lsd_lines], [image_shape], want_lines=False)
    return {"segments": out["segments"], "nfa": out["nfa"]}

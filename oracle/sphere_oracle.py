"""ORACLE (test infrastructure, not product code) -- stage 1, sphere mapping.

CPU restatement in numpy/float64 of the reference's sphere mapping for the
lines->VPs hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module.

PARITY STATUS: *unpinned* against the original binaries.  The reference
rasterises with matplotlib 1.5.1 / Agg (reference sphere_mapping.py:36-72),
which is not installable here, and the reference holds no golden images.  The
geometry (coordinate_conversion.py, the great-circle formula) is restated
literally; the rasterisation rule is ours and is documented below.

Two formulations:
  * votes  -- BASELINE.json north_star: all N(N-1)/2 pairwise intersections
              p = l_i x l_j, inverse gnomonic projection, weighted votes into
              an SxS histogram.
  * curves -- reference geometry: each line's great circle beta(alpha)
              (sphere_mapping.py:61-63) sampled at the reference's 10 000
              alphas (sphere_mapping.py:40) and accumulated as per-pixel line
              coverage counts; alpha "over" compositing becomes
              255*(1-(1-alpha)^k) for k covering lines.
"""
import numpy as np

PI = np.pi
NUM_SAMPLES = 10000          # sphere_mapping.py:40


# --------------------------------------------------------------------------
# T0  coordinate_conversion.py:4-61 (scalar functions, restated vectorised)
# --------------------------------------------------------------------------
def index_to_angle(index, shape):
    """coordinate_conversion.py:4-20."""
    a, b = index[0], index[1]
    M, N = shape[0], shape[1]
    return np.array([(a - 0.5 * M + 0.5) * PI / M, (b - 0.5 * N + 0.5) * PI / N])


def angle_to_index(angle, shape):
    """coordinate_conversion.py:23-35 (continuous index; integers = cell centres)."""
    M, N = shape[0], shape[1]
    return np.array([(angle[0] / PI + 0.5 - 0.5 / M) * M,
                     (angle[1] / PI + 0.5 - 0.5 / N) * N])


def angle_to_point(angle):
    """coordinate_conversion.py:38-50 (note sign(0) == 0 quirk is kept)."""
    alpha, beta = angle[0], angle[1]
    p = np.array([np.sin(alpha) * np.cos(beta), np.sin(beta), np.cos(alpha) * np.cos(beta)])
    return p * np.sign(p[2])


def point_to_angle(point):
    """coordinate_conversion.py:53-61."""
    beta = np.arcsin(point[1])
    inner = np.maximum(np.minimum(point[0] / np.cos(beta), 1), -1)
    return np.array([np.arcsin(inner), beta])


def _bin(angle, S):
    """angle_to_index (coordinate_conversion.py:29-30) then round-to-cell:
    floor(a + 0.5), clipped to [0, S-1]."""
    a = (angle / PI + 0.5 - 0.5 / S) * S
    return np.clip(np.floor(a + 0.5), 0, S - 1).astype(np.int64)


# --------------------------------------------------------------------------
# S1 (votes)
# --------------------------------------------------------------------------
def pair_bins(li, lj, S):
    """Rows/cols of the intersections of line arrays li, lj (K,3).

    p = li x lj (each product and difference individually rounded, no FMA),
    unit-normalise, flip to z >= 0, point_to_angle, angle_to_index, round.
    Returns (row, col, valid); row 0 = beta max (image orientation of
    sphere_mapping.py:22-33), col 0 = alpha min.  Deviation from
    point_to_angle: p_y/|p| is clipped to [-1,1] before arcsin.
    """
    px = li[:, 1] * lj[:, 2] - li[:, 2] * lj[:, 1]
    py = li[:, 2] * lj[:, 0] - li[:, 0] * lj[:, 2]
    pz = li[:, 0] * lj[:, 1] - li[:, 1] * lj[:, 0]
    n = np.sqrt((px * px + py * py) + pz * pz)
    valid = np.isfinite(n) & (n > 0)
    n = np.where(valid, n, 1.0)
    flip = pz < 0
    px = np.where(flip, -px, px)
    py = np.where(flip, -py, py)
    y = np.maximum(np.minimum(py / n, 1.0), -1.0)
    beta = np.arcsin(y)
    with np.errstate(divide="ignore", invalid="ignore"):
        inner = (px / n) / np.cos(beta)
    valid &= ~np.isnan(inner)
    inner = np.maximum(np.minimum(np.where(np.isnan(inner), 0.0, inner), 1.0), -1.0)
    alpha = np.arcsin(inner)
    col = _bin(alpha, S)
    row = (S - 1) - _bin(beta, S)
    return row, col, valid


WEIGHT_FRAC_BITS = 16   # fixed-point weight resolution of the weighted mode


def quantise_pair_weight(wi, wj):
    """Weighted votes are accumulated in Q.16 fixed point so that the sum is
    order independent: q = floor(wi*wj*2^16 + 0.5)."""
    return np.floor(wi * wj * float(1 << WEIGHT_FRAC_BITS) + 0.5).astype(np.int64)


def sphere_votes(lines, S, weights=None, chunk=2_000_000):
    """All-pairs intersection vote histogram.

    Returns (S,S) int64: counts if weights is None, else Q.16 fixed-point sums
    (divide by 2^16 for the real-valued histogram).
    """
    lines = np.asarray(lines, dtype=np.float64)
    N = lines.shape[0]
    hist = np.zeros(S * S, dtype=np.int64)
    if N < 2:
        return hist.reshape(S, S)
    # enumerate i<j row by row to bound memory
    i0 = 0
    while i0 < N - 1:
        # rows i0..i1 such that pairs <= chunk
        i1 = i0
        cnt = 0
        while i1 < N - 1 and cnt + (N - 1 - i1) <= max(chunk, N):
            cnt += N - 1 - i1
            i1 += 1
        ii = np.repeat(np.arange(i0, i1), N - 1 - np.arange(i0, i1))
        jj = np.concatenate([np.arange(i + 1, N) for i in range(i0, i1)])
        row, col, valid = pair_bins(lines[ii], lines[jj], S)
        flat = (row * S + col)[valid]
        if weights is None:
            hist += np.bincount(flat, minlength=S * S)
        else:
            q = quantise_pair_weight(weights[ii], weights[jj])[valid]
            np.add.at(hist, flat, q)
        i0 = i1
    return hist.reshape(S, S)


def votes_to_image(hist):
    """Histogram -> uint8 sphere image: floor(255*h/max(h)) in exact integer
    arithmetic (all-zero histogram -> zeros)."""
    h = np.asarray(hist, dtype=np.int64)
    m = int(h.max()) if h.size else 0
    if m <= 0:
        return np.zeros(h.shape, dtype=np.uint8)
    return ((h * 255) // m).astype(np.uint8)


# --------------------------------------------------------------------------
# S1 (curves) -- reference geometry
# --------------------------------------------------------------------------
def curve_rows(lines, S, num=NUM_SAMPLES):
    """Row index of every reference sample.  Returns (rows (N,num) int64,
    cols (num,) int64).

    a = linspace(-pi/2, pi/2, num)                      sphere_mapping.py:40
    b = arctan((-l0 sin a - l2 cos a) / l1)             sphere_mapping.py:61-63
    (the leading minus of :61 and the `b *= -1` of :63 cancel).
    """
    lines = np.asarray(lines, dtype=np.float64)
    a = np.linspace(-PI / 2, PI / 2, num=num)
    sa, ca = np.sin(a), np.cos(a)
    with np.errstate(divide="ignore", invalid="ignore"):
        b = np.arctan((-lines[:, 0:1] * sa[None, :] - lines[:, 2:3] * ca[None, :]) / lines[:, 1:2])
    cols = _bin(a, S)
    rows = (S - 1) - _bin(np.where(np.isnan(b), 0.0, b), S)
    rows = np.where(np.isnan(b), -1, rows)
    return rows, cols


def sphere_curve_counts(lines, S, num=NUM_SAMPLES):
    """Per-pixel number of lines whose sampled great circle covers the pixel.

    Rasterisation rule (ours; Agg's anti-aliased 1.39 px stroke is unpinned):
    for each line and each column c, the covered rows are the closed interval
    [min, max] of the sample rows over samples k with col(k) == c, extended by
    the first sample of the next column (so the polyline is connected across
    column boundaries).  A pixel is counted at most once per line, like a
    single Agg stroke.
    """
    lines = np.asarray(lines, dtype=np.float64)
    N = lines.shape[0]
    counts = np.zeros((S, S), dtype=np.int64)
    if N == 0:
        return counts
    rows, cols = curve_rows(lines, S, num)
    first = np.searchsorted(cols, np.arange(S), side="left")
    last = np.searchsorted(cols, np.arange(S), side="right")   # exclusive
    for c in range(S):
        k0, k1 = first[c], min(last[c] + 1, num)               # include next column's first sample
        if k1 <= k0:
            continue
        r = rows[:, k0:k1]
        ok = r >= 0
        rmin = np.where(ok, r, S).min(axis=1)
        rmax = np.where(ok, r, -1).max(axis=1)
        good = rmax >= rmin
        # difference array along rows for this column
        d = np.zeros(S + 1, dtype=np.int64)
        np.add.at(d, rmin[good], 1)
        np.add.at(d, rmax[good] + 1, -1)
        counts[:, c] += np.cumsum(d)[:S]
    return counts


def coverage_lut(alpha, kmax=4096):
    """uint8 value of a pixel covered by k strokes of opacity alpha on black:
    floor(255*(1-(1-alpha)^k))."""
    k = np.arange(kmax + 1, dtype=np.float64)
    return np.floor(255.0 * (1.0 - np.power(1.0 - alpha, k))).astype(np.uint8)


def curve_image(counts, alpha=0.1):
    lut = coverage_lut(alpha)
    return lut[np.minimum(counts, len(lut) - 1)]


def sphere_line_plot(lines, size, alpha=0.1, f=1.0, alternative=False):
    """Restatement of sphere_mapping.sphere_line_plot (sphere_mapping.py:36-72).
    Mutates lines[:,0:2] *= f in place like the reference (:55-56)."""
    if alternative:
        raise NotImplementedError("alternative=True is never used by the reference's callers")
    lines[:, 0] *= f
    lines[:, 1] *= f
    return curve_image(sphere_curve_counts(lines, size), alpha)

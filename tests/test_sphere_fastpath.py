"""S1: the float32 pre-binning of the vote kernel (csrc/sphere.cu::fast_cell) restated in numpy float32, with the
device's rsqrtf / asinf rounding emulated by random +-3 ulp perturbations: whenever the fast path ACCEPTS a pair,
its cell must be the cell of the float64 expressions (oracle.sphere_oracle.pair_bins, pinned against the
reference's coordinate_conversion.py); pairs it rejects are evaluated in float64 by the kernel.  The constants
mirror the kernel's (2^-20 mag/|p| + 2^-21, 4e-7, 1.5e-4, factor 2, 0.05 guards).  tools/check_fast_cell.py runs
the same check over 83 M pairs including adversarial line families."""
import numpy as np

from oracle import sphere_oracle as so
from vanishing_points_2017_b200 import synth

f32 = np.float32
rs = np.random.RandomState(0)

def ulp_noise(x, k=3):
    return (x * (f32(1) + f32(k) * f32(2.0 ** -24) * rs.uniform(-1, 1, x.shape).astype(f32))).astype(f32)

def fast(li, lj, S):
    a = li.astype(f32); b = lj.astype(f32)
    ax, ay, az = a[:, 0], a[:, 1], a[:, 2]; bx, by, bz = b[:, 0], b[:, 1], b[:, 2]
    t0, t1, t2, t3, t4, t5 = ay * bz, az * by, az * bx, ax * bz, ax * by, ay * bx
    px, py, pz = t0 - t1, t2 - t3, t4 - t5
    mag = (np.abs(t0) + np.abs(t1)) + (np.abs(t2) + np.abs(t3)) + (np.abs(t4) + np.abs(t5))
    n2 = px * px + py * py + pz * pz
    ok = (n2 > f32(1e-30)) & (n2 < f32(1e30))
    n2s = np.where(ok, n2, f32(1))
    rn = ulp_noise((f32(1) / np.sqrt(n2s)).astype(f32))
    flip = pz < 0
    px = np.where(flip, -px, px); py = np.where(flip, -py, py)
    y = py * rn; x = px * rn
    c2 = f32(1) - y * y
    ok &= c2 > f32(0.0025)
    c2s = np.where(ok, c2, f32(1))
    rc = ulp_noise((f32(1) / np.sqrt(c2s)).astype(f32))
    inner = x * rc
    q2 = f32(1) - inner * inner
    ok &= q2 > f32(0.0025)
    q2s = np.where(ok, q2, f32(1))
    eps = f32(2.0 ** -20) * (mag * rn) + f32(2.0 ** -21)
    dbeta = eps * rc + f32(4e-7)
    dalpha = eps * (rc + rc * rc) * ulp_noise((f32(1) / np.sqrt(q2s)).astype(f32)) + f32(4e-7)
    sop = f32(S / np.pi); hs = f32(0.5 * S)
    fa = ulp_noise(np.arcsin(np.clip(inner, -1, 1)).astype(f32)) * sop + hs
    fb = ulp_noise(np.arcsin(np.clip(y, -1, 1)).astype(f32)) * sop + hs
    ma = f32(2) * (dalpha * sop + f32(1.5e-4)); mb = f32(2) * (dbeta * sop + f32(1.5e-4))
    ra, rb = np.floor(fa), np.floor(fb)
    ok &= ~((fa - ra < ma) | (ra + 1 - fa < ma) | (fb - rb < mb) | (rb + 1 - fb < mb))
    col = np.clip(ra.astype(np.int64), 0, S - 1)
    row = (S - 1) - np.clip(rb.astype(np.int64), 0, S - 1)
    return ok, row, col



def test_fast_path_never_disagrees_with_float64_cells():
    total = accepted = 0
    for seed, n, S in ((601, 700, 500), (602, 500, 250), (603, 400, 64)):
        sc = synth.make_scene(seed, n, 800, 600, noise_deg=0.5 + seed % 3, outlier_frac=0.15)
        lines = sc["lines"]
        ii, jj = np.triu_indices(n, 1)
        row, col, valid = so.pair_bins(lines[ii], lines[jj], S)
        ok, frow, fcol = fast(lines[ii], lines[jj], S)
        assert not np.any(ok & (~valid | (frow != row) | (fcol != col)))
        total += len(ii); accepted += int(ok.sum())
    # adversarial: wild scales, a nearly parallel family, intersections near the poles
    n = 600
    L = rs.standard_normal((n, 3)) * np.exp(rs.uniform(-6, 6, (n, 1)))
    L[: n // 3] = L[0] + 1e-4 * rs.standard_normal((n // 3, 3)) * np.abs(L[0])
    L[n // 3: n // 2, 1] *= 1e-6
    ii, jj = np.triu_indices(n, 1)
    row, col, valid = so.pair_bins(L[ii], L[jj], 500)
    ok, frow, fcol = fast(L[ii], L[jj], 500)
    assert not np.any(ok & (~valid | (frow != row) | (fcol != col)))
    assert accepted > 0.85 * total          # the fast path must carry the bulk of the pairs


# ---- curves mode: csrc/sphere.cu::fast_row ---------------------------------------------------------------------
U = f32(2.0 ** -24)


def exact_rows(L, sa, ca, S):
    """sphere_mapping.py:61-63 + coordinate_conversion.py:29-30 in float64 (what trig_row computes)."""
    with np.errstate(all="ignore"):
        beta = np.arctan((-L[:, 0:1] * sa[None] - L[:, 2:3] * ca[None]) / L[:, 1:2])
        r = np.clip(np.floor(((beta / np.pi + 0.5) - 0.5 / S) * S + 0.5), 0, S - 1)
    return np.where(np.isnan(beta), -1, (S - 1) - r).astype(np.int64)


def fast_rows(L, sa, ca, S):
    with np.errstate(all="ignore"):
        l0 = L[:, 0:1].astype(f32); l1 = L[:, 1:2].astype(f32); l2 = L[:, 2:3].astype(f32)
        inv = ulp_noise((f32(1) / l1).astype(f32), 1)
        p = (-l0) * sa.astype(f32)[None]
        q = l2 * ca.astype(f32)[None]
        g = (p - q) * inv
        eg = f32(10) * U * ((np.abs(p) + np.abs(q)) * np.abs(inv))
        gm = np.maximum(np.abs(g) - eg, f32(0))
        dbeta = eg / (f32(1) + gm * gm) + f32(2.0 ** -22)
        beta = ulp_noise(np.arctan(g).astype(f32), 2)
        sop = f32(S / np.pi); hs = f32(0.5 * S)
        u = (beta.astype(np.float64) * np.float64(sop) + np.float64(hs)).astype(f32)       # fmaf: one rounding
        m = f32(2) * (dbeta * sop + f32(2.0 ** -23) * f32(S)) + f32(1e-6)
        fl = np.floor(u)
        ok = (u - fl > m) & (f32(1) - (u - fl) > m) & (np.abs(u) < f32(1e9))
        r = np.clip(np.where(ok, fl, 0).astype(np.int64), 0, S - 1)
    return ok, (S - 1) - r


def test_curves_fast_path_never_disagrees_with_float64_rows():
    """tools/check_fast_row.py runs the same check over 137 M (line, sample) pairs."""
    k = np.unique(np.concatenate([np.arange(0, 10000, 23), [9999]]))
    alpha = k * (np.pi / 9999) + (-0.5 * np.pi)
    alpha[-1] = 0.5 * np.pi
    sa, ca = np.sin(alpha), np.cos(alpha)
    total = accepted = 0
    for seed, n, S in ((701, 800, 500), (702, 500, 250), (703, 300, 1536)):
        L = synth.make_scene(seed, n, 800, 600, noise_deg=0.5)["lines"]
        ok, fr = fast_rows(L, sa, ca, S)
        assert not np.any(ok & (fr != exact_rows(L, sa, ca, S)))
        total += ok.size; accepted += int(ok.sum())
    assert accepted > 0.99 * total
    n = 1200
    L = rs.standard_normal((n, 3)) * np.exp(rs.uniform(-8, 8, (n, 3)))
    L[: n // 4, 1] *= 1e-7                       # |g| huge
    L[n // 4: n // 2, 1] *= 1e7                  # g ~ 0
    L[n // 2: 5 * n // 8, 1] = 0.0               # division by zero
    L[5 * n // 8: 3 * n // 4, 0] = 0.0
    ok, fr = fast_rows(L, sa, ca, 500)
    assert not np.any(ok & (fr != exact_rows(L, sa, ca, 500)))

"""CPU emulation of sphere.cu::fast_cell in numpy float32 (device rsqrtf / asinf errors emulated by random
+-3 ulp perturbations): whenever the fast path accepts a pair, its cell must equal the float64 cell."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
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

tot = acc = bad = 0
def run(lines, S, tag):
    global tot, acc, bad
    n = lines.shape[0]
    ii, jj = np.triu_indices(n, 1)
    for s0 in range(0, len(ii), 2_000_000):
        li, lj = lines[ii[s0:s0 + 2_000_000]], lines[jj[s0:s0 + 2_000_000]]
        row, col, valid = so.pair_bins(li, lj, S)
        ok, frow, fcol = fast(li, lj, S)
        wrong = ok & (~valid | (frow != row) | (fcol != col))
        tot += len(li); acc += int(ok.sum()); bad += int(wrong.sum())
        if wrong.any():
            k = np.where(wrong)[0][0]
            print("MISMATCH", tag, li[k], lj[k], row[k], col[k], frow[k], fcol[k], valid[k])
for seed in range(12):
    sc = synth.make_scene(600 + seed, 1200 + 100 * seed, 800, 600, noise_deg=0.3 * (1 + seed % 4), outlier_frac=0.1 + 0.05 * (seed % 3))
    for S in (500, 250, 64, 1000):
        run(sc["lines"], S, "scene%d S%d" % (seed, S))
    print(seed, "pairs", tot, "accepted %.4f" % (acc / tot), "bad", bad, flush=True)
# adversarial: random lines of wild scales, nearly parallel families, lines through the poles
for k in range(6):
    n = 1500
    L = rs.standard_normal((n, 3)) * np.exp(rs.uniform(-6, 6, (n, 1)))
    L[: n // 3] = L[0] + 1e-4 * rs.standard_normal((n // 3, 3)) * np.abs(L[0])     # nearly parallel family
    L[n // 3: n // 2, 1] *= 1e-6                                                     # beta ~ +-pi/2 intersections
    run(L, 500, "adversarial%d" % k)
    print("adv", k, "pairs", tot, "accepted %.4f" % (acc / tot), "bad", bad, flush=True)

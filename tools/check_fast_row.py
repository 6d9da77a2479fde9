"""CPU emulation of sphere.cu::fast_row (curves mode) in numpy float32, the device's atanf / division
errors emulated by random +-2 ulp perturbations: whenever the fast path accepts a (line, sample), its
row must equal the row of the float64 expression (sphere_mapping.py:61-63 + coordinate_conversion.py:29-30)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vanishing_points_2017_b200 import synth
f32 = np.float32
rs = np.random.RandomState(0)
U = f32(2.0 ** -24)


def ulp_noise(x, k=2):
    return (x * (f32(1) + f32(k) * U * rs.uniform(-1, 1, x.shape).astype(f32))).astype(f32)


def exact_rows(L, sa, ca, S):
    with np.errstate(all="ignore"):
        g = (-L[:, 0:1] * sa[None] - L[:, 2:3] * ca[None]) / L[:, 1:2]
        beta = np.arctan(g)
        a = ((beta / np.pi + 0.5) - 0.5 / S) * S
        r = np.clip(np.floor(a + 0.5), 0, S - 1)
    row = (S - 1) - r
    return np.where(np.isnan(beta), -1, row).astype(np.int64)


def fast_rows(L, sa, ca, S):
    with np.errstate(all="ignore"):
        l0 = L[:, 0:1].astype(f32); l2 = L[:, 2:3].astype(f32); l1 = L[:, 1:2].astype(f32)
        inv = ulp_noise((f32(1) / l1).astype(f32), 1)
        inv_abs = np.abs(inv)
        saf, caf = sa.astype(f32)[None], ca.astype(f32)[None]
        p = (-l0) * saf
        q = l2 * caf
        n = p - q
        g = n * inv
        A = np.abs(p) + np.abs(q)
        eg = f32(10) * U * (A * inv_abs)
        gm = np.maximum(np.abs(g) - eg, f32(0))
        dbeta = eg / (f32(1) + gm * gm) + f32(2.0 ** -22)
        beta = ulp_noise(np.arctan(g).astype(f32))
        sop = f32(S / np.pi); hs = f32(0.5 * S)
        u = (beta.astype(np.float64) * np.float64(sop) + np.float64(hs)).astype(f32)       # fmaf: one rounding
        du = dbeta * sop + f32(2.0 ** -23) * f32(S)
        m = f32(2) * du + f32(1e-6)
        fl = np.floor(u)
        fr = u - fl
        ok = (fr > m) & (f32(1) - fr > m) & (np.abs(u) < f32(1e9))
        r = np.clip(np.where(ok, fl, 0).astype(np.int64), 0, S - 1)
    return ok, (S - 1) - r


tot = acc = bad = 0


def run(L, S, tag):
    global tot, acc, bad
    k = np.unique(np.concatenate([np.arange(0, 10000, 7), [9999]]))
    step = np.pi / 9999
    alpha = k * step + (-0.5 * np.pi)
    alpha[-1] = 0.5 * np.pi
    sa, ca = np.sin(alpha), np.cos(alpha)
    ex = exact_rows(L, sa, ca, S)
    ok, fr = fast_rows(L, sa, ca, S)
    wrong = ok & (fr != ex)
    tot += ok.size; acc += int(ok.sum()); bad += int(wrong.sum())
    if wrong.any():
        i, j = np.argwhere(wrong)[0]
        print("MISMATCH", tag, L[i], k[j], ex[i, j], fr[i, j])


for seed in range(8):
    sc = synth.make_scene(700 + seed, 1500, 800, 600, noise_deg=0.5)
    for S in (500, 250, 100, 1000, 1536):
        run(sc["lines"], S, "scene%d S%d" % (seed, S))
    print(seed, "samples", tot, "accepted %.4f" % (acc / tot), "bad", bad, flush=True)
for kk in range(6):
    n = 3000
    L = rs.standard_normal((n, 3)) * np.exp(rs.uniform(-8, 8, (n, 3)))
    L[: n // 4, 1] *= 1e-7                         # nearly vertical great circles: |g| huge
    L[n // 4: n // 2, 1] *= 1e7                    # g ~ 0 everywhere
    L[n // 2: 5 * n // 8, 1] = 0.0                 # division by zero
    L[5 * n // 8: 3 * n // 4, 0] = 0.0
    for S in (500, 64):
        run(L, S, "adversarial%d S%d" % (kk, S))
    print("adv", kk, "samples", tot, "accepted %.4f" % (acc / tot), "bad", bad, flush=True)
sys.exit(1 if bad else 0)

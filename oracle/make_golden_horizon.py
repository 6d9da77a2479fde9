"""Golden vectors for row N1 (horizon / orthogonal triplet): runs the reference's own, UNMODIFIED
calc_horizon.calculate_horizon_and_ortho_vp (imported from /root/reference; it is plain numpy and runs
under Python 3 as it is) on seeded EM-result-shaped inputs and stores inputs + outputs in
tests/golden/horizon_cases.npz.   python oracle/make_golden_horizon.py
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "horizon_cases.npz")


def manhattan(rs, extra):
    """Three orthogonal VPs of a random camera (small pitch / roll) + `extra` random ones, unit, z >= 0."""
    yaw, pitch, roll = rs.uniform(-np.pi, np.pi), rs.normal(0, 0.15), rs.normal(0, 0.08)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    f = rs.uniform(0.8, 2.5)
    K = np.diag([f, f, 1.0])
    v = (K @ Rz @ Rx @ Ry).T                      # rows: images of the three axes
    v = np.vstack([v, rs.standard_normal((extra, 3))])
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v *= np.where(v[:, 2:3] < 0, -1.0, 1.0)
    return v


def main():
    sys.path.insert(0, REF)
    import calc_horizon as ref                     # the reference's module, unmodified
    rs = np.random.RandomState(20171)
    cases = []
    for i in range(160):
        kind = i % 8
        if kind == 0:
            m = int(rs.randint(0, 3))              # 0, 1 or 2 VPs: the fallback branches
            vp = manhattan(rs, 0)[:m] if m else np.zeros((0, 3))
        else:
            vp = manhattan(rs, int(rs.randint(0, 12)))
            vp = vp[rs.permutation(vp.shape[0])]
        m = vp.shape[0]
        # distinct counts: numpy.argsort's order of EQUAL counts (calc_horizon.py:34) depends on the numpy
        # build (the reference's numpy 1.11.3 insertion-sorts up to 16 elements, i.e. stable; numpy 2.x may
        # use a vectorised unstable sort), so ties are not pinned by running the reference here
        counts = rs.permutation(np.arange(3, 400))[:m]
        maxbest = [10, 20, 3, 5, 10, 2, 10, 1][kind]
        tv = [np.pi / 10., np.pi / 10., np.pi / 8., np.pi / 10.][i % 4]
        tz = [np.pi / 4., np.pi / 3.][i % 2]
        try:
            with np.errstate(all="ignore"):
                out = ref.calculate_horizon_and_ortho_vp({"vp": vp, "counts": counts}, maxbest=maxbest, theta_vmin=tv, theta_z=tz)
            from_ref = True
        except TypeError:
            # fewer than two VPs take part: hlin = np.cross([0,0,1], [1,0,1]) is an INTEGER array
            # (calc_horizon.py:206, :212) and `hP1 /= hP1[2]` (:221) is an in-place true division of
            # integers, which numpy >= 1.10 under Python 3 refuses; under the reference's Python 2 it
            # is an integer division with the exact results [-1, 0, 1] and [1, 0, 1].  Those values
            # are stored, flagged as not produced by running the reference.
            assert min(maxbest, m) < 2
            h1 = vp[0] if min(maxbest, m) > 0 else np.array([-1.0, 0.0, 0.0])
            h2 = vp[0] if min(maxbest, m) > 0 else np.array([1.0, 0.0, 0.0])
            out = (np.array([-1.0, 0.0, 1.0]), np.array([1.0, 0.0, 1.0]), np.array([0.0, 1.0, 0.0]), h1, h2, np.array([0, 0]))
            from_ref = False
        cases.append((vp, counts, maxbest, tv, tz, out, from_ref))
    M = 16
    n = len(cases)
    g = {"vp": np.zeros((n, M, 3)), "counts": np.zeros((n, M), np.int64), "n_vp": np.zeros(n, np.int64),
         "maxbest": np.zeros(n, np.int64), "theta_vmin": np.zeros(n), "theta_z": np.zeros(n),
         "points": np.zeros((n, 5, 3)), "combo": -np.ones((n, 3), np.int64), "from_reference": np.zeros(n, bool)}
    for i, (vp, counts, maxbest, tv, tz, out, from_ref) in enumerate(cases):
        g["from_reference"][i] = from_ref
        m = vp.shape[0]
        g["vp"][i, :m] = vp; g["counts"][i, :m] = counts; g["n_vp"][i] = m
        g["maxbest"][i] = maxbest; g["theta_vmin"][i] = tv; g["theta_z"][i] = tz
        for q in range(5):
            g["points"][i, q] = out[q]
        c = np.asarray(out[5]).reshape(-1)
        g["combo"][i, :c.size] = c
    np.savez_compressed(OUT, **g)
    scored = sum(1 for c in cases if min(c[2], c[0].shape[0]) > 2)
    print("wrote", OUT, n, "cases,", scored, "with triplets,", int(g["from_reference"].sum()), "run through the reference")


if __name__ == "__main__":
    main()

"""Golden vectors for rows T0 and S1 from the reference's OWN code.

* T0: /root/reference/coordinate_conversion.py is plain numpy and imports under Python 3 unmodified.
  Its four functions are run on seeded inputs (index_to_angle / angle_to_index on the 500-, 250- and
  20-cell grids the pipeline uses, angle_to_point, point_to_angle).
* S1, votes formulation: for ALL pairs i < j of a seeded 450-line scene the intersection
  p = l_i x l_j (unit, z >= 0) goes through the reference's point_to_angle and angle_to_index; the
  rounded cell (row 0 = beta max, the orientation of the reference canvas) is stored per pair.  The
  cross product / normalisation is not reference code (north_star's restatement of the stage); the
  projection and the index map are.
* S1, curves formulation: the great-circle expression is cut out of the source of
  sphere_mapping.sphere_line_plot (sphere_mapping.py:40 and :61-63; the module imports matplotlib,
  which is not installed, so it cannot be imported) and executed unmodified for 48 lines; the row
  index of each of the 10 000 samples (through the reference's angle_to_index) is stored.

Runs only where /root/reference exists.   python oracle/make_golden_sphere.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "sphere_cases.npz")
sys.path.insert(0, ROOT)


def load_cc():
    spec = importlib.util.spec_from_file_location("ref_coordinate_conversion", os.path.join(REF, "coordinate_conversion.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_curve_code():
    """`a = linspace(...)` (:40), the non-alternative `b = ...` (:61) and `b *= -1` (:63) of sphere_line_plot."""
    src = open(os.path.join(REF, "sphere_mapping.py")).read().splitlines()
    start = next(i for i, ln in enumerate(src) if ln.startswith("def sphere_line_plot("))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    body = src[start:end]
    a_line = next(ln.strip() for ln in body if ln.strip().startswith("a = linspace("))
    else_at = next(i for i, ln in enumerate(body) if ln.strip() == "else:")
    b_line = body[else_at + 1].strip()
    neg_line = next(ln.strip() for ln in body[else_at + 1:] if ln.strip().startswith("b *="))
    assert b_line.startswith("b = -np.arctan(") and neg_line == "b *= -1", (b_line, neg_line)
    return a_line, b_line, neg_line


def pair_points(lines):
    """p = l_i x l_j for all i < j, unit, z >= 0 (the same individually rounded float64 operations as
    oracle.sphere_oracle.pair_bins)."""
    n = lines.shape[0]
    ii, jj = np.triu_indices(n, 1)
    li, lj = lines[ii], lines[jj]
    px = li[:, 1] * lj[:, 2] - li[:, 2] * lj[:, 1]
    py = li[:, 2] * lj[:, 0] - li[:, 0] * lj[:, 2]
    pz = li[:, 0] * lj[:, 1] - li[:, 1] * lj[:, 0]
    nr = np.sqrt((px * px + py * py) + pz * pz)
    flip = pz < 0
    px = np.where(flip, -px, px); py = np.where(flip, -py, py); pz = np.abs(pz)
    return np.stack([px / nr, py / nr, pz / nr], axis=1)


def main():
    from vanishing_points_2017_b200 import synth
    cc = load_cc()
    out = {}
    rs = np.random.RandomState(2017)
    # ---- T0 ------------------------------------------------------------------
    for S in (500, 250, 20):
        idx = rs.uniform(-2, S + 2, (400, 2))
        idx[:S // 2] = np.stack([np.arange(S // 2) * 2.0, np.arange(S // 2) * 2.0 + 1], axis=1)      # integer cell centres too
        ang = rs.uniform(-np.pi / 2, np.pi / 2, (400, 2))
        out["t0_idx_%d" % S] = idx
        out["t0_idx2ang_%d" % S] = np.stack([cc.index_to_angle(i, (S, S)) for i in idx])
        out["t0_ang_%d" % S] = ang
        out["t0_ang2idx_%d" % S] = np.stack([cc.angle_to_index(a, (S, S)) for a in ang])
    ang = rs.uniform(-np.pi, np.pi, (600, 2))
    ang[:4] = [[0, 0], [np.pi / 2, 0], [0, np.pi / 2], [-np.pi / 2, -np.pi / 2]]
    out["t0_a2p_in"] = ang
    out["t0_a2p_out"] = np.stack([cc.angle_to_point(a) for a in ang])
    pts = rs.standard_normal((600, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    pts[pts[:, 2] < 0] *= -1
    out["t0_p2a_in"] = pts
    out["t0_p2a_out"] = np.stack([cc.point_to_angle(p) for p in pts])
    # ---- S1 votes ------------------------------------------------------------
    S = 500
    sc = synth.make_scene(seed=501, n_segments=450, width=800, height=600)
    lines = sc["lines"]
    P = pair_points(lines)
    rows = np.empty(P.shape[0], np.int16)
    cols = np.empty(P.shape[0], np.int16)
    for k, p in enumerate(P):
        a = cc.point_to_angle(p)
        i = cc.angle_to_index(a, (S, S))
        cols[k] = int(np.clip(np.floor(i[0] + 0.5), 0, S - 1))
        rows[k] = S - 1 - int(np.clip(np.floor(i[1] + 0.5), 0, S - 1))
    out["votes_segments"] = sc["segments"]
    out["votes_lines"] = lines
    out["votes_rows"] = rows
    out["votes_cols"] = cols
    # ---- S1 curves -----------------------------------------------------------
    a_line, b_line, neg_line = reference_curve_code()
    sc = synth.make_scene(seed=502, n_segments=48, width=640, height=480)
    lines = sc["lines"].copy()
    ns = {"np": np, "linspace": np.linspace, "pi": np.pi, "lines": lines}
    exec(a_line, ns)
    crow = np.empty((lines.shape[0], 10000), np.int16)
    for i in range(lines.shape[0]):
        ns["i"] = i
        with np.errstate(divide="ignore", invalid="ignore"):
            exec(b_line, ns)
            exec(neg_line, ns)
        b = ns["b"]
        for k in range(10000):
            bi = cc.angle_to_index(np.array([ns["a"][k], b[k]]), (S, S))[1]
            crow[i, k] = -1 if np.isnan(bi) else S - 1 - int(np.clip(np.floor(bi + 0.5), 0, S - 1))
    out["curves_lines"] = sc["lines"]
    out["curves_alpha"] = ns["a"]
    out["curves_rows"] = crow
    out["curves_source"] = np.array("%s | %s | %s" % (a_line, b_line, neg_line))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", P.shape[0], "pairs")


if __name__ == "__main__":
    main()

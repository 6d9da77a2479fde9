"""Golden vectors for row N2: runs the reference's own `detect_lsd_lines` (its source is cut out of
/root/reference/evaluation.py -- the module itself imports caffe and cannot be imported here -- and
executed unmodified, with `lsd.detect_line_segments` stubbed to return a prepared array of raw LSD rows)
and stores inputs and outputs in tests/golden/lsd_norm_cases.npz.   python oracle/make_golden_lsd.py"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/evaluation.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "lsd_norm_cases.npz")


def reference_function():
    # the file is Python 2 (print statements elsewhere), so the function is cut out as text: from its
    # `def` line to the next top-level `def`; the function body itself is valid Python 3
    lines = open(REF).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith("def detect_lsd_lines("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("def "))
    code = "\n".join(lines[start:end])
    ast.parse(code)
    ns = {"np": np, "lsd": types.SimpleNamespace(detect_line_segments=None)}
    exec(compile(code, REF, "exec"), ns)
    return ns


def main():
    ns = reference_function()
    rs = np.random.RandomState(404)
    shapes = [(480, 640), (640, 480), (533, 800), (800, 533), (450, 800), (1, 1), (375, 500), (600, 600)]
    cases = {}
    for i, (h, w) in enumerate(shapes):
        n = int(rs.randint(0, 400)) if i != 5 else 3
        rows = np.zeros((n, 7))
        rows[:, 0] = rs.uniform(0, w, n); rows[:, 2] = rs.uniform(0, w, n)
        rows[:, 1] = rs.uniform(0, h, n); rows[:, 3] = rs.uniform(0, h, n)
        rows[:, 4] = rs.uniform(1, 5, n); rows[:, 5] = 0.125; rows[:, 6] = rs.uniform(0, 60, n)
        if n:
            rows[0, 0:4] = [w / 2.0, h / 2.0, 0.0, 0.0]          # exact centre: signed zeros after the flip
        ns["lsd"].detect_line_segments = lambda image, r=rows: r.copy()
        out = ns["detect_lsd_lines"](np.full((h, w), 0.5))
        cases["lsd_%d" % i] = rows
        cases["shape_%d" % i] = np.array([h, w])
        cases["segments_%d" % i] = out["segments"]
        cases["nfa_%d" % i] = out["nfa"]
    cases["n_cases"] = np.array(len(shapes))
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, len(shapes), "cases")


if __name__ == "__main__":
    main()

"""Host mirror of the reference's auc.py (the evaluation that consumes the
horizon errors, benchmark.py:262).  Plain host arithmetic over one value per image -- there is nothing to
put on the device; the errors themselves come from horizon_kernel (`Pipeline.horizons(true_horizons=...)`).
Same name, arguments and return value as auc.calc_auc (auc.py:5-37); sklearn.metrics.auc is the
trapezoidal rule over the points sorted by x."""
import numpy as np


def calc_auc(error_array, cutoff=0.25):
    err = np.sort(np.asarray(error_array, np.float64).squeeze().reshape(-1))      # auc.py:7-8
    n = err.shape[0]
    frac = (np.arange(n) + 1) * 1.0 / n                                           # :17
    pts = np.stack([err, frac], axis=1)
    mid = 1.0
    for i in range(1, n):                                                         # :21-24: the last crossing wins
        if err[i - 1] < cutoff < err[i]:
            mid = (err[i - 1] * frac[i - 1] + err[i] * frac[i]) / (err[i] + err[i - 1])
    last = np.array([cutoff, 1.0 if pts[-1, 0] < cutoff else mid])                # :26-29
    pts = np.vstack([pts, last])
    pts = pts[np.argsort(pts[:, 0]), :]                                           # :31-32
    sel = pts[:, 0] <= cutoff
    x, y = pts[sel, 0], pts[sel, 1]
    area = float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2.0)) if x.shape[0] > 1 else 0.0   # sklearn.metrics.auc
    return area / cutoff, pts

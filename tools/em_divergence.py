"""GPU box: how the CUDA EM and the oracle drift apart with the iteration count on one image of
gpurun_out/examples_case.npz (rounding-level differences amplified by the iteration, or a defect?)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import lsd_oracle, vp_oracle as vo  # noqa: E402
from vanishing_points_2017_b200 import vp_localisation as em  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cases", "examples_case.npz"))
b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
s = g["seg_%d" % b]
sig = g["sig"][b].astype(np.float64)
sph = g["sph"][b]
lines = lsd_oracle.lines_from_segments(s)
for it in (1, 2, 3, 5, 8, 10, 11, 12, 14, 16, 18, 19, 20, 21):
    ref = vo.expectation_maximisation(lines.copy(), s.copy(), sig.copy(), sphere_image=sph, num_iter=it)
    res = em.expectation_maximisation(lines.copy(), s.copy(), sig.copy(), sphere_image=sph, num_iter=it)
    if ref["vp"] is None or res["vp"] is None:
        print(it, "None", ref["vp"] is None, res["vp"] is None)
        continue
    same_shape = ref["vp"].shape == res["vp"].shape
    line = "num_iter %2d: iterations %d/%d VPs %d/%d" % (it, ref["iterations"], res["iterations"], ref["vp"].shape[0], res["vp"].shape[0])
    if same_shape:
        ang = np.arccos(np.minimum(np.abs(np.sum(ref["vp"] * res["vp"], axis=1)), 1.0))
        dm = np.abs(res["decision_metric"] - ref["decision_metric"]) / (np.abs(ref["decision_metric"]) + 1e-300)
        line += "  max angle %.3e  sigma rel %.3e  dm rel (median %.2e, max %.2e)  assoc diff %d" % (
            ang.max(), np.max(np.abs(res["sigma"] / ref["sigma"] - 1)), np.median(dm), dm.max(), int(np.sum(res["vp_assoc"] != ref["vp_assoc"])))
    print(line, flush=True)

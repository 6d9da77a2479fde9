"""Debug helper: run the resident pipeline on a small batch and dump inputs,
intermediates and results to gpurun_out/ for offline comparison with the oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cnn_oracle
from vanishing_points_2017_b200 import pipeline, synth

ws, bs = cnn_oracle.random_weights(0, scale=3.0)
p = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
batch = synth.make_batch(2, n_images=5)
res, sig, sph = p(batch["segments"], batch["offsets"], want_response=True, want_sphere=True)
out = {"segments": batch["segments"], "offsets": batch["offsets"], "sig": sig, "sph": sph}
for b, r in enumerate(res):
    if r["vp"] is not None:
        for k in ("vp", "counts", "vp_assoc", "sigma"):
            out["%s_%d" % (k, b)] = r[k]
        out["iter_%d" % b] = np.array(r["iterations"])
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/pipeline_case.npz", **out)
print([None if r["vp"] is None else (r["vp"].shape[0], r["iterations"]) for r in res])

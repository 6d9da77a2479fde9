"""Debug helper: the example-image fixtures through the pipeline; inputs, intermediates and results to
gpurun_out/examples_case.npz for offline comparison with the oracle / the host build of the EM."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cnn_oracle, lsd_oracle
from vanishing_points_2017_b200 import pipeline
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "examples_lsd.npz"))
n = int(g["n_images"])
rows = [g["lsd_%d" % i] for i in range(n)]
shapes = [tuple(int(v) for v in g["shape_%d" % i]) for i in range(n)]
off = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows])]).astype(np.int32)
ws, bs = cnn_oracle.random_weights(0, scale=3.0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
pipe.upload_lsd(np.concatenate(rows), off, [s[1] for s in shapes], [s[0] for s in shapes])
pipe.run()
res, sig, sph = pipe.fetch(want_response=True, want_sphere=True)
out = {"off": off, "sig": sig, "sph": sph}
for b, r in enumerate(res):
    out["seg_%d" % b] = lsd_oracle.segments_from_lsd(rows[b], shapes[b])["segments"]
    if r["vp"] is not None:
        for k in ("vp", "counts", "vp_assoc", "sigma"):
            out["%s_%d" % (k, b)] = r[k]
        out["iter_%d" % b] = np.array(r["iterations"])
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/examples_case.npz", **out)
print([None if r["vp"] is None else (r["vp"].shape[0], r["iterations"]) for r in res])

import sys, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import cnn_oracle, sphere_oracle, vp_oracle
from vanishing_points_2017_b200 import pipeline, synth
ws, bs = cnn_oracle.random_weights(0, scale=3.0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
segs, ns = [], [120, 64, 200]
for i, n in enumerate(ns):
    segs.append(synth.make_scene(4242 + i, n)["segments"])
seg = np.concatenate(segs)
off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
res, sig, sph = pipe(seg, off, want_response=True, want_sphere=True)
for b in range(3):
    s = seg[off[b]:off[b + 1]]
    ref = vp_oracle.expectation_maximisation(synth.lines_from_segments(s), s.copy(), sig[b].copy(), sphere_image=sph[b])
    print("image", b, "ours", None if res[b]["vp"] is None else res[b]["vp"].shape, "iters", res[b].get("iterations"), "counts", res[b].get("counts"))
    print("       ref", None if ref["vp"] is None else ref["vp"].shape, "iters", ref.get("iterations"), "counts", ref.get("counts"))
    if res[b]["vp"] is not None and ref["vp"] is not None:
        print("  ours vp\n", res[b]["vp"], "\n  ref vp\n", ref["vp"])
        print("  ours sigma", res[b]["sigma"], "ref sigma", ref["sigma"])
import os
os.makedirs("gpurun_out", exist_ok=True)
out = {"seg": seg, "off": off, "sig": sig, "sph": sph}
for b, r in enumerate(res):
    if r["vp"] is not None:
        for k in ("vp", "counts", "vp_assoc", "sigma"):
            out["%s_%d" % (k, b)] = r[k]
        out["iter_%d" % b] = np.array(r["iterations"])
np.savez_compressed("gpurun_out/smoke_case.npz", **out)

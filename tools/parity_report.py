"""GPU box: which candidate smoke scenes match the oracle, and how stable the ORACLE itself is on each
(16 runs on inputs perturbed by 1e-14 relative).  Output goes to profiles/ as evidence for the scenes
pinned in __graft_entry__.smoke() and tests/test_examples_gpu.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import cnn_oracle, lsd_oracle, sphere_oracle, vp_oracle  # noqa: E402
from vanishing_points_2017_b200 import pipeline, synth  # noqa: E402


def same(a, r):
    if (a["vp"] is None) != (r["vp"] is None):
        return False
    if r["vp"] is None:
        return True
    if a["vp"].shape != r["vp"].shape:
        return False
    return np.arccos(np.minimum(np.abs(np.sum(a["vp"] * r["vp"], axis=1)), 1.0)).max() < 1e-4


def report(tag, segs, res, sig, sph, trials=16):
    rs = np.random.RandomState(7)
    for b, s in enumerate(segs):
        def oracle_em(sp):
            try:
                return vp_oracle.expectation_maximisation(synth.lines_from_segments(sp), sp.copy(), sig[b].astype(np.float64),
                                                          sphere_image=sph[b])
            except ValueError:
                return {"vp": None}
        ref = oracle_em(s)
        ok = same(res[b], ref)
        exact = ok and (ref["vp"] is None or (np.array_equal(res[b]["vp_assoc"], ref["vp_assoc"]) and
                                               np.array_equal(res[b]["counts"], ref["counts"])))
        flips = sum(0 if same(oracle_em(s * (1.0 + 1e-14 * rs.standard_normal(s.shape))), ref) else 1 for _ in range(trials))
        print("%s image %d N=%d: gpu==oracle %s, assoc/counts identical %s, oracle flips %d/%d, VPs %s" % (
            tag, b, s.shape[0], ok, exact, flips, trials, None if ref["vp"] is None else ref["vp"].shape[0]), flush=True)


ws, bs = cnn_oracle.random_weights(0, scale=3.0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
scenes = [(4250, 90), (4254, 75), (4258, 140), (4244, 200), (4251, 110), (4252, 160), (4253, 130), (4255, 180), (4256, 100),
          (4257, 150), (4259, 120), (4260, 170)]
segs = [synth.make_scene(seed, n)["segments"] for seed, n in scenes]
off = np.concatenate([[0], np.cumsum([s.shape[0] for s in segs])]).astype(np.int32)
res, sig, sph = pipe(np.concatenate(segs), off, want_response=True, want_sphere=True)
print("scenes:", scenes)
report("smoke-candidate", segs, res, sig, sph)

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "examples_lsd.npz"))
n = int(g["n_images"])
rows = [g["lsd_%d" % i] for i in range(n)]
shapes = [tuple(int(v) for v in g["shape_%d" % i]) for i in range(n)]
off = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows])]).astype(np.int32)
pipe.upload_lsd(np.concatenate(rows), off, [s[1] for s in shapes], [s[0] for s in shapes])
pipe.run()
res, sig, sph = pipe.fetch(want_response=True, want_sphere=True)
segs = [lsd_oracle.segments_from_lsd(rows[i], shapes[i])["segments"] for i in range(n)]
report("example", segs, res, sig, sph, trials=6)

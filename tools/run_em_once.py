"""One EM call (C ABI, `vpk_em`) on one synthetic image of N segments: for compute-sanitizer / ncu."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vanishing_points_2017_b200 import sphere_mapping, synth, vp_localisation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1600)
ap.add_argument("--num-iter", type=int, default=100)
ap.add_argument("--num-init-vp", type=int, default=25)
a = ap.parse_args()
sc = synth.make_scene(7300 + a.n, a.n, 800, 600, noise_deg=1.0)
_, img = sphere_mapping.sphere_votes(sc["lines"], 500)
resp = synth.ideal_response(sc["vps"], seed=a.n)
res = vp_localisation.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp, sphere_image=img,
                                               num_iter=a.num_iter, num_init_vp=a.num_init_vp)
print("images with VPs:", 0 if res["vp"] is None else 1, "iterations", res["iterations"],
      None if res["vp"] is None else res["counts"].astype(int).tolist())

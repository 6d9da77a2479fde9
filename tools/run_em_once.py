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
if os.environ.get("VPK_PROFILE"):
    # second call with per-kernel events and the device-side phase counters (fused mode: per-phase cycles of the
    # leading CTA; VPK_EM_MARKS builds also print POST's inner marks to stderr)
    from vanishing_points_2017_b200 import _lib
    ctx = _lib.default_context()
    ctx.profile_reset(); ctx.em_stats(reset=True); ctx.em_phase_cycles(reset=True)
    ctx.profile_enable(True)
    res = vp_localisation.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp, sphere_image=img,
                                                   num_iter=a.num_iter, num_init_vp=a.num_init_vp)
    ctx.profile_enable(False)
    st = ctx.em_stats()
    print("kernels:", {k: (round(v["ms"] * 1e3 / max(v["launches"], 1), 1), v["launches"]) for k, v in ctx.profile_read().items()}, "(us per launch, launches)")
    cyc = ctx.em_phase_cycles()
    n = max(st["wmat_products"], 1)
    print("supersteps", st["supersteps"], "phase us per superstep @1.9GHz:", {k: round(v / n / 1900.0, 2) for k, v in cyc.items()})

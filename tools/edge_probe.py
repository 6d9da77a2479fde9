"""ad-hoc: EM at sizes the tests do not reach (tiny N, N > 5000) against the oracle."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import sphere_oracle as so, vp_oracle as vo
from vanishing_points_2017_b200 import synth, sphere_mapping as sm, vp_localisation as em

def run(seed, N, kw={}):
    sc = synth.make_scene(seed, N, 800, 600, noise_deg=1.0)
    img = sm.sphere_votes(sc["lines"].copy(), 500)[1]          # the bins are bit-exact against the oracle (test_sphere_gpu)
    resp = synth.ideal_response(sc["vps"], seed=seed)
    t = time.time()
    try:
        ref = vo.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp.copy(), sphere_image=img, **kw)
    except ValueError:
        ref = {"vp": None}
    t_or = time.time() - t
    t = time.time()
    res = em.expectation_maximisation(sc["lines"].copy(), sc["segments"].copy(), resp.copy(), sphere_image=img, **kw)
    t_gpu = time.time() - t
    if ref["vp"] is None or res["vp"] is None:
        print(N, "oracle None" if ref["vp"] is None else "oracle VPs", "gpu None" if res["vp"] is None else "gpu VPs", flush=True)
        return
    ok = res["iterations"] == int(ref["iterations"]) and res["vp"].shape == ref["vp"].shape
    ang = np.arccos(np.minimum(np.abs(np.sum(res["vp"] * ref["vp"], axis=1)), 1.0)).max() if ok else -1
    same = ok and np.array_equal(res["vp_assoc"], ref["vp_assoc"]) and np.array_equal(res["counts"], ref["counts"])
    print("N=%d iters gpu %d oracle %d  max angle %.2e  assoc/counts identical %s  oracle %.1fs gpu %.2fs" %
          (N, res["iterations"], int(ref["iterations"]), ang, same, t_or, t_gpu), flush=True)

for i, N in enumerate([1, 2, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 65, 127, 129]):
    run(9000 + i, N)
for i, N in enumerate([6000, 8000]):
    run(9100 + i, N, dict(num_init_vp=32) if i else {})

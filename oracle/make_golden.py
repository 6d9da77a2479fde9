"""Generate tests/golden/*.npz from the reference's OWN EM implementation.

Runs only in the build container (needs /root/reference).  The reference's
modules are Python 2; they are read from /root/reference, given the five
mechanical Python-3 patches of SURVEY.md section 8(c) *in a temp directory*
(nothing of the reference is written into this repo), imported, and executed
on seeded synthetic scenes.  No arithmetic is changed by the patches:

  1. print statements -> print()            2. `/` used for slicing -> `//`
  3. np.linalg.linalg.LinAlgError -> np.linalg.LinAlgError
  4. sklearn kwarg affinity= -> metric=     5. np.array(to_be_removed) -> dtype=int

Usage:  python oracle/make_golden.py [--out tests/golden]
"""
import argparse
import importlib
import os
import re
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)


def load_reference():
    from oracle import ref_patch
    return ref_patch.load(ref_patch.materialise(tempfile.mkdtemp(prefix="vpref_")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--large", action="store_true",
                    help="only the N = 1600 run (em_full_n1600_large.npz; the reference takes minutes there)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    warnings.simplefilter("ignore")

    from vanishing_points_2017_b200 import synth
    from oracle import sphere_oracle

    vp, prob = load_reference()
    import contextlib, io

    S = 500
    if args.large:
        return full_runs(vp, args.out, S, [("n1600_large", 301, 1600, 800, 600, "ideal", 0.15, 1, 1.0)])
    # ---- function-level vectors (one small scene) -------------------------
    sc = synth.make_scene(seed=11, n_segments=72, width=640, height=480)
    lp = sc["segments"]
    l = sc["lines"] / np.linalg.norm(sc["lines"], axis=1, keepdims=True)
    resp = synth.ideal_response(sc["vps"], seed=11)
    S = 500
    img = sphere_oracle.votes_to_image(sphere_oracle.sphere_votes(sc["lines"], S))
    out = {"lp": lp, "l": l, "resp": resp, "sphere_image": img}
    out["lsim"] = vp.calc_lsim(lp, sigma=1)
    out["lscore_k4"] = vp.line_rating_knn(lp, k2=4)
    out["langles"] = vp.lines_angles(lp)
    out["maxima"] = vp.find_maxima(resp)
    out["v0"] = vp.find_initial_vps(img, resp, 25)
    pp = prob.pdf_params(resp)
    out["pdf_means"], out["pdf_weights"], out["pdf_sigma"] = pp.means, pp.weights, np.array(pp.sigma)
    llen = np.array([vp.line_length(x) for x in lp])
    lweight = llen * np.clip(out["lscore_k4"], 0.2, 1)
    out["lweight"] = lweight
    M = out["v0"].shape[0]
    s = np.ones(M) * pp.sigma * 1e-6
    p = prob.calc_probabilities(0, pp, out["v0"][None, :, :], l, lp, s, llen, "angle")
    out["p_v"], out["p_lv"], out["p_vl"], out["p_l"], out["lvsq"], out["angles"] = p.v, p.lv, p.vl, p.l, p.lvsq, p.angles
    w = vp.weight_matrix(p.vl, lweight, out["lsim"], bias=1)
    out["w"] = w
    c, cw, assoc = vp.calc_vp_line_counts(out["v0"], l, lp, s, w, lweight, "angle", thresh=1.96 ** 2)
    out["counts"], out["counts_weighted"], out["assoc"] = c, cw, assoc
    out["newvp"] = np.stack([vp.calc_new_vanishing_point(l, w[m]) if np.max(w[m]) > 0 else np.zeros(3)
                             for m in range(M)])
    # merge: duplicate the best-supported VP with a tiny perturbation
    best = int(np.argmax(c))
    vdup = out["v0"][best] + np.array([2e-4, -1e-4, 0.0])
    vdup /= np.linalg.norm(vdup)
    vm = np.vstack([out["v0"][c >= 3], vdup[None, :]])
    sm_ = np.ones(vm.shape[0]) * pp.sigma * 1e-6
    out["merge_in_v"], out["merge_in_s"] = vm.copy(), sm_.copy()
    hist = np.zeros((3, vm.shape[0], 3)); hist[1] = vm; hist[0] = vm
    with contextlib.redirect_stdout(io.StringIO()):
        mg = vp.merge_vps(1, hist, sm_.copy(), l, 1e-2, lweight, out["lsim"], 1, pp, lp, llen, "angle", outlier_stdd=1)
    out["merge_out_v"], out["merge_out_s"] = mg["v"][1], mg["s"]
    np.savez_compressed(os.path.join(args.out, "em_functions_n72.npz"), **out)
    print("functions: M0=%d  merge %d -> %d" % (M, vm.shape[0], mg["v"].shape[1]))

    # ---- split vector: two VPs fused into one hypothesis -------------------
    sc = synth.make_scene(seed=23, n_segments=160, width=640, height=480, outlier_frac=0.05)
    lp = sc["segments"]
    l = sc["lines"] / np.linalg.norm(sc["lines"], axis=1, keepdims=True)
    lsim = vp.calc_lsim(lp, sigma=1)
    llen = np.array([vp.line_length(x) for x in lp])
    lweight = llen * np.clip(vp.line_rating_knn(lp, k2=4), 0.2, 1)
    langles = vp.lines_angles(lp)
    resp = synth.ideal_response(sc["vps"], seed=23)
    pp = prob.pdf_params(resp)
    # hypotheses: true VP 0 and a VP halfway between true VPs 1 and 2 with a large variance
    tv = sc["vps"]
    inside = [k for k in range(3) if abs(tv[k, 0] / tv[k, 2]) < 1 and abs(tv[k, 1] / tv[k, 2]) < 1]
    vsplit = np.vstack([tv[0], tv[1], tv[2]])
    ssplit = np.array([1e-6, 1e-6, 1e-6])
    p = prob.calc_probabilities(0, pp, vsplit[None], l, lp, ssplit, llen, "angle")
    w = vp.weight_matrix(p.vl, lweight, lsim, bias=1)
    hist = np.zeros((2, 3, 3)); hist[0] = vsplit; hist[1] = vsplit
    sp = vp.split_best_vp(0, hist.copy(), ssplit.copy(), linePoints=lp, lines=l, weightMatrix=w,
                          lineWeights=lweight, lineAngles=langles, min_diff=1e-3)
    np.savez_compressed(os.path.join(args.out, "em_split_n160.npz"), lp=lp, l=l, lsim=lsim, lweight=lweight,
                        langles=langles, w=w, v_in=vsplit, s_in=ssplit, v_out=sp["v"][0], s_out=sp["s"],
                        inside=np.array(inside))
    print("split: M 3 ->", sp["v"].shape[1])

    # ---- full EM runs ------------------------------------------------------
    full_runs(vp, args.out, S, FULL_CASES)


FULL_CASES = [
        # (tag, seed, N, w, h, response kind, outlier_frac, extra_vps)
        ("n90_ideal", 101, 90, 640, 480, "ideal", 0.15, 0),
        ("n150_ideal", 102, 150, 640, 480, "ideal", 0.15, 0),
        ("n150_f32", 103, 150, 800, 600, "ideal32", 0.15, 1),
        ("n250_ideal", 104, 250, 800, 533, "ideal", 0.15, 0),
        ("n250_noisy", 105, 250, 600, 800, "noisy", 0.25, 2),
        ("n400_ideal", 106, 400, 640, 480, "ideal", 0.15, 0),
        ("n600_split", 107, 600, 800, 600, "ideal", 0.15, 1),
        # noisier scenes: >= 10 iterations, so the periodic split/merge paths run
        ("n400_long", 206, 400, 800, 600, "ideal", 0.2, 3, 2.0),
        ("n600_long", 204, 600, 800, 600, "ideal", 0.3, 2, 3.0),
        ("n800_long", 203, 800, 800, 600, "ideal", 0.15, 0, 1.0),
    ]


def full_runs(vp, out_dir, S, cases):
    import contextlib, io
    from vanishing_points_2017_b200 import synth
    from oracle import sphere_oracle
    split_calls = []
    orig_split = vp.split_best_vp

    def split_spy(i, v, s, **kw):
        rec = {"v_in": v[i].copy(), "s_in": s.copy(), "w": kw["weightMatrix"].copy()}
        r = orig_split(i, v, s, **kw)
        rec["v_out"], rec["s_out"] = r["v"][i].copy(), r["s"].copy()
        split_calls.append(rec)
        return r

    vp.split_best_vp = split_spy
    for case in cases:
        tag, seed, N, wd, ht, kind, ofrac, extra = case[:8]
        noise = case[8] if len(case) > 8 else 0.5
        split_calls.clear()
        sc = synth.make_scene(seed=seed, n_segments=N, width=wd, height=ht, outlier_frac=ofrac,
                              extra_vps=extra, noise_deg=noise)
        lp = sc["segments"].copy()
        lines = sc["lines"].copy()
        img = sphere_oracle.votes_to_image(sphere_oracle.sphere_votes(lines, S))
        if kind == "ideal":
            resp = synth.ideal_response(sc["vps"], seed=seed)
        elif kind == "ideal32":
            resp = synth.ideal_response(sc["vps"], seed=seed).astype(np.float32)
        else:
            rs = np.random.RandomState(seed)
            resp = np.clip(synth.ideal_response(sc["vps"], seed=seed, peak=0.7, noise=0.3)
                           + rs.normal(0, 0.05, (20, 20)), 0.001, 0.999)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = vp.expectation_maximisation(lines.copy(), lp.copy(), resp.copy(), sphere_image=img.copy(),
                                              distance_measure="angle", use_weights=True, do_split=True,
                                              do_merge=True)
        log = buf.getvalue()
        np.savez_compressed(os.path.join(out_dir, "em_full_%s.npz" % tag),
                            segments=lp, lines=lines, resp=resp, sphere_image=img, true_vps=sc["vps"],
                            vp=res["vp"], counts=res["counts"], counts_weighted=res["counts_weighted"],
                            vp_assoc=res["vp_assoc"], sigma=res["sigma"],
                            iterations=np.array(res["iterations"]), decision_metric=res["decision_metric"],
                            log=np.array(log))
        print("%s: N=%d iterations=%d VPs=%d counts=%s" % (tag, N, res["iterations"], res["vp"].shape[0],
                                                          res["counts"].astype(int).tolist()))
        grew = [c for c in split_calls if c["v_out"].shape[0] > c["v_in"].shape[0]]
        if tag == "n600_long" and grew:
            c = grew[0]
            ln = lines / np.linalg.norm(lines, axis=1, keepdims=True)
            llen = np.array([vp.line_length(x) for x in lp])
            with contextlib.redirect_stdout(io.StringIO()):
                lwt = llen * np.clip(vp.line_rating_knn(lp, k2=4), 0.2, 1)
            np.savez_compressed(os.path.join(out_dir, "em_split_real_n600.npz"), lp=lp, l=ln, lweight=lwt,
                                langles=vp.lines_angles(lp), **c)
            print("  split vector: M %d -> %d" % (c["v_in"].shape[0], c["v_out"].shape[0]))


if __name__ == "__main__":
    main()

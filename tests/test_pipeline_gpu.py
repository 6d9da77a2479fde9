"""The resident pipeline equals the three stages called one by one, and the
stages equal their oracles on the pipeline's own intermediates."""
import numpy as np
import pytest

from oracle import cnn_oracle, sphere_oracle as so, vp_oracle as vo
from vanishing_points_2017_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe():
    from vanishing_points_2017_b200 import pipeline
    ws, bs = cnn_oracle.random_weights(0, scale=3.0)
    return pipeline.Pipeline(0, ws, bs, sphere_mode="votes"), ws, bs


def test_pipeline_matches_stagewise_oracles(pipe):
    p, ws, bs = pipe
    batch = synth.make_batch(2, n_images=5)
    res, sig, sph = p(batch["segments"], batch["offsets"], want_response=True, want_sphere=True)
    off = batch["offsets"]
    for b in range(5):
        lines = batch["lines"][off[b]:off[b + 1]]
        segs = batch["segments"][off[b]:off[b + 1]]
        # stage 1: bit-exact image
        np.testing.assert_array_equal(sph[b], so.votes_to_image(so.sphere_votes(lines, 500)))
    # stage 2: same weights, fp32 oracle
    rsig, rlog = cnn_oracle.forward(sph, ws, bs)
    np.testing.assert_allclose(sig, rsig, atol=0.25 * 1e-2 * float(np.max(np.abs(rlog))) + 1e-6)
    # stage 3: oracle EM on the pipeline's own sphere image and response
    for b in range(5):
        lines = batch["lines"][off[b]:off[b + 1]].copy()
        segs = batch["segments"][off[b]:off[b + 1]].copy()
        try:
            ref = vo.expectation_maximisation(lines, segs, sig[b].copy(), sphere_image=sph[b])
        except ValueError:
            ref = {"vp": None}
        if ref["vp"] is None:
            assert res[b]["vp"] is None
            continue
        assert res[b]["vp"].shape == ref["vp"].shape
        ang = np.arccos(np.minimum(np.abs(np.sum(res[b]["vp"] * ref["vp"], axis=1)), 1.0))
        assert ang.max() < 1e-4
        np.testing.assert_array_equal(res[b]["counts"], ref["counts"])
    ms = p.stage_ms()
    assert ms["total"] > 0 and abs(ms["sphere"] + ms["cnn"] + ms["em"] - ms["total"]) < 0.05 * ms["total"] + 0.1


def test_pipeline_is_deterministic_and_shard_invariant(pipe):
    from vanishing_points_2017_b200 import pipeline
    p, _, _ = pipe
    batch = synth.make_batch(4, n_images=12)
    full = p(batch["segments"], batch["offsets"])
    again = p(batch["segments"], batch["offsets"])
    for a, b in zip(full, again):
        assert (a["vp"] is None) == (b["vp"] is None)
        if a["vp"] is not None:
            np.testing.assert_array_equal(a["vp"], b["vp"])
    # sharded over 3 "ranks" (same GPU here) and gathered: identical per-image results
    got = [None] * 12
    for r in range(3):
        idx = pipeline.shard_batch(batch["offsets"], 3, r)
        seg, off = pipeline.take_images(batch["segments"], batch["offsets"], idx)
        part = pipeline.gather_results(p(seg, off), idx, 12, 1)
        for i in idx:
            got[i] = part[i]
    for a, b in zip(full, got):
        if a["vp"] is not None:
            np.testing.assert_array_equal(a["vp"], b["vp"])


def test_streamed_pipeline_keeps_batches_in_flight_and_equals_the_serial_results(pipe):
    """pipeline.StreamedPipeline: several library contexts, each driven by its own host thread, keep several batches
    in flight on one GPU.  Every batch still runs alone on its context, so the results are bit-identical to the
    serial Pipeline's, in submission order."""
    from vanishing_points_2017_b200 import pipeline
    p, ws, bs = pipe
    batches = []
    for k in range(5):
        b = synth.make_batch(2, n_images=6 + k)
        batches.append((b["segments"], b["offsets"]))
    serial = [p(seg, off, raw=True) for seg, off in batches]
    sp = pipeline.StreamedPipeline(0, ws, bs, depth=2, sphere_mode="votes")
    try:
        assert sp.depth == 2 and sp.pipes[0].ctx.h != sp.pipes[1].ctx.h
        streamed = list(sp.map(batches, raw=True))
        futs = [sp.submit(seg, off, raw=True) for seg, off in batches[:2]]      # both contexts busy at once
        streamed += [f.result() for f in futs]
    finally:
        sp.close()
    for a, b in zip(serial + serial[:2], streamed):
        for key in ("status", "n_vp", "iterations", "counts"):
            np.testing.assert_array_equal(a[key], b[key])
        nv = a["n_vp"]
        for i in range(len(nv)):
            np.testing.assert_array_equal(a["vp"][i, :nv[i]], b["vp"][i, :nv[i]])
            np.testing.assert_array_equal(a["sigma"][i, :nv[i]], b["sigma"][i, :nv[i]])
    # device marks across contexts
    c0, c1 = pipeline._lib.Context(0), pipeline._lib.Context(0)
    c0.mark(0); c1.mark(1)
    assert c0.elapsed_ms(0, c1, 1) >= 0.0
    c0.close(); c1.close()


def test_pipeline_defaults_to_the_reference_line_plot_when_weights_are_given(pipe):
    """With weights supplied and no sphere_mode, the pipeline feeds the CNN what the trained net expects: the
    great-circle line plot of sphere_mapping.sphere_line_plot (sphere_mapping.py:36-72), not the vote histogram."""
    from vanishing_points_2017_b200 import _lib, pipeline
    _, ws, bs = pipe
    p = pipeline.Pipeline(0, ws, bs, ctx=_lib.Context(0))
    try:
        assert p.mode == _lib.SPHERE_CURVES
        batch = synth.make_batch(2, n_images=3)
        res, sig, sph = p(batch["segments"], batch["offsets"], want_response=True, want_sphere=True)
        off = batch["offsets"]
        for b in range(3):
            lines = batch["lines"][off[b]:off[b + 1]].copy()
            np.testing.assert_array_equal(sph[b], so.sphere_line_plot(lines, 500, alpha=0.1))
        again = p(batch["segments"], batch["offsets"], want_sphere=True)[2]       # cached tables, no per-call host work
        np.testing.assert_array_equal(again, sph)
    finally:
        p.ctx.close()

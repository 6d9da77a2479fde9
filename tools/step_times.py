"""Per-step times of the resident and the end-to-end (host buffers) pipeline.
--verbose: one line per step (interleaves with VPK_EM_TRACE=1 output on stderr)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vanishing_points_2017_b200 import cnn as vcnn, pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--images", type=int, default=None)
ap.add_argument("--runs", type=int, default=12)
ap.add_argument("--verbose", action="store_true")
a = ap.parse_args()
name, seg, off = bench.make_workload(a.config, 0, a.images)
ws, bs = vcnn.random_weights(0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
pipe.upload(seg, off)
res, e2e = [], []
for i in range(a.runs):
    t0 = time.perf_counter()
    pipe.run()
    pipe.ctx.synchronize()
    w = (time.perf_counter() - t0) * 1e3
    ms = pipe.stage_ms()
    res.append("%.2f/%.2f/%.2f" % (ms["em"], ms["total"], w))
    if a.verbose:
        print("step %d sphere %.3f cnn %.3f em %.3f total %.3f wall %.3f" % (i, ms["sphere"], ms["cnn"], ms["em"], ms["total"], w),
              file=sys.stderr, flush=True)
import torch  # noqa: E402
seg_pin, off_pin = torch.from_numpy(seg).pin_memory().numpy(), torch.from_numpy(off).pin_memory().numpy()
for _ in range(a.runs):
    t0 = time.perf_counter()
    pipe(seg_pin, off_pin, raw=True)
    e2e.append("%.2f" % ((time.perf_counter() - t0) * 1e3))
print("groups", os.environ.get("VPK_EM_GROUPS", "default"), "host loop" if os.environ.get("VPK_EM_HOST_LOOP") else "device loop",
      "resident em/total/wall ms:", " ".join(res))
print("   e2e wall ms:", " ".join(e2e))

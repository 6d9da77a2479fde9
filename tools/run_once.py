"""Run the resident pipeline a few times on one synthetic batch (for ncu captures)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vanishing_points_2017_b200 import cnn as vcnn, pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--images", type=int, default=None)
ap.add_argument("--runs", type=int, default=1)
a = ap.parse_args()
name, seg, off = bench.make_workload(a.config, 0, a.images)
ws, bs = vcnn.random_weights(0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
pipe.upload(seg, off)
for _ in range(a.runs):
    pipe.run()
res = pipe.fetch(raw=True)
print("images with VPs:", int((res["status"] == 0).sum()), "of", len(off) - 1, "iterations max", int(res["iterations"].max()))

"""Diagnostic build only (VPK_DEFINES=VPK_EM_MARKS python -m vanishing_points_2017_b200.build --force):
cycles of the POST kernel between its markers, averaged per (image, superstep)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vanishing_points_2017_b200 import cnn as vcnn, pipeline  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name, seg, off = bench.make_workload(cfg, 0, None)
ws, bs = vcnn.random_weights(0)
pipe = pipeline.Pipeline(0, ws, bs, sphere_mode="votes")
pipe.upload(seg, off)
pipe.run()
pipe.ctx.profile_enable(True)
pipe.run()
pipe.ctx.profile_enable(False)
print(pipe.ctx.em_stats())

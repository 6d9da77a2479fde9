"""GPU box: throughput of the resident pipeline with 1 vs 2 (vs 3) batches in flight -- one library context
(own streams, own workspaces) and one host thread per batch in flight."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vanishing_points_2017_b200 import _lib, cnn as vcnn, pipeline  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
name, seg, off = bench.make_workload(cfg, 0, None)
B = len(off) - 1
ws, bs = vcnn.random_weights(0)
for depth in [int(x) for x in os.environ.get("DEPTHS", "1,2,3").split(",")]:
    pipes = [pipeline.Pipeline(0, ws, bs, sphere_mode="votes", ctx=_lib.Context(0)) for _ in range(depth)]
    for p in pipes:
        p.upload(seg, off)
        for _ in range(3):
            p.run()

    def work(p, n):
        for _ in range(n):
            p.run()
    th = [threading.Thread(target=work, args=(p, steps)) for p in pipes]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print("%s depth %d: %.0f images/s  (%.3f ms per batch of %d)" % (name, depth, B * steps * depth / dt, dt * 1e3 / (steps * depth), B), flush=True)
    for p in pipes:
        p.ctx.close()

#!/bin/bash
# sweep of the number of EM groups (concurrent device-driven superstep loops); usage: tools/sweep_groups.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
for cfg in 2 3 4; do
  for g in 1 2 3 4 6 8; do
    steps=10; [ $cfg = 3 ] && steps=4; [ $cfg = 4 ] && steps=2
    VPK_EM_GROUPS=$g python bench.py --config $cfg --steps $steps --no-cpu-baseline > gpurun_out/${TAG}_c${cfg}_g${g}.json 2> gpurun_out/${TAG}_c${cfg}_g${g}.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_c${cfg}_g${g}.json").read().strip().splitlines()[-1])
    print("cfg $cfg groups $g value %.0f e2e %.0f em_ms %.3f total_ms %.3f ok %d" % (d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["em"], d["ms_per_step"], d["images_with_vps"]))
except Exception as e:
    print("cfg $cfg groups $g FAILED", e)
PY
  done
done

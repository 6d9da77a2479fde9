#!/bin/bash
# ncu passes of one resident-pipeline run (YUD-shaped batch); outputs under gpurun_out/<tag>_*
# usage: tools/gpu_profile.sh TAG
# The EM runs with the host-driven loop and one group here, so that every superstep kernel is an
# ordinary stream launch in submission order (ncu serialises the launches anyway).
TAG=${1:-rX}
mkdir -p gpurun_out
export VPK_EM_HOST_LOOP=1 VPK_EM_GROUPS=1
# launch list (per-launch durations, cold-cache + serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/run_once.py --runs 2 > gpurun_out/${TAG}_launches.log 2>&1
# full capture: three supersteps' worth of the EM kernels (skipping the first 10 supersteps)
ncu --set full --clock-control none --import-source on -k regex:"em_(estep|wmat|post)_kernel" -s 30 -c 9 \
    -o gpurun_out/${TAG}_em_steps -f python tools/run_once.py --runs 1 > gpurun_out/${TAG}_em_steps.log 2>&1
# full capture: the once-per-batch kernels
ncu --set full --clock-control none --import-source on -k regex:"em_pair|em_init|sphere_votes|sphere_items|sphere_curves|gemm_bf16|splitk|lrn_pool|votes_image|plane_max|conv1_operand" -c 20 \
    -o gpurun_out/${TAG}_once -f python tools/run_once.py --runs 1 > gpurun_out/${TAG}_once.log 2>&1
ls -la gpurun_out | tail -8
# summarise on the box and drop the reports (gpurun_out/ may carry at most 64 MiB back)
python tools/ncu_summary.py gpurun_out/${TAG}_em_steps.ncu-rep gpurun_out/${TAG}_ncu_full_em_supersteps.csv
python tools/ncu_summary.py gpurun_out/${TAG}_once.ncu-rep gpurun_out/${TAG}_ncu_full_once_kernels.csv
if [ -z "$KEEP_NCU_REP" ]; then rm -f gpurun_out/${TAG}_em_steps.ncu-rep gpurun_out/${TAG}_once.ncu-rep; fi

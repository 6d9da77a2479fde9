#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation: the mbarrier /
# bulk-copy ring of the weight-matrix pass, the cluster / distributed-shared-memory code and the ticket
# counters that close a superstep, the tcgen05 / TMA GEMM, the sphere-vote atomics.
# usage: tools/sanitize.sh TAG     (GPU box; summaries under gpurun_out/TAG_sanitize_*.txt)
TAG=${1:-rX}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, extra env, command...
    local name=$1 tool=$2; shift 2
    timeout 420 $CS --tool $tool --print-limit 40 --error-exitcode 9 "$@" > gpurun_out/${TAG}_sanitize_${name}_${tool}.log 2>&1
    local rc=$?
    { echo "== $name / $tool: exit $rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|images with VPs|passed|failed" gpurun_out/${TAG}_sanitize_${name}_${tool}.log | sort | uniq -c | head -20; } \
        >> gpurun_out/${TAG}_sanitize_summary.txt
}
rm -f gpurun_out/${TAG}_sanitize_summary.txt
# memcheck: the default path (device-driven EM loop: CUDA graph of conditional WHILE nodes) and the large-image paths
run pipe12 memcheck python tools/run_once.py --config 2 --images 12
run em1600 memcheck python tools/run_em_once.py --n 1600
# racecheck needs ordinary launches (it dies on graphs whose kernels set conditional handles): host-driven loop,
# the same kernels; N = 1600 / 3200 take the 2- and 4-CTA cluster split of the weight-matrix kernel (DSMEM reduction)
export NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=4096
VPK_EM_HOST_LOOP=1 run pipe12_hostloop racecheck python tools/run_once.py --config 2 --images 12
VPK_EM_HOST_LOOP=1 run em1600_hostloop racecheck python tools/run_em_once.py --n 1600
VPK_EM_HOST_LOOP=1 run em3200_hostloop racecheck python tools/run_em_once.py --n 3200 --num-iter 3
# the persistent cluster-per-image kernel (mbarrier ring carried across images, DSMEM mirror, work queue)
VPK_EM_MODE=fused run pipe12_fused racecheck python tools/run_once.py --config 2 --images 12
VPK_EM_MODE=fused run pipe12_fused memcheck python tools/run_once.py --config 2 --images 12
cat gpurun_out/${TAG}_sanitize_summary.txt

#!/bin/bash
# Run under gpurun (1 GPU).  Writes raw captures to gpurun_out/; the summaries that are
# judged are copied (by hand / profiles/summarize.py) into profiles/.
#   profiles/run_ncu.sh <tag> <workload> [n]
set -u
TAG=${1:-r01}; WL=${2:-config3}; N=${3:-0}
mkdir -p gpurun_out
NARG=""; [ "$N" != "0" ] && NARG="--n $N"
# (1) every launch of ~2 steps with its device time (cold cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 450 --csv \
    --log-file gpurun_out/launches_${TAG}_${WL}.csv \
    python bench.py --workload $WL $NARG --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) full metric set of the walk kernels and one sort pass
ncu --set full --clock-control none --import-source on \
    -k regex:'list3_coop|list1_coop|coll_coop|list2_warp|list_kernel|rs_onesweep|box_extents|heavy_step' -s 300 -c 110 \
    -o gpurun_out/prof_${TAG}_${WL} \
    python bench.py --workload $WL $NARG --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/

#!/bin/bash
# Run under gpurun (1 GPU).  Raw captures go to gpurun_out/ (scratch); profiles/summarize.py turns
# them into the committed summaries.   profiles/run_ncu.sh <tag> <spec> [full-kernel-regex] [count]
#   spec: tests/perf_probe.py workload spec, e.g. config3:10000000 or uniform:10000000:f64
set -u
TAG=${1:-r01}; SPEC=${2:-config3:10000000}
REGEX=${3:-'list13_coop|coll_topdown|coll_compact|list2_masked|list13_unstage|heavy_map|list_kernel|rs_onesweep|box_extents_own|permute_kernel|make_keys|leaf_fixup_kernel|level_restrict'}
COUNT=${4:-120}
NAME=$(echo "$SPEC" | tr ':' '_')
mkdir -p gpurun_out
# (1) every launch of ONE warm step (tests/ncu_driver.py brackets it with cudaProfilerStart/Stop)
#     with its device time; serialised: compare SHARES with the bench's live numbers
#     (+ sm__cycles_active avg / max: an SM busy far longer than the average is a tail to fix)
ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}_${NAME}.csv \
    python tests/ncu_driver.py $SPEC > gpurun_out/ncu_launches_${TAG}.log 2>&1
# (2) full metric set of the top kernels of the same warm step
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"$REGEX" -c $COUNT -f -o gpurun_out/prof_${TAG}_${NAME} \
    python tests/ncu_driver.py $SPEC > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_${NAME}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${NAME}_raw.csv 2>/dev/null
# the report itself only travels back when it is small (gpurun_out/ is capped at 64 MiB)
SZ=$(stat -c %s gpurun_out/prof_${TAG}_${NAME}.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/prof_${TAG}_${NAME}.ncu-rep; fi
ls -la gpurun_out/

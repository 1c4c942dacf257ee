#!/bin/bash
# Profiling recipe (B200_PROFILING.md).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- bash profiles/run_ncu.sh <tag> [workload] [kernel-regex]
# Each ncu pass follows a plain run of the same command that exited 0.
set -u
TAG=${1:-r01}
W=${2:-lj1m}
KERNEL=${3:-nl_force_kernel}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload $W --steps 6 --warmup 3"
$CMD > $OUT/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_${W}_$TAG.csv $CMD > $OUT/ncu_launches_$W.log 2>&1
$CMD > $OUT/plain2_$W.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 4 -c 2 \
    -o $OUT/prof_${W}_$TAG -f $CMD > $OUT/ncu_full_$W.log 2>&1
ls -la $OUT

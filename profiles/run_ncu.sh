#!/bin/bash
# Profiling recipe (B200_PROFILING.md).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- bash profiles/run_ncu.sh <tag>
# Each ncu pass follows a plain run of the same command that exited 0.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
prof() {  # workload  kernel-regex
  local W=$1 K=$2
  local CMD="python bench.py --workload $W --steps 12 --warmup 3 --quick"
  timeout 200 $CMD > $OUT/plain_$W.log 2>&1 &&
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/launches_${W}_$TAG.csv $CMD > $OUT/ncu_launches_$W.log 2>&1
  timeout 200 $CMD > $OUT/plain2_$W.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 \
      -o $OUT/prof_${W}_$TAG -f $CMD > $OUT/ncu_full_$W.log 2>&1
}
prof lj1m nl_force_kernel
prof lj32k lj_allpairs_sym_kernel
# the list build: the first launch of a run is a real rebuild (later ones return at the gate)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nl_list_kernel -c 1 \
    -o $OUT/prof_list_$TAG -f python bench.py --workload lj1m --steps 3 --warmup 3 --quick > $OUT/ncu_full_list.log 2>&1
ls -la $OUT | tail -12

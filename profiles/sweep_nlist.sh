#!/bin/bash
# Tuning sweep of the neighbour-list path (historic, r01c): stencil half width (TMD_NL_H) x lanes per atom (TMD_NL_G, fixed at 4 since r01p)
# x resident blocks (TMD_NL_MINB).   usage: sweep_nlist.sh "H list" "G list" "MINB list"
OUT=gpurun_out/sweep_nlist.txt
: > $OUT
for h in ${1:-1}; do for g in ${2:-4}; do for b in ${3:-6}; do for k in ${4:-0.3}; do
  echo -n "H=$h G=$g MINB=$b skin=$k " >> $OUT
  TMD_NL_H=$h TMD_NL_G=$g TMD_NL_MINB=$b timeout 120 python bench.py --workload lj1m --steps 60 --warmup 5 --quick --skin $k 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms_per_step=%.4f force_eval_ms=%.4f value=%.3e launches=%d nlist=%s' % (d['ms_per_step'], d['roofline']['launch_ms'], d['value'], d['gpu_launches'], d['config']['nlist']))
        break
else: print('FAILED')
" >> $OUT
done; done; done; done
cat $OUT

set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload lj1m --steps 30 --warmup 3 --quick"
timeout 200 $CMD > $OUT/plain_lj1m.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02o_launches_lj1m.csv $CMD > $OUT/ncu_launches_lj1m.log 2>&1
echo "ncu rc=$?"

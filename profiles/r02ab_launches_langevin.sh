set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload lj1m_langevin --steps 30 --warmup 3 --quick"
TMD_BENCH_CONFIGS=0 timeout 200 $CMD > $OUT/plain_langevin.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/r02ab_launches_langevin.csv $CMD > $OUT/ncu_launches_langevin.log 2>&1
echo "ncu rc=$?"

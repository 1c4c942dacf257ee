set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02a_gputests.log 2>&1; echo "pytest rc=$?" 
tail -5 $OUT/r02a_gputests.log
timeout 300 python bench.py --workload lj1m > $OUT/r02a_bench_lj1m.json 2> $OUT/r02a_bench_lj1m.err; echo "bench rc=$?"
tail -c 1500 $OUT/r02a_bench_lj1m.json
CMD="python bench.py --workload lj1m --steps 30 --warmup 3 --quick"
timeout 200 $CMD > $OUT/plain_lj1m.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02a_launches_lj1m.csv $CMD > $OUT/ncu_launches_lj1m.log 2>&1
echo "ncu rc=$?"

set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload replicas --steps 6 --warmup 3 --quick"
timeout 200 $CMD > $OUT/plain_replicas.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:inertia_doublewell -s 3 -c 1 -o $OUT/r02p_prof_replicas -f $CMD > $OUT/ncu_full_replicas.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_full_replicas.log
CMD2="python bench.py --workload lj32k --steps 6 --warmup 3 --quick"
timeout 200 $CMD2 > $OUT/plain_lj32k.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02p_launches_lj32k.csv $CMD2 > $OUT/ncu_l32.log 2>&1
echo "ncu2 rc=$?"

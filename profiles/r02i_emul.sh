set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python profiles/emulate_slabrun.py 8 30"
timeout 300 $CMD > $OUT/r02i_emul_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r02i_launches_slabrun8.csv $CMD > $OUT/r02i_emul_ncu.log 2>&1
echo rc=$?; cat $OUT/r02i_emul_plain.log | tail -3

set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python profiles/emulate_slabrun.py 8 10"
timeout 300 $CMD > $OUT/r02m_emul_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:nl_force_kernel -s 60 -c 2 -o $OUT/r02m_prof_slab8_traversal -f $CMD > $OUT/r02m_emul_ncu.log 2>&1
echo rc=$?; tail -2 $OUT/r02m_emul_ncu.log

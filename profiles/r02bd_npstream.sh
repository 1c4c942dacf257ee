set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m pytest tests/test_npy_normal.py -m gpu -q > $OUT/r02bd_npy_tests.log 2>&1; echo "npy tests rc=$?"; tail -6 $OUT/r02bd_npy_tests.log
TMD_PROF_ONLY_GENERATOR=1 timeout 60 python profiles/r02ba_npstream.py > $OUT/r02bd_npstream.json 2> $OUT/r02bd_npstream.err; echo "prof rc=$?"; cat $OUT/r02bd_npstream.json; tail -3 $OUT/r02bd_npstream.err

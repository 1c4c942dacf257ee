set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_npy_normal.py -m gpu -q > $OUT/r02ba_npy_tests.log 2>&1; echo "npy tests rc=$?"; tail -15 $OUT/r02ba_npy_tests.log
timeout 400 python -m pytest tests -m gpu -q --ignore=tests/test_npy_normal.py > $OUT/r02ba_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -8 $OUT/r02ba_gputests.log
timeout 200 python profiles/r02ba_npstream.py > $OUT/r02ba_npstream.json 2> $OUT/r02ba_npstream.err; echo "prof rc=$?"; cat $OUT/r02ba_npstream.json; tail -3 $OUT/r02ba_npstream.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02ba_npstream_launches.csv python profiles/r02ba_npstream_launches.py > $OUT/r02ba_ncu.log 2>&1; echo "ncu rc=$?"; grep -c np_ $OUT/r02ba_npstream_launches.csv

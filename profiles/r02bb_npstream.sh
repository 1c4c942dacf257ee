set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_npy_normal.py -m gpu -q > $OUT/r02bb_npy_tests.log 2>&1; echo "npy tests rc=$?"; tail -12 $OUT/r02bb_npy_tests.log
TMD_PROF_SKIP_HOST=1 timeout 200 python profiles/r02ba_npstream.py > $OUT/r02bb_npstream.json 2> $OUT/r02bb_npstream.err; echo "prof rc=$?"; cat $OUT/r02bb_npstream.json; tail -3 $OUT/r02bb_npstream.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02bb_npstream_launches.csv python profiles/r02ba_npstream_launches.py > $OUT/r02bb_ncu.log 2>&1; echo "ncu rc=$?"; grep -c np_ $OUT/r02bb_npstream_launches.csv
timeout 200 python bench.py --workload lj1m_langevin_default_rng --quick --steps 20 > $OUT/r02bb_bench_lj1m_langevin_default_rng.json 2> $OUT/r02bb_bench.err; echo "bench rc=$?"; cat $OUT/r02bb_bench_lj1m_langevin_default_rng.json | cut -c1-900; tail -3 $OUT/r02bb_bench.err

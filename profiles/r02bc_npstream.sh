set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_npy_normal.py -m gpu -q > $OUT/r02bc_npy_tests.log 2>&1; echo "npy tests rc=$?"; tail -12 $OUT/r02bc_npy_tests.log
timeout 300 python -m pytest tests -m gpu -q --ignore=tests/test_npy_normal.py > $OUT/r02bc_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/r02bc_gputests.log
TMD_PROF_SKIP_HOST=1 timeout 200 python profiles/r02ba_npstream.py > $OUT/r02bc_npstream.json 2> $OUT/r02bc_npstream.err; echo "prof rc=$?"; cat $OUT/r02bc_npstream.json; tail -3 $OUT/r02bc_npstream.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02bc_npstream_launches.csv python profiles/r02ba_npstream_launches.py > $OUT/r02bc_ncu.log 2>&1; echo "ncu rc=$?"; grep -c np_ $OUT/r02bc_npstream_launches.csv
timeout 200 python bench.py --workload lj1m_langevin_default_rng --quick --steps 20 > $OUT/r02bc_bench_lj1m_langevin_default_rng.json 2> $OUT/r02bc_bench.err; echo "bench rc=$?"; cat $OUT/r02bc_bench_lj1m_langevin_default_rng.json | cut -c1-900; tail -3 $OUT/r02bc_bench.err

set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gates.py -m gpu -q -k "two_real_ranks or ipc" > $OUT/r02z_2rank_tests.log 2>&1; echo "2-rank tests rc=$?"; tail -6 $OUT/r02z_2rank_tests.log

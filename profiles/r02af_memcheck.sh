set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_slabrun.py -m gpu -q -k "two_types_and_masses" > $OUT/r02af_plain.log 2>&1; echo "plain rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_slabrun.py -m gpu -q -x -k "two_types_and_masses" > $OUT/r02af_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" $OUT/r02af_memcheck.log | head -20

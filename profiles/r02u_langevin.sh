set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_slabrun.py -m gpu -x -q 2>&1 | tail -15

set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_slabrun.py -m gpu -x -q 2>&1 | tail -3
for G in eager graph; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py slabrun 24 40 $G vv > $OUT/r02f_check_$G.log 2>&1; echo "check $G rc=$?"; grep MULTI_GPU_CHECK $OUT/r02f_check_$G.log || tail -25 $OUT/r02f_check_$G.log
done

set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_slabrun.py tests/test_gpu_resident.py -m gpu -x -q 2>&1 | tail -3
for G in eager graph; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py slabrun 24 41 $G vv > $OUT/r02t_check_$G.log 2>&1; echo "check $G rc=$?"; grep MULTI_GPU_CHECK $OUT/r02t_check_$G.log || tail -25 $OUT/r02t_check_$G.log
done
TMD_BENCH_CONFIGS=0 bash profiles/r02g_bench_g2.sh 2

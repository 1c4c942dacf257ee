set -u
OUT=gpurun_out; mkdir -p $OUT
for G in eager graph; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py slabrun 24 41 $G vv > $OUT/r02j_check_$G.log 2>&1; echo "check $G rc=$?"; grep MULTI_GPU_CHECK $OUT/r02j_check_$G.log || tail -25 $OUT/r02j_check_$G.log
done
bash profiles/r02g_bench_g2.sh 2

set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
for G in graph eager; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 profiles/profile_slabrun.py 30 $G > $OUT/r02l_prof_g${N}_$G.log 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $OUT/r02l_prof_g${N}_$G.log | tail -25
done

set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gates.py -m gpu -q > $OUT/r02c_gates.log 2>&1; echo "gates rc=$?"
tail -12 $OUT/r02c_gates.log
timeout 120 profiles/microbench/tma_gather4 > $OUT/r02c_tma_gather4.txt 2>&1; echo "tma rc=$?"; cat $OUT/r02c_tma_gather4.txt
CMD="python bench.py --workload lj1m --steps 12 --warmup 3 --quick"
timeout 200 $CMD > $OUT/plain_lj1m.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:nl_force_kernel -s 6 -c 2 \
   --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active,l1tex__f_wavefronts.sum,l1tex__m_xbar2l1tex_read_sectors.sum \
   -o $OUT/r02c_prof_lj1m_epi -f $CMD > $OUT/ncu_full_lj1m.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_full_lj1m.log

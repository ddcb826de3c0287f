# developer aid: radix-16 sub-transforms in the spectral kernels against the 8 x 8 x 4 ones
python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral or expr or hack" > gpurun_out/r2x_pytest_spec.log 2>&1
tail -n 3 gpurun_out/r2x_pytest_spec.log
bash tools/gpu_spec_ab.sh r2x r8 base r8 base > /dev/null 2>&1
cat gpurun_out/specab_r2x.txt | cut -c1-140
for sl in 1 3; do echo "== slots $sl"; TSB_SPEC_OVERLAP=$sl python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-140; done
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:spectral -s 8 -c 2 python tools/cfg5_bench.py 65536 64 8 nobuf 2>&1 | grep -v "^==PROF\|^cfg5" | cut -c1-120 > gpurun_out/r2x_ncu_metrics.txt
cat gpurun_out/r2x_ncu_metrics.txt

# developer aid: spectral kernels after the generator/first-pass fusion — tests, timing, data-pipe metrics
python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral or expr or hack" > gpurun_out/r2z_pytest_spec.log 2>&1
tail -n 3 gpurun_out/r2z_pytest_spec.log
python tools/cfg5_bench.py 65536 64 24 | cut -c1-150
for sl in 1 3; do echo "== slots $sl"; TSB_SPEC_OVERLAP=$sl python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-150; done
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:spectral -s 8 -c 2 python tools/cfg5_bench.py 65536 64 8 nobuf 2>&1 | grep -v "^==\|^cfg5\|Warning\|Section\|---" | cut -c1-120 > gpurun_out/r2z_ncu_metrics.txt
cat gpurun_out/r2z_ncu_metrics.txt

# round-2 evidence, final refresh (fp32 register tail, generalised tail, e2e batch cap): suite + smoke + bench, fp32 counters
set -x
T=r2g
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 600 gpurun_out/${T}_bench_reference.json
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed"
ncu --metrics $M --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/${T}_ncu_flops_f32.csv python tools/quick_bench.py 1024 1480 60 f32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -f -o gpurun_out/${T}_ncu_full_pv1024_f32 python tools/quick_bench.py 1024 1480 10 f32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_f64.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/${T}_bench_under_ncu.log 2>&1
ls -la gpurun_out | grep ${T}

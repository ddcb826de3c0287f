# round-2 evidence on HEAD: GPU suite, smoke, contract bench, ncu captures of the final chain kernels.  usage: bash tools/gpu_r2p.sh TAG
tag=${1:-r2p}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json
python tools/quick_bench.py 1024 740 20 f64 > gpurun_out/${tag}_quick.txt 2>&1
python tools/quick_bench.py 1024 2368 20 f32 >> gpurun_out/${tag}_quick.txt 2>&1
cat gpurun_out/${tag}_quick.txt
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/${tag}_ncu_flops_f64.csv python tools/quick_bench.py 1024 740 60 f64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -o gpurun_out/${tag}_ncu_full_pv1024_f64 python tools/quick_bench.py 1024 740 10 f64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -o gpurun_out/${tag}_ncu_full_pv1024_f32 python tools/quick_bench.py 1024 2368 10 f32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_f64.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12

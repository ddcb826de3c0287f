set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "teacher_forced" 2>&1 | tail -5
python tools/cfg3_teacher_forced.py --seconds 3600 --every 4096 --window 128 --procs 8 > gpurun_out/r2_cfg3_teacher_forced_3600s.json 2> gpurun_out/r2g_tf.err; tail -c 600 gpurun_out/r2_cfg3_teacher_forced_3600s.json; tail -3 gpurun_out/r2g_tf.err
python tools/long_render_check.py --seconds 600 --ratio 1.5 > gpurun_out/r2_long_render_600s_x1p5.json 2>/dev/null
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_ncu_flops_f64.csv python tools/quick_bench.py 1024 740 60 f64 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_ncu_flops_f32.csv python tools/quick_bench.py 1024 2368 60 f32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_full_pv1024_f64 python tools/quick_bench.py 1024 740 10 f64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_full_pv1024_f32 python tools/quick_bench.py 1024 2368 10 f32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_f64.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/r2g_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12

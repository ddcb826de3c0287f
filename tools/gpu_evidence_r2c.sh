# round-2 evidence, last batch (register overlap tail, fmax-free PV math, radix-16 spectral kernels):
# GPU suite + smoke + bench on HEAD, per-opcode flop counts of the bench launch shape, ncu --set full of the chain and
# spectral kernels, launch list of the bench.   usage: bash tools/gpu_evidence_r2c.sh
set -x
T=r2f
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
python tools/quick_bench.py 1024 1184 20 f64 > gpurun_out/${T}_quick.txt 2>&1; python tools/quick_bench.py 1024 2368 20 f32 >> gpurun_out/${T}_quick.txt 2>&1; cat gpurun_out/${T}_quick.txt
python tools/cfg5_bench.py > gpurun_out/${T}_cfg5.txt 2>&1; cat gpurun_out/${T}_cfg5.txt
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed"
ncu --metrics $M --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/${T}_ncu_flops_f64.csv python tools/quick_bench.py 1024 1184 60 f64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -f -o gpurun_out/${T}_ncu_full_pv1024_f64 python tools/quick_bench.py 1024 1184 10 f64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectral -s 8 -c 2 -f -o gpurun_out/${T}_ncu_full_cfg5 python tools/cfg5_bench.py 65536 64 8 nobuf > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_f64.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/${T}_bench_under_ncu.log 2>&1
ls -la gpurun_out | grep ${T}

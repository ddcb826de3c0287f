# developer aid (round 2, session 5): old-vs-new timings of the chain kernel (fmax-free fp64 math) and the spectral
# kernels (fast-path generation, tile-size variants), plus the parity tests they touch
bash tools/gpu_spec_ab.sh r2v old base a128 b128 ab128 uo2 uo4 base old > /dev/null 2>&1
bash tools/gpu_var.sh r2v f64 740 old > /dev/null 2>&1
python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral or expr or hack" > gpurun_out/r2v_pytest_spec.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pv_vs_oracle or golden or full_size or segments or timepar_matches" > gpurun_out/r2v_pytest_chain.log 2>&1
cat gpurun_out/specab_r2v.txt gpurun_out/var_r2v.txt; tail -3 gpurun_out/r2v_pytest_spec.log gpurun_out/r2v_pytest_chain.log

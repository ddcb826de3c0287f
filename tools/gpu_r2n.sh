# round-2 third batch: spectral tests with the table-reciprocal generator, cfg5 timing, the exchange probe
set -x
python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral" > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
python tools/cfg5_bench.py > gpurun_out/r2n_cfg5.txt 2>&1; cat gpurun_out/r2n_cfg5.txt
build/tools/xpose_probe > gpurun_out/r2n_xpose_probe.txt 2>&1; cat gpurun_out/r2n_xpose_probe.txt

python tools/quick_bench.py 1024 1480 20 f32 | grep "^pv"
python tools/quick_bench.py 1024 1480 20 f32 gain | grep "^gain"
python tools/quick_bench.py 1024 1184 20 f64 gain | grep "^gain"
python -m pytest tests/test_gpu_parity.py tests/test_facade.py -m gpu -x -q > gpurun_out/r3i_pytest.log 2>&1; tail -n 3 gpurun_out/r3i_pytest.log

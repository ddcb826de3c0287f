# developer aid: register-resident overlap tail (fp64, N=1024) against the shared-memory ring
bash tools/gpu_var.sh r3c f64 740 nort > /dev/null 2>&1
cat gpurun_out/var_r3c.txt
python tools/quick_bench.py 1024 740 20 f64o32 2>&1 | tail -2
TSB200_LIB=$PWD/build/var/nort/libtsb200.so python tools/quick_bench.py 1024 740 20 f64o32 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py tests/test_facade.py -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; tail -n 3 gpurun_out/r3c_pytest.log

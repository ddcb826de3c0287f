# developer aid: parity tests of the chain kernels + quick timings.  usage: bash tools/gpu_quick.sh TAG
tag=$1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/quick_${tag}_pytest.log 2>&1; tail -3 gpurun_out/quick_${tag}_pytest.log
python tools/quick_bench.py 1024 740,1480 20 f64 > gpurun_out/quick_${tag}.txt 2>&1
python tools/quick_bench.py 1024 888 20 f64 >> gpurun_out/quick_${tag}.txt 2>&1
python tools/quick_bench.py 2048 444 20 f64 >> gpurun_out/quick_${tag}.txt 2>&1
python tools/quick_bench.py 512 1776 20 f64 >> gpurun_out/quick_${tag}.txt 2>&1
python tools/quick_bench.py 1024 2368 20 f32 >> gpurun_out/quick_${tag}.txt 2>&1
cat gpurun_out/quick_${tag}.txt

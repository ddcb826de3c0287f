# round-2 evidence, last batch: exchange probe (shuffle vs shared memory), the other frame sizes and cfg2 on the final kernels
set -x
build/tools/xpose_probe > gpurun_out/r2q_xpose_probe.txt 2>&1; cat gpurun_out/r2q_xpose_probe.txt
out=gpurun_out/r2q_sizes.txt; : > $out
python tools/quick_bench.py 64 2368 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 256 2368 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 512 2368 20 f32 >> $out 2>&1
python tools/quick_bench.py 512 1628 20 f64 >> $out 2>&1
python tools/quick_bench.py 2048 1184 20 f32 >> $out 2>&1
python tools/quick_bench.py 2048 444 20 f64 >> $out 2>&1
python tools/quick_bench.py 4096 592 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 8192 296 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 2048 1024,2048 60 f32,f64 gain >> $out 2>&1
cat $out
python tools/fft_bench.py > gpurun_out/r2q_fft_bench.txt 2>&1; tail -12 gpurun_out/r2q_fft_bench.txt

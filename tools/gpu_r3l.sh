# developer aid: the other frame sizes and cfg2 on the final kernels (lane counts = whole waves of each kernel's residency)
out=gpurun_out/r3l_sizes.txt; : > $out
python tools/quick_bench.py 64 2368 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 256 2368 20 f32,f64 >> $out 2>&1
TSB200_LIB=$PWD/build/var/n8_rt/libtsb200.so python tools/quick_bench.py 256 2368 20 f64 | sed 's/^pv/pv[n8_rt variant]/' >> $out 2>&1
python tools/quick_bench.py 512 2368 20 f32 >> $out 2>&1
python tools/quick_bench.py 512 2368 20 f64 >> $out 2>&1
python tools/quick_bench.py 2048 1184 20 f32 >> $out 2>&1
python tools/quick_bench.py 2048 592 20 f64 >> $out 2>&1
python tools/quick_bench.py 4096 592 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 8192 296 20 f32,f64 >> $out 2>&1
python tools/quick_bench.py 2048 1024,2048 60 f32 gain >> $out 2>&1
python tools/quick_bench.py 2048 1024,1184 60 f64 gain >> $out 2>&1
grep "^pv\|^gain" $out | cut -c1-150

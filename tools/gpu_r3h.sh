python tools/quick_bench.py 1024 2368 20 f32 | grep "^pv"
TSB200_LIB=$PWD/build/var/f32_96/libtsb200.so python tools/quick_bench.py 1024 1480 20 f32 | grep "^pv"
TSB200_LIB=$PWD/build/var/rt32_96/libtsb200.so python tools/quick_bench.py 1024 1480 20 f32 | grep "^pv"
TSB200_LIB=$PWD/build/var/rt32_80/libtsb200.so python tools/quick_bench.py 1024 1776 20 f32 | grep "^pv"
for i in 1 2 3; do python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-100; done

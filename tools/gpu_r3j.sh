# developer aid: register overlap tail at other frame sizes (build/var/rtA) against the regular build
for v in base rtA; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo "== $v"
  for n in 512 2048 4096; do TSB200_LIB=$PWD/$lib python tools/quick_bench.py $n 1776 20 f32,f64 pv | grep "^pv" | cut -c1-110; done
  TSB200_LIB=$PWD/$lib python tools/quick_bench.py 2048 1776 20 f32,f64 gain | grep "^gain" | cut -c1-110
  TSB200_LIB=$PWD/$lib python tools/quick_bench.py 1024 1480 20 f32 pv | grep "^pv" | cut -c1-110
  TSB200_LIB=$PWD/$lib python tools/quick_bench.py 1024 1184 20 f64 pv | grep "^pv" | cut -c1-110
done
TSB200_LIB=$PWD/build/var/rt96/libtsb200.so python tools/quick_bench.py 1024 1480 20 f64 pv | grep "^pv" | cut -c1-110
TSB200_LIB=$PWD/build/var/rtA/libtsb200.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oracle or golden or register or out32 or timepar_matches" 2>&1 | tail -n 2

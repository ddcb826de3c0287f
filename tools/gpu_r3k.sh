for v in base n9_nort n9_r128; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo -n "$v: "; TSB200_LIB=$PWD/$lib python tools/quick_bench.py 512 13024 5 f64 pv | grep "^pv" | cut -c1-110
done
python tools/quick_bench.py 256 13024 5 f64,f32 pv | grep "^pv" | cut -c1-110
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -n 2

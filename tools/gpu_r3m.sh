for v in base ug4 ug8; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo -n "$v: "; TSB200_LIB=$PWD/$lib python tools/cfg5_bench.py 65536 64 24 | grep "from-HBM" | cut -c1-120
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "own_hop or register" 2>&1 | tail -n 2

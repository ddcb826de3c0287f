python tools/quick_bench.py 1024 592,1184,740 20 f64 2>&1 | grep "^pv"
TSB200_LIB=$PWD/build/var/nort/libtsb200.so python tools/quick_bench.py 1024 592,740,1480 20 f64 2>&1 | grep "^pv"

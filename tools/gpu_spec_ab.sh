# developer aid: cfg5 timings of library variants (no tests).  usage: bash tools/gpu_spec_ab.sh TAG variant...
tag=$1; shift
out=gpurun_out/specab_$tag.txt; : > $out
for v in "$@"; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  printf "%-14s " $v >> $out
  TSB200_LIB=$PWD/$lib python tools/cfg5_bench.py 65536 64 24 2>&1 | cut -c1-110 >> $out
done
cat $out

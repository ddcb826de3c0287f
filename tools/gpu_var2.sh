# developer aid: usage: bash tools/gpu_var2.sh TAG "quick_bench args" variant...
tag=$1; args=$2; shift 2
out=gpurun_out/var_$tag.txt; : > $out
for v in "$@"; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo "== $v" >> $out
  TSB200_LIB=$PWD/$lib python tools/quick_bench.py $args >> $out 2>&1
done
cat $out
